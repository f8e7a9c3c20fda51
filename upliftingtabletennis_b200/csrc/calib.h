// Camera calibration from table keypoints: the arithmetic of
//   dataprocessing/my_dlt.py:5-161                    (normalised DLT, RQ decomposition)
//   dataprocessing/regress_cameramatrices.py:38-116   (8-parameter fit of fx, fy, t, euler angles by SciPy BFGS on the
//                                                      summed reprojection distance)
// as one host/device header (the CUDA kernels in calib.cu and the host unit test compile the same code).
//
// scipy.optimize.minimize(method='BFGS') is restated, not substituted: _minimize_bfgs with its defaults (gtol 1e-5 on the
// max-norm, maxiter 200 n, inverse-Hessian update with the rhok = 1000 guard), the 2-point finite-difference gradient
// with absolute step sqrt(eps) (approx_derivative: dx = (x + h) - x), _line_search_wolfe12 = MINPACK-2 DCSRCH
// (c1 1e-4, c2 0.9, xtol 1e-14, step range [1e-100, 1e100], 100 trials) falling back to the Nocedal-Wright search with
// cubic/quadratic zoom (line_search_wolfe2, 10 + 10 trials), and the "precision loss" exit when both fail.
// The objective is piecewise smooth (a sum of Euclidean norms), so iterates agree with SciPy's up to the rounding noise
// the finite differences amplify; tests compare stopping points, not bits.
#pragma once
#include <math.h>

#include "lbfgsb4.h"      // TTK_HD, lb_dcstep (MINPACK-2 dcstep)

#define CB_N 8
#ifndef CB_TRACE                 // test hook: the host unit test records the iterates
#define CB_TRACE(k, x, f, alpha)
#endif
#define CB_MAXPTS 13
#define CB_PI 3.141592653589793

struct CalibProblem {
  int n;                        // points in the fit
  double X[CB_MAXPTS][3];       // world points
  double u[CB_MAXPTS][2];       // detections (px)
  double px, py;                // principal point (WIDTH // 2, HEIGHT // 2)
};

struct CalibResult {
  double x[CB_N];
  double f;
  int nit, status;              // status 0: converged, 1: maxiter, 2: precision loss, 3: nan
};

// Rotation.from_euler('xyz', [a, b, c]).as_matrix(): extrinsic rotations about x, then y, then z: R = Rz(c) Ry(b) Rx(a)
TTK_HD static inline void cb_rotation(double a, double b, double c, double R[3][3]) {
  const double sa = sin(a), ca = cos(a), sb = sin(b), cb = cos(b), sc = sin(c), cc = cos(c);
  R[0][0] = cb * cc;
  R[0][1] = sa * sb * cc - ca * sc;
  R[0][2] = ca * sb * cc + sa * sc;
  R[1][0] = cb * sc;
  R[1][1] = sa * sb * sc + ca * cc;
  R[1][2] = ca * sb * sc - sa * cc;
  R[2][0] = -sb;
  R[2][1] = sa * cb;
  R[2][2] = ca * cb;
}

// Rotation.from_matrix(R).as_euler('xyz') followed by the wrap of regress_cameramatrices.py:92.  Returns 0 when the
// determinant is not positive (SciPy raises ValueError and the reference falls back to zero angles, :86-89).
TTK_HD static inline int cb_euler_xyz(const double R[3][3], double* ang) {
  const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                     R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  int ok = 1;
  if (!(det > 0.0)) {
    ang[0] = ang[1] = ang[2] = 0.0;
    ok = 0;
  } else {
    double sb = -R[2][0];
    sb = sb > 1.0 ? 1.0 : (sb < -1.0 ? -1.0 : sb);
    ang[1] = asin(sb);
    if (fabs(sb) < 1.0 - 1e-14) {
      ang[0] = atan2(R[2][1], R[2][2]);
      ang[2] = atan2(R[1][0], R[0][0]);
    } else {                       // gimbal lock: SciPy sets the third angle to zero
      ang[2] = 0.0;
      ang[0] = atan2(-sb * R[0][1], R[1][1]);
    }
  }
  for (int i = 0; i < 3; ++i) {    // np.mod(x + pi, 2 pi) - pi
    double m = fmod(ang[i] + CB_PI, 2.0 * CB_PI);
    if (m < 0.0) m += 2.0 * CB_PI;
    ang[i] = m - CB_PI;
  }
  return ok;
}

// np.sum over a contiguous float64 vector (pairwise summation: < 8 terms sequential, else 8 accumulators + tail)
TTK_HD static inline double cb_np_sum(const double* a, int n) {
  if (n < 8) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i];
    return s;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) s += a[i];
  return s;
}

// reprojection distance of point i under parameters x = (fx, fy, tx, ty, tz, a, b, c)
TTK_HD static inline double cb_point_error(const CalibProblem* P, const double R[3][3], const double* x, int i) {
  const double* X = P->X[i];
  const double cx = R[0][0] * X[0] + R[0][1] * X[1] + R[0][2] * X[2] + x[2];
  const double cy = R[1][0] * X[0] + R[1][1] * X[1] + R[1][2] * X[2] + x[3];
  const double cz = R[2][0] * X[0] + R[2][1] * X[1] + R[2][2] * X[2] + x[4];
  const double iu = x[0] * cx + P->px * cz, iv = x[1] * cy + P->py * cz;
  const double du = iu / cz - P->u[i][0], dv = iv / cz - P->u[i][1];
  return sqrt(du * du + dv * dv);
}

// opt_func (regress_cameramatrices.py:73-75)
TTK_HD static inline double cb_loss(const CalibProblem* P, const double* x) {
  double R[3][3], r[CB_MAXPTS];
  cb_rotation(x[5], x[6], x[7], R);
  for (int i = 0; i < P->n; ++i) r[i] = cb_point_error(P, R, x, i);
  return cb_np_sum(r, P->n);
}

// approx_derivative(fun, x, method='2-point', abs_step=sqrt(eps), f0=f0)
TTK_HD static inline void cb_grad(const CalibProblem* P, const double* x, double f0, double* g) {
  const double h0 = 1.4901161193847656e-08;
#if defined(__CUDA_ARCH__)
  // the warp runs the optimiser in lock step; lanes 0..7 each evaluate one forward difference
  const int lane = threadIdx.x & 31;
  double gi = 0.0;
  if (lane < CB_N) {
    double xx[CB_N];
    for (int j = 0; j < CB_N; ++j) xx[j] = x[j];
    double h = h0;
    if ((x[lane] + h) - x[lane] == 0.0) h = h0 * (x[lane] >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x[lane]));
    xx[lane] = x[lane] + h;
    gi = (cb_loss(P, xx) - f0) / (xx[lane] - x[lane]);
  }
  for (int i = 0; i < CB_N; ++i) g[i] = __shfl_sync(0xffffffffu, gi, i);
#else
  double xx[CB_N];
  for (int j = 0; j < CB_N; ++j) xx[j] = x[j];
  for (int i = 0; i < CB_N; ++i) {
    double h = h0;
    if ((x[i] + h) - x[i] == 0.0) h = h0 * (x[i] >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x[i]));
    xx[i] = x[i] + h;
    g[i] = (cb_loss(P, xx) - f0) / (xx[i] - x[i]);
    xx[i] = x[i];
  }
#endif
}

TTK_HD static inline double cb_dot(const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < CB_N; ++i) s += a[i] * b[i];
  return s;
}

struct CbLine {                 // phi(s) = f(xk + s pk)
  const CalibProblem* P;
  const double* xk;
  const double* pk;
  double g[CB_N];               // gradient at the last derphi point
};

TTK_HD static inline double cb_phi(const CbLine* L, double s) {
  double x[CB_N];
  for (int i = 0; i < CB_N; ++i) x[i] = L->xk[i] + s * L->pk[i];
  return cb_loss(L->P, x);
}

TTK_HD static inline double cb_derphi(CbLine* L, double s, double f_at_s) {
  double x[CB_N];
  for (int i = 0; i < CB_N; ++i) x[i] = L->xk[i] + s * L->pk[i];
  cb_grad(L->P, x, f_at_s, L->g);
  return cb_dot(L->g, L->pk);
}

// ---- scalar_search_wolfe1: DCSRCH (scipy/optimize/_dcsrch.py) -----------------------------------------------------
// returns 1 and *stp, *phi1 on success (gradient at stp in L->g), 0 on failure
TTK_HD static inline int cb_wolfe1(CbLine* L, double phi0, double old_phi0, double derphi0, double* stp_out, double* phi_out) {
  const double ftol = 1e-4, gtol = 0.9, xtol = 1e-14, stpmin = 1e-100, stpmax = 1e100;
  double alpha1 = 1.0;
  if (derphi0 != 0.0) {
    alpha1 = fmin(1.0, 1.01 * 2.0 * (phi0 - old_phi0) / derphi0);
    if (alpha1 < 0.0) alpha1 = 1.0;
  }
  // START
  if (alpha1 < stpmin || alpha1 > stpmax || derphi0 >= 0.0 || !(alpha1 == alpha1)) return 0;
  int brackt = 0, stage = 1;
  const double finit = phi0, ginit = derphi0, gtest = ftol * ginit;
  double width = stpmax - stpmin, width1 = width / 0.5;
  double stx = 0.0, fx = finit, gx = ginit, sty = 0.0, fy = finit, gy = ginit, stmin = 0.0, stmax = alpha1 + 4.0 * alpha1;
  double stp = alpha1;
  double f = cb_phi(L, stp);
  double g = cb_derphi(L, stp, f);
  for (int it = 1; it < 100; ++it) {
    const double ftest = finit + stp * gtest;
    if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
    int warn = 0, conv = 0;
    if (brackt && (stp <= stmin || stp >= stmax)) warn = 1;
    if (brackt && stmax - stmin <= xtol * stmax) warn = 1;
    if (stp == stpmax && f <= ftest && g <= gtest) warn = 1;
    if (stp == stpmin && (f > ftest || g >= gtest)) warn = 1;
    if (f <= ftest && fabs(g) <= gtol * -ginit) conv = 1;
    if (conv) {
      *stp_out = stp;
      *phi_out = f;
      return 1;
    }
    if (warn) return 0;
    if (stage == 1 && f <= fx && f > ftest) {
      const double fm = f - stp * gtest, gm = g - gtest;
      double fxm = fx - stx * gtest, fym = fy - sty * gtest, gxm = gx - gtest, gym = gy - gtest;
      lb_dcstep(&stx, &fxm, &gxm, &sty, &fym, &gym, &stp, fm, gm, &brackt, stmin, stmax);
      fx = fxm + stx * gtest;
      fy = fym + sty * gtest;
      gx = gxm + gtest;
      gy = gym + gtest;
    } else {
      lb_dcstep(&stx, &fx, &gx, &sty, &fy, &gy, &stp, f, g, &brackt, stmin, stmax);
    }
    if (brackt) {
      if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
      width1 = width;
      width = fabs(sty - stx);
      stmin = fmin(stx, sty);
      stmax = fmax(stx, sty);
    } else {
      stmin = stp + 1.1 * (stp - stx);
      stmax = stp + 4.0 * (stp - stx);
    }
    stp = fmin(fmax(stp, stpmin), stpmax);
    if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
    if (!isfinite(stp)) return 0;
    f = cb_phi(L, stp);
    g = cb_derphi(L, stp, f);
  }
  return 0;       // maxiter trials without convergence
}

// ---- scalar_search_wolfe2 (scipy/optimize/_linesearch.py) ---------------------------------------------------------
TTK_HD static inline int cb_cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc, double* xmin) {
  const double C = fpa, db = b - a, dc = c - a;
  const double denom = (db * dc) * (db * dc) * (db - dc);
  if (denom == 0.0 || !isfinite(denom)) return 0;
  const double r0 = fb - fa - C * db, r1 = fc - fa - C * dc;
  double A = dc * dc * r0 + (-(db * db)) * r1;
  double B = (-(dc * dc * dc)) * r0 + (db * db * db) * r1;
  A /= denom;
  B /= denom;
  const double radical = B * B - 3.0 * A * C;
  if (!(radical >= 0.0) || A == 0.0) return 0;
  *xmin = a + (-B + sqrt(radical)) / (3.0 * A);
  return isfinite(*xmin) ? 1 : 0;
}

TTK_HD static inline int cb_quadmin(double a, double fa, double fpa, double b, double fb, double* xmin) {
  const double D = fa, C = fpa, db = b - a * 1.0;
  if (db * db == 0.0) return 0;
  const double B = (fb - D - C * db) / (db * db);
  if (B == 0.0 || !isfinite(B)) return 0;
  *xmin = a - C / (2.0 * B);
  return isfinite(*xmin) ? 1 : 0;
}

// returns 1 on success with a_star, val_star (gradient at a_star in L->g)
TTK_HD static inline int cb_zoom(CbLine* L, double a_lo, double a_hi, double phi_lo, double phi_hi, double derphi_lo, double phi0,
                                 double derphi0, double c1, double c2, double* a_star, double* val_star) {
  const double delta1 = 0.2, delta2 = 0.1;
  double phi_rec = phi0, a_rec = 0.0;
  for (int i = 0;; ++i) {
    const double dalpha = a_hi - a_lo;
    double a, b;
    if (dalpha < 0.0) a = a_hi, b = a_lo; else a = a_lo, b = a_hi;
    double a_j = 0.0;
    int have = 0;
    if (i > 0) {
      const double cchk = delta1 * dalpha;
      have = cb_cubicmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, a_rec, phi_rec, &a_j);
      if (have && (a_j > b - cchk || a_j < a + cchk)) have = 0;
    }
    if (!have) {
      const double qchk = delta2 * dalpha;
      have = cb_quadmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, &a_j);
      if (!have || a_j > b - qchk || a_j < a + qchk) a_j = a_lo + 0.5 * dalpha;
    }
    const double phi_aj = cb_phi(L, a_j);
    if (phi_aj > phi0 + c1 * a_j * derphi0 || phi_aj >= phi_lo) {
      phi_rec = phi_hi;
      a_rec = a_hi;
      a_hi = a_j;
      phi_hi = phi_aj;
    } else {
      const double derphi_aj = cb_derphi(L, a_j, phi_aj);
      if (fabs(derphi_aj) <= -c2 * derphi0) {
        *a_star = a_j;
        *val_star = phi_aj;
        return 1;
      }
      if (derphi_aj * (a_hi - a_lo) >= 0.0) {
        phi_rec = phi_hi;
        a_rec = a_hi;
        a_hi = a_lo;
        phi_hi = phi_lo;
      } else {
        phi_rec = phi_lo;
        a_rec = a_lo;
      }
      a_lo = a_j;
      phi_lo = phi_aj;
      derphi_lo = derphi_aj;
    }
    if (i + 1 > 10) return 0;
  }
}

// returns 0: failed, 1: converged (gradient in L->g), 2: step accepted without a gradient (10 doublings, no bracket)
TTK_HD static inline int cb_wolfe2(CbLine* L, double phi0, double old_phi0, double derphi0, double* a_star, double* val_star) {
  const double c1 = 1e-4, c2 = 0.9, amax = 1e100;
  double alpha0 = 0.0, alpha1 = 1.0;
  if (derphi0 != 0.0) alpha1 = fmin(1.0, 1.01 * 2.0 * (phi0 - old_phi0) / derphi0);
  if (alpha1 < 0.0) alpha1 = 1.0;
  alpha1 = fmin(alpha1, amax);
  double phi_a1 = cb_phi(L, alpha1), phi_a0 = phi0, derphi_a0 = derphi0;
  for (int i = 0; i < 10; ++i) {
    if (alpha1 == 0.0 || alpha0 > amax) return 0;
    if (phi_a1 > phi0 + c1 * alpha1 * derphi0 || (phi_a1 >= phi_a0 && i > 0))
      return cb_zoom(L, alpha0, alpha1, phi_a0, phi_a1, derphi_a0, phi0, derphi0, c1, c2, a_star, val_star);
    const double derphi_a1 = cb_derphi(L, alpha1, phi_a1);
    if (fabs(derphi_a1) <= -c2 * derphi0) {
      *a_star = alpha1;
      *val_star = phi_a1;
      return 1;
    }
    if (derphi_a1 >= 0.0) return cb_zoom(L, alpha1, alpha0, phi_a1, phi_a0, derphi_a1, phi0, derphi0, c1, c2, a_star, val_star);
    const double alpha2 = fmin(2.0 * alpha1, amax);
    alpha0 = alpha1;
    alpha1 = alpha2;
    phi_a0 = phi_a1;
    phi_a1 = cb_phi(L, alpha1);
    derphi_a0 = derphi_a1;
  }
  *a_star = alpha1;
  *val_star = phi_a1;
  return 2;
}

// ---- _minimize_bfgs (scipy/optimize/_optimize.py) -----------------------------------------------------------------
TTK_HD static inline CalibResult cb_bfgs(const CalibProblem* P, const double* x0) {
  CalibResult res;
  double xk[CB_N], gfk[CB_N], pk[CB_N], sk[CB_N], yk[CB_N], H[CB_N][CB_N], T1[CB_N][CB_N];
  for (int i = 0; i < CB_N; ++i) xk[i] = x0[i];
  double old_fval = cb_loss(P, xk);
  cb_grad(P, xk, old_fval, gfk);
  for (int i = 0; i < CB_N; ++i)
    for (int j = 0; j < CB_N; ++j) H[i][j] = i == j ? 1.0 : 0.0;
  double old_old_fval = old_fval + sqrt(cb_dot(gfk, gfk)) / 2.0;
  int k = 0, warnflag = 0;
  const int maxiter = CB_N * 200;
  double gnorm = 0.0;
  for (int i = 0; i < CB_N; ++i) gnorm = fmax(gnorm, fabs(gfk[i]));
  bool nan_g = false;
  for (int i = 0; i < CB_N; ++i) nan_g |= !(gfk[i] == gfk[i]);
  if (nan_g) gnorm = NAN;
  while (gnorm > 1e-5 && k < maxiter) {
    for (int i = 0; i < CB_N; ++i) {
      double s = 0.0;
      for (int j = 0; j < CB_N; ++j) s += H[i][j] * gfk[j];
      pk[i] = -s;
    }
    CbLine L;
    L.P = P;
    L.xk = xk;
    L.pk = pk;
    const double derphi0 = cb_dot(gfk, pk);
    double alpha = 0.0, fnew = 0.0;
    int have_g = 1;
    int ok = cb_wolfe1(&L, old_fval, old_old_fval, derphi0, &alpha, &fnew);
    if (!ok) {
      const int r = cb_wolfe2(&L, old_fval, old_old_fval, derphi0, &alpha, &fnew);
      ok = r != 0;
      have_g = r == 1;
    }
    if (!ok) {
      warnflag = 2;
      break;
    }
    old_old_fval = old_fval;
    old_fval = fnew;
    for (int i = 0; i < CB_N; ++i) {
      sk[i] = alpha * pk[i];
      xk[i] = xk[i] + sk[i];
    }
    if (!have_g) cb_grad(P, xk, cb_loss(P, xk), L.g);
    for (int i = 0; i < CB_N; ++i) {
      yk[i] = L.g[i] - gfk[i];
      gfk[i] = L.g[i];
    }
    ++k;
    CB_TRACE(k, xk, old_fval, alpha);
    gnorm = 0.0;
    nan_g = false;
    for (int i = 0; i < CB_N; ++i) {
      gnorm = fmax(gnorm, fabs(gfk[i]));
      nan_g |= !(gfk[i] == gfk[i]);
    }
    if (nan_g) gnorm = NAN;
    if (gnorm <= 1e-5) break;
    double pn = 0.0;
    for (int i = 0; i < CB_N; ++i) pn += pk[i] * pk[i];
    if (alpha * sqrt(pn) <= 0.0) break;
    if (!isfinite(old_fval)) {
      warnflag = 2;
      break;
    }
    const double rhok_inv = cb_dot(yk, sk);
    const double rhok = rhok_inv == 0.0 ? 1000.0 : 1.0 / rhok_inv;
    // Hk = (I - rho s y^T) Hk (I - rho y s^T) + rho s s^T
    for (int i = 0; i < CB_N; ++i)
      for (int j = 0; j < CB_N; ++j) {          // T1 = Hk A2
        double s = 0.0;
        for (int l = 0; l < CB_N; ++l) s += H[i][l] * ((l == j ? 1.0 : 0.0) - yk[l] * sk[j] * rhok);
        T1[i][j] = s;
      }
    for (int i = 0; i < CB_N; ++i)
      for (int j = 0; j < CB_N; ++j) {          // Hk = A1 T1 + rho s s^T
        double s = 0.0;
        for (int l = 0; l < CB_N; ++l) s += ((i == l ? 1.0 : 0.0) - sk[i] * yk[l] * rhok) * T1[l][j];
        H[i][j] = s + rhok * sk[i] * sk[j];
      }
  }
  bool nan_x = false;
  for (int i = 0; i < CB_N; ++i) {
    res.x[i] = xk[i];
    nan_x |= !(xk[i] == xk[i]);
  }
  res.f = old_fval;
  res.nit = k;
  if (warnflag == 2) res.status = 2;
  else if (k >= maxiter) res.status = 1;
  else if (!(gnorm == gnorm) || !(old_fval == old_fval) || nan_x) res.status = 3;
  else res.status = 0;
  return res;
}

// ---- DLT start (dataprocessing/my_dlt.py) ---------------------------------------------------------------------------
// normalize_points: per-axis mean / population std (np.mean, np.std), zero std -> 1e-10
TTK_HD static inline void cb_mean_std(const double* v, int n, int stride, double* mean, double* sd) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i * stride];
  const double m = s / n;
  double q = 0.0;
  for (int i = 0; i < n; ++i) {
    const double d = v[i * stride] - m;
    q += d * d;
  }
  double sdv = sqrt(q / n);
  if (sdv == 0.0) sdv = 1e-10;
  *mean = m;
  *sd = sdv;
}

// right singular vector of the smallest singular value of A (m x 12, m <= 26): one-sided Jacobi (Hestenes)
TTK_HD static inline void cb_null_vector(double (*A)[12], int m, double* v) {
  double V[12][12];
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < 11; ++p)
      for (int q = p + 1; q < 12; ++q) {
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int i = 0; i < m; ++i) {
          al += A[i][p] * A[i][p];
          be += A[i][q] * A[i][q];
          ga += A[i][p] * A[i][q];
        }
        if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
        rotated = 1;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < m; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq;
          A[i][q] = s * ap + c * aq;
        }
        for (int i = 0; i < 12; ++i) {
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  int best = 0;
  double bn = INFINITY;
  for (int j = 0; j < 12; ++j) {
    double nn = 0.0;
    for (int i = 0; i < m; ++i) nn += A[i][j] * A[i][j];
    if (nn < bn) bn = nn, best = j;
  }
  for (int i = 0; i < 12; ++i) v[i] = V[i][best];
}

// dlt_calib: K (3x3, K[2][2] = 1), R (3x3), t (3).  Returns 0 on a degenerate decomposition (K[2][2] == 0).
TTK_HD static inline int cb_dlt(const CalibProblem* P, double K[3][3], double R[3][3], double* t) {
  const int n = P->n;
  double m3[3], s3[3], m2[2], s2[2];
  for (int d = 0; d < 3; ++d) cb_mean_std(&P->X[0][d], n, 3, &m3[d], &s3[d]);
  for (int d = 0; d < 2; ++d) cb_mean_std(&P->u[0][d], n, 2, &m2[d], &s2[d]);
  double A[2 * CB_MAXPTS][12];
  for (int i = 0; i < n; ++i) {
    double Xn[3], xn[2];
    for (int d = 0; d < 3; ++d) Xn[d] = (1.0 / s3[d]) * P->X[i][d] + (-m3[d] / s3[d]);
    for (int d = 0; d < 2; ++d) xn[d] = (1.0 / s2[d]) * P->u[i][d] + (-m2[d] / s2[d]);
    double* r0 = A[2 * i];
    double* r1 = A[2 * i + 1];
    r0[0] = -Xn[0], r0[1] = -Xn[1], r0[2] = -Xn[2], r0[3] = -1.0, r0[4] = r0[5] = r0[6] = r0[7] = 0.0;
    r0[8] = xn[0] * Xn[0], r0[9] = xn[0] * Xn[1], r0[10] = xn[0] * Xn[2], r0[11] = xn[0];
    r1[0] = r1[1] = r1[2] = r1[3] = 0.0, r1[4] = -Xn[0], r1[5] = -Xn[1], r1[6] = -Xn[2], r1[7] = -1.0;
    r1[8] = xn[1] * Xn[0], r1[9] = xn[1] * Xn[1], r1[10] = xn[1] * Xn[2], r1[11] = xn[1];
  }
  double v[12];
  cb_null_vector(A, 2 * n, v);
  // P = inv(T_2d) @ P_norm @ T_3d with T = [diag(1/s) | -m/s; 0 1]
  double Q[3][4], Pm[3][4];
  for (int j = 0; j < 4; ++j) {
    Q[0][j] = s2[0] * v[j] + m2[0] * v[8 + j];
    Q[1][j] = s2[1] * v[4 + j] + m2[1] * v[8 + j];
    Q[2][j] = v[8 + j];
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Pm[i][j] = Q[i][j] * (1.0 / s3[j]);
    Pm[i][3] = Q[i][0] * (-m3[0] / s3[0]) + Q[i][1] * (-m3[1] / s3[1]) + Q[i][2] * (-m3[2] / s3[2]) + Q[i][3];
  }
  double scale = Pm[2][3];
  if (scale == 0.0) {
    double f = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) f += Pm[i][j] * Pm[i][j];
    scale = sqrt(f);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) Pm[i][j] /= scale;
  // RQ of M = P[:, :3] with a positive diagonal of K (= scipy.linalg.rq followed by the sign fix of my_dlt.py:120-122)
  double k22 = sqrt(Pm[2][0] * Pm[2][0] + Pm[2][1] * Pm[2][1] + Pm[2][2] * Pm[2][2]);
  if (k22 == 0.0) return 0;
  for (int j = 0; j < 3; ++j) R[2][j] = Pm[2][j] / k22;
  const double k12 = Pm[1][0] * R[2][0] + Pm[1][1] * R[2][1] + Pm[1][2] * R[2][2];
  double w[3];
  for (int j = 0; j < 3; ++j) w[j] = Pm[1][j] - k12 * R[2][j];
  const double k11 = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int j = 0; j < 3; ++j) R[1][j] = w[j] / k11;
  const double k02 = Pm[0][0] * R[2][0] + Pm[0][1] * R[2][1] + Pm[0][2] * R[2][2];
  const double k01 = Pm[0][0] * R[1][0] + Pm[0][1] * R[1][1] + Pm[0][2] * R[1][2];
  for (int j = 0; j < 3; ++j) w[j] = Pm[0][j] - k02 * R[2][j] - k01 * R[1][j];
  const double k00 = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int j = 0; j < 3; ++j) R[0][j] = w[j] / k00;
  K[0][0] = k00 / k22, K[0][1] = k01 / k22, K[0][2] = k02 / k22;
  K[1][0] = 0.0, K[1][1] = k11 / k22, K[1][2] = k12 / k22;
  K[2][0] = 0.0, K[2][1] = 0.0, K[2][2] = 1.0;
  const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                     R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  if (det < 0.0)
    for (int i = 0; i < 3; ++i) R[i][2] = -R[i][2];
  // t = solve(K, p4): back substitution
  t[2] = Pm[2][3] / K[2][2];
  t[1] = (Pm[1][3] - K[1][2] * t[2]) / K[1][1];
  t[0] = (Pm[0][3] - K[0][1] * t[1] - K[0][2] * t[2]) / K[0][0];
  return 1;
}

// x0 of regress_cameramatrices.py:84-92 from a start (fx, fy, R, t)
TTK_HD static inline void cb_start(double fx, double fy, const double R[3][3], const double* t, double* x0) {
  x0[0] = fx, x0[1] = fy, x0[2] = t[0], x0[3] = t[1], x0[4] = t[2];
  cb_euler_xyz(R, x0 + 5);
}
