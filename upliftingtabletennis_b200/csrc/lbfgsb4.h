// L-BFGS-B for the 4-parameter Gaussian window fit of the heatmap decode, written so that it follows
// what the reference executes step for step:
//   scipy.optimize.minimize(loss, [1,1,1,1], method='L-BFGS-B', bounds=...)            (SciPy defaults)
//   (tabledetection/helper_tabledetection.py:117-126, balldetection/helper_balldetection.py:84-87)
// i.e. L-BFGS-B 3.0 (Zhu, Byrd, Lu, Nocedal; Morales & Nocedal 2011 subspace step) with m = 10,
// factr = 1e7, pgtol = 1e-5, maxls = 20, driven by SciPy's 2-point finite-difference gradient
// (absolute step 1e-8, flipped at an upper bound; scipy/optimize/_numdiff.py).  SciPy is third party and
// its compiled core is not part of the reference tree; the published algorithm is restated here.
//
// Because n = 4 the limited-memory matrix B = theta*I - W M W^T is formed densely from the stored
// (s, y) pairs (BFGS recursion from theta*I, algebraically identical to the compact representation);
// the generalized Cauchy point, the subspace minimisation and the More-Thuente line search (dcsrch /
// dcstep) then follow the original control flow, so the iterates and the stopping point agree with
// SciPy's to rounding noise.  Used by the decode kernel (device) and by the host unit test.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TTK_HD __host__ __device__
#else
#define TTK_HD
#endif

#define LB_N 4
#define LB_M 10

struct TtkLbfgsbResult {
  double x[LB_N];
  double f;
  int nit;
  int nfev;
  int success;      // 1: converged (projected gradient or relative reduction test), 0: abnormal termination
  int reason;       // 1 pgtol, 2 factr, -1 abnormal line search, -2 iteration limit
};

struct TtkGaussObjective {
  double w[9];
  int clamp;        // table variant clamps sigma at 0.5 inside the loss (helper_tabledetection.py:80-81)
};

// loss(params): mean over the 3x3 window of (gaussian - w)^2, summed in numpy's pairwise order for 9 terms
TTK_HD static inline double ttk_gauss_loss(const TtkGaussObjective* o, const double* p) {
  const double x0 = p[0], y0 = p[1];
  double sx = p[2], sy = p[3];
  if (o->clamp) {
    sx = sx > 0.5 ? sx : 0.5;
    sy = sy > 0.5 ? sy : 0.5;
  }
  const double dsx = 2.0 * (sx * sx), dsy = 2.0 * (sy * sy);
#if defined(__CUDA_ARCH__) && defined(TTK_WARP_COOPERATIVE_LOSS)
  // Device: the whole warp runs the optimiser in lock step (identical data, no divergence); lane l < 9 evaluates window
  // element l and the 9 terms are summed in the same pairwise order as the scalar path below.
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double t = 0.0;
  if (lane < 9) {
    const double dx = (double)(lane % 3) - x0, dy = (double)(lane / 3) - y0;
    const double g = exp(-((dx * dx) / dsx + (dy * dy) / dsy));
    const double r = g - o->w[lane];
    t = r * r;
  }
  const double a = t + __shfl_down_sync(full, t, 1);
  const double b = a + __shfl_down_sync(full, a, 2);
  const double c = b + __shfl_down_sync(full, b, 4);
  const double s = __shfl_sync(full, c, 0) + __shfl_sync(full, t, 8);
  return s / 9.0;
#else
  double t[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      const double dx = (double)i - x0, dy = (double)j - y0;
      const double g = exp(-((dx * dx) / dsx + (dy * dy) / dsy));
      const double r = g - o->w[j * 3 + i];
      t[j * 3 + i] = r * r;
    }
  const double s = (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]))) + t[8];
  return s / 9.0;
#endif
}

// f and SciPy's 2-point finite-difference gradient (5 evaluations)
TTK_HD static inline double ttk_gauss_fg(const TtkGaussObjective* o, const double* x, const double* lo, const double* hi, double* g,
                                         int* nfev) {
  const double f0 = ttk_gauss_loss(o, x);
  for (int i = 0; i < LB_N; ++i) {
    double h = 1e-8;
    const double xt = x[i] + h;
    if (xt < lo[i] || xt > hi[i]) h = -h;          // _adjust_scheme_to_bounds, '1-sided'
    double xp[LB_N] = {x[0], x[1], x[2], x[3]};
    xp[i] = x[i] + h;
    const double dx = xp[i] - x[i];
    g[i] = (ttk_gauss_loss(o, xp) - f0) / dx;
  }
  *nfev += 1;       // SciPy counts the finite-difference evaluations separately (ngev); nfev follows fun calls
  return f0;
}

TTK_HD static inline double lb_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }

TTK_HD static inline void lb_matvec(const double (*B)[LB_N], const double* v, double* out) {
  for (int i = 0; i < LB_N; ++i) out[i] = B[i][0] * v[0] + B[i][1] * v[1] + B[i][2] * v[2] + B[i][3] * v[3];
}

TTK_HD static inline double lb_projgr(const double* x, const double* lo, const double* hi, const double* g) {
  double nrm = 0.0;
  for (int i = 0; i < LB_N; ++i) {
    double gi = g[i];
    if (gi < 0.0)
      gi = fmax(x[i] - hi[i], gi);
    else
      gi = fmin(x[i] - lo[i], gi);
    nrm = fmax(nrm, fabs(gi));
  }
  return nrm;
}

// ---- More-Thuente line search (MINPACK-2 dcsrch / dcstep) ------------------------------------------
struct LbSearch {
  int brackt, stage;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

TTK_HD static inline void lb_dcstep(double* stx, double* fx, double* dx, double* sty, double* fy, double* dy, double* stp, double fp,
                                    double dp, int* brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (*dx / fabs(*dx));
  double stpf;
  if (fp > *fx) {
    const double theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
    if (*stp < *stx) gamma = -gamma;
    const double p = (gamma - *dx) + theta, q = ((gamma - *dx) + gamma) + dp, r = p / q;
    const double stpc = *stx + r * (*stp - *stx);
    const double stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
    if (fabs(stpc - *stx) < fabs(stpq - *stx))
      stpf = stpc;
    else
      stpf = stpc + (stpq - stpc) / 2.0;
    *brackt = 1;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
    if (*stp > *stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + *dx, r = p / q;
    const double stpc = *stp + r * (*stx - *stp);
    const double stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
    stpf = fabs(stpc - *stp) > fabs(stpq - *stp) ? stpc : stpq;
    *brackt = 1;
  } else if (fabs(dp) < fabs(*dx)) {
    const double theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
    if (*stp > *stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = (gamma + (*dx - dp)) + gamma, r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0)
      stpc = *stp + r * (*stx - *stp);
    else if (*stp > *stx)
      stpc = stpmax;
    else
      stpc = stpmin;
    const double stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
    if (*brackt) {
      stpf = fabs(stpc - *stp) < fabs(stpq - *stp) ? stpc : stpq;
      if (*stp > *stx)
        stpf = fmin(*stp + 0.66 * (*sty - *stp), stpf);
      else
        stpf = fmax(*stp + 0.66 * (*sty - *stp), stpf);
    } else {
      stpf = fabs(stpc - *stp) > fabs(stpq - *stp) ? stpc : stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (*brackt) {
      const double theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
      const double s = fmax(fabs(theta), fmax(fabs(*dy), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
      if (*stp > *sty) gamma = -gamma;
      const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + *dy, r = p / q;
      stpf = *stp + r * (*sty - *stp);
    } else if (*stp > *stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > *fx) {
    *sty = *stp;
    *fy = fp;
    *dy = dp;
  } else {
    if (sgnd < 0.0) {
      *sty = *stx;
      *fy = *fx;
      *dy = *dx;
    }
    *stx = *stp;
    *fx = fp;
    *dx = dp;
  }
  *stp = stpf;
}

// returns 0: evaluate f,g at the new stp ('FG'); 1: converged; 2: warning (also ends the search)
TTK_HD static inline int lb_dcsrch(LbSearch* S, int start, double f, double g, double* stp, double stpmax) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0;
  if (start) {
    S->brackt = 0;
    S->stage = 1;
    S->finit = f;
    S->ginit = g;
    S->gtest = ftol * S->ginit;
    S->width = stpmax - stpmin;
    S->width1 = S->width / 0.5;
    S->stx = 0.0;
    S->fx = S->finit;
    S->gx = S->ginit;
    S->sty = 0.0;
    S->fy = S->finit;
    S->gy = S->ginit;
    S->stmin = 0.0;
    S->stmax = *stp + 4.0 * *stp;
    return 0;
  }
  const double ftest = S->finit + *stp * S->gtest;
  if (S->stage == 1 && f <= ftest && g >= 0.0) S->stage = 2;
  int status = 0;
  if (S->brackt && (*stp <= S->stmin || *stp >= S->stmax)) status = 2;
  if (S->brackt && S->stmax - S->stmin <= xtol * S->stmax) status = 2;
  if (*stp == stpmax && f <= ftest && g <= S->gtest) status = 2;
  if (*stp == stpmin && (f > ftest || g >= S->gtest)) status = 2;
  if (f <= ftest && fabs(g) <= gtol * (-S->ginit)) status = 1;
  if (status) return status;
  if (S->stage == 1 && f <= S->fx && f > ftest) {
    const double fm = f - *stp * S->gtest;
    double fxm = S->fx - S->stx * S->gtest, fym = S->fy - S->sty * S->gtest;
    const double gm = g - S->gtest;
    double gxm = S->gx - S->gtest, gym = S->gy - S->gtest;
    lb_dcstep(&S->stx, &fxm, &gxm, &S->sty, &fym, &gym, stp, fm, gm, &S->brackt, S->stmin, S->stmax);
    S->fx = fxm + S->stx * S->gtest;
    S->fy = fym + S->sty * S->gtest;
    S->gx = gxm + S->gtest;
    S->gy = gym + S->gtest;
  } else {
    lb_dcstep(&S->stx, &S->fx, &S->gx, &S->sty, &S->fy, &S->gy, stp, f, g, &S->brackt, S->stmin, S->stmax);
  }
  if (S->brackt) {
    if (fabs(S->sty - S->stx) >= 0.66 * S->width1) *stp = S->stx + 0.5 * (S->sty - S->stx);
    S->width1 = S->width;
    S->width = fabs(S->sty - S->stx);
  }
  if (S->brackt) {
    S->stmin = fmin(S->stx, S->sty);
    S->stmax = fmax(S->stx, S->sty);
  } else {
    S->stmin = *stp + 1.1 * (*stp - S->stx);
    S->stmax = *stp + 4.0 * (*stp - S->stx);
  }
  *stp = fmax(*stp, stpmin);
  *stp = fmin(*stp, stpmax);
  if ((S->brackt && (*stp <= S->stmin || *stp >= S->stmax)) || (S->brackt && S->stmax - S->stmin <= xtol * S->stmax)) *stp = S->stx;
  return 0;
}

// ---- dense limited-memory matrix -----------------------------------------------------------------------
struct LbMemory {
  double s[LB_M][LB_N], y[LB_M][LB_N];
  int col, head;      // ring buffer: pairs head .. head+col-1 (mod m), oldest first
  double theta;
};

TTK_HD static inline void lb_build_B(const LbMemory* M, double (*B)[LB_N]) {
  for (int i = 0; i < LB_N; ++i)
    for (int j = 0; j < LB_N; ++j) B[i][j] = i == j ? M->theta : 0.0;
  for (int k = 0; k < M->col; ++k) {
    const int p = (M->head + k) % LB_M;
    double Bs[LB_N];
    lb_matvec(B, M->s[p], Bs);
    const double sBs = lb_dot(M->s[p], Bs), ys = lb_dot(M->y[p], M->s[p]);
    for (int i = 0; i < LB_N; ++i)
      for (int j = 0; j < LB_N; ++j) B[i][j] += -Bs[i] * Bs[j] / sBs + M->y[p][i] * M->y[p][j] / ys;
  }
}

// Generalized Cauchy point along the projected steepest-descent path (subroutine cauchy).
TTK_HD static inline void lb_cauchy(const double* x, const double* lo, const double* hi, const double* g, const double (*B)[LB_N],
                                    double sbgnrm, int* iwhere, double* xcp) {
  const double epsmch = 2.220446049250313e-16;
  for (int i = 0; i < LB_N; ++i) xcp[i] = x[i];
  if (sbgnrm <= 0.0) return;
  double d[LB_N], tb[LB_N];
  int order[LB_N], nbreak = 0;
  double f1 = 0.0;
  for (int i = 0; i < LB_N; ++i) {
    const double neggi = -g[i];
    const double tl = x[i] - lo[i], tu = hi[i] - x[i];
    if (iwhere[i] != 3 && iwhere[i] != -1) {
      const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
      iwhere[i] = 0;
      if (xlower) {
        if (neggi <= 0.0) iwhere[i] = 1;
      } else if (xupper) {
        if (neggi >= 0.0) iwhere[i] = 2;
      } else if (fabs(neggi) <= 0.0) {
        iwhere[i] = -3;
      }
    }
    if (iwhere[i] != 0 && iwhere[i] != -1) {
      d[i] = 0.0;
    } else {
      d[i] = neggi;
      f1 -= neggi * neggi;
      if (neggi < 0.0) {
        order[nbreak] = i;
        tb[nbreak++] = tl / (-neggi);
      } else if (neggi > 0.0) {
        order[nbreak] = i;
        tb[nbreak++] = tu / neggi;
      }
    }
  }
  if (nbreak == 0) return;
  // ascending breakpoints (n <= 4: insertion sort; the original uses a heap)
  for (int a = 1; a < nbreak; ++a)
    for (int b = a; b > 0 && tb[b] < tb[b - 1]; --b) {
      const double tt = tb[b];
      tb[b] = tb[b - 1];
      tb[b - 1] = tt;
      const int oo = order[b];
      order[b] = order[b - 1];
      order[b - 1] = oo;
    }
  double Bd[LB_N];
  lb_matvec(B, d, Bd);
  double f2 = lb_dot(d, Bd);
  const double f2_org = f2;
  double dtm = -f1 / f2, tsum = 0.0, tj = 0.0;
  double z[LB_N] = {0.0, 0.0, 0.0, 0.0};
  int nleft = nbreak;
  bool done = false;
  for (int k = 0; k < nbreak; ++k) {
    const double tj0 = tj;
    tj = tb[k];
    const int ibp = order[k];
    const double dt = tj - tj0;
    if (dtm < dt) break;
    tsum += dt;
    --nleft;
    for (int i = 0; i < LB_N; ++i) z[i] += dt * d[i];
    const double dibp = d[ibp];
    d[ibp] = 0.0;
    if (dibp > 0.0) {
      z[ibp] = hi[ibp] - x[ibp];
      xcp[ibp] = hi[ibp];
      iwhere[ibp] = 2;
    } else {
      z[ibp] = lo[ibp] - x[ibp];
      xcp[ibp] = lo[ibp];
      iwhere[ibp] = 1;
    }
    if (nleft == 0 && nbreak == LB_N) {
      done = true;      // every variable sits on a bound
      break;
    }
    double Bz[LB_N];
    lb_matvec(B, z, Bz);
    lb_matvec(B, d, Bd);
    f1 = lb_dot(g, d) + lb_dot(d, Bz);
    f2 = lb_dot(d, Bd);
    f2 = fmax(epsmch * f2_org, f2);
    if (nleft > 0) {
      dtm = -f1 / f2;
    } else {
      dtm = 0.0;        // all moving variables were bounded
    }
  }
  if (done) return;
  if (dtm <= 0.0) dtm = 0.0;
  tsum += dtm;
  for (int i = 0; i < LB_N; ++i) xcp[i] += tsum * d[i];
}

// Solve the nf x nf SPD system A v = rhs in place (Cholesky).  Returns 0 when A is not positive definite.
TTK_HD static inline int lb_chol_solve(double (*A)[LB_N], double* rhs, int nf) {
  for (int c = 0; c < nf; ++c) {
    double dg = A[c][c];
    for (int k = 0; k < c; ++k) dg -= A[c][k] * A[c][k];
    if (!(dg > 0.0)) return 0;
    dg = sqrt(dg);
    A[c][c] = dg;
    for (int r = c + 1; r < nf; ++r) {
      double sv = A[r][c];
      for (int k = 0; k < c; ++k) sv -= A[r][k] * A[c][k];
      A[r][c] = sv / dg;
    }
  }
  for (int r = 0; r < nf; ++r) {
    double sv = rhs[r];
    for (int k = 0; k < r; ++k) sv -= A[r][k] * rhs[k];
    rhs[r] = sv / A[r][r];
  }
  for (int r = nf - 1; r >= 0; --r) {
    double sv = rhs[r];
    for (int k = r + 1; k < nf; ++k) sv -= A[k][r] * rhs[k];
    rhs[r] = sv / A[r][r];
  }
  return 1;
}

// Subspace minimisation over the variables free at the Cauchy point (cmprlb + subsm, L-BFGS-B 3.0).
// z enters as the Cauchy point and leaves as the subspace minimiser.  Returns 0 on a singular system.
TTK_HD static inline int lb_subsm(const double* x, const double* lo, const double* hi, const double* g, const double (*B)[LB_N],
                                  const int* iwhere, double* z) {
  int ind[LB_N], nf = 0;
  for (int i = 0; i < LB_N; ++i)
    if (iwhere[i] <= 0) ind[nf++] = i;
  if (nf == 0) return 1;
  double dz[LB_N], Bdz[LB_N];
  for (int i = 0; i < LB_N; ++i) dz[i] = z[i] - x[i];
  lb_matvec(B, dz, Bdz);
  double A[LB_N][LB_N], du[LB_N];
  for (int r = 0; r < nf; ++r) {
    for (int c = 0; c < nf; ++c) A[r][c] = B[ind[r]][ind[c]];
    du[r] = -(g[ind[r]] + Bdz[ind[r]]);
  }
  if (!lb_chol_solve(A, du, nf)) return 0;
  double xp[LB_N];
  for (int i = 0; i < LB_N; ++i) xp[i] = z[i];
  int iword = 0;
  for (int r = 0; r < nf; ++r) {
    const int k = ind[r];
    const double xk = z[k] + du[r];
    z[k] = fmin(hi[k], fmax(lo[k], xk));
    if (z[k] == lo[k] || z[k] == hi[k]) iword = 1;
  }
  if (iword == 0) return 1;
  double dd_p = 0.0;
  for (int i = 0; i < LB_N; ++i) dd_p += (z[i] - x[i]) * g[i];
  if (dd_p > 0.0) {
    // the projected point is not a descent direction: fall back to the truncated step
    for (int i = 0; i < LB_N; ++i) z[i] = xp[i];
    double alpha = 1.0, temp1 = 1.0;
    int ibd = -1;
    for (int r = 0; r < nf; ++r) {
      const int k = ind[r];
      const double dk = du[r];
      if (dk < 0.0) {
        const double temp2 = lo[k] - z[k];
        if (temp2 >= 0.0)
          temp1 = 0.0;
        else if (dk * alpha < temp2)
          temp1 = temp2 / dk;
      } else if (dk > 0.0) {
        const double temp2 = hi[k] - z[k];
        if (temp2 <= 0.0)
          temp1 = 0.0;
        else if (dk * alpha > temp2)
          temp1 = temp2 / dk;
      }
      if (temp1 < alpha) {
        alpha = temp1;
        ibd = r;
      }
    }
    if (alpha < 1.0 && ibd >= 0) {
      const double dk = du[ibd];
      const int k = ind[ibd];
      if (dk > 0.0) {
        z[k] = hi[k];
        du[ibd] = 0.0;
      } else if (dk < 0.0) {
        z[k] = lo[k];
        du[ibd] = 0.0;
      }
    }
    for (int r = 0; r < nf; ++r) z[ind[r]] += alpha * du[r];
  }
  return 1;
}

// variant 0: table (sigma in [0.5, 3], clamped loss), 1: ball (sigma in [0.5, 50])
TTK_HD static inline TtkLbfgsbResult ttk_lbfgsb_gauss(const double* w9, int variant) {
  const double epsmch = 2.220446049250313e-16, factr = 1e7, pgtol = 1e-5;
  const double tol = factr * epsmch;
  const int maxls = 20, maxiter = 15000;
  TtkGaussObjective obj;
  for (int i = 0; i < 9; ++i) obj.w[i] = w9[i];
  obj.clamp = variant == 0;
  const double smax = variant == 0 ? 3.0 : 50.0;
  const double lo[LB_N] = {0.0, 0.0, 0.5, 0.5}, hi[LB_N] = {3.0, 3.0, smax, smax};
  TtkLbfgsbResult R;
  double x[LB_N] = {1.0, 1.0, 1.0, 1.0}, g[LB_N];
  int iwhere[LB_N] = {0, 0, 0, 0};
  LbMemory M;
  M.col = 0;
  M.head = 0;
  M.theta = 1.0;
  R.nfev = 0;
  R.nit = 0;
  R.success = 0;
  R.reason = 0;
  double f = ttk_gauss_fg(&obj, x, lo, hi, g, &R.nfev);
  double sbgnrm = lb_projgr(x, lo, hi, g);
  bool finished = false;
  if (!(sbgnrm == sbgnrm) || !(f == f)) {           // NaN objective: the reference's fit fails
    finished = true;
    R.reason = -1;
  } else if (sbgnrm <= pgtol) {
    finished = true;
    R.success = 1;
    R.reason = 1;
  }
  int iter = 0;
  while (!finished) {
    double B[LB_N][LB_N], z[LB_N];
    lb_build_B(&M, B);
    lb_cauchy(x, lo, hi, g, B, sbgnrm, iwhere, z);
    int nfree = 0;
    for (int i = 0; i < LB_N; ++i)
      if (iwhere[i] <= 0) ++nfree;
    if (nfree != 0 && M.col != 0) {
      if (!lb_subsm(x, lo, hi, g, B, iwhere, z)) {    // singular: refresh the memory and restart the iteration
        M.col = 0;
        M.head = 0;
        M.theta = 1.0;
        continue;
      }
    }
    // ---- line search along d = z - x (lnsrlb) ----
    double d[LB_N], t[LB_N], r[LB_N];
    for (int i = 0; i < LB_N; ++i) d[i] = z[i] - x[i];
    const double dtd = lb_dot(d, d);
    (void)dtd;
    double stpmx = 1e10;
    if (iter == 0) {
      stpmx = 1.0;
    } else {
      for (int i = 0; i < LB_N; ++i) {
        const double a1 = d[i];
        if (a1 < 0.0) {
          const double a2 = lo[i] - x[i];
          if (a2 >= 0.0)
            stpmx = 0.0;
          else if (a1 * stpmx < a2)
            stpmx = a2 / a1;
        } else if (a1 > 0.0) {
          const double a2 = hi[i] - x[i];
          if (a2 <= 0.0)
            stpmx = 0.0;
          else if (a1 * stpmx > a2)
            stpmx = a2 / a1;
        }
      }
    }
    double stp = 1.0;                                  // all four variables are boxed
    for (int i = 0; i < LB_N; ++i) {
      t[i] = x[i];
      r[i] = g[i];
    }
    const double fold = f;
    int ifun = 0, iback = 0, info = 0;
    double gd = 0.0, gdold = 0.0;
    LbSearch S;
    bool ls_done = false;
    while (!ls_done) {
      gd = lb_dot(g, d);
      int status;
      if (ifun == 0) {
        gdold = gd;
        if (gd >= 0.0) {          // ascent direction in projection
          info = -4;
          break;
        }
        status = lb_dcsrch(&S, 1, f, gd, &stp, stpmx);
      } else {
        status = lb_dcsrch(&S, 0, f, gd, &stp, stpmx);
      }
      if (status == 0) {
        ++ifun;
        iback = ifun - 1;
        if (iback >= maxls) {
          info = -3;              // scipy's core stops the search after maxls trial points
          break;
        }
        if (stp == 1.0) {
          for (int i = 0; i < LB_N; ++i) x[i] = z[i];
        } else {
          for (int i = 0; i < LB_N; ++i) x[i] = stp * d[i] + t[i];
        }
        f = ttk_gauss_fg(&obj, x, lo, hi, g, &R.nfev);
      } else {
        ls_done = true;
      }
    }
    if (info != 0) {
      for (int i = 0; i < LB_N; ++i) {
        x[i] = t[i];
        g[i] = r[i];
      }
      f = fold;
      if (M.col == 0) {           // ABNORMAL_TERMINATION_IN_LNSRCH
        R.reason = -1;
        ++iter;
        break;
      }
      M.col = 0;                  // refresh the memory and restart from the steepest-descent model
      M.head = 0;
      M.theta = 1.0;
      continue;
    }
    ++iter;
    sbgnrm = lb_projgr(x, lo, hi, g);
    if (iter >= maxiter) {
      R.reason = -2;
      break;
    }
    if (sbgnrm <= pgtol) {
      R.success = 1;
      R.reason = 1;
      break;
    }
    const double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
    if (fold - f <= tol * ddum) {
      R.success = 1;
      R.reason = 2;
      break;
    }
    // ---- BFGS pair (s = stp*d, y = g - g_old) ----
    double yv[LB_N], sv[LB_N];
    for (int i = 0; i < LB_N; ++i) yv[i] = g[i] - r[i];
    const double rr = lb_dot(yv, yv);
    double dr, ddm;
    if (stp == 1.0) {
      dr = gd - gdold;
      ddm = -gdold;
      for (int i = 0; i < LB_N; ++i) sv[i] = d[i];
    } else {
      dr = (gd - gdold) * stp;
      for (int i = 0; i < LB_N; ++i) sv[i] = d[i] * stp;
      ddm = -gdold * stp;
    }
    if (dr <= epsmch * ddm) continue;      // skip the update
    int slot;
    if (M.col < LB_M) {
      slot = (M.head + M.col) % LB_M;
      ++M.col;
    } else {
      slot = M.head;
      M.head = (M.head + 1) % LB_M;
    }
    for (int i = 0; i < LB_N; ++i) {
      M.s[slot][i] = sv[i];
      M.y[slot][i] = yv[i];
    }
    M.theta = rr / dr;
  }
  for (int i = 0; i < LB_N; ++i) R.x[i] = x[i];
  R.f = f;
  R.nit = iter;
  return R;
}
