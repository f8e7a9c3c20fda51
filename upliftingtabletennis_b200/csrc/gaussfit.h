// 3x3 Gaussian sub-pixel fit shared by the decode kernel (device) and its host unit test.
//
// Objective and bounds are the reference's (tabledetection/helper_tabledetection.py:76-126,
// balldetection/helper_balldetection.py:70-87):
//   L(x0,y0,sx,sy) = mean_{i,j in {0,1,2}} (exp(-((i-x0)^2/(2 sx^2) + (j-y0)^2/(2 sy^2))) - w[j][i])^2
//   start (1,1,1,1); x0,y0 in [0,3]; sigma in [0.5,3] (table variant) or [0.5,50] (ball variant).
// The reference minimises L with SciPy's L-BFGS-B (finite-difference gradient, pgtol 1e-5,
// factr 1e7).  Here: a bounded Newton / Levenberg-Marquardt iteration in float64 with the exact
// gradient and Hessian, converged to a projected-gradient norm far below SciPy's stopping
// threshold, so the two agree to within SciPy's own stopping error (decode).
// The sigma clamp max(0.5, sigma) of the table variant is inactive inside the bounds.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TTK_HD __host__ __device__ __forceinline__
#else
#define TTK_HD static inline
#endif

struct TtkFit {
  double p[4];     // x0, y0, sigma_x, sigma_y
  double f;        // objective at p
  int iters;
  int ok;          // 0: non-finite window (reference fallback applies)
};

// f, gradient g[4], Hessian h[4][4] (symmetric, full) of L at p.
TTK_HD double ttk_gauss_fgh(const double* p, const double* w, double* g, double (*h)[4]) {
  const double x0 = p[0], y0 = p[1], sx = p[2], sy = p[3];
  const double isx2 = 1.0 / (sx * sx), isy2 = 1.0 / (sy * sy);
  const double isx = 1.0 / sx, isy = 1.0 / sy;
  double f = 0.0;
  for (int a = 0; a < 4; ++a) {
    g[a] = 0.0;
    for (int b = 0; b < 4; ++b) h[a][b] = 0.0;
  }
  for (int j = 0; j < 3; ++j) {
    for (int i = 0; i < 3; ++i) {
      const double dx = (double)i - x0, dy = (double)j - y0;
      const double u = dx * dx * isx2, v = dy * dy * isy2;
      const double e = exp(-0.5 * (u + v));
      const double r = e - w[j * 3 + i];
      f += r * r;
      // first derivatives of e:  e * a_k
      const double a0 = dx * isx2, a1 = dy * isy2, a2 = u * isx, a3 = v * isy;
      const double a[4] = {a0, a1, a2, a3};
      // second derivatives of e: e * (a_k a_l + b_kl), b = d a_k / d p_l
      double b[4][4] = {{0}};
      b[0][0] = -isx2;
      b[1][1] = -isy2;
      b[0][2] = b[2][0] = -2.0 * a0 * isx;
      b[1][3] = b[3][1] = -2.0 * a1 * isy;
      b[2][2] = -3.0 * u * isx2;
      b[3][3] = -3.0 * v * isy2;
      for (int k = 0; k < 4; ++k) {
        g[k] += r * e * a[k];
        for (int l = 0; l < 4; ++l) h[k][l] += e * a[k] * e * a[l] + r * e * (a[k] * a[l] + b[k][l]);
      }
    }
  }
  const double s = 2.0 / 9.0;
  for (int k = 0; k < 4; ++k) {
    g[k] *= s;
    for (int l = 0; l < 4; ++l) h[k][l] *= s;
  }
  return f / 9.0;
}

TTK_HD double ttk_gauss_f(const double* p, const double* w) {
  const double isx2 = 1.0 / (p[2] * p[2]), isy2 = 1.0 / (p[3] * p[3]);
  double f = 0.0;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      const double dx = (double)i - p[0], dy = (double)j - p[1];
      const double r = exp(-0.5 * (dx * dx * isx2 + dy * dy * isy2)) - w[j * 3 + i];
      f += r * r;
    }
  return f / 9.0;
}

// Solve A d = rhs for the nf x nf leading system (Cholesky).  Returns 0 if A is not positive definite.
TTK_HD int ttk_chol_solve(double (*A)[4], double* rhs, int nf) {
  for (int c = 0; c < nf; ++c) {
    double d = A[c][c];
    for (int k = 0; k < c; ++k) d -= A[c][k] * A[c][k];
    if (!(d > 1e-300)) return 0;
    d = sqrt(d);
    A[c][c] = d;
    for (int r = c + 1; r < nf; ++r) {
      double s = A[r][c];
      for (int k = 0; k < c; ++k) s -= A[r][k] * A[c][k];
      A[r][c] = s / d;
    }
  }
  for (int r = 0; r < nf; ++r) {
    double s = rhs[r];
    for (int k = 0; k < r; ++k) s -= A[r][k] * rhs[k];
    rhs[r] = s / A[r][r];
  }
  for (int r = nf - 1; r >= 0; --r) {
    double s = rhs[r];
    for (int k = r + 1; k < nf; ++k) s -= A[k][r] * rhs[k];
    rhs[r] = s / A[r][r];
  }
  return 1;
}

// variant 0: table (sigma <= 3), 1: ball (sigma <= 50)
TTK_HD TtkFit ttk_gauss_fit(const double* w, int variant) {
  TtkFit out;
  const double lo[4] = {0.0, 0.0, 0.5, 0.5};
  const double smax = variant == 0 ? 3.0 : 50.0;
  const double hi[4] = {3.0, 3.0, smax, smax};
  double p[4] = {1.0, 1.0, 1.0, 1.0};
  out.ok = 1;
  for (int i = 0; i < 9; ++i)
    if (!(fabs(w[i]) <= 1.79e308)) out.ok = 0;   // NaN or inf
  double g[4], h[4][4];
  double f = ttk_gauss_fgh(p, w, g, h);
  double lam = 1e-3;
  int it = 0;
  if (out.ok) {
    for (; it < 200; ++it) {
      int idx[4], nf = 0;
      double pg = 0.0;
      for (int k = 0; k < 4; ++k) {
        const bool act = (p[k] <= lo[k] && g[k] > 0.0) || (p[k] >= hi[k] && g[k] < 0.0);
        if (!act) {
          idx[nf++] = k;
          pg = fmax(pg, fabs(g[k]));
        }
      }
      if (nf == 0 || pg < 1e-14) break;
      bool accepted = false;
      double step = 0.0;
      for (int tries = 0; tries < 60 && !accepted; ++tries) {
        double A[4][4], d[4];
        for (int r = 0; r < nf; ++r) {
          for (int c = 0; c < nf; ++c) A[r][c] = h[idx[r]][idx[c]];
          A[r][r] += lam * (fabs(h[idx[r]][idx[r]]) + 1e-12);
          d[r] = -g[idx[r]];
        }
        if (ttk_chol_solve(A, d, nf)) {
          double q[4] = {p[0], p[1], p[2], p[3]};
          for (int r = 0; r < nf; ++r) q[idx[r]] = fmin(hi[idx[r]], fmax(lo[idx[r]], p[idx[r]] + d[r]));
          const double fq = ttk_gauss_f(q, w);
          if (fq < f) {
            step = 0.0;
            for (int k = 0; k < 4; ++k) {
              step = fmax(step, fabs(q[k] - p[k]));
              p[k] = q[k];
            }
            accepted = true;
            lam = fmax(lam * 0.1, 1e-15);
            break;
          }
        }
        lam *= 10.0;
        if (lam > 1e15) break;
      }
      if (!accepted) break;
      f = ttk_gauss_fgh(p, w, g, h);
      if (step < 1e-15) break;
    }
  }
  for (int k = 0; k < 4; ++k) out.p[k] = p[k];
  out.f = f;
  out.iters = it;
  return out;
}
