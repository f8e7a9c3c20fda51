// Camera calibration on the device: calibrate_camera (inference/utils.py:312-329) =
// calc_cameramatrices(use_ransac=True) (dataprocessing/regress_cameramatrices.py:199-231):
//   1. DLT on all visible keypoints (my_dlt.py) -> start parameters                         calib_start_kernel, 1 thread / clip
//   2. 100 RANSAC hypotheses (keys 10 and 11 + 4 sampled keys), each an 8-parameter SciPy-BFGS fit from that start,
//      scored by the number of keypoints within 3.5 px (regress_cameramatrices.py:119-165)  calib_ransac_kernel, 1 warp / hypothesis
//   3. first hypothesis with the most inliers, refit on its inliers (:170-178)              calib_refine_kernel, 1 warp / clip
// The reference runs the 101 fits one after the other on the CPU (~24 s per clip); here the 100
// hypotheses run concurrently, one warp each: the optimiser's control flow is replicated across the warp and lanes 0..7
// evaluate the 8 forward differences of every gradient in parallel (calib.h).  The hypothesis table itself is data
// independent (numpy Generator(42).choice over the visible keys) and is computed by the host exactly as the reference does.
#include "calib.h"
#include "ttk_internal.h"

namespace {

constexpr int N_KP = CB_MAXPTS;
constexpr int HYP_STRIDE = 12;        // doubles per hypothesis record: x[8], inlier count, inlier mask, bfgs status, nit

struct ClipView {
  const double* kp;                   // 13 x 3 (x, y, v)
  const double* world;                // 13 x 3
};

__device__ __forceinline__ bool visible(const double* kp, int i) { return kp[i * 3 + 2] == 1.0; }

__device__ void add_point(CalibProblem* P, const ClipView& c, int i) {
  const int n = P->n++;
  for (int d = 0; d < 3; ++d) P->X[n][d] = c.world[i * 3 + d];
  P->u[n][0] = c.kp[i * 3];
  P->u[n][1] = c.kp[i * 3 + 1];
}

__global__ void calib_start_kernel(const double* __restrict__ kps, const double* __restrict__ world, int width, int height,
                                   double* __restrict__ start, int32_t* __restrict__ info) {
  const int clip = blockIdx.x;
  ClipView c{kps + (size_t)clip * N_KP * 3, world};
  CalibProblem P;
  P.n = 0;
  P.px = (double)(width / 2);
  P.py = (double)(height / 2);
  for (int i = 0; i < N_KP; ++i)
    if (visible(c.kp, i)) add_point(&P, c, i);
  double K[3][3], R[3][3], t[3];
  const int ok = P.n >= 6 ? cb_dlt(&P, K, R, t) : 0;
  double* x0 = start + (size_t)clip * CB_N;
  if (ok) {
    cb_start(K[0][0], K[1][1], R, t, x0);
  } else {
    for (int i = 0; i < CB_N; ++i) x0[i] = NAN;
  }
  info[clip * 4 + 3] = ok;
}

__global__ void __launch_bounds__(32) calib_ransac_kernel(const double* __restrict__ kps, const double* __restrict__ world,
                                                          const int32_t* __restrict__ samples, int n_hyp, int n_sample,
                                                          int width, int height, double threshold,
                                                          const double* __restrict__ start, double* __restrict__ hyp) {
  const int h = blockIdx.x, clip = blockIdx.y;
  ClipView c{kps + (size_t)clip * N_KP * 3, world};
  const int32_t* smp = samples + ((size_t)clip * n_hyp + h) * n_sample;
  CalibProblem P;
  P.n = 0;
  P.px = (double)(width / 2);
  P.py = (double)(height / 2);
  // subset_points2d = [*fixed_points2d, *sampled_points2d], both in keypoint order (regress_cameramatrices.py:136-146)
  for (int i = 0; i < N_KP; ++i)
    if (visible(c.kp, i) && (i + 1 == 10 || i + 1 == 11)) add_point(&P, c, i);
  for (int i = 0; i < N_KP; ++i) {
    if (!visible(c.kp, i) || i + 1 == 10 || i + 1 == 11) continue;
    bool in = false;
    for (int s = 0; s < n_sample; ++s) in |= smp[s] == i + 1;
    if (in) add_point(&P, c, i);
  }
  double x0[CB_N];
  for (int i = 0; i < CB_N; ++i) x0[i] = start[(size_t)clip * CB_N + i];
  const CalibResult r = cb_bfgs(&P, x0);
  // inliers over all visible keypoints (:152-158)
  CalibProblem all;
  all.n = 0;
  all.px = P.px;
  all.py = P.py;
  int idx[N_KP];
  for (int i = 0; i < N_KP; ++i)
    if (visible(c.kp, i)) {
      idx[all.n] = i;
      add_point(&all, c, i);
    }
  double R[3][3];
  cb_rotation(r.x[5], r.x[6], r.x[7], R);
  int count = 0;
  unsigned mask = 0;
  for (int j = 0; j < all.n; ++j)
    if (cb_point_error(&all, R, r.x, j) < threshold) {
      ++count;
      mask |= 1u << idx[j];
    }
  if (threadIdx.x == 0) {
    double* o = hyp + ((size_t)clip * n_hyp + h) * HYP_STRIDE;
    for (int i = 0; i < CB_N; ++i) o[i] = r.x[i];
    o[8] = (double)count;
    o[9] = (double)mask;
    o[10] = (double)r.status;
    o[11] = (double)r.nit;
  }
}

__global__ void __launch_bounds__(32) calib_refine_kernel(const double* __restrict__ kps, const double* __restrict__ world, int n_hyp,
                                                          int width, int height, const double* __restrict__ hyp,
                                                          double* __restrict__ mint_out, double* __restrict__ mext_out,
                                                          int32_t* __restrict__ info) {
  const int clip = blockIdx.x;
  ClipView c{kps + (size_t)clip * N_KP * 3, world};
  const double* H = hyp + (size_t)clip * n_hyp * HYP_STRIDE;
  int best = -1, best_count = -1;
  for (int h = 0; h < n_hyp; ++h) {               // "if best_inliers is None or len(inliers) > len(best_inliers)"
    const int cnt = (int)H[h * HYP_STRIDE + 8];
    if (cnt > best_count) best_count = cnt, best = h;
  }
  const double* xb = H + (size_t)best * HYP_STRIDE;
  const unsigned mask = (unsigned)xb[9];
  CalibProblem P;
  P.n = 0;
  P.px = (double)(width / 2);
  P.py = (double)(height / 2);
  for (int i = 0; i < N_KP; ++i)
    if (mask >> i & 1u) add_point(&P, c, i);
  // start of the refit: the best hypothesis' matrices, angles re-extracted from its rotation (:84-92)
  double R[3][3], x0[CB_N];
  cb_rotation(xb[5], xb[6], xb[7], R);
  const double t[3] = {xb[2], xb[3], xb[4]};
  cb_start(xb[0], xb[1], R, t, x0);
  CalibResult r;
  if (P.n > 0) {
    r = cb_bfgs(&P, x0);
  } else {                                         // the reference fails on an empty inlier set; report it
    for (int i = 0; i < CB_N; ++i) r.x[i] = NAN;
    r.status = 3;
    r.nit = 0;
  }
  if (threadIdx.x == 0) {
    cb_rotation(r.x[5], r.x[6], r.x[7], R);
    double* mi = mint_out + (size_t)clip * 12;
    double* me = mext_out + (size_t)clip * 16;
    const double Mi[12] = {r.x[0], 0, P.px, 0, 0, r.x[1], P.py, 0, 0, 0, 1, 0};
    for (int i = 0; i < 12; ++i) mi[i] = Mi[i];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) me[i * 4 + j] = R[i][j];
      me[i * 4 + 3] = r.x[2 + i];
    }
    me[12] = me[13] = me[14] = 0.0;
    me[15] = 1.0;
    info[clip * 4 + 0] = best_count;
    info[clip * 4 + 1] = best;
    info[clip * 4 + 2] = r.status;
  }
}

}  // namespace

extern "C" size_t ttk_calibrate_workspace_bytes(int n_clips, int n_hypotheses) {
  if (n_clips <= 0 || n_hypotheses <= 0) return 0;
  return (size_t)n_clips * (CB_N + (size_t)n_hypotheses * HYP_STRIDE) * sizeof(double);
}

extern "C" int ttk_calibrate_camera(const double* keypoints_dev, const double* world_points_dev, const int32_t* samples_dev,
                                    int n_clips, int n_hypotheses, int n_sample, int image_width, int image_height,
                                    double inlier_threshold_px, double* mint_out_dev, double* mext_out_dev, int32_t* info_out_dev,
                                    void* workspace_dev, size_t workspace_bytes, void* stream) {
  TTK_CHECK_ARG(n_clips >= 0 && n_clips <= 65535 && n_hypotheses > 0 && n_sample > 0 && n_sample <= N_KP, "ttk_calibrate_camera: bad sizes");
  if (n_clips == 0) return TTK_OK;
  TTK_CHECK_ARG(keypoints_dev && world_points_dev && samples_dev && mint_out_dev && mext_out_dev && info_out_dev && workspace_dev,
                "ttk_calibrate_camera: null pointer");
  TTK_CHECK_ARG(workspace_bytes >= ttk_calibrate_workspace_bytes(n_clips, n_hypotheses), "ttk_calibrate_camera: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  double* start = (double*)workspace_dev;
  double* hyp = start + (size_t)n_clips * CB_N;
  calib_start_kernel<<<n_clips, 1, 0, s>>>(keypoints_dev, world_points_dev, image_width, image_height, start, info_out_dev);
  TTK_LAUNCH_CHECK();
  calib_ransac_kernel<<<dim3(n_hypotheses, n_clips), 32, 0, s>>>(keypoints_dev, world_points_dev, samples_dev, n_hypotheses, n_sample,
                                                                image_width, image_height, inlier_threshold_px, start, hyp);
  TTK_LAUNCH_CHECK();
  calib_refine_kernel<<<n_clips, 32, 0, s>>>(keypoints_dev, world_points_dev, n_hypotheses, image_width, image_height, hyp,
                                             mint_out_dev, mext_out_dev, info_out_dev);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}
