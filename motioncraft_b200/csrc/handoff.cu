// motioncraft_b200 -- result hand-off kernels: what the reference does to the sampled
// motion on the HOST after the sampler returns, done on the device before the one device->host copy.
//
//   handoff_smplx_kernel  tools/visualize.py:219-249 (motionx): x = pred * std + mean; repack the 322-dim vector into
//                         SMPL-X poses (165) | expressions (100) | translation (3); per-column Gaussian temporal filter
//                         = scipy.ndimage.gaussian_filter(col, sigma, mode="nearest") with sigma 3.5 / 2.0 / 3.0.
//                         HBM-bound gather + a (2 r + 1)-tap stencil over T; float64 like the reference's numpy code and in
//                         scipy's accumulation order (ni_filters.c, symmetric branch: centre tap, then
//                         tmp += (x[l + j] + x[l - j]) * w[j] from the farthest tap inwards; no FMA contraction), so the
//                         result is BIT-identical to the reference's.  One block per (sample, output column): the
//                         column is staged in shared memory once (edge-replicated), algorithmic bytes = 4 T read + 8 T
//                         written per column.
//   align_faces_kernel    mogen/datasets/base_dataset.py:121-125: the evaluator replaces the face / shape columns
//                         [156, 309) and [312, 322) of the prediction by the ground truth.
#include <algorithm>

#include "../../include/mcm_b200.h"
#include "common.cuh"

namespace mcm {
namespace {

constexpr int N_POSE = 165, N_EXPR = 100, N_TRANS = 3, N_OUT = N_POSE + N_EXPR + N_TRANS, FEATS = 322;

struct HandoffParams {
  const float* pred;                 // [B, T, 322]
  const int* lengths;                // [B] valid frames per sample (device), or nullptr = T
  const double *mean, *stdv;         // [322]
  const double *w_pose, *w_expr, *w_trans;   // normalised Gaussian taps [2 r + 1]
  int r_pose, r_expr, r_trans;
  int B, T, denorm_f32;
  double *pose, *expr, *trans;       // [B, T, 165] | [B, T, 100] | [B, T, 3]
};

// source column of the 322-vector for output column o (-1: a column that stays zero), visualize.py:241-246
__device__ __forceinline__ int src_col(int o) {
  if (o < 66) return o;                               // global orientation + 21 body joints
  if (o < 69) return 156 + (o - 66);                  // jaw
  if (o < 75) return -1;                              // eye poses: not predicted
  if (o < N_POSE) return 66 + (o - 75);               // hands
  if (o < N_POSE + N_EXPR) return 209 + (o - N_POSE); // expressions
  return 309 + (o - N_POSE - N_EXPR);                 // translation
}

__global__ void __launch_bounds__(256)
handoff_smplx_kernel(const HandoffParams p) {
  extern __shared__ double ext[];                      // [len + 2 r] edge-replicated column
  const int o = blockIdx.x, b = blockIdx.y;
  const int len = p.lengths ? min(max(p.lengths[b], 0), p.T) : p.T;
  const double* w;
  int r, n_cols, oc;
  double* dst;
  if (o < N_POSE) { w = p.w_pose; r = p.r_pose; dst = p.pose; n_cols = N_POSE; oc = o; }
  else if (o < N_POSE + N_EXPR) { w = p.w_expr; r = p.r_expr; dst = p.expr; n_cols = N_EXPR; oc = o - N_POSE; }
  else { w = p.w_trans; r = p.r_trans; dst = p.trans; n_cols = N_TRANS; oc = o - N_POSE - N_EXPR; }
  dst += ((size_t)b * p.T) * n_cols + oc;
  const int sc = src_col(o);
  if (sc < 0 || len == 0) {
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) dst[(size_t)t * n_cols] = 0.0;
    return;
  }
  const float* src = p.pred + ((size_t)b * p.T) * FEATS + sc;
  const double m = p.mean[sc], s = p.stdv[sc];
  for (int i = threadIdx.x; i < len + 2 * r; i += blockDim.x) {
    const int t = min(max(i - r, 0), len - 1);         // mode="nearest"
    const float x = src[(size_t)t * FEATS];
    ext[i] = p.denorm_f32 ? (double)__fadd_rn(__fmul_rn(x, (float)s), (float)m) : __dadd_rn(__dmul_rn((double)x, s), m);
  }
  __syncthreads();
  for (int l = threadIdx.x; l < p.T; l += blockDim.x) {
    double tmp = 0.0;
    if (l < len) {
      const double* c = ext + l + r;
      tmp = __dmul_rn(c[0], w[r]);
      for (int j = -r; j < 0; ++j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(c[j], c[-j]), w[r + j]));
    }
    // expressions / translation are slices of the de-normalised array in the reference: with float32 mean / std they are
    // float32 arrays, and scipy casts the double result of a line to the array's dtype
    if (p.denorm_f32 && o >= N_POSE) tmp = (double)(float)tmp;
    dst[(size_t)l * n_cols] = tmp;
  }
}

// out = pred * std + mean per feature column, as numpy evaluates `pred_motion * std + mean` (tools/visualize.py:221,
// tools/m2d_test.py:203, tools/s2g_test.py:216): float32 arithmetic when both arrays are float32, float64 otherwise; the
// float64 result and its float32 rounding (what `torch.tensor(pred_motion)` assigned into a float32 tensor gives) are
// both written when requested.
__global__ void __launch_bounds__(256)
denorm_kernel(const float* __restrict__ pred, const double* __restrict__ mean, const double* __restrict__ stdv, size_t rows,
              int feats, int f32, double* __restrict__ out64, float* __restrict__ out32) {
  const size_t total = rows * (size_t)feats;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % feats);
    const double v = f32 ? (double)__fadd_rn(__fmul_rn(pred[i], (float)stdv[c]), (float)mean[c])
                         : __dadd_rn(__dmul_rn((double)pred[i], stdv[c]), mean[c]);
    if (out64) out64[i] = v;
    if (out32) out32[i] = (float)v;
  }
}

__global__ void __launch_bounds__(256)
align_faces_kernel(float* __restrict__ pred, const float* __restrict__ motion, size_t rows, int feats) {
  const int lo0 = 156, hi0 = 309, lo1 = 312;
  const int per_row = (hi0 - lo0) + (feats - lo1);
  const size_t total = rows * (size_t)per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / per_row;
    const int k = (int)(i - r * per_row);
    const int c = k < hi0 - lo0 ? lo0 + k : lo1 + (k - (hi0 - lo0));
    pred[r * feats + c] = motion[r * feats + c];
  }
}

}  // namespace
}  // namespace mcm

using namespace mcm;

extern "C" {

int mcm_handoff_smplx(const float* pred, int B, int T, const int* lengths_dev, const double* mean_dev, const double* std_dev,
                      int denorm_f32, const double* w_pose_dev, int r_pose, const double* w_expr_dev, int r_expr,
                      const double* w_trans_dev, int r_trans, double* pose_out, double* expr_out, double* trans_out,
                      void* stream) {
  MCM_CHECK(pred && mean_dev && std_dev && w_pose_dev && w_expr_dev && w_trans_dev && pose_out && expr_out && trans_out,
            "mcm_handoff_smplx: null argument");
  MCM_CHECK(B >= 1 && T >= 1 && r_pose >= 0 && r_expr >= 0 && r_trans >= 0, "mcm_handoff_smplx: bad sizes");
  const int r_max = std::max(r_pose, std::max(r_expr, r_trans));
  const size_t smem = (size_t)(T + 2 * r_max) * sizeof(double);
  MCM_CHECK(smem <= 200 * 1024, "mcm_handoff_smplx: sequence too long for one shared-memory column (T + 2 r <= 25600)");
  if (smem > 48 * 1024)
    MCM_CUDA(cudaFuncSetAttribute(handoff_smplx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HandoffParams p;
  p.pred = pred; p.lengths = lengths_dev; p.mean = mean_dev; p.stdv = std_dev;
  p.w_pose = w_pose_dev; p.w_expr = w_expr_dev; p.w_trans = w_trans_dev;
  p.r_pose = r_pose; p.r_expr = r_expr; p.r_trans = r_trans;
  p.B = B; p.T = T; p.denorm_f32 = denorm_f32;
  p.pose = pose_out; p.expr = expr_out; p.trans = trans_out;
  handoff_smplx_kernel<<<dim3(N_OUT, B), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  MCM_CUDA(cudaGetLastError());
  return 0;
}

int mcm_handoff_denorm(const float* pred, const double* mean_dev, const double* std_dev, long long rows, int feats,
                       int denorm_f32, double* out64, float* out32, void* stream) {
  MCM_CHECK(pred && mean_dev && std_dev && rows >= 1 && feats >= 1 && (out64 || out32), "mcm_handoff_denorm: bad argument");
  const size_t total = (size_t)rows * (size_t)feats;
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  denorm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, mean_dev, std_dev, (size_t)rows, feats,
                                                                          denorm_f32, out64, out32);
  MCM_CUDA(cudaGetLastError());
  return 0;
}

int mcm_handoff_align_faces(float* pred, const float* motion, long long rows, int feats, void* stream) {
  MCM_CHECK(pred && motion && rows >= 1, "mcm_handoff_align_faces: bad argument");
  MCM_CHECK(feats == FEATS, "mcm_handoff_align_faces: the face / shape column ranges are those of the 322-dim SMPL-X vector");
  const size_t total = (size_t)rows * (size_t)((309 - 156) + (feats - 312));
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  align_faces_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, motion, (size_t)rows, feats);
  MCM_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
