// 3D end-point-error loss on the pose pyramid (reference code/models/criterions.py:101-136 with EPE3D_loss :22-46,
// geometry.py:376-399 batch_transform_xyz, :429-445 batch_inverse_project) -- the step that follows the solver in
// every training iteration.  The reference back-projects a 60x80 depth map, transforms the cloud by the ground
// truth and by each of the N estimated poses, and averages the point distances per sample with a Python loop over
// the batch and boolean indexing (one host sync per sample and pose).  Here: one CTA per frame pair, the cloud is
// never materialised, the N sums and the valid count are reduced in a fixed order (deterministic), no host sync.
//
//   loss_b = sum_n  mean_{valid px} || (R_n p + t_n) - (R_gt p + t_gt) ||,   p = [(u-cx)/fx, (v-cy)/fy, 1] depth
//   valid  = target flow has no NaN  and  not (invalid > 0)                          (criterions.py:31-36)
//   a sample without a valid pixel contributes 0                                    (criterions.py:41-42)
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"

namespace dpft {

constexpr int kLossThreads = 256;
constexpr int kLossMaxPoses = 8;

struct PoseLossParams {
  const float *depth, *invalid, *K, *R_gt, *t_gt, *R_est, *t_est;
  const float* g_loss;          // backward only
  float *loss, *g_R, *g_t;      // forward: loss (B); backward: (B,N,3,3), (B,N,3)
  int B, N, h, w;
};

// one row of t + R p; the target and the estimates go through the SAME operation sequence, so an estimate equal to
// the target gives a distance of exactly zero (and a zero gradient, as torch.norm's backward does)
__device__ __forceinline__ float rigid(const float* row, float t, const float* P) {
  return fmaf(row[2], P[2], fmaf(row[1], P[1], fmaf(row[0], P[0], t)));
}

// BWD false: loss.  BWD true: d loss / d (R_est, t_est); every pose gets 12 sums per pair.
template <bool BWD>
__global__ void __launch_bounds__(kLossThreads) pose_epe_kernel(const PoseLossParams p) {
  constexpr int NV = BWD ? 12 : 1;
  __shared__ float s_red[kLossThreads / 32][kLossMaxPoses * NV + 1];
  const int b = blockIdx.x, N = p.N, plane = p.h * p.w;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  float Rg[9], tg[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rg[i] = __ldg(p.R_gt + (size_t)b * 9 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) tg[i] = __ldg(p.t_gt + (size_t)b * 3 + i);
  float acc[kLossMaxPoses * NV];
#pragma unroll
  for (int i = 0; i < kLossMaxPoses * NV; ++i) acc[i] = 0.f;
  float count = 0.f;
  for (int pix = threadIdx.x; pix < plane; pix += kLossThreads) {
    const int y = pix / p.w, x = pix - y * p.w;
    const float z = __ldg(p.depth + (size_t)b * plane + pix);
    const float P[3] = {((float)x - cx) / fx * z, ((float)y - cy) / fy * z, z};
    float G[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) G[r] = rigid(Rg + 3 * r, tg[r], P);
    bool valid = !(G[0] != G[0] || G[1] != G[1] || G[2] != G[2]);
    if (p.invalid) valid = valid && !(__ldg(p.invalid + (size_t)b * plane + pix) > 0.f);
    if (!valid) continue;
    count += 1.f;
#pragma unroll
    for (int n = 0; n < kLossMaxPoses; ++n) {
      if (n >= N) break;
      const float* R = p.R_est + ((size_t)b * N + n) * 9;
      const float* t = p.t_est + ((size_t)b * N + n) * 3;
      float d[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float row[3] = {__ldg(R + 3 * r), __ldg(R + 3 * r + 1), __ldg(R + 3 * r + 2)};
        d[r] = G[r] - rigid(row, __ldg(t + r), P);
      }
      const float e = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (!BWD) {
        acc[n] += e;
      } else {
        // d ||d|| / d(est point) = -d / ||d||  (0 where the points coincide, as torch.norm's backward)
        const float s = e > 0.f ? -1.f / e : 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float gr = s * d[r];
          acc[n * 12 + 3 * r] = fmaf(gr, P[0], acc[n * 12 + 3 * r]);
          acc[n * 12 + 3 * r + 1] = fmaf(gr, P[1], acc[n * 12 + 3 * r + 1]);
          acc[n * 12 + 3 * r + 2] = fmaf(gr, P[2], acc[n * 12 + 3 * r + 2]);
          acc[n * 12 + 9 + r] += gr;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = N * NV;
#pragma unroll
  for (int i = 0; i < kLossMaxPoses * NV; ++i) {
    if (i >= nv) break;
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  {
    float v = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][kLossMaxPoses * NV] = v;
  }
  __syncthreads();
  float n_valid = 0.f;
#pragma unroll
  for (int w = 0; w < kLossThreads / 32; ++w) n_valid += s_red[w][kLossMaxPoses * NV];
  const float scale = n_valid > 0.f ? 1.f / n_valid : 0.f;
  if (!BWD) {
    if (threadIdx.x == 0) {
      float total = 0.f;
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int w = 0; w < kLossThreads / 32; ++w) s += s_red[w][n];
        total += s * scale;
      }
      p.loss[b] = total;
    }
  } else if (threadIdx.x < nv) {
    float s = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) s += s_red[w][threadIdx.x];
    s *= scale * __ldg(p.g_loss + b);
    const int n = threadIdx.x / 12, k = threadIdx.x - 12 * n;
    if (k < 9) p.g_R[((size_t)b * N + n) * 9 + k] = s;
    else p.g_t[((size_t)b * N + n) * 3 + (k - 9)] = s;
  }
}

static int loss_check(const PoseLossParams& p) {
  if (p.B < 1 || p.N < 1 || p.N > kLossMaxPoses || p.h < 1 || p.w < 1)
    return set_error(DPFT_EINVAL, "1 <= B, 1 <= N <= %d, 1 <= h, w", kLossMaxPoses);
  if (!p.depth || !p.K || !p.R_gt || !p.t_gt || !p.R_est || !p.t_est)
    return set_error(DPFT_EINVAL, "depth, K, R_gt, t_gt, R_est and t_est are required");
  return 0;
}

}  // namespace dpft

using namespace dpft;

extern "C" int dpft_pose_epe_loss(const float* depth, const float* invalid, const float* K, const float* R_gt,
                                  const float* t_gt, const float* R_est, const float* t_est, int B, int N, int h, int w,
                                  float* loss, void* stream) {
  PoseLossParams p{};
  p.depth = depth; p.invalid = invalid; p.K = K; p.R_gt = R_gt; p.t_gt = t_gt; p.R_est = R_est; p.t_est = t_est;
  p.loss = loss; p.B = B; p.N = N; p.h = h; p.w = w;
  if (int e = loss_check(p)) return e;
  if (!loss) return set_error(DPFT_EINVAL, "loss is required");
  pose_epe_kernel<false><<<B, kLossThreads, 0, (cudaStream_t)stream>>>(p);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "pose_epe_loss launch: %s", cudaGetErrorString(err));
  return 0;
}

extern "C" int dpft_pose_epe_loss_backward(const float* depth, const float* invalid, const float* K, const float* R_gt,
                                           const float* t_gt, const float* R_est, const float* t_est, int B, int N, int h,
                                           int w, const float* g_loss, float* g_R_est, float* g_t_est, void* stream) {
  PoseLossParams p{};
  p.depth = depth; p.invalid = invalid; p.K = K; p.R_gt = R_gt; p.t_gt = t_gt; p.R_est = R_est; p.t_est = t_est;
  p.g_loss = g_loss; p.g_R = g_R_est; p.g_t = g_t_est; p.B = B; p.N = N; p.h = h; p.w = w;
  if (int e = loss_check(p)) return e;
  if (!g_loss || !g_R_est || !g_t_est) return set_error(DPFT_EINVAL, "g_loss, g_R_est and g_t_est are required");
  pose_epe_kernel<true><<<B, kLossThreads, 0, (cudaStream_t)stream>>>(p);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "pose_epe_loss_backward launch: %s", cudaGetErrorString(err));
  return 0;
}
