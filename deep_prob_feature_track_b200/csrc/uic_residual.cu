// forward_residuals of the U_IC tracker (reference code/models/algorithms.py:725-786, 2119-2137): one warp +
// residual evaluation at a given pose and, per frame pair, sum over valid pixels of the squared weighted
// residuals divided by the number of valid pixels.  Not a hot path (the convergence-basin study calls it),
// so it is two plain passes: the batch-global extremes of the warped sigma first (only with
// remove_tru_sigma), then the masked sums.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"
#include "dpft_kernels.h"

namespace dpft {

struct ResParams {
  const float *x0, *x1, *s0, *s1, *d0, *d1, *K, *pose;
  const uint8_t *m0, *m1;
  const float* icp_r;        // (B,H,W) point-to-plane residual / sigma_icp (1e-6 where its own mask is set) or nullptr
  const uint8_t* icp_occ;    // (B,H,W)
  uint32_t* mm;              // [0..1] sigma0 min/max, [2..3] warped sigma min/max (order-encoded)
  float* sums;               // (B,2): sum of squares, number of masked pixels
  float* norm_out;           // optional (B,H,W): sqrt(sum_c wm_c^2), wm = 1e-6 where masked (ScaleNet's view of the residual)
  float w_icp;
  int H, W, B, C;
};

// PASS 0: extremes of the warped sigma.  PASS 1: masked sums.
template <int PASS, bool TRU>
__global__ void __launch_bounds__(128) uic_residual_kernel(const ResParams p) {
  __shared__ float s_red[4][2];
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, C = p.C, plane = H * W;
  const int pix = blockIdx.x * 128 + threadIdx.x;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const Pose pose = load_pose(p.pose + (size_t)b * 12);
  const size_t pair_off = (size_t)b * C * plane;
  float ssq = 0.f, ninv = 0.f, lo = CUDART_INF_F, hi = -CUDART_INF_F;
  if (pix < plane) {
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float d0 = __ldg(p.d0 + (size_t)b * plane + pix);
    float u, v, inv_z;
    warp_pixel(pose, px, py, d0, fx, fy, cx, cy, u, v, inv_z);
    const Tap tap = make_tap(u, v, H, W);
    if (PASS == 0) {
      for (int c = 0; c < C; ++c) {
        const float sr = sample_exact(p.s1 + pair_off + (size_t)c * plane, tap, W);
        lo = fminf(lo, sr);
        hi = fmaxf(hi, sr);
      }
    } else {
      const float d1w = sample_exact(p.d1 + (size_t)b * plane, tap, W);
      bool occ = occluded(u, v, inv_z, d1w, H, W);
      if (p.m0) occ = occ || (__ldg(p.m0 + (size_t)b * plane + pix) == 0);
      if (p.m1) occ = occ || !(sample_mask(p.m1 + (size_t)b * plane, tap, W) > 0.f);
      if (TRU) {
        const float s0c0 = __ldg(p.s0 + pair_off + pix);
        const float sr0 = sample_exact(p.s1 + pair_off, tap, W);
        occ = occ || (s0c0 == ord2f(p.mm[0])) || (s0c0 == ord2f(p.mm[1])) || (sr0 == ord2f(p.mm[2])) || (sr0 == ord2f(p.mm[3]));
      }
      if (p.icp_occ) occ = occ || (__ldg(p.icp_occ + (size_t)b * plane + pix) != 0);
      if (p.norm_out && occ) p.norm_out[(size_t)b * plane + pix] = sqrtf((float)C * 1e-12f);
      if (!occ) {
        for (int c = 0; c < C; ++c) {
          const size_t k0 = pair_off + (size_t)c * plane;
          const float* q1 = p.x1 + k0 + tap.o;
          const float fr = blend_fast(__ldg(q1), __ldg(q1 + 1), __ldg(q1 + W), __ldg(q1 + W + 1), tap);
          const float sr = sample_exact(p.s1 + k0, tap, W);
          const float s0v = __ldg(p.s0 + k0 + pix);
          const float res = fr - __ldg(p.x0 + k0 + pix);
          const float wres = res * rsqrtf(fmaf(sr, sr, s0v * s0v));
          ssq = fmaf(wres, wres, ssq);
        }
        if (p.norm_out) p.norm_out[(size_t)b * plane + pix] = sqrtf(ssq);
        if (p.icp_r) {
          const float r = p.w_icp * __ldg(p.icp_r + (size_t)b * plane + pix);
          ssq = fmaf(r, r, ssq);
        }
      } else {
        ninv = 1.f;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (PASS == 0) {
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (lane == 0 && lo <= hi) {
      atomicMin(p.mm + 2, f2ord(lo));
      atomicMax(p.mm + 3, f2ord(hi));
    }
  } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
      ninv += __shfl_xor_sync(0xffffffffu, ninv, o);
    }
    if (lane == 0) {
      s_red[warp][0] = ssq;
      s_red[warp][1] = ninv;
    }
    __syncthreads();
    if (threadIdx.x < 2)
      atomicAdd(p.sums + 2 * b + threadIdx.x,
                s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
  }
}

__global__ void residual_prepare_kernel(uint32_t* mm, float* sums, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6) mm[i] = (i & 1) ? 0u : 0xffffffffu;   // [min,max] x {sigma0, warped sigma, depth1}
  if (i < n) sums[i] = 0.f;
}

__global__ void residual_finish_kernel(const float* __restrict__ sums, float* __restrict__ loss, int B, float plane) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) loss[b] = sums[2 * b] / (plane - sums[2 * b + 1]);
}

// defined in uic_forward.cu
void launch_minmax(const float* v, size_t n, uint32_t* mm, cudaStream_t stream);

struct ResPlan {
  size_t off_mm, off_sums, off_vn, off_icp_r, off_icp_occ, off_rec, total;
};

static ResPlan make_res_plan(const dpft_level_t& L, int B, uint32_t flags) {
  ResPlan pl{};
  const size_t plane = (size_t)L.H * L.W;
  const bool icp = flags & DPFT_COMBINE_ICP;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  pl.off_mm = take(8 * sizeof(uint32_t));
  pl.off_sums = take((size_t)B * 2 * sizeof(float));
  pl.off_vn = take(icp ? 6 * (size_t)B * plane * sizeof(float) : 0);
  pl.off_icp_r = take(icp ? (size_t)B * plane * sizeof(float) : 0);
  pl.off_icp_occ = take(icp ? (size_t)B * plane : 0);
  pl.off_rec = take(icp ? (size_t)B * 28 * sizeof(float) : 0);
  pl.total = off;
  return pl;
}

}  // namespace dpft

using namespace dpft;

// ScaleNet's inputs at `pose` (see context.cu): icp_r (B,1,H,W) and feat_norm (B,1,H,W).  Same passes as
// dpft_uic_residual_loss, with the two maps as outputs; the feature mask does NOT include the ICP mask here
// (compose_residuals, alg:1960-1989, knows nothing of the ICP term).  Workspace: dpft_uic_residual_workspace_bytes
// with DPFT_COMBINE_ICP.
extern "C" int dpft_uic_icp_context(const dpft_level_t* level, int B, int C, uint32_t flags, const float* pose,
                                    float* icp_r, float* feat_norm, void* workspace, size_t workspace_bytes,
                                    void* stream_) {
  if (!level || B < 1 || B > 65535 || C < 1 || !pose || !icp_r || !feat_norm || !workspace)
    return set_error(DPFT_EINVAL, "level, pose, icp_r, feat_norm and workspace are required");
  const dpft_level_t& L = *level;
  if (!L.x0 || !L.x1 || !L.sigma0 || !L.sigma1 || !L.invd0 || !L.invd1 || !L.K || !L.depth0 || !L.depth1 || L.H < 2 || L.W < 2)
    return set_error(DPFT_EINVAL, "x0, x1, sigma0, sigma1, invd0, invd1, depth0, depth1 and K are required");
  const bool tru = flags & DPFT_REMOVE_TRU_SIGMA;
  const ResPlan pl = make_res_plan(L, B, flags | DPFT_COMBINE_ICP);
  if (workspace_bytes < pl.total) return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, pl.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  const size_t plane = (size_t)L.H * L.W;
  ResParams p{};
  p.x0 = L.x0; p.x1 = L.x1; p.s0 = L.sigma0; p.s1 = L.sigma1; p.d0 = L.invd0; p.d1 = L.invd1; p.K = L.K; p.pose = pose;
  p.m0 = L.obj_mask0; p.m1 = L.obj_mask1;
  p.mm = (uint32_t*)(ws + pl.off_mm);
  p.sums = (float*)(ws + pl.off_sums);
  p.norm_out = feat_norm;
  p.H = L.H; p.W = L.W; p.B = B; p.C = C;
  residual_prepare_kernel<<<(2 * B + 255) / 256, 256, 0, stream>>>(p.mm, p.sums, 2 * B);
  float* vn = (float*)(ws + pl.off_vn);
  launch_minmax(L.depth1, (size_t)B * plane, p.mm + 4, stream);
  launch_vertex_normal(L.depth1, L.K, p.mm + 4, vn, vn + 3 * (size_t)B * plane, B, L.H, L.W, stream);
  // inside the solver loop the ICP term honours the object masks (alg:668-672)
  launch_icp_term(L.depth0, L.K, vn, vn + 3 * (size_t)B * plane, pose, L.obj_mask0, L.obj_mask1, (float*)(ws + pl.off_rec),
                  nullptr, icp_r, nullptr, B, L.H, L.W, stream);
  const dim3 grid((unsigned)((plane + 127) / 128), B);
  if (tru) {
    launch_minmax(L.sigma0, (size_t)B * C * plane, p.mm, stream);
    uic_residual_kernel<0, true><<<grid, 128, 0, stream>>>(p);
    uic_residual_kernel<1, true><<<grid, 128, 0, stream>>>(p);
  } else {
    uic_residual_kernel<1, false><<<grid, 128, 0, stream>>>(p);
  }
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "context launch: %s", cudaGetErrorString(err));
  return 0;
}

extern "C" size_t dpft_uic_residual_workspace_bytes(const dpft_level_t* level, int B, int C, uint32_t flags) {
  (void)C;
  if (!level || B < 1) {
    set_error(DPFT_EINVAL, "bad arguments");
    return 0;
  }
  return make_res_plan(*level, B, flags).total;
}

extern "C" int dpft_uic_residual_loss(const dpft_level_t* level, int B, int C, uint32_t flags, float w_icp,
                                      const float* pose, float* loss, void* workspace, size_t workspace_bytes,
                                      void* stream_) {
  if (!level || B < 1 || B > 65535 || C < 1 || !pose || !loss || !workspace)
    return set_error(DPFT_EINVAL, "level, pose, loss and workspace are required");
  const dpft_level_t& L = *level;
  if (!L.x0 || !L.x1 || !L.sigma0 || !L.sigma1 || !L.invd0 || !L.invd1 || !L.K || L.H < 2 || L.W < 2)
    return set_error(DPFT_EINVAL, "x0, x1, sigma0, sigma1, invd0, invd1 and K are required");
  const bool icp = flags & DPFT_COMBINE_ICP, tru = flags & DPFT_REMOVE_TRU_SIGMA;
  if (icp && (!L.depth0 || !L.depth1)) return set_error(DPFT_EINVAL, "DPFT_COMBINE_ICP needs depth0 and depth1");
  const ResPlan pl = make_res_plan(L, B, flags);
  if (workspace_bytes < pl.total) return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, pl.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  const size_t plane = (size_t)L.H * L.W;
  ResParams p{};
  p.x0 = L.x0; p.x1 = L.x1; p.s0 = L.sigma0; p.s1 = L.sigma1; p.d0 = L.invd0; p.d1 = L.invd1; p.K = L.K; p.pose = pose;
  p.m0 = L.obj_mask0; p.m1 = L.obj_mask1;
  p.mm = (uint32_t*)(ws + pl.off_mm);
  p.sums = (float*)(ws + pl.off_sums);
  p.w_icp = w_icp; p.H = L.H; p.W = L.W; p.B = B; p.C = C;
  residual_prepare_kernel<<<(2 * B + 255) / 256, 256, 0, stream>>>(p.mm, p.sums, 2 * B);
  if (icp) {
    float* vn = (float*)(ws + pl.off_vn);
    float* icp_r = (float*)(ws + pl.off_icp_r);
    uint8_t* icp_occ = (uint8_t*)(ws + pl.off_icp_occ);
    launch_minmax(L.depth1, (size_t)B * plane, p.mm + 4, stream);
    launch_vertex_normal(L.depth1, L.K, p.mm + 4, vn, vn + 3 * (size_t)B * plane, B, L.H, L.W, stream);
    // the reference leaves the object masks out of this ICP evaluation (algorithms.py:768-769)
    launch_icp_term(L.depth0, L.K, vn, vn + 3 * (size_t)B * plane, pose, nullptr, nullptr, (float*)(ws + pl.off_rec),
                    icp_occ, icp_r, nullptr, B, L.H, L.W, stream);
    p.icp_r = icp_r;
    p.icp_occ = icp_occ;
  }
  const dim3 grid((unsigned)((plane + 127) / 128), B);
  if (tru) {
    launch_minmax(L.sigma0, (size_t)B * C * plane, p.mm, stream);
    uic_residual_kernel<0, true><<<grid, 128, 0, stream>>>(p);
    uic_residual_kernel<1, true><<<grid, 128, 0, stream>>>(p);
  } else {
    uic_residual_kernel<1, false><<<grid, 128, 0, stream>>>(p);
  }
  residual_finish_kernel<<<(B + 127) / 128, 128, 0, stream>>>(p.sums, loss, B, (float)plane);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "residual launch: %s", cudaGetErrorString(err));
  return 0;
}
