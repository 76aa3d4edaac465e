// Depth stage of LeastSquareTracking._preprocess (reference code/models/LeastSquareTracking.py:656-661,
// 668-674, ImagePyramids algorithms.py:1201-1219): inverse depth clamp(1/d, 0, 10), the pixels sitting on the
// batch-global minimum and then on the batch-global maximum of that tensor set to zero (invalid), and the
// max-pooled pyramids (kernel = stride = 2^l, floor sizes) of the inverse depth and, for the ICP term, of the
// metric depth.  The reference spends ~10 launches and two host-visible reductions per frame on this; here it
// is one reduction pass and one pass that writes every level.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"

namespace dpft {

__device__ __forceinline__ float inv_depth(float d) { return fminf(fmaxf(__fdiv_rn(1.f, d), 0.f), 10.f); }

__global__ void __launch_bounds__(256) invdepth_minmax_kernel(const float* __restrict__ depth, size_t n,
                                                              uint32_t* __restrict__ mm) {
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = inv_depth(__ldg(depth + i));
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, f2ord(lo));
    atomicMax(mm + 1, f2ord(hi));
  }
}

struct PyrParams {
  const float* depth;
  float* invd[DPFT_MAX_LEVELS];
  float* dpt[DPFT_MAX_LEVELS];   // optional
  const uint32_t* mm;
  int B, H, W, n_levels;
};

// One thread per 2^(L-1) x 2^(L-1) block of the finest level: writes the block at level 0 and its maxima at
// every coarser level the block fully covers (floor sizes: partial blocks at the right / bottom edge belong to
// no coarse pixel, exactly like nn.MaxPool2d without ceil_mode).
__global__ void __launch_bounds__(128) depth_pyramid_kernel(const PyrParams p) {
  const int S = 1 << (p.n_levels - 1);
  const int bw = (p.W + S - 1) / S, bh = (p.H + S - 1) / S;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= bw * bh) return;
  const int b = blockIdx.y;
  const int bx = idx % bw, by = idx / bw;
  const float lo = ord2f(p.mm[0]);
  const float hi_raw = ord2f(p.mm[1]);
  const float hi = (hi_raw > lo) ? hi_raw : 0.f;   // max AFTER the minima were zeroed (LST:658-661 runs in sequence)
  const size_t plane = (size_t)p.H * p.W;
  for (int l = p.n_levels - 1; l >= 0; --l) {
    // pixels of level l inside this block: (S >> l) per side, each the max over (1 << l)^2 finest pixels
    const int k = 1 << l, per = S >> l;
    const int Hl = p.H >> l, Wl = p.W >> l;
    for (int j = 0; j < per; ++j)
      for (int i = 0; i < per; ++i) {
        const int yl = by * per + j, xl = bx * per + i;
        if (yl >= Hl || xl >= Wl) continue;
        float mi = -CUDART_INF_F, md = -CUDART_INF_F;
        for (int dy = 0; dy < k; ++dy)
          for (int dx = 0; dx < k; ++dx) {
            const float d = __ldg(p.depth + (size_t)b * plane + (size_t)(yl * k + dy) * p.W + (xl * k + dx));
            float v = inv_depth(d);
            if (v == lo || v == hi) v = 0.f;
            mi = fmaxf(mi, v);
            md = fmaxf(md, d);
          }
        p.invd[l][(size_t)b * Hl * Wl + (size_t)yl * Wl + xl] = mi;
        if (p.dpt[l]) p.dpt[l][(size_t)b * Hl * Wl + (size_t)yl * Wl + xl] = md;
      }
  }
}

__global__ void mm_init_kernel(uint32_t* mm) {
  mm[0] = 0xffffffffu;
  mm[1] = 0u;
}

}  // namespace dpft

using namespace dpft;

extern "C" int dpft_preprocess_depth(const float* depth, int B, int H, int W, int n_levels, float* const* invd_out,
                                     float* const* depth_out, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!depth || !invd_out || B < 1 || B > 65535 || H < 1 || W < 1 || n_levels < 1 || n_levels > DPFT_MAX_LEVELS)
    return set_error(DPFT_EINVAL, "depth, invd_out required; 1 <= n_levels <= %d", DPFT_MAX_LEVELS);
  if ((H >> (n_levels - 1)) < 1 || (W >> (n_levels - 1)) < 1) return set_error(DPFT_EINVAL, "image too small for %d levels", n_levels);
  if (!workspace || workspace_bytes < 2 * sizeof(uint32_t)) return set_error(DPFT_ENOSPACE, "workspace of 8 bytes needed");
  cudaStream_t stream = (cudaStream_t)stream_;
  PyrParams p{};
  p.depth = depth; p.mm = (uint32_t*)workspace; p.B = B; p.H = H; p.W = W; p.n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    if (!invd_out[l]) return set_error(DPFT_EINVAL, "invd_out[%d] is NULL", l);
    p.invd[l] = invd_out[l];
    p.dpt[l] = depth_out ? depth_out[l] : nullptr;
  }
  mm_init_kernel<<<1, 1, 0, stream>>>((uint32_t*)workspace);
  const size_t n = (size_t)B * H * W;
  invdepth_minmax_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(depth, n, (uint32_t*)workspace);
  const int S = 1 << (n_levels - 1);
  const int blocks = ((W + S - 1) / S) * ((H + S - 1) / S);
  depth_pyramid_kernel<<<dim3((blocks + 127) / 128, B), 128, 0, stream>>>(p);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "preprocess launch: %s", cudaGetErrorString(err));
  return 0;
}
