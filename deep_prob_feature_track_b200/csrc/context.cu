// Context tensors of the learned networks the reference calls inside the solver loop: the
// kernels that already evaluate the residual write what the CNN wants, once, instead of a residual map that torch
// then takes the absolute value of, concatenates with copies of the feature maps and an upsampled prior.
//
//   dpft_ic_context        DeepRobustEstimator('MultiScale2w') input (algorithms.py:1471-1474): (B,4,H,W) =
//                          [ |r| , x0 , x1 , bilinear_up(wPrior, align_corners=True) ] for a one-channel IC level
//                          (r = x1(warp) - x0, 1e-3 where masked: compute_warped_residual, alg:1919-1957)
//   dpft_uic_icp_context   ScaleNet inputs (algorithms.py:1535-1567, called at alg:677-680): the point-to-plane
//                          residual map (B,1,H,W) (1e-6 where its own mask is set) and n = sqrt(sum_c wres_c^2)
//                          (B,1,H,W) of the masked, uncertainty-weighted feature residual -- ScaleNet only ever uses
//                          compute_rtr(.) = sum over channels of squares, so n carries all it reads of the (B,C,H,W) map.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"
#include "dpft_kernels.h"

namespace dpft {

struct IcCtxParams {
  const float *x0, *x1, *d0, *d1, *K, *pose, *prior;
  const uint8_t *m0, *m1;
  float* ctx;          // (B,4,H,W)
  uint8_t* occ_out;    // (B,H,W) or nullptr
  int H, W, B, hp, wp;
};

// torch's upsample_bilinear2d(align_corners=True): src = dst * (in - 1) / (out - 1)
__device__ __forceinline__ float upsample_prior(const float* __restrict__ q, int hp, int wp, int H, int W, int y, int x) {
  const float rh = (H > 1) ? (float)(hp - 1) / (float)(H - 1) : 0.f, rw = (W > 1) ? (float)(wp - 1) / (float)(W - 1) : 0.f;
  const float h1r = rh * (float)y, w1r = rw * (float)x;
  const int h1 = (int)h1r, w1 = (int)w1r;
  const int h1p = (h1 < hp - 1) ? 1 : 0, w1p = (w1 < wp - 1) ? 1 : 0;
  const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
  const float* r0 = q + h1 * wp + w1;
  const float* r1 = r0 + h1p * wp;
  return h0l * (w0l * __ldg(r0) + w1l * __ldg(r0 + w1p)) + h1l * (w0l * __ldg(r1) + w1l * __ldg(r1 + w1p));
}

__global__ void __launch_bounds__(256) ic_context_kernel(const IcCtxParams p) {
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, plane = H * W;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= plane) return;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const Pose pose = load_pose(p.pose + (size_t)b * 12);
  const int y = pix / W, x = pix - y * W;
  const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
  const float d0 = __ldg(p.d0 + (size_t)b * plane + pix);
  float u, v, inv_z;
  warp_pixel(pose, px, py, d0, fx, fy, cx, cy, u, v, inv_z);
  const Tap tap = make_tap(u, v, H, W);
  const float d1w = sample_exact(p.d1 + (size_t)b * plane, tap, W);
  bool occ = occluded(u, v, inv_z, d1w, H, W);
  if (p.m0) occ = occ || (__ldg(p.m0 + (size_t)b * plane + pix) == 0);
  if (p.m1) occ = occ || !(sample_mask(p.m1 + (size_t)b * plane, tap, W) > 0.f);
  const float* q1 = p.x1 + (size_t)b * plane + tap.o;
  const float f0 = __ldg(p.x0 + (size_t)b * plane + pix);
  const float fr = blend_exact(__ldg(q1), __ldg(q1 + 1), __ldg(q1 + W), __ldg(q1 + W + 1), tap);
  const float r = occ ? 1e-3f : xsub(fr, f0);
  float* c = p.ctx + (size_t)b * 4 * plane + pix;
  c[0] = fabsf(r);
  c[plane] = f0;
  c[2 * (size_t)plane] = __ldg(p.x1 + (size_t)b * plane + pix);
  c[3 * (size_t)plane] = p.prior ? upsample_prior(p.prior + (size_t)b * p.hp * p.wp, p.hp, p.wp, H, W, y, x) : 1.f;
  if (p.occ_out) p.occ_out[(size_t)b * plane + pix] = occ ? 1 : 0;
}

}  // namespace dpft

using namespace dpft;

extern "C" int dpft_ic_context(const dpft_level_t* level, int B, const float* pose, const float* w_prior, int hp, int wp,
                               float* context, uint8_t* occ_out, void* stream) {
  if (!level || B < 1 || B > 65535 || !pose || !context) return set_error(DPFT_EINVAL, "level, pose and context are required");
  const dpft_level_t& L = *level;
  if (!L.x0 || !L.x1 || !L.invd0 || !L.invd1 || !L.K || L.H < 2 || L.W < 2)
    return set_error(DPFT_EINVAL, "x0, x1, invd0, invd1 and K are required (one feature channel)");
  if (w_prior && (hp < 1 || wp < 1)) return set_error(DPFT_EINVAL, "w_prior needs its size");
  IcCtxParams p{};
  p.x0 = L.x0; p.x1 = L.x1; p.d0 = L.invd0; p.d1 = L.invd1; p.K = L.K; p.pose = pose; p.prior = w_prior;
  p.m0 = L.obj_mask0; p.m1 = L.obj_mask1; p.ctx = context; p.occ_out = occ_out;
  p.H = L.H; p.W = L.W; p.B = B; p.hp = hp; p.wp = wp;
  const dim3 grid((unsigned)(((size_t)L.H * L.W + 255) / 256), B);
  ic_context_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "ic_context launch: %s", cudaGetErrorString(err));
  return 0;
}
