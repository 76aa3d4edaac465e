// IC tracker (reference TrustRegionBase, code/models/algorithms.py:45-139, with DirectSolverNet :1604-1691).
//
// The inverse-compositional property is used as the reference uses it: J = grad(x0) . d(u,v)/d(xi) and
// J^T W J are formed once per level, only J^T W r changes with the pose.  Because the reference calls learned
// networks in the middle of the loop (the convolutional M-estimator after the first residual, the damping MLP
// on the "residual volume" of 10 trial poses in every iteration), the path is split into four entry points
// that the Python module strings together around those cuDNN / cuBLAS calls:
//
//   dpft_ic_residual        r = x1(warp(pose)) - x0, 1e-3 where masked, and the mask   (alg:1919-1957)
//   dpft_ic_normal_matrix   A = sum_c,pix w J J^T                                        (alg:71-75, 116-121)
//   dpft_ic_rhs             b_s = sum w J r(pose_s) for S poses in one launch            (alg:1623-1624, 1682-1683)
//   dpft_ic_update          pose_s = pose (+) solve(A + damping_s, b)                    (alg:1629, 1678-1680, 1689-1691)
//
// J is never materialised: per pixel J_c = gx_c ju + gy_c jv, so only sum_c w gx^2, sum w gx gy, sum w gy^2,
// sum w gx r and sum w gy r are needed before the 6-vectors ju, jv come in.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"
#include "dpft_kernels.h"

namespace dpft {

struct IcParams {
  const float *x0, *x1, *d0, *d1, *K, *gx, *gy, *w;   // w: (B,C,H,W) weights or nullptr (= 1)
  const uint8_t *m0, *m1;
  const float* pose;   // (S,B,12)
  float* r_out;        // (B,C,H,W)
  uint8_t* occ_out;    // (B,H,W)
  float* out;          // normal matrix: (B,21); rhs: (S,B,6)
  int H, W, B, C, ppt;
};

// MODE 0: residual map + mask.  MODE 1: J^T W J.  MODE 2: J^T W r (blockIdx.z = pose sample).
template <int MODE>
__global__ void __launch_bounds__(128, 4) ic_kernel(const IcParams p) {
  __shared__ float s_red[4][21];
  const int b = blockIdx.y, s = blockIdx.z;
  const int H = p.H, W = p.W, C = p.C, plane = H * W;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const size_t pair_off = (size_t)b * C * plane;
  Pose pose;
  if (MODE != 1) pose = load_pose(p.pose + ((size_t)s * p.B + b) * 12);
  constexpr int NACC = (MODE == 1) ? 21 : 6;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  for (int i = 0; i < p.ppt; ++i) {
    const int pix = (blockIdx.x * p.ppt + i) * 128 + threadIdx.x;
    if (pix >= plane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float d0 = __ldg(p.d0 + (size_t)b * plane + pix);
    float ju[6], jv[6];
    warp_rows(px, py, d0, fx, fy, ju, jv);
    if (MODE == 1) {
      float saa = 0.f, sab = 0.f, sbb = 0.f;
      for (int c = 0; c < C; ++c) {
        const size_t k0 = pair_off + (size_t)c * plane + pix;
        const float gx = __ldg(p.gx + k0), gy = __ldg(p.gy + k0), w = p.w ? __ldg(p.w + k0) : 1.f;
        saa = fmaf(w * gx, gx, saa);
        sab = fmaf(w * gx, gy, sab);
        sbb = fmaf(w * gy, gy, sbb);
      }
      accumulate_system(acc, ju, jv, saa, sab, sbb, 0.f, 0.f);
      continue;
    }
    float u, v, inv_z;
    warp_pixel(pose, px, py, d0, fx, fy, cx, cy, u, v, inv_z);
    const Tap tap = make_tap(u, v, H, W);
    const float d1w = sample_exact(p.d1 + (size_t)b * plane, tap, W);
    bool occ = occluded(u, v, inv_z, d1w, H, W);
    if (p.m0) occ = occ || (__ldg(p.m0 + (size_t)b * plane + pix) == 0);
    if (p.m1) occ = occ || !(sample_mask(p.m1 + (size_t)b * plane, tap, W) > 0.f);
    float sar = 0.f, sbr = 0.f;
    for (int c = 0; c < C; ++c) {
      const size_t k0 = pair_off + (size_t)c * plane;
      const float* q1 = p.x1 + k0 + tap.o;
      // the residual map feeds a CNN in the reference, keep it to the oracle's rounding
      const float fr = blend_exact(__ldg(q1), __ldg(q1 + 1), __ldg(q1 + W), __ldg(q1 + W + 1), tap);
      const float r = occ ? 1e-3f : xsub(fr, __ldg(p.x0 + k0 + pix));
      if (MODE == 0) {
        p.r_out[k0 + pix] = r;
      } else {
        const float w = p.w ? __ldg(p.w + k0 + pix) : 1.f;
        const float wr = w * r;
        sar = fmaf(__ldg(p.gx + k0 + pix), wr, sar);
        sbr = fmaf(__ldg(p.gy + k0 + pix), wr, sbr);
      }
    }
    if (MODE == 0) {
      p.occ_out[(size_t)b * plane + pix] = occ ? 1 : 0;
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        float a = acc[k];
        if (k != 4) a = fmaf(sar, ju[k], a);
        if (k != 3) a = fmaf(sbr, jv[k], a);
        acc[k] = a;
      }
    }
  }
  if (MODE == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    const float v = s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x];
    float* dst = (MODE == 1) ? p.out + (size_t)b * 21 : p.out + ((size_t)s * p.B + b) * 6;
    atomicAdd(dst + threadIdx.x, v);
  }
}

// pose_out[s][b] = pose_in[b] (+) solve(H_s, rhs[b]) with
//   mode 0: H = A + eps I                       (lev_mar_H, alg:2094-2103; S = 1)
//   mode 1: H = A + lambda_s diag(A) + eps I     (trial poses of the residual volume, alg:1676-1680)
//   mode 2: H = A + diag(damp[b]) + eps I        (learned damping, alg:1688-1691; S = 1)
// eps = 1e-6 trace(A).
__global__ void ic_update_kernel(const float* __restrict__ A21, const float* __restrict__ rhs, const float* __restrict__ lambdas,
                                 const float* __restrict__ damp, const float* __restrict__ pose_in, float* __restrict__ pose_out,
                                 float* __restrict__ H_out, int32_t* __restrict__ status, int B, int S, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S) return;
  const int s = i / B, b = i - s * B;
  double A[21], r[6];
  for (int k = 0; k < 21; ++k) A[k] = (double)A21[(size_t)b * 21 + k];
  for (int k = 0; k < 6; ++k) r[k] = (double)rhs[(size_t)b * 6 + k];
  double tr = 0.0;
  for (int k = 0; k < 6; ++k) tr += A[tri(k, k)];
  for (int k = 0; k < 6; ++k) {
    double d = tr * 1e-6;
    if (mode == 1) d += (double)lambdas[s] * A[tri(k, k)];
    if (mode == 2) d += (double)damp[(size_t)b * 6 + k];
    A[tri(k, k)] += d;
  }
  if (H_out)
    for (int k = 0; k < 21; ++k) H_out[((size_t)s * B + b) * 21 + k] = (float)A[k];
  double xi[6];
  const bool ok = solve_and_update(A, r, false, pose_in + (size_t)b * 12, pose_out + ((size_t)s * B + b) * 12, xi);
  if (!ok && status) atomicOr(status, DPFT_ST_SINGULAR);
}

static int ic_check(const dpft_level_t* L, int B, int C) {
  if (!L || B < 1 || B > 65535 || C < 1) return set_error(DPFT_EINVAL, "bad problem size");
  if (!L->x0 || !L->x1 || !L->invd0 || !L->invd1 || !L->K || L->H < 2 || L->W < 2)
    return set_error(DPFT_EINVAL, "x0, x1, invd0, invd1 and K are required");
  return 0;
}

static IcParams ic_params(const dpft_level_t& L, int B, int C, long work_items) {
  IcParams p{};
  p.x0 = L.x0; p.x1 = L.x1; p.d0 = L.invd0; p.d1 = L.invd1; p.K = L.K; p.m0 = L.obj_mask0; p.m1 = L.obj_mask1;
  p.H = L.H; p.W = L.W; p.B = B; p.C = C;
  const long want_threads = 148L * 2048 * 2;
  long ppt = (work_items + want_threads - 1) / want_threads;
  p.ppt = (int)std::max(1L, std::min(ppt, 8L));
  return p;
}

}  // namespace dpft

using namespace dpft;

// a launch that failed (bad configuration, a sticky error of an earlier call) must not read as success
static int launched(const char* what) {
  const cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : dpft::set_error((int)err, "%s launch: %s", what, cudaGetErrorString(err));
}

extern "C" int dpft_ic_gradients(const dpft_level_t* level, int B, int C, float* gx, float* gy, void* stream) {
  if (int e = ic_check(level, B, C)) return e;
  if (!gx || !gy) return set_error(DPFT_EINVAL, "gx and gy are required");
  launch_sobel_unit(level->x0, gx, gy, B * C, level->H, level->W, (cudaStream_t)stream);
  return launched("ic_gradients");
}

extern "C" int dpft_ic_residual(const dpft_level_t* level, int B, int C, const float* pose, float* r_out,
                                uint8_t* occ_out, void* stream) {
  if (int e = ic_check(level, B, C)) return e;
  if (!pose || !r_out || !occ_out) return set_error(DPFT_EINVAL, "pose, r_out and occ_out are required");
  const long plane = (long)level->H * level->W;
  IcParams p = ic_params(*level, B, C, (long)B * plane);
  p.pose = pose; p.r_out = r_out; p.occ_out = occ_out;
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B, 1);
  ic_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  return launched("ic_residual");
}

extern "C" int dpft_ic_normal_matrix(const dpft_level_t* level, int B, int C, const float* gx, const float* gy,
                                     const float* weights, float* A21, void* stream_) {
  if (int e = ic_check(level, B, C)) return e;
  if (!gx || !gy || !A21) return set_error(DPFT_EINVAL, "gx, gy and A21 are required");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long plane = (long)level->H * level->W;
  IcParams p = ic_params(*level, B, C, (long)B * plane);
  p.gx = gx; p.gy = gy; p.w = weights; p.out = A21;
  cudaMemsetAsync(A21, 0, (size_t)B * 21 * sizeof(float), stream);
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B, 1);
  ic_kernel<1><<<grid, 128, 0, stream>>>(p);
  return launched("ic_normal_matrix");
}

extern "C" int dpft_ic_rhs(const dpft_level_t* level, int B, int C, const float* gx, const float* gy,
                           const float* weights, const float* poses, int S, float* rhs, void* stream_) {
  if (int e = ic_check(level, B, C)) return e;
  if (!gx || !gy || !poses || !rhs || S < 1 || S > 65535) return set_error(DPFT_EINVAL, "gx, gy, poses, rhs required; 1 <= S");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long plane = (long)level->H * level->W;
  IcParams p = ic_params(*level, B, C, (long)B * plane * S);
  p.gx = gx; p.gy = gy; p.w = weights; p.pose = poses; p.out = rhs;
  cudaMemsetAsync(rhs, 0, (size_t)S * B * 6 * sizeof(float), stream);
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B, S);
  ic_kernel<2><<<grid, 128, 0, stream>>>(p);
  return launched("ic_rhs");
}

extern "C" int dpft_ic_update(int B, int S, int mode, const float* A21, const float* rhs, const float* lambdas,
                              const float* damp, const float* pose_in, float* pose_out, float* H_out, int32_t* status,
                              void* stream) {
  if (B < 1 || S < 1 || !A21 || !rhs || !pose_in || !pose_out) return set_error(DPFT_EINVAL, "A21, rhs, pose_in, pose_out required");
  if (mode < 0 || mode > 2 || (mode == 1 && !lambdas) || (mode == 2 && !damp) || (mode != 1 && S != 1))
    return set_error(DPFT_EINVAL, "mode 0/2 need S == 1; mode 1 needs lambdas; mode 2 needs damp");
  ic_update_kernel<<<(B * S + 63) / 64, 64, 0, (cudaStream_t)stream>>>(A21, rhs, lambdas, damp, pose_in, pose_out, H_out,
                                                                        status, B, S, mode);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "ic_update launch: %s", cudaGetErrorString(err));
  return 0;
}
