// IC tracker backward: the adjoint of every dpft_ic_* entry point of ic_path.cu, so that torch.autograd can
// chain them around the learned networks the reference calls inside the loop (convolutional M-estimator,
// damping MLP) -- what autograd derives from reference code/models/algorithms.py:45-121, 1604-1691, 1919-1957.
//
//   dpft_ic_gradients_backward      d/d(gx,gy) -> d/d(x0)                  (unit Sobel adjoint)
//   dpft_ic_residual_backward       d/d(r)     -> d/d(x0), d/d(x1), d/d(pose)
//   dpft_ic_normal_matrix_backward  d/d(A21)   -> d/d(gx), d/d(gy), d/d(w)
//   dpft_ic_rhs_backward            d/d(rhs_s) -> d/d(gx), d/d(gy), d/d(w), d/d(x0), d/d(x1), d/d(pose_s)
//   dpft_ic_update_backward         d/d(pose_out_s) -> d/d(A21), d/d(rhs), d/d(damp), d/d(pose_in)
//
// The mask and the 1e-3 fill are stop-gradients: a masked pixel passes nothing back through r, but its constant
// residual still multiplies J in J^T W r, so gx, gy and w do receive a gradient there.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"
#include "dpft_kernels.h"
#include "pose_adjoint.cuh"

namespace dpft {

struct IcBwdParams {
  const float *x0, *x1, *d0, *d1, *K, *gx, *gy, *w;
  const uint8_t *m0, *m1;
  const float* pose;    // (S,B,12)
  const float* g_out;   // residual: d/d(r) (B,C,H,W); rhs: d/d(rhs) (S,B,6); normal matrix: d/d(A21) (B,21)
  float *g_gx, *g_gy, *g_w, *g_x0, *g_x1, *g_pose;
  int H, W, B, C, ppt;
};

// FROM_RHS false: adjoint of ic_kernel<0> (residual map).  FROM_RHS true: adjoint of ic_kernel<2> (J^T W r),
// blockIdx.z = pose sample.  All outputs are accumulated with red.global.add (several samples share a pixel).
template <bool FROM_RHS>
__global__ void __launch_bounds__(128, 4) ic_bwd_warp_kernel(const IcBwdParams p) {
  __shared__ float s_red[4][12];
  const int b = blockIdx.y, s = blockIdx.z;
  const int H = p.H, W = p.W, C = p.C, plane = H * W;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const size_t pair_off = (size_t)b * C * plane;
  const Pose pose = load_pose(p.pose + ((size_t)s * p.B + b) * 12);
  float gb[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) gb[k] = FROM_RHS ? __ldg(p.g_out + ((size_t)s * p.B + b) * 6 + k) : 0.f;
  float gR[9], gt[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) gR[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) gt[i] = 0.f;

  for (int i = 0; i < p.ppt; ++i) {
    const int pix = (blockIdx.x * p.ppt + i) * 128 + threadIdx.x;
    if (pix >= plane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float d0 = __ldg(p.d0 + (size_t)b * plane + pix);
    // forward geometry in the forward kernels' arithmetic: the mask must come out identical
    float w3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      w3[k] = xadd(xadd(xadd(xmul(pose.r[3 * k], px), xmul(pose.r[3 * k + 1], py)), pose.r[3 * k + 2]), xmul(pose.t[k], d0));
    const float u = xadd(xmul(xdiv(w3[0], w3[2]), fx), cx);
    const float v = xadd(xmul(xdiv(w3[1], w3[2]), fy), cy);
    const float inv_z = xdiv(d0, w3[2]);
    const Tap tap = make_tap(u, v, H, W);
    const float d1w = sample_exact(p.d1 + (size_t)b * plane, tap, W);
    bool occ = occluded(u, v, inv_z, d1w, H, W);
    if (p.m0) occ = occ || (__ldg(p.m0 + (size_t)b * plane + pix) == 0);
    if (p.m1) occ = occ || !(sample_mask(p.m1 + (size_t)b * plane, tap, W) > 0.f);
    float pu = 0.f, pv = 0.f;
    if (FROM_RHS) {
      float ju[6], jv[6];
      warp_rows(px, py, d0, fx, fy, ju, jv);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        pu = fmaf(gb[k], ju[k], pu);
        pv = fmaf(gb[k], jv[k], pv);
      }
    }
    const float tyn = tap.wa + tap.wb, tys = tap.wc + tap.wd, txl = tap.wa + tap.wc, txr = tap.wb + tap.wd;
    float g_ix = 0.f, g_iy = 0.f;
    for (int c = 0; c < C; ++c) {
      const size_t k0 = pair_off + (size_t)c * plane + pix;
      const size_t k1 = pair_off + (size_t)c * plane + tap.o;
      float gres;
      float xa = 0.f, xb = 0.f, xc = 0.f, xd = 0.f;
      if (!occ || FROM_RHS) {
        xa = __ldg(p.x1 + k1); xb = __ldg(p.x1 + k1 + 1); xc = __ldg(p.x1 + k1 + W); xd = __ldg(p.x1 + k1 + W + 1);
      }
      if (FROM_RHS) {
        const float gxc = __ldg(p.gx + k0), gyc = __ldg(p.gy + k0), wc = p.w ? __ldg(p.w + k0) : 1.f;
        const float r = occ ? 1e-3f : xsub(blend_exact(xa, xb, xc, xd, tap), __ldg(p.x0 + k0));
        const float jp = fmaf(gxc, pu, gyc * pv);      // (J_c . d/d(rhs)) at this pixel
        atomicAdd(p.g_gx + k0, wc * r * pu);
        atomicAdd(p.g_gy + k0, wc * r * pv);
        if (p.g_w) atomicAdd(p.g_w + k0, r * jp);
        gres = wc * jp;
      } else {
        gres = __ldg(p.g_out + k0);
      }
      if (occ) continue;
      atomicAdd(p.g_x0 + k0, -gres);
      atomicAdd(p.g_x1 + k1, tap.wa * gres);
      atomicAdd(p.g_x1 + k1 + 1, tap.wb * gres);
      atomicAdd(p.g_x1 + k1 + W, tap.wc * gres);
      atomicAdd(p.g_x1 + k1 + W + 1, tap.wd * gres);
      g_ix = fmaf(gres, fmaf(xb - xa, tyn, (xd - xc) * tys), g_ix);
      g_iy = fmaf(gres, fmaf(xc - xa, txl, (xd - xb) * txr), g_iy);
    }
    // grid_sampler gives zero coordinate gradient on and outside the border
    const float ixu = xmul(xmul(xadd(xsub(xdiv(u, 0.5f * (float)(W - 1)), 1.f), 1.f), 0.5f), (float)(W - 1));
    const float iyu = xmul(xmul(xadd(xsub(xdiv(v, 0.5f * (float)(H - 1)), 1.f), 1.f), 0.5f), (float)(H - 1));
    const float gu = (ixu > 0.f && ixu < (float)(W - 1)) ? g_ix : 0.f;
    const float gv = (iyu > 0.f && iyu < (float)(H - 1)) ? g_iy : 0.f;
    const float iz = 1.f / w3[2];
    const float gw0 = gu * fx * iz, gw1 = gv * fy * iz;
    const float gw[3] = {gw0, gw1, -(gw0 * w3[0] + gw1 * w3[1]) * iz};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      gR[3 * r] = fmaf(gw[r], px, gR[3 * r]);
      gR[3 * r + 1] = fmaf(gw[r], py, gR[3 * r + 1]);
      gR[3 * r + 2] += gw[r];
      gt[r] = fmaf(gw[r], d0, gt[r]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float sum = i < 9 ? gR[i] : gt[i - 9];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_red[warp][i] = sum;
  }
  __syncthreads();
  if (threadIdx.x < 12)
    atomicAdd(p.g_pose + ((size_t)s * p.B + b) * 12 + threadIdx.x,
              s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
}

// Adjoint of ic_kernel<1>: A21[i<=j] = sum w J_i J_j with J = gx ju + gy jv.  Plain stores (one writer per element).
__global__ void __launch_bounds__(128, 4) ic_bwd_normal_kernel(const IcBwdParams p) {
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, C = p.C, plane = H * W;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const size_t pair_off = (size_t)b * C * plane;
  float G[21];
#pragma unroll
  for (int i = 0; i < 21; ++i) G[i] = __ldg(p.g_out + (size_t)b * 21 + i);
  for (int i = 0; i < p.ppt; ++i) {
    const int pix = (blockIdx.x * p.ppt + i) * 128 + threadIdx.x;
    if (pix >= plane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float d0 = __ldg(p.d0 + (size_t)b * plane + pix);
    float ju[6], jv[6];
    warp_rows(px, py, d0, fx, fy, ju, jv);
    float quu = 0.f, quv = 0.f, qvv = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) {
        const float g = G[tri(r, c)];
        quu = fmaf(g, ju[r] * ju[c], quu);
        quv = fmaf(g, fmaf(ju[r], jv[c], jv[r] * ju[c]), quv);
        qvv = fmaf(g, jv[r] * jv[c], qvv);
      }
    for (int c = 0; c < C; ++c) {
      const size_t k0 = pair_off + (size_t)c * plane + pix;
      const float gx = __ldg(p.gx + k0), gy = __ldg(p.gy + k0), w = p.w ? __ldg(p.w + k0) : 1.f;
      p.g_gx[k0] = w * fmaf(2.f * gx, quu, gy * quv);
      p.g_gy[k0] = w * fmaf(gx, quv, 2.f * gy * qvv);
      if (p.g_w) p.g_w[k0] = fmaf(gx * gx, quu, fmaf(gx * gy, quv, gy * gy * qvv));
    }
  }
}

// Adjoint of ic_update_kernel, one thread per pair looping over the S samples.
__global__ void ic_update_bwd_kernel(const float* __restrict__ A21, const float* __restrict__ rhs, const float* __restrict__ lambdas,
                                     const float* __restrict__ damp, const float* __restrict__ pose_in,
                                     const float* __restrict__ g_pose_out, float* __restrict__ g_A21, float* __restrict__ g_rhs,
                                     float* __restrict__ g_damp, float* __restrict__ g_pose_in, int B, int S, int mode) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double A[21], r[6], gA[21], gr[6], gd[6], gp[12];
  for (int k = 0; k < 21; ++k) {
    A[k] = (double)A21[(size_t)b * 21 + k];
    gA[k] = 0.0;
  }
  for (int k = 0; k < 6; ++k) {
    r[k] = (double)rhs[(size_t)b * 6 + k];
    gr[k] = gd[k] = 0.0;
  }
  for (int k = 0; k < 12; ++k) gp[k] = 0.0;
  double tr = 0.0;
  for (int k = 0; k < 6; ++k) tr += A[tri(k, k)];
  for (int s = 0; s < S; ++s) {
    double Hd[21];
    for (int k = 0; k < 21; ++k) Hd[k] = A[k];
    for (int k = 0; k < 6; ++k) {
      double d = tr * 1e-6;
      if (mode == 1) d += (double)lambdas[s] * A[tri(k, k)];
      if (mode == 2) d += (double)damp[(size_t)b * 6 + k];
      Hd[tri(k, k)] += d;
    }
    double xi[6], lam[6];
    solve_update_adjoint(Hd, r, pose_in + (size_t)b * 12, g_pose_out + ((size_t)s * B + b) * 12, xi, lam, gp);
    // d/d(Hd) = -lam xi^T; Hd_ij = Hd_ji = A21[tri(i,j)], the diagonal also carries the damping and 1e-6 tr(A)
    double trH = 0.0;
    for (int k = 0; k < 6; ++k) trH += -lam[k] * xi[k];
    for (int i = 0; i < 6; ++i) {
      gr[i] += lam[i];
      for (int j = i + 1; j < 6; ++j) gA[tri(i, j)] += -(lam[i] * xi[j] + lam[j] * xi[i]);
      const double hii = -lam[i] * xi[i];
      gA[tri(i, i)] += hii * (mode == 1 ? 1.0 + (double)lambdas[s] : 1.0) + 1e-6 * trH;
      if (mode == 2) gd[i] += hii;
    }
  }
  for (int k = 0; k < 21; ++k) g_A21[(size_t)b * 21 + k] = (float)gA[k];
  for (int k = 0; k < 6; ++k) g_rhs[(size_t)b * 6 + k] = (float)gr[k];
  if (g_damp)
    for (int k = 0; k < 6; ++k) g_damp[(size_t)b * 6 + k] = (float)gd[k];
  for (int k = 0; k < 12; ++k) g_pose_in[(size_t)b * 12 + k] = (float)gp[k];
}

static int icb_check(const dpft_level_t* L, int B, int C) {
  if (!L || B < 1 || B > 65535 || C < 1) return set_error(DPFT_EINVAL, "bad problem size");
  if (!L->x0 || !L->x1 || !L->invd0 || !L->invd1 || !L->K || L->H < 2 || L->W < 2)
    return set_error(DPFT_EINVAL, "x0, x1, invd0, invd1 and K are required");
  return 0;
}

static IcBwdParams icb_params(const dpft_level_t& L, int B, int C, long work_items) {
  IcBwdParams p{};
  p.x0 = L.x0; p.x1 = L.x1; p.d0 = L.invd0; p.d1 = L.invd1; p.K = L.K; p.m0 = L.obj_mask0; p.m1 = L.obj_mask1;
  p.H = L.H; p.W = L.W; p.B = B; p.C = C;
  const long want_threads = 148L * 2048 * 2;
  const long ppt = (work_items + want_threads - 1) / want_threads;
  p.ppt = (int)std::max(1L, std::min(ppt, 8L));
  return p;
}

static int launch_status(const char* what) {
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "%s launch: %s", what, cudaGetErrorString(err));
  return 0;
}

}  // namespace dpft

using namespace dpft;

extern "C" int dpft_ic_gradients_backward(const dpft_level_t* level, int B, int C, const float* g_gx, const float* g_gy,
                                          float* g_x0, void* stream) {
  if (int e = icb_check(level, B, C)) return e;
  if (!g_gx || !g_gy || !g_x0) return set_error(DPFT_EINVAL, "g_gx, g_gy and g_x0 are required");
  launch_sobel_unit_bwd(level->x0, g_gx, g_gy, g_x0, B * C, level->H, level->W, (cudaStream_t)stream);
  return launch_status("ic_gradients_backward");
}

extern "C" int dpft_ic_residual_backward(const dpft_level_t* level, int B, int C, const float* pose, const float* g_r,
                                         float* g_x0, float* g_x1, float* g_pose, void* stream) {
  if (int e = icb_check(level, B, C)) return e;
  if (!pose || !g_r || !g_x0 || !g_x1 || !g_pose) return set_error(DPFT_EINVAL, "pose, g_r, g_x0, g_x1 and g_pose are required");
  const long plane = (long)level->H * level->W;
  IcBwdParams p = icb_params(*level, B, C, (long)B * plane);
  p.pose = pose; p.g_out = g_r; p.g_x0 = g_x0; p.g_x1 = g_x1; p.g_pose = g_pose;
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B, 1);
  ic_bwd_warp_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  return launch_status("ic_residual_backward");
}

extern "C" int dpft_ic_normal_matrix_backward(const dpft_level_t* level, int B, int C, const float* gx, const float* gy,
                                              const float* weights, const float* g_A21, float* g_gx, float* g_gy,
                                              float* g_weights, void* stream) {
  if (int e = icb_check(level, B, C)) return e;
  if (!gx || !gy || !g_A21 || !g_gx || !g_gy) return set_error(DPFT_EINVAL, "gx, gy, g_A21, g_gx and g_gy are required");
  const long plane = (long)level->H * level->W;
  IcBwdParams p = icb_params(*level, B, C, (long)B * plane);
  p.gx = gx; p.gy = gy; p.w = weights; p.g_out = g_A21; p.g_gx = g_gx; p.g_gy = g_gy; p.g_w = g_weights;
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B, 1);
  ic_bwd_normal_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  return launch_status("ic_normal_matrix_backward");
}

extern "C" int dpft_ic_rhs_backward(const dpft_level_t* level, int B, int C, const float* gx, const float* gy,
                                    const float* weights, const float* poses, int S, const float* g_rhs, float* g_gx,
                                    float* g_gy, float* g_weights, float* g_x0, float* g_x1, float* g_poses, void* stream) {
  if (int e = icb_check(level, B, C)) return e;
  if (!gx || !gy || !poses || !g_rhs || !g_gx || !g_gy || !g_x0 || !g_x1 || !g_poses || S < 1 || S > 65535)
    return set_error(DPFT_EINVAL, "gx, gy, poses, g_rhs and the gradient outputs are required; 1 <= S");
  const long plane = (long)level->H * level->W;
  IcBwdParams p = icb_params(*level, B, C, (long)B * plane * S);
  p.gx = gx; p.gy = gy; p.w = weights; p.pose = poses; p.g_out = g_rhs;
  p.g_gx = g_gx; p.g_gy = g_gy; p.g_w = g_weights; p.g_x0 = g_x0; p.g_x1 = g_x1; p.g_pose = g_poses;
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B, S);
  ic_bwd_warp_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  return launch_status("ic_rhs_backward");
}

extern "C" int dpft_ic_update_backward(int B, int S, int mode, const float* A21, const float* rhs, const float* lambdas,
                                       const float* damp, const float* pose_in, const float* g_pose_out, float* g_A21,
                                       float* g_rhs, float* g_damp, float* g_pose_in, void* stream) {
  if (B < 1 || S < 1 || !A21 || !rhs || !pose_in || !g_pose_out || !g_A21 || !g_rhs || !g_pose_in)
    return set_error(DPFT_EINVAL, "A21, rhs, pose_in, g_pose_out, g_A21, g_rhs and g_pose_in are required");
  if (mode < 0 || mode > 2 || (mode == 1 && !lambdas) || (mode == 2 && !damp) || (mode != 1 && S != 1))
    return set_error(DPFT_EINVAL, "mode 0/2 need S == 1; mode 1 needs lambdas; mode 2 needs damp");
  ic_update_bwd_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(A21, rhs, lambdas, damp, pose_in, g_pose_out, g_A21,
                                                                       g_rhs, g_damp, g_pose_in, B, S, mode);
  return launch_status("ic_update_backward");
}
