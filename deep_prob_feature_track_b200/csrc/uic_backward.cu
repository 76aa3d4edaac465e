// U_IC backward: reverse-mode through the unrolled Gauss-Newton iterations, recompute-based.
//
// The forward saves only the pose before every iteration (pose_hist), the reduced normal equations
// (sys_hist) and the batch-global sigma extremes (aux_hist); everything per-pixel is recomputed here from
// the inputs.  For iteration k, walking k = last .. 0:
//
//   pose_bwd_kernel   (one thread per pair, fp64)  d/d(pose_{k+1}) -> d/d(xi) through compose, Rodrigues
//                     and the damped solve; emits lambda = H^-1 d(xi) and M = Abar + Abar^T with
//                     Abar = -lambda xi^T + 1e-6 tr(.) I  (+ an external gradient on J^T W J), and the
//                     part of d/d(pose_k) that flows through R <- R dR, t <- R dt + t.
//   uic_bwd_px_kernel (one thread per pixel)       with J_c = a_c ju + b_c jv the gradient of the loss
//                     w.r.t. (a_c, b_c, wres_c) needs only five per-pixel scalars of (M, lambda, ju, jv);
//                     from there back through the residual / sigma / bilinear lookup into x0, sigma0,
//                     their unit Sobel gradients (accumulated, turned into x0 / sigma0 gradients once per
//                     level by sobel_unit_bwd_kernel), scattered into x1 / sigma1 (red.global.add), and
//                     through (u,v) into the pose (12 sums per pair).
//
// Mirrors what torch.autograd derives from reference code/models/algorithms.py:611-723:
// masks and the 1e-6 fill are stop-gradients (masked pixels pass nothing through wres but still through J).
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

constexpr int kBT = 128;   // threads per CTA of the pixel kernel

struct BwdParams {
  const float *x0, *x1, *s0, *s1, *d0, *d1, *K;
  const uint8_t *m0, *m1;
  const float *gfx, *gfy, *gsx, *gsy;        // unit Sobel gradients of x0 / sigma0 (recomputed per level)
  float *g_x0, *g_x1, *g_s0, *g_s1;          // OUT (accumulated): gradients of the four maps
  float *g_gfx, *g_gfy, *g_gsx, *g_gsy;      // accumulated gradients w.r.t. the unit Sobel gradients
  int H, W, B, C, ppt;
  const float* pose;     // (B,12) pose the iteration linearised at
  const float* mlam;     // (B,27) 21 upper-triangular entries of M, then lambda
  const float* aux;      // [4] gmin, gmax of the warped sigma, min, max of sigma0 (this iteration)
  float* gpose;          // (B,12) accumulated d/d(pose_k)
};

#ifndef DPFT_BWD_CTAS
#define DPFT_BWD_CTAS 4
#endif
template <int CH, bool TRU>
__global__ void __launch_bounds__(kBT, DPFT_BWD_CTAS) uic_bwd_px_kernel(const BwdParams p) {
  __shared__ float s_red[kBT / 32][12];
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, C = p.C;
  const int iplane = H * W;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1);
  const float cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const size_t pair_off = (size_t)b * C * iplane;
  const float* d0p = p.d0 + (size_t)b * iplane;
  const float* d1p = p.d1 + (size_t)b * iplane;
  const uint8_t* m0p = p.m0 ? p.m0 + (size_t)b * iplane : nullptr;
  const uint8_t* m1p = p.m1 ? p.m1 + (size_t)b * iplane : nullptr;
  const Pose pose = load_pose(p.pose + (size_t)b * 12);
  float M[21], lam[6];
#pragma unroll
  for (int i = 0; i < 21; ++i) M[i] = __ldg(p.mlam + (size_t)b * 27 + i);
#pragma unroll
  for (int i = 0; i < 6; ++i) lam[i] = __ldg(p.mlam + (size_t)b * 27 + 21 + i);
  float gmin = 0.f, gmax = 0.f, s0lo = 0.f, s0hi = 0.f;
  if (TRU) {
    gmin = __ldg(p.aux);
    gmax = __ldg(p.aux + 1);
    s0lo = __ldg(p.aux + 2);
    s0hi = __ldg(p.aux + 3);
  }
  float gR[9], gt[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) gR[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) gt[i] = 0.f;

  for (int i = 0; i < p.ppt; ++i) {
    const int pix = (blockIdx.x * p.ppt + i) * kBT + threadIdx.x;
    if (pix >= iplane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float d0 = __ldg(d0p + pix);
    // forward geometry, same arithmetic as the forward kernels (the mask must come out identical)
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      w[k] = xadd(xadd(xadd(xmul(pose.r[3 * k], px), xmul(pose.r[3 * k + 1], py)), pose.r[3 * k + 2]), xmul(pose.t[k], d0));
    const float u = xadd(xmul(xdiv(w[0], w[2]), fx), cx);
    const float v = xadd(xmul(xdiv(w[1], w[2]), fy), cy);
    const float inv_z = xdiv(d0, w[2]);
    const Tap tap = make_tap(u, v, H, W);
    const float d1w = sample_exact(d1p, tap, W);
    bool occ = occluded(u, v, inv_z, d1w, H, W);
    if (m0p) occ = occ || (__ldg(m0p + pix) == 0);
    if (m1p) occ = occ || !(sample_mask(m1p, tap, W) > 0.f);
    if (TRU) {
      const float s0c0 = __ldg(p.s0 + pair_off + pix);
      occ = occ || (s0c0 == s0lo) || (s0c0 == s0hi);
      const float sr0 = sample_exact(p.s1 + pair_off, tap, W);
      occ = occ || (sr0 == gmin) || (sr0 == gmax);
    }
    // five scalars carry (M, lambda) to the per-channel level: with J = a ju + b jv,
    //   d/da = a ju'M ju + b ju'M jv + wm lambda.ju,  d/db = a ju'M jv + b jv'M jv + wm lambda.jv,
    //   d/dwm = a lambda.ju + b lambda.jv
    float ju[6], jv[6];
    warp_rows(px, py, d0, fx, fy, ju, jv);
    float Mu[6], Mv[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      float su = 0.f, sv = 0.f;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float m = M[r <= c ? tri(r, c) : tri(c, r)];
        su = fmaf(m, ju[c], su);
        sv = fmaf(m, jv[c], sv);
      }
      Mu[r] = su;
      Mv[r] = sv;
    }
    float muu = 0.f, muv = 0.f, mvv = 0.f, lu = 0.f, lv = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      muu = fmaf(ju[r], Mu[r], muu);
      muv = fmaf(ju[r], Mv[r], muv);
      mvv = fmaf(jv[r], Mv[r], mvv);
      lu = fmaf(lam[r], ju[r], lu);
      lv = fmaf(lam[r], jv[r], lv);
    }
    // is the sample position strictly inside the clip range?  (grid_sampler gives zero coordinate
    // gradient on and outside the border)
    const float ixu = xmul(xmul(xadd(xsub(xdiv(u, 0.5f * (float)(W - 1)), 1.f), 1.f), 0.5f), (float)(W - 1));
    const float iyu = xmul(xmul(xadd(xsub(xdiv(v, 0.5f * (float)(H - 1)), 1.f), 1.f), 0.5f), (float)(H - 1));
    const bool in_x = ixu > 0.f && ixu < (float)(W - 1), in_y = iyu > 0.f && iyu < (float)(H - 1);
    // fractional weights along each axis (wa = txl*tyn, wb = txr*tyn, wc = txl*tys, wd = txr*tys)
    const float tyn = tap.wa + tap.wb, tys = tap.wc + tap.wd, txl = tap.wa + tap.wc, txr = tap.wb + tap.wd;
    float g_ix = 0.f, g_iy = 0.f;

    // Channels go in groups of G: every load of the group (keyframe maps, unit gradients, the eight lookups AND the
    // running sums that are about to be updated) is issued before the first dependent instruction, so a thread has
    // 20 G independent loads in flight per round trip instead of one channel's worth -- the kernel is bound by
    // the latency of exactly these loads (ncu: 88 % of the stall samples are long-scoreboard waits).
#ifndef DPFT_BWD_GROUP
#define DPFT_BWD_GROUP 2
#endif
    constexpr int G = CH >= DPFT_BWD_GROUP ? DPFT_BWD_GROUP : (CH >= 2 ? 2 : 1);
    for (int c0 = 0; c0 < C; c0 += G) {
      float f0[G], s0v[G], gfx[G], gfy[G], gsx[G], gsy[G], xa[G], xb[G], xc[G], xd[G], za[G], zb[G], zc[G], zd[G];
      float o_gfx[G], o_gfy[G], o_gsx[G], o_gsy[G], o_x0[G], o_s0[G];
#pragma unroll
      for (int c = 0; c < G; ++c) {
        const size_t k0 = pair_off + (size_t)(c0 + c) * iplane + pix;
        const size_t k1 = pair_off + (size_t)(c0 + c) * iplane + tap.o;
        f0[c] = __ldg(p.x0 + k0); s0v[c] = __ldg(p.s0 + k0);
        gfx[c] = __ldg(p.gfx + k0); gfy[c] = __ldg(p.gfy + k0); gsx[c] = __ldg(p.gsx + k0); gsy[c] = __ldg(p.gsy + k0);
        xa[c] = __ldg(p.x1 + k1); xb[c] = __ldg(p.x1 + k1 + 1); xc[c] = __ldg(p.x1 + k1 + W); xd[c] = __ldg(p.x1 + k1 + W + 1);
        za[c] = __ldg(p.s1 + k1); zb[c] = __ldg(p.s1 + k1 + 1); zc[c] = __ldg(p.s1 + k1 + W); zd[c] = __ldg(p.s1 + k1 + W + 1);
        o_gfx[c] = p.g_gfx[k0]; o_gfy[c] = p.g_gfy[k0]; o_gsx[c] = p.g_gsx[k0]; o_gsy[c] = p.g_gsy[k0];
        o_x0[c] = p.g_x0[k0]; o_s0[c] = p.g_s0[k0];
      }
#pragma unroll
      for (int c = 0; c < G; ++c) {
        const size_t k0 = pair_off + (size_t)(c0 + c) * iplane + pix;
        const size_t k1 = pair_off + (size_t)(c0 + c) * iplane + tap.o;
        // forward quantities
        const float fr = blend_fast(xa[c], xb[c], xc[c], xd[c], tap);
        const float sr = TRU ? blend_exact(za[c], zb[c], zc[c], zd[c], tap) : blend_fast(za[c], zb[c], zc[c], zd[c], tap);
        const float res = fr - f0[c];
        const float rs = rsqrtf(fmaf(sr, sr, s0v[c] * s0v[c]));
        const float rs3 = rs * rs * rs;
        const float wres = res * rs;
        const float q = res * s0v[c] * rs3;
        const float a = fmaf(gfx[c], rs, q * gsx[c]);
        const float bq = fmaf(gfy[c], rs, q * gsy[c]);
        const float wm = occ ? 1e-6f : wres;
        // reverse
        const float ga = fmaf(a, muu, fmaf(bq, muv, wm * lu));
        const float gb = fmaf(a, muv, fmaf(bq, mvv, wm * lv));
        const float gwres = occ ? 0.f : fmaf(a, lu, bq * lv);
        const float gq = fmaf(ga, gsx[c], gb * gsy[c]);
        float grs = fmaf(ga, gfx[c], gb * gfy[c]);                    // via a, b
        float gres = fmaf(gq, s0v[c] * rs3, gwres * rs);              // via q, wres
        float gs0 = gq * res * rs3;                                   // via q
        grs = fmaf(3.f * gq, res * s0v[c] * rs * rs, grs);            // q ~ rs^3
        grs = fmaf(gwres, res, grs);                                  // wres = res rs
        const float gsr = -grs * rs3 * sr;                            // rs = (sr^2 + s0^2)^-1/2
        gs0 = fmaf(-grs * rs3, s0v[c], gs0);
        // own-pixel accumulations (this thread is the only writer of these elements in this launch)
        p.g_gfx[k0] = fmaf(ga, rs, o_gfx[c]);
        p.g_gfy[k0] = fmaf(gb, rs, o_gfy[c]);
        p.g_gsx[k0] = fmaf(ga, q, o_gsx[c]);
        p.g_gsy[k0] = fmaf(gb, q, o_gsy[c]);
        p.g_x0[k0] = o_x0[c] - gres;
        p.g_s0[k0] = o_s0[c] + gs0;
        // bilinear adjoint: scatter into the live frame's maps
        atomicAdd(p.g_x1 + k1, tap.wa * gres);
        atomicAdd(p.g_x1 + k1 + 1, tap.wb * gres);
        atomicAdd(p.g_x1 + k1 + W, tap.wc * gres);
        atomicAdd(p.g_x1 + k1 + W + 1, tap.wd * gres);
        atomicAdd(p.g_s1 + k1, tap.wa * gsr);
        atomicAdd(p.g_s1 + k1 + 1, tap.wb * gsr);
        atomicAdd(p.g_s1 + k1 + W, tap.wc * gsr);
        atomicAdd(p.g_s1 + k1 + W + 1, tap.wd * gsr);
        // and into the sample position
        g_ix = fmaf(gres, fmaf(xb[c] - xa[c], tyn, (xd[c] - xc[c]) * tys), g_ix);
        g_iy = fmaf(gres, fmaf(xc[c] - xa[c], txl, (xd[c] - xb[c]) * txr), g_iy);
        g_ix = fmaf(gsr, fmaf(zb[c] - za[c], tyn, (zd[c] - zc[c]) * tys), g_ix);
        g_iy = fmaf(gsr, fmaf(zc[c] - za[c], txl, (zd[c] - zb[c]) * txr), g_iy);
      }
    }
    // (u,v) -> w = R ray + t d0 -> pose
    const float gu = in_x ? g_ix : 0.f, gv = in_y ? g_iy : 0.f;
    const float iz = 1.f / w[2];
    const float gw0 = gu * fx * iz, gw1 = gv * fy * iz;
    const float gw2 = -(gw0 * w[0] + gw1 * w[1]) * iz;
    const float gw[3] = {gw0, gw1, gw2};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      gR[3 * r] = fmaf(gw[r], px, gR[3 * r]);
      gR[3 * r + 1] = fmaf(gw[r], py, gR[3 * r + 1]);
      gR[3 * r + 2] += gw[r];
      gt[r] = fmaf(gw[r], d0, gt[r]);
    }
  }
  // 12 pose-gradient sums: warp shuffle, then one atomic per CTA and entry
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float s = i < 9 ? gR[i] : gt[i - 9];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < kBT / 32; ++wq) s += s_red[wq][threadIdx.x];
    atomicAdd(p.gpose + (size_t)b * 12 + threadIdx.x, s);
  }
}

// d/d(pose_{k+1}) -> (M, lambda) of iteration k and the compose part of d/d(pose_k).  One thread per pair.
__global__ void pose_bwd_kernel(const float* __restrict__ sys, const float* __restrict__ pose_k,
                                const float* __restrict__ gpose_next, const float* __restrict__ gA_ext,
                                float* __restrict__ gpose_k, float* __restrict__ mlam, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double A[21], rhs[6];
  for (int i = 0; i < 21; ++i) A[i] = (double)sys[(size_t)b * 27 + i];
  for (int i = 0; i < 6; ++i) rhs[i] = (double)sys[(size_t)b * 27 + 21 + i];
  double tr = 0.0;
  for (int i = 0; i < 6; ++i) tr += A[tri(i, i)];
  for (int i = 0; i < 6; ++i) A[tri(i, i)] += tr * 1e-6;
  double xi[6], lam[6], gk[12];
  for (int i = 0; i < 12; ++i) gk[i] = 0.0;
  solve_update_adjoint(A, rhs, pose_k + (size_t)b * 12, gpose_next + (size_t)b * 12, xi, lam, gk);
  // xi = H^-1 rhs:  Hbar = -lambda xi^T;  H = A + 1e-6 tr(A) I
  double trH = 0;
  for (int i = 0; i < 6; ++i) trH += -lam[i] * xi[i];
  float* out = mlam + (size_t)b * 27;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      double m = -(lam[i] * xi[j] + lam[j] * xi[i]);
      if (i == j) m += 2.0 * 1e-6 * trH;
      if (gA_ext) m += (double)gA_ext[(size_t)b * 36 + 6 * i + j] + (double)gA_ext[(size_t)b * 36 + 6 * j + i];
      out[tri(i, j)] = (float)m;
    }
  for (int i = 0; i < 6; ++i) out[21 + i] = (float)lam[i];
  for (int i = 0; i < 12; ++i) gpose_k[(size_t)b * 12 + i] += (float)gk[i];
}

// Adjoint of g = S / sqrt(|S|^2 + 1e-8), S = replicate-padded Sobel of img, ADDED to g_img.  Gather form: a CTA owns
// 32 x 8 pixels of one plane, evaluates bs = n gg - (S . gg) n^3 S on the tile plus a one-pixel halo in shared memory
// and every thread sums the transposed stencil for its own pixel -- one plain read-modify-write per element instead
// of eight scattered atomicAdd per pixel.  With D_r[x] = bsx_r[x-1] - bsx_r[x+1], T_r[x] = bsy_r[x-1] + 2 bsy_r[x] +
// bsy_r[x+1]:  g[y][x] += (D_{y-1} + T_{y-1}) + 2 D_y + (D_{y+1} - T_{y+1}).  The replicate padding folds what would
// leave the image back onto the border, which is the same as extending bs by reflection: bsx odd / bsy even across a
// vertical border, bsx even / bsy odd across a horizontal one.
__global__ void __launch_bounds__(256) sobel_unit_bwd_kernel(const float* __restrict__ img, const float* __restrict__ ggx,
                                                             const float* __restrict__ ggy, float* __restrict__ g_img,
                                                             int planes, int H, int W) {
  __shared__ float sbx[10][35], sby[10][35];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const int x = x0 + tx, y = y0 + ty;
  for (int pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const size_t base = (size_t)pl * H * W;
    const float* q = img + base;
    for (int i = threadIdx.x; i < 10 * 34; i += 256) {
      const int hy = i / 34, hx = i - hy * 34;
      const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
      const int cy = min(max(gy, 0), H - 1), cx = min(max(gx, 0), W - 1);
      float bx = 0.f, by = 0.f;
      if (gy <= H && gx <= W) {      // (positions further out than the one-pixel fold feed no pixel of the image)
        const int xl = max(cx - 1, 0), xr = min(cx + 1, W - 1), yt = max(cy - 1, 0) * W, ym = cy * W, yb = min(cy + 1, H - 1) * W;
        const float a = __ldg(q + yt + xl), b = __ldg(q + yt + cx), c = __ldg(q + yt + xr);
        const float d = __ldg(q + ym + xl), f = __ldg(q + ym + xr);
        const float g = __ldg(q + yb + xl), h = __ldg(q + yb + cx), k = __ldg(q + yb + xr);
        const float sx = (c - a) + 2.f * (f - d) + (k - g);
        const float sy = (g - a) + 2.f * (h - b) + (k - c);
        const float n = rsqrtf(fmaf(sx, sx, fmaf(sy, sy, 1e-8f)));
        const float gx_ = __ldg(ggx + base + ym + cx), gy_ = __ldg(ggy + base + ym + cx);
        const float dot = (sx * gx_ + sy * gy_) * n * n * n;
        bx = fmaf(-dot, sx, n * gx_);
        by = fmaf(-dot, sy, n * gy_);
        if (gx != cx) bx = -bx;      // reflection across a vertical border
        if (gy != cy) by = -by;      // ... across a horizontal border
      }
      sbx[hy][hx] = bx;
      sby[hy][hx] = by;
    }
    __syncthreads();
    if (x < W && y < H) {
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float D = sbx[ty + r][tx] - sbx[ty + r][tx + 2];
        const float T = sby[ty + r][tx] + 2.f * sby[ty + r][tx + 1] + sby[ty + r][tx + 2];
        acc += (r == 0) ? (D + T) : (r == 1) ? 2.f * D : (D - T);
      }
      g_img[base + (size_t)y * W + x] += acc;
    }
    __syncthreads();
  }
}

void launch_sobel_unit_bwd(const float* img, const float* ggx, const float* ggy, float* g_img, int planes, int H, int W,
                           cudaStream_t stream) {
  const dim3 sg((W + 31) / 32, (H + 7) / 8, std::min(planes, 4096));
  sobel_unit_bwd_kernel<<<sg, 256, 0, stream>>>(img, ggx, ggy, g_img, planes, H, W);
}

__global__ void bwd_mm_init_kernel(uint32_t* mm) {
  mm[0] = 0xffffffffu;
  mm[1] = 0u;
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[i] : 0.f;
}

struct BwdPlan {
  size_t max_plane, grad_elems;
  size_t off_unit, off_gunit, off_gpose, off_mlam, off_vn, off_dmm, total;
};

static BwdPlan make_bwd_plan(const dpft_level_t* lv, int n_levels, int B, int C, int iters, bool icp = false) {
  BwdPlan pl{};
  for (int l = 0; l < n_levels; ++l) pl.max_plane = std::max(pl.max_plane, (size_t)lv[l].H * lv[l].W);
  pl.grad_elems = (size_t)B * C * pl.max_plane;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  pl.off_unit = take(4 * pl.grad_elems * sizeof(float));
  pl.off_gunit = take(4 * pl.grad_elems * sizeof(float));
  pl.off_gpose = take((size_t)(n_levels * iters + 1) * B * 12 * sizeof(float));
  pl.off_mlam = take((size_t)B * 27 * sizeof(float));
  pl.off_vn = take(icp ? 6 * (size_t)B * pl.max_plane * sizeof(float) : 0);
  pl.off_dmm = take(2 * sizeof(uint32_t));
  pl.total = off;
  return pl;
}

template <int CH>
static void launch_bwd(const BwdParams& prm, dim3 grid, bool tru, cudaStream_t stream) {
  if (tru) uic_bwd_px_kernel<CH, true><<<grid, kBT, 0, stream>>>(prm);
  else uic_bwd_px_kernel<CH, false><<<grid, kBT, 0, stream>>>(prm);
}

}  // namespace dpft

using namespace dpft;

extern "C" size_t dpft_uic_backward_workspace_bytes(const dpft_level_t* levels, int n_levels, int B, int C,
                                                    int iters, uint32_t flags) {
  if (!levels || n_levels < 1 || n_levels > DPFT_MAX_LEVELS || B < 1 || C < 1 || iters < 1) {
    set_error(DPFT_EINVAL, "bad problem size");
    return 0;
  }
  return make_bwd_plan(levels, n_levels, B, C, iters, flags & DPFT_COMBINE_ICP).total;
}

extern "C" int dpft_uic_backward(const dpft_level_t* levels, const dpft_level_grad_t* grads, int n_levels, int B,
                                 int C, int iters, uint32_t flags, float w_icp, const float* pose_hist, const float* sys_hist,
                                 const float* aux_hist, const float* grad_pose_hist, const float* grad_A,
                                 float* grad_pose_in, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!levels || !grads || n_levels < 1 || n_levels > DPFT_MAX_LEVELS || B < 1 || B > 65535 || C < 1 || iters < 1)
    return set_error(DPFT_EINVAL, "bad problem size");
  if (!pose_hist || !sys_hist || !aux_hist || !grad_pose_hist || !grad_pose_in || !workspace)
    return set_error(DPFT_EINVAL, "pose_hist, sys_hist, aux_hist, grad_pose_hist, grad_pose_in and workspace are required");
  const bool icp = flags & DPFT_COMBINE_ICP;
  if (icp)
    for (int l = 0; l < n_levels; ++l)
      if (!levels[l].depth0 || !levels[l].depth1) return set_error(DPFT_EINVAL, "level %d: DPFT_COMBINE_ICP needs depth0 and depth1", l);
  for (int l = 0; l < n_levels; ++l)
    if (!grads[l].g_x0 || !grads[l].g_x1 || !grads[l].g_sigma0 || !grads[l].g_sigma1)
      return set_error(DPFT_EINVAL, "level %d: all four gradient maps are required", l);
  const BwdPlan pl = make_bwd_plan(levels, n_levels, B, C, iters, icp);
  if (workspace_bytes < pl.total) return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, pl.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  float* unit = (float*)(ws + pl.off_unit);
  float* gunit = (float*)(ws + pl.off_gunit);
  float* gpose = (float*)(ws + pl.off_gpose);
  float* mlam = (float*)(ws + pl.off_mlam);
  float* vn = (float*)(ws + pl.off_vn);
  uint32_t* dmm = (uint32_t*)(ws + pl.off_dmm);
  const bool tru = flags & DPFT_REMOVE_TRU_SIGMA;
  const int n_it = n_levels * iters;
  const int CH = (C % 8 == 0) ? 8 : (C % 4 == 0) ? 4 : (C % 2 == 0) ? 2 : 1;
  {
    const size_t n = (size_t)(n_it + 1) * B * 12;
    copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(grad_pose_hist, gpose, n);
  }
  for (int l = n_levels - 1; l >= 0; --l) {
    const dpft_level_t& L = levels[l];
    const size_t plane = (size_t)L.H * L.W;
    const size_t ne = (size_t)B * C * plane;
    float* u0 = unit; float* u1 = unit + pl.grad_elems; float* u2 = unit + 2 * pl.grad_elems; float* u3 = unit + 3 * pl.grad_elems;
    float* g0 = gunit; float* g1 = gunit + pl.grad_elems; float* g2 = gunit + 2 * pl.grad_elems; float* g3 = gunit + 3 * pl.grad_elems;
    launch_sobel_unit(L.x0, u0, u1, B * C, L.H, L.W, stream);
    launch_sobel_unit(L.sigma0, u2, u3, B * C, L.H, L.W, stream);
    cudaMemsetAsync(gunit, 0, 4 * pl.grad_elems * sizeof(float), stream);
    if (icp) {
      bwd_mm_init_kernel<<<1, 1, 0, stream>>>(dmm);
      launch_minmax(L.depth1, (size_t)B * plane, dmm, stream);
      launch_vertex_normal(L.depth1, L.K, dmm, vn, vn + 3 * (size_t)B * plane, B, L.H, L.W, stream);
    }
    const long want_threads = 148L * 2048 * 2;
    long ppt = ((long)B * (long)plane + want_threads - 1) / want_threads;
    ppt = std::max(1L, std::min(ppt, 8L));
    const dim3 grid((unsigned)((plane + (size_t)kBT * ppt - 1) / ((size_t)kBT * ppt)), B);
    for (int it = iters - 1; it >= 0; --it) {
      const int k = l * iters + it;
      const float* gA = (grad_A && it == iters - 1) ? grad_A + (size_t)l * B * 36 : nullptr;
      pose_bwd_kernel<<<(B + 63) / 64, 64, 0, stream>>>(sys_hist + (size_t)k * B * 27, pose_hist + (size_t)k * B * 12,
                                                         gpose + (size_t)(k + 1) * B * 12, gA,
                                                         gpose + (size_t)k * B * 12, mlam, B);
      BwdParams prm{};
      prm.x0 = L.x0; prm.x1 = L.x1; prm.s0 = L.sigma0; prm.s1 = L.sigma1; prm.d0 = L.invd0; prm.d1 = L.invd1; prm.K = L.K;
      prm.m0 = L.obj_mask0; prm.m1 = L.obj_mask1;
      prm.gfx = u0; prm.gfy = u1; prm.gsx = u2; prm.gsy = u3;
      prm.g_x0 = grads[l].g_x0; prm.g_x1 = grads[l].g_x1; prm.g_s0 = grads[l].g_sigma0; prm.g_s1 = grads[l].g_sigma1;
      prm.g_gfx = g0; prm.g_gfy = g1; prm.g_gsx = g2; prm.g_gsy = g3;
      prm.H = L.H; prm.W = L.W; prm.B = B; prm.C = C; prm.ppt = (int)ppt;
      prm.pose = pose_hist + (size_t)k * B * 12;
      prm.mlam = mlam;
      prm.aux = aux_hist + 4 * k;
      prm.gpose = gpose + (size_t)k * B * 12;
      if (icp)
        launch_icp_bwd(L.depth0, L.K, vn, vn + 3 * (size_t)B * plane, prm.pose, mlam, L.obj_mask0, L.obj_mask1,
                       prm.gpose, w_icp * w_icp, B, L.H, L.W, stream);
      switch (CH) {
        case 8: launch_bwd<8>(prm, grid, tru, stream); break;
        case 4: launch_bwd<4>(prm, grid, tru, stream); break;
        case 2: launch_bwd<2>(prm, grid, tru, stream); break;
        default: launch_bwd<1>(prm, grid, tru, stream); break;
      }
    }
    // unit-gradient adjoints of the level -> x0 / sigma0
    launch_sobel_unit_bwd(L.x0, g0, g1, grads[l].g_x0, B * C, L.H, L.W, stream);
    launch_sobel_unit_bwd(L.sigma0, g2, g3, grads[l].g_sigma0, B * C, L.H, L.W, stream);
    (void)ne;
  }
  copy_kernel<<<(unsigned)(((size_t)B * 12 + 255) / 256), 256, 0, stream>>>(gpose, grad_pose_in, (size_t)B * 12);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "backward launch: %s", cudaGetErrorString(err));
  return 0;
}
