// Device-side building blocks shared by the solver kernels (sm_100a).
//
// Arithmetic that feeds a validity mask is written with the *_rn intrinsics so that nvcc can not
// contract it into FMAs: every product and sum is rounded on its own, in the order the oracle
// (oracle/ic_oracle.py: project, _unnormalise, sample_border) spells out.  Everything else is
// ordinary fp32 and free to fuse.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpft {

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
#ifdef DPFT_EXPERIMENT_FAST_DIV   // timing experiment only: NOT mask-exact
__device__ __forceinline__ float xdiv(float a, float b) { return __fdividef(a, b); }
#else
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
#endif

// a / b given r = RN(1/b) (__frcp_rn): q = RN(a r) followed by two exact-remainder corrections gives the IEEE
// quotient (Markstein); checked against a / b on 2.4e8 random operand pairs of the ranges used here
// (profiles/micro/check_division.py) and by the bit-exact mask tests.  Three quotients by the same divisor
// cost one reciprocal, and a divisor that is constant over a tile costs it once.
__device__ __forceinline__ float div_by(float a, float b, float r) {
  float q = __fmul_rn(a, r);
  q = fmaf(fmaf(-q, b, a), r, q);
  q = fmaf(fmaf(-q, b, a), r, q);
  return q;
}

// 1/sqrt(x) as ONE MUFU.RSQ.  rsqrtf() wraps the same instruction in a scale-up / scale-down for denormal
// inputs (three more instructions, 24 times per pixel row here); the arguments on this path are sums of squares
// with a 1e-8 floor or squared uncertainties, never denormal.  Relative error 2^-22.9, as rsqrtf.
__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Order-preserving map float -> uint32 so atomicMin/atomicMax work on floats of either sign.
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Pose of one pair, held in registers.
struct Pose {
  float r[9];
  float t[3];
};

__device__ __forceinline__ Pose load_pose(const float* __restrict__ row) {
  Pose p;
#pragma unroll
  for (int i = 0; i < 9; ++i) p.r[i] = row[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) p.t[i] = row[9 + i];
  return p;
}

// SE(3) warp of the ray (x, y, 1) at inverse depth d (reference geometry.py:291-323).
// w = ((r0*x + r1*y) + r2) + t*d, each step rounded; u = (w.x/w.z)*fx + cx; inv_z = d / w.z
__device__ __forceinline__ void warp_pixel(const Pose& p, float x, float y, float d, float fx, float fy,
                                           float cx, float cy, float& u, float& v, float& inv_z) {
  float w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    w[i] = xadd(xadd(xadd(xmul(p.r[3 * i], x), xmul(p.r[3 * i + 1], y)), p.r[3 * i + 2]), xmul(p.t[i], d));
  u = xadd(xmul(xdiv(w[0], w[2]), fx), cx);
  v = xadd(xmul(xdiv(w[1], w[2]), fy), cy);
  inv_z = xdiv(d, w[2]);
}

// Bilinear footprint with border padding: warp_features (geometry.py:353-365) followed by torch's
// grid_sampler_2d(align_corners=True, padding_mode='border') un-normalisation, clip and weights.
//
// The four texels are addressed as o, o+1, o+W, o+W+1 with o = yi*W + xi, xi <= W-2, yi <= H-2.  torch
// clamps the east/south index at the border instead (where its weight is exactly 0); here the footprint
// is shifted one texel inwards and the two weights swap places, which gives the same sum bit for bit
// (x*w + y*0 == y*0 + x*w up to the sign of a zero) and keeps every load at a fixed offset from one base.
struct Tap {
  int o;
  float wa, wb, wc, wd;   // weights of texels o, o+1, o+W, o+W+1
};

__device__ __forceinline__ float unnormalise(float coord, float half_span, float span) {
  float g = xsub(xdiv(coord, half_span), 1.f);          // u / ((W-1)/2) - 1
  float pix = xmul(xmul(xadd(g, 1.f), 0.5f), span);     // ((g+1)/2) * (W-1)
  return fminf(fmaxf(pix, 0.f), span);                  // clip to [0, W-1]
}

__device__ __forceinline__ Tap make_tap_at(float ix, float iy, int H, int W);

__device__ __forceinline__ Tap make_tap(float u, float v, int H, int W) {
  const float ix = unnormalise(u, 0.5f * (float)(W - 1), (float)(W - 1));
  const float iy = unnormalise(v, 0.5f * (float)(H - 1), (float)(H - 1));
  return make_tap_at(ix, iy, H, W);
}

// same, with the reciprocals of the half spans supplied (hoisted out of the row loop)
__device__ __forceinline__ Tap make_tap_r(float u, float v, int H, int W, float rcp_half_w, float rcp_half_h) {
  const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
  const float gx = xsub(div_by(u, hw, rcp_half_w), 1.f), gy = xsub(div_by(v, hh, rcp_half_h), 1.f);
  const float ix = fminf(fmaxf(xmul(xmul(xadd(gx, 1.f), 0.5f), (float)(W - 1)), 0.f), (float)(W - 1));
  const float iy = fminf(fmaxf(xmul(xmul(xadd(gy, 1.f), 0.5f), (float)(H - 1)), 0.f), (float)(H - 1));
  return make_tap_at(ix, iy, H, W);
}

__device__ __forceinline__ Tap make_tap_at(float ix, float iy, int H, int W) {
  const float xw = floorf(ix), yn = floorf(iy);
  float txr = xsub(ix, xw), tys = xsub(iy, yn);                       // east / south weights
  float txl = xsub(xadd(xw, 1.f), ix), tyn = xsub(xadd(yn, 1.f), iy); // west / north weights
  int xi = (int)xw, yi = (int)yn;
  if (xi > W - 2) { xi = W - 2; const float t = txl; txl = txr; txr = t; }
  if (yi > H - 2) { yi = H - 2; const float t = tyn; tyn = tys; tys = t; }
  xi = max(xi, 0);
  yi = max(yi, 0);
  Tap t;
  t.wa = xmul(txl, tyn);
  t.wb = xmul(txr, tyn);
  t.wc = xmul(txl, tys);
  t.wd = xmul(txr, tys);
  t.o = yi * W + xi;
  return t;
}

// ((a*wa + b*wb) + c*wc) + d*wd with every step rounded (mask-grade).
__device__ __forceinline__ float blend_exact(float a, float b, float c, float d, const Tap& t) {
  return xadd(xadd(xadd(xmul(a, t.wa), xmul(b, t.wb)), xmul(c, t.wc)), xmul(d, t.wd));
}
__device__ __forceinline__ float blend_fast(float a, float b, float c, float d, const Tap& t) {
  return fmaf(d, t.wd, fmaf(c, t.wc, fmaf(b, t.wb, a * t.wa)));
}
__device__ __forceinline__ float sample_exact(const float* __restrict__ plane, const Tap& t, int W) {
  const float* q = plane + t.o;
  const float* r = q + W;
  return blend_exact(__ldg(q), __ldg(q + 1), __ldg(r), __ldg(r + 1), t);
}
__device__ __forceinline__ float sample_mask(const uint8_t* __restrict__ plane, const Tap& t, int W) {
  const uint8_t* q = plane + t.o;
  return blend_exact((float)__ldg(q), (float)__ldg(q + 1), (float)__ldg(q + W), (float)__ldg(q + W + 1), t);
}

// z-buffer + in-view test (geometry.py:334-350). true = excluded.
__device__ __forceinline__ bool occluded(float u, float v, float inv_z, float d1w, int H, int W) {
  const bool ok = (inv_z > xsub(d1w, 0.1f)) && (u > 0.f) && (u < (float)W) && (v > 0.f) && (v < (float)H);
  return !ok;
}

// Rows of d(u,v)/d(xi) at the identity (algorithms.py:1884-1917), twist = [rot, trs].
// ju[4] and jv[3] are structurally zero and never touched by the callers.
__device__ __forceinline__ void warp_rows(float x, float y, float d, float fx, float fy, float ju[6], float jv[6]) {
  const float xy = x * y;
  ju[0] = fx * (-xy);
  ju[1] = fx * (1.f + x * x);
  ju[2] = fx * (-y);
  ju[3] = fx * d;
  ju[4] = 0.f;
  ju[5] = fx * (-d * x);
  jv[0] = fy * (-1.f - y * y);
  jv[1] = fy * xy;
  jv[2] = fy * x;
  jv[3] = 0.f;
  jv[4] = fy * d;
  jv[5] = fy * (-d * y);
}

// index of (i,j), i<=j, in the 21-entry row-major upper triangle of a symmetric 6x6
__host__ __device__ constexpr int tri(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

// A += sum over the pixel's channels of J_c J_c^T and b += sum J_c r_c, where J_c = a_c*ju + b_c*jv:
// with saa = sum a^2, sab = sum a b, sbb = sum b^2, sar = sum a r, sbr = sum b r this is
//   A += ju (saa ju + sab jv)^T + jv (sab ju + sbb jv)^T ,  b += sar ju + sbr jv
// (the C x 6 Jacobian of algorithms.py:867-887 is never formed).
__device__ __forceinline__ void accumulate_system(float acc[27], const float ju[6], const float jv[6], float saa,
                                                  float sab, float sbb, float sar, float sbr) {
  float P[6], Q[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (j == 4) { P[j] = sab * jv[j]; Q[j] = sbb * jv[j]; }
    else if (j == 3) { P[j] = saa * ju[j]; Q[j] = sab * ju[j]; }
    else { P[j] = fmaf(saa, ju[j], sab * jv[j]); Q[j] = fmaf(sab, ju[j], sbb * jv[j]); }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) {
      float a = acc[tri(i, j)];
      if (i != 4) a = fmaf(ju[i], P[j], a);
      if (i != 3) a = fmaf(jv[i], Q[j], a);
      acc[tri(i, j)] = a;
    }
    float r = acc[21 + i];
    if (i != 4) r = fmaf(sar, ju[i], r);
    if (i != 3) r = fmaf(sbr, jv[i], r);
    acc[21 + i] = r;
  }
}

// ------------------------------------------------------------------ 6x6 solve and pose update (fp64)
// H = A + 1e-6 trace(A) I (algorithms.py:2094-2103), xi = H^-1 b by Cholesky (H is SPD), then the
// inverse-compositional update of algorithms.py:2035-2054 / geometry.py:105-123,163-185:
//   dR = exp(-xi_w) (Rodrigues, no small-angle guard), dt = -dR xi_v, R <- R dR, t <- R dt + t.
// A: 21 upper-triangular entries. Returns false if a pivot is not positive.
__device__ inline bool solve_and_update(const double A[21], const double rhs[6], bool damp_trace,
                                        const float* __restrict__ pose_in, float* __restrict__ pose_out,
                                        double xi_out[6]) {
  // Cholesky with the reciprocal pivots kept (one rsqrt per column, no division, no sqrt: the solve sits on the
  // serial tail of every launch and fp64 sqrt / div are long dependent sequences)
  double L[6][6], invd[6];
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) tr += A[tri(i, i)];
  const double eps = damp_trace ? tr * 1e-6 : 0.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[tri(j, j)] + eps;
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    ok = ok && (s > 0.0);
    const double inv = rsqrt(s);
    invd[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = A[tri(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v * inv;
    }
  }
  double z[6], xi[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = rhs[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= L[i][k] * z[k];
    z[i] = v * invd[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = z[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= L[k][i] * xi[k];
    xi[i] = v * invd[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) xi_out[i] = xi[i];

  const double wx = -xi[0], wy = -xi[1], wz = -xi[2];
  const double th = sqrt(wx * wx + wy * wy + wz * wz);
  const double ith = 1.0 / th;
  const double kx = wx * ith, ky = wy * ith, kz = wz * ith;   // NaN at th == 0, as in the reference
  double s, c;
  sincos(th, &s, &c);
  const double c1 = 1.0 - c;
  const double Kx[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
  double dR[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double kk = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) kk += Kx[3 * i + k] * Kx[3 * k + j];
      dR[3 * i + j] = (i == j ? 1.0 : 0.0) + Kx[3 * i + j] * s + kk * c1;
    }
  double dt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) dt[i] = -(dR[3 * i] * xi[3] + dR[3 * i + 1] * xi[4] + dR[3 * i + 2] * xi[5]);
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = (double)__ldcg(pose_in + i);   // L2: in the persistent kernels another SM wrote it
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = (double)__ldcg(pose_in + 9 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      pose_out[3 * i + j] = (float)(R[3 * i] * dR[j] + R[3 * i + 1] * dR[3 + j] + R[3 * i + 2] * dR[6 + j]);
    pose_out[9 + i] = (float)(R[3 * i] * dt[0] + R[3 * i + 1] * dt[1] + R[3 * i + 2] * dt[2] + t[i]);
  }
  return ok;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace dpft
