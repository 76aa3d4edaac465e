// One warp tile of the fused U_IC iteration: 30 output columns x `rows` rows of one frame pair.
//
// Lanes 0 and 31 are halo columns (clamped at the image border = replicate padding), so the horizontal Sobel
// taps come from __shfl of the neighbouring lanes and the vertical taps from a 3-row register window that
// slides down the tile: every x0 / sigma0 element is loaded once per tile (coalesced rows), the unit Sobel
// gradients are recomputed instead of stored, and nothing per-pixel is written back.
// Restates reference code/models/algorithms.py:611-723 with the arithmetic of oracle/ic_oracle.py.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "dpft_device.cuh"

namespace dpft {

constexpr int kTileCols = 30;   // output columns per warp tile

struct PairView {               // one frame pair of one level, pointers already offset to the pair
  const float *x0, *x1, *s0, *s1, *d0, *d1;
  const uint8_t *m0, *m1;       // optional object masks
  uint8_t* occ_out;             // optional debug outputs (per pair)
  float* sr0_dbg;
  int H, W, C;
  unsigned splane;              // channel stride of sigma0 / sigma1 in elements: H*W, or 0 when the uncertainty is ONE
                                // map per frame that the reference would have repeated to C channels (alg:1425-1427)
  int scm;                      // channels of sigma0 / sigma1 IN MEMORY (C, or 1): with splane == 0 and scm == C the
                                // tensors are full but their channels were found to be copies of channel 0
  float fx, fy, cx, cy;
  float s0lo, s0hi;             // extremes of sigma0 over the whole level tensor (remove_tru_sigma)
  const void *tm_x1, *tm_s1, *tm_d1;   // tensor maps of the live frame's level tensors (TMA ring staging only)
  int b;                        // pair index (the maps cover the whole batch)
};

struct TileSums {
  float acc[27];                // 21 upper-triangular J^T J entries, 6 J^T r entries
  float vmin, vmax;             // running extremes of the warped sigma (their corrections live in shared memory)
#ifdef DPFT_DEBUG_STAMPS
  int ndirect = 0;              // rows of the staged routine that took the direct-load body
  int nrestart = 0, nlanes = 0; // ring restarts; lanes (summed over rows) whose footprint was not resident
  int nstaged = 0, ntru = 0;    // source rows staged; rows that took the sigma-extreme bookkeeping branch
  long long c_front = 0, c_wait = 0, c_body = 0;   // clock cycles: geometry + ring logic | waiting for the ring | the rest of the row
#endif
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.f;
    vmin = CUDART_INF_F;
    vmax = -CUDART_INF_F;
  }
};

// base + idx (elements) as ONE IMAD.WIDE.U32: the base is made opaque (a per-thread 64-bit register pair) so
// the compiler cannot fall back to uniform-register bases with 64-bit byte offsets per plane, which costs four
// integer instructions per load
__device__ __forceinline__ const float* opaque(const float* p) {
  asm volatile("" : "+l"(p));
  return p;
}
// keyframe-side window loads: every element is read once per tile, so there is nothing for L1 to keep
// (-DDPFT_WINDOW_LDCG: ld.global.cg, no L1 line allocated; a tuning hook, see profiles/r2/)
#ifdef DPFT_WINDOW_LDCG
#define DPFT_LDW(p) __ldcg(p)
#else
#define DPFT_LDW(p) __ldg(p)
#endif
__device__ __forceinline__ float ldf(const float* __restrict__ base, unsigned idx) {
  const float* q;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(q) : "r"(idx), "l"(base));
  return DPFT_LDW(q);
}
__device__ __forceinline__ void ldf2(const float* __restrict__ base, unsigned idx, float& a, float& b) {
  const float* q;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(q) : "r"(idx), "l"(base));
  a = __ldg(q);
  b = __ldg(q + 1);
}

// 16-byte shared-memory load that the compiler may not hoist out of the row loop: the pose is read where it
// is used instead of pinning 12 registers across the whole tile
__device__ __forceinline__ float4 lds_v4(const float* p) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "r"((unsigned)__cvta_generic_to_shared(p)));
  return r;
}

// `spose`: this warp's pose in SHARED memory, 12 floats R|t, 16-byte aligned.
// `scorr`: this warp's [12][33] shared scratch for the sigma-extreme corrections (remove_tru_sigma), zeroed by
//          the caller; row i < 6 belongs to the minimum, row 6 + i to the maximum, one column per lane.
//
// GW, GH > 0: the level's width and height are compile-time constants (the reference's pyramid sizes get their
// own instantiations).  Every channel plane and footprint tap is then an immediate offset from ONE address
// register per map and row, instead of a 64-bit multiply-add per load; GW = GH = 0 is the generic routine.
//
// RES ("resident"): `live` is a SHARED-memory copy of the live frame of this pair -- x1 (C planes), sigma1 (C planes, or
// one with a single uncertainty map), invd1 -- which the CTA staged before the walk; every footprint tap is then an
// LDS, whatever the warp field does.  What the small pyramid levels run (a 30x40 level is 82 KB), see
// uic_iter_kernel.
template <int CH, bool TRU, int GW = 0, int GH = 0, bool RES = false>
__device__ __forceinline__ void process_tile(const PairView& g, const float* spose, float (*scorr)[33],
                                             const int seg, const int y0, const int y1, const int lane,
                                             TileSums& S, const float* live = nullptr) {
  static_assert(!RES || (GW == 0 && GH == 0), "the resident routine is generic in the level size");
  constexpr bool FIXED = GW > 0 && GH > 0;
  constexpr int PLANE = GW * GH;
  const int H = FIXED ? GH : g.H, W = FIXED ? GW : g.W, C = g.C;
  const unsigned iplane = (unsigned)(H * W), Wu = (unsigned)W;
  const int x = seg * kTileCols - 1 + lane;
  const int xc = min(max(x, 0), W - 1);
  const bool col_out = lane >= 1 && lane <= kTileCols && x < W;
  const float fx = g.fx, fy = g.fy, cx = g.cx, cy = g.cy;
  const float px = xdiv(xsub((float)xc, cx), fx);
  // reciprocals of the divisors that do not change over the tile (see div_by)
  const float rcp_fy = __frcp_rn(fy);
  const float rcp_hw = __frcp_rn(0.5f * (float)(W - 1)), rcp_hh = __frcp_rn(0.5f * (float)(H - 1));

  for (int c0 = 0; c0 < C; c0 += CH) {
    const float* X0 = g.x0 + (size_t)c0 * iplane;
    const unsigned splane = g.splane;
    const float* S0 = g.s0 + (size_t)c0 * splane;
    const float* X1 = RES ? live + (unsigned)c0 * iplane : g.x1 + (size_t)c0 * iplane;
    const float* S1 = RES ? live + (unsigned)C * iplane + (unsigned)c0 * splane : g.s1 + (size_t)c0 * splane;
    const float* D1 = RES ? live + (unsigned)(C + g.scm) * iplane : g.d1;
    if (!FIXED) { X0 = opaque(X0); S0 = opaque(S0); }
    if (!FIXED && !RES) { X1 = opaque(X1); S1 = opaque(S1); }

    // 3-row sliding windows of the keyframe maps (own column): top / mid / (bot loaded per row)
    float ft[CH], fm[CH], st[CH], sm[CH];
    {
      const unsigned ot = (unsigned)(max(y0 - 1, 0) * W + xc), om = (unsigned)(min(y0, H - 1) * W + xc);
      if (FIXED) {
        const float *xt = X0 + ot, *xm = X0 + om, *zt = S0 + ot, *zm = S0 + om;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          ft[c] = __ldg(xt + c * PLANE);
          fm[c] = __ldg(xm + c * PLANE);
          st[c] = __ldg(zt + c * splane);
          sm[c] = __ldg(zm + c * splane);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          ft[c] = ldf(X0, ot + c * iplane);
          fm[c] = ldf(X0, om + c * iplane);
          st[c] = ldf(S0, ot + c * splane);
          sm[c] = ldf(S0, om + c * splane);
        }
      }
    }
    for (int y = y0; y < y1; ++y) {
      const unsigned ob = (unsigned)(min(y + 1, H - 1) * W + xc);
      float fb[CH], sb[CH];
      if (FIXED) {
        const float *xr = X0 + ob, *zr = S0 + ob;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          fb[c] = __ldg(xr + c * PLANE);
          sb[c] = __ldg(zr + c * splane);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          fb[c] = ldf(X0, ob + c * iplane);
          sb[c] = ldf(S0, ob + c * splane);
        }
      }
      const unsigned o = (unsigned)(y * W + xc);
      const float d0 = __ldg(g.d0 + o);
      const float py = div_by(xsub((float)y, cy), fy, rcp_fy);

      float u, v, inv_z;
      {
        // SE(3) warp (geometry.py:291-323): w = ((r0 x + r1 y) + r2) + t d, every step rounded on its own
        const float4 ra = lds_v4(spose), rb = lds_v4(spose + 4), rc = lds_v4(spose + 8);
        const float wx = xadd(xadd(xadd(xmul(ra.x, px), xmul(ra.y, py)), ra.z), xmul(rc.y, d0));
        const float wy = xadd(xadd(xadd(xmul(ra.w, px), xmul(rb.x, py)), rb.y), xmul(rc.z, d0));
        const float wz = xadd(xadd(xadd(xmul(rb.z, px), xmul(rb.w, py)), rc.x), xmul(rc.w, d0));
        const float rz = __frcp_rn(wz);
        u = xadd(xmul(div_by(wx, wz, rz), fx), cx);
        v = xadd(xmul(div_by(wy, wz, rz), fy), cy);
        inv_z = div_by(d0, wz, rz);
      }
      const Tap tap = make_tap_r(u, v, H, W, rcp_hw, rcp_hh);
      float d1w;
      if (RES) {
        const float* q = D1 + tap.o;
        d1w = blend_exact(q[0], q[1], q[Wu], q[Wu + 1], tap);
      } else {
        d1w = sample_exact(D1, tap, W);
      }
      bool occ = occluded(u, v, inv_z, d1w, H, W);
      if (g.m0) occ = occ || (__ldg(g.m0 + o) == 0);
      if (g.m1) occ = occ || !(sample_mask(g.m1, tap, W) > 0.f);
      if (TRU) {
        const float s0c0 = (c0 == 0) ? sm[0] : __ldg(g.s0 + o);
        occ = occ || (s0c0 == g.s0lo) || (s0c0 == g.s0hi);
      }

      float saa = 0.f, sab = 0.f, sbb = 0.f, sar = 0.f, sbr = 0.f, sca = 0.f, scb = 0.f;
      float pmin = CUDART_INF_F, pmax = -CUDART_INF_F, sr0 = 0.f;
#ifndef DPFT_GATHER_GROUP
#define DPFT_GATHER_GROUP 4
#endif
      constexpr int G = CH < DPFT_GATHER_GROUP ? CH : DPFT_GATHER_GROUP;   // channels whose 8*G lookups are in flight together
#pragma unroll
      for (int g0 = 0; g0 < CH; g0 += G) {
        float xa[G], xb[G], xc_[G], xd[G], za[G], zb[G], zc[G], zd[G];
        if (FIXED) {
          const float *xq = X1 + tap.o, *zq = S1 + tap.o;
#pragma unroll
          for (int c = 0; c < G; ++c) {
            const int k = (g0 + c) * PLANE;
            const unsigned kz = (unsigned)(g0 + c) * splane;
            xa[c] = __ldg(xq + k); xb[c] = __ldg(xq + k + 1); xc_[c] = __ldg(xq + k + GW); xd[c] = __ldg(xq + k + GW + 1);
            za[c] = __ldg(zq + kz); zb[c] = __ldg(zq + kz + 1); zc[c] = __ldg(zq + kz + GW); zd[c] = __ldg(zq + kz + GW + 1);
          }
        } else if (RES) {
          const float *xq = X1 + tap.o, *zq = S1 + tap.o;
#pragma unroll
          for (int c = 0; c < G; ++c) {
            const unsigned k = (unsigned)(g0 + c) * iplane, kz = (unsigned)(g0 + c) * splane;
            xa[c] = xq[k]; xb[c] = xq[k + 1]; xc_[c] = xq[k + Wu]; xd[c] = xq[k + Wu + 1];
            za[c] = zq[kz]; zb[c] = zq[kz + 1]; zc[c] = zq[kz + Wu]; zd[c] = zq[kz + Wu + 1];
          }
        } else {
#pragma unroll
          for (int c = 0; c < G; ++c) {
            const unsigned ia = (unsigned)tap.o + (unsigned)(g0 + c) * iplane, ic = ia + Wu;
            const unsigned ja = (unsigned)tap.o + (unsigned)(g0 + c) * splane, jc = ja + Wu;
            ldf2(X1, ia, xa[c], xb[c]); ldf2(X1, ic, xc_[c], xd[c]);
            ldf2(S1, ja, za[c], zb[c]); ldf2(S1, jc, zc[c], zd[c]);
          }
        }
        float gfx[G], gfy[G], gsx[G], gsy[G];
#pragma unroll
        for (int c = 0; c < G; ++c) {
          const int k = g0 + c;
          // unit Sobel gradient of x0 and sigma0 (algorithms.py:1844-1865), separable form:
          // Sx = vs(x+1) - vs(x-1), Sy = vd(x-1) + 2 vd(x) + vd(x+1), vs = t+2m+b, vd = b-t
          const float fvs = ft[k] + 2.f * fm[k] + fb[k], fvd = fb[k] - ft[k];
          const float svs = st[k] + 2.f * sm[k] + sb[k], svd = sb[k] - st[k];
          const float fSx = __shfl_down_sync(0xffffffffu, fvs, 1) - __shfl_up_sync(0xffffffffu, fvs, 1);
          const float fSy = __shfl_up_sync(0xffffffffu, fvd, 1) + 2.f * fvd + __shfl_down_sync(0xffffffffu, fvd, 1);
          const float sSx = __shfl_down_sync(0xffffffffu, svs, 1) - __shfl_up_sync(0xffffffffu, svs, 1);
          const float sSy = __shfl_up_sync(0xffffffffu, svd, 1) + 2.f * svd + __shfl_down_sync(0xffffffffu, svd, 1);
          const float fin = rsqrt_fast(fmaf(fSx, fSx, fmaf(fSy, fSy, 1e-8f)));
          const float sin_ = rsqrt_fast(fmaf(sSx, sSx, fmaf(sSy, sSy, 1e-8f)));
          gfx[c] = fSx * fin; gfy[c] = fSy * fin; gsx[c] = sSx * sin_; gsy[c] = sSy * sin_;
        }
#pragma unroll
        for (int c = 0; c < G; ++c) {
          const int k = g0 + c;
          const float fr = blend_fast(xa[c], xb[c], xc_[c], xd[c], tap);
          // sigma is compared for equality against its batch extremes -> mask-grade arithmetic
          const float sr = TRU ? blend_exact(za[c], zb[c], zc[c], zd[c], tap) : blend_fast(za[c], zb[c], zc[c], zd[c], tap);
          // residual, its uncertainty and the 2-vector d(wres)/d(u,v) (algorithms.py:1969-1972, :872)
          const float res = fr - fm[k];
          const float s0v = sm[k];
          const float rs = rsqrt_fast(fmaf(sr, sr, s0v * s0v));   // 1 / sigma
          const float wres = res * rs;
          const float q = wres * (s0v * (rs * rs));           // res * sigma0 / sigma^3
          const float a = fmaf(gfx[c], rs, q * gsx[c]);
          const float bq = fmaf(gfy[c], rs, q * gsy[c]);
          const float wm = occ ? 1e-6f : wres;
          saa = fmaf(a, a, saa);
          sab = fmaf(a, bq, sab);
          sbb = fmaf(bq, bq, sbb);
          sar = fmaf(a, wm, sar);
          sbr = fmaf(bq, wm, sbr);
          if (TRU) {
            const float dw = wres - 1e-6f;
            sca = fmaf(a, dw, sca);
            scb = fmaf(bq, dw, scb);
            pmin = fminf(pmin, sr);
            pmax = fmaxf(pmax, sr);
            if (c0 == 0 && k == 0) sr0 = sr;
          }
        }
      }
      if (TRU && c0 != 0) {
        // channel 0 decides the mask; seeing it in every pass keeps the running extremes comparable
        if (RES) {
          const float* q = live + (unsigned)C * iplane + tap.o;
          sr0 = blend_exact(q[0], q[1], q[Wu], q[Wu + 1], tap);
        } else {
          sr0 = sample_exact(g.s1, tap, W);
        }
        pmin = fminf(pmin, sr0);
        pmax = fmaxf(pmax, sr0);
      }
      // halo lanes and columns past the image contribute nothing
      if (!col_out) { saa = sab = sbb = sar = sbr = 0.f; }
      float ju[6], jv[6];
      warp_rows(px, py, d0, fx, fy, ju, jv);
      accumulate_system(S.acc, ju, jv, saa, sab, sbb, sar, sbr);
      if (TRU) {
        // running extremes of the warped sigma and what their pixels added to J^T r.  New extremes and
        // ties are rare after the first rows, so the bookkeeping sits behind one warp-uniform branch.
        const bool lo = col_out && (pmin < S.vmin), hi = col_out && (pmax > S.vmax);
        const float nmin = lo ? pmin : S.vmin, nmax = hi ? pmax : S.vmax;
        const bool tmin = col_out && !occ && (sr0 == nmin), tmax = col_out && !occ && (sr0 == nmax);
        if (__any_sync(0xffffffffu, lo || hi || tmin || tmax)) {
          S.vmin = nmin;
          S.vmax = nmax;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            float cc = 0.f;
            if (i != 4) cc = fmaf(sca, ju[i], cc);
            if (i != 3) cc = fmaf(scb, jv[i], cc);
            if (lo || tmin) scorr[i][lane] = (lo ? 0.f : scorr[i][lane]) + (tmin ? cc : 0.f);
            if (hi || tmax) scorr[6 + i][lane] = (hi ? 0.f : scorr[6 + i][lane]) + (tmax ? cc : 0.f);
          }
        }
      }
      if (g.occ_out && c0 == 0 && col_out) {
        g.occ_out[(size_t)y * W + x] = occ ? 1 : 0;
        if (TRU) g.sr0_dbg[(size_t)y * W + x] = sr0;
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        ft[c] = fm[c];
        fm[c] = fb[c];
        st[c] = sm[c];
        sm[c] = sb[c];
      }
    }
  }
}

}  // namespace dpft
