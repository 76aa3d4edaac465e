// Software-pipelined variant of the warp tile (uic_tile.cuh): the 68 bilinear lookups of a tile row
// (x1, sigma1: 8 channels x 4 texels each, plus the 4 texels of the live inverse depth) are issued with
// cp.async one row AHEAD into a two-stage shared-memory buffer, and the keyframe rows / inverse depth are
// prefetched one row ahead into registers.  The measured limiter of the synchronous routine is the time a
// warp sits on the scoreboard waiting for exactly these loads (profiles/r1_uic_iter_kernel_level0.txt);
// here row y+1's loads are in flight while row y is being computed, and they do not occupy registers.
//
// Same arithmetic, same rounding, same results as process_tile (tests run both).  8 channels per pass.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "uic_tile.cuh"

namespace dpft {

constexpr int kAsyncVals = 68;                         // per lane and row: 32 x1 + 32 sigma1 + 4 invd1 texels
constexpr int kAsyncStageFloats = kAsyncVals * 32;     // one stage of one warp
constexpr int kAsyncWarpFloats = 2 * kAsyncStageFloats;

__device__ __forceinline__ void cp_async4(unsigned smem_addr, const float* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct RowGeom {      // pose-dependent per-pixel state of one row, carried from issue to consume
  float u, v, inv_z, py, d0;
  Tap tap;
};

// stage layout (floats): [k][lane], k = 4*c + {0..3} for x1 channel c texels a,b,c,d; 32 + 4*c + {0..3} for
// sigma1; 64..67 for invd1.
template <bool TRU>
__device__ __forceinline__ void issue_row(const PairView& g, const float* X1, const float* S1, const unsigned iplane,
                                          const unsigned Wu, const Tap& tap, float* stage, const int lane) {
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage) + 4u * lane;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const unsigned ia = (unsigned)tap.o + (unsigned)c * iplane, ic = ia + Wu;
    const float *qa, *qc, *za, *zc;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(qa) : "r"(ia), "l"(X1));
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(qc) : "r"(ic), "l"(X1));
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(za) : "r"(ia), "l"(S1));
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(zc) : "r"(ic), "l"(S1));
    cp_async4(sbase + 128u * (4 * c + 0), qa);
    cp_async4(sbase + 128u * (4 * c + 1), qa + 1);
    cp_async4(sbase + 128u * (4 * c + 2), qc);
    cp_async4(sbase + 128u * (4 * c + 3), qc + 1);
    cp_async4(sbase + 128u * (32 + 4 * c + 0), za);
    cp_async4(sbase + 128u * (32 + 4 * c + 1), za + 1);
    cp_async4(sbase + 128u * (32 + 4 * c + 2), zc);
    cp_async4(sbase + 128u * (32 + 4 * c + 3), zc + 1);
  }
  const float* d = g.d1 + tap.o;
  cp_async4(sbase + 128u * 64, d);
  cp_async4(sbase + 128u * 65, d + 1);
  cp_async4(sbase + 128u * 66, d + g.W);
  cp_async4(sbase + 128u * 67, d + g.W + 1);
}

template <bool TRU>
__device__ __forceinline__ void process_tile_async(const PairView& g, const float* spose, float (*scorr)[33],
                                                   float* wstage /* kAsyncWarpFloats of this warp */, const int seg,
                                                   const int y0, const int y1, const int lane, TileSums& S) {
  constexpr int CH = 8;
  const int H = g.H, W = g.W, C = g.C;
  const unsigned iplane = (unsigned)(H * W), Wu = (unsigned)W;
  const int x = seg * kTileCols - 1 + lane;
  const int xc = min(max(x, 0), W - 1);
  const bool col_out = lane >= 1 && lane <= kTileCols && x < W;
  const float fx = g.fx, fy = g.fy, cx = g.cx, cy = g.cy;
  const float px = xdiv(xsub((float)xc, cx), fx);

  auto geometry = [&](const int y, const float d0) {
    RowGeom r;
    r.d0 = d0;
    r.py = xdiv(xsub((float)y, cy), fy);
    const float4 ra = lds_v4(spose), rb = lds_v4(spose + 4), rc = lds_v4(spose + 8);
    const float wx = xadd(xadd(xadd(xmul(ra.x, px), xmul(ra.y, r.py)), ra.z), xmul(rc.y, d0));
    const float wy = xadd(xadd(xadd(xmul(ra.w, px), xmul(rb.x, r.py)), rb.y), xmul(rc.z, d0));
    const float wz = xadd(xadd(xadd(xmul(rb.z, px), xmul(rb.w, r.py)), rc.x), xmul(rc.w, d0));
    r.u = xadd(xmul(xdiv(wx, wz), fx), cx);
    r.v = xadd(xmul(xdiv(wy, wz), fy), cy);
    r.inv_z = xdiv(d0, wz);
    r.tap = make_tap(r.u, r.v, H, W);
    return r;
  };

  for (int c0 = 0; c0 < C; c0 += CH) {
    const float* X0 = opaque(g.x0 + (size_t)c0 * iplane);
    const float* S0 = opaque(g.s0 + (size_t)c0 * iplane);
    const float* X1 = opaque(g.x1 + (size_t)c0 * iplane);
    const float* S1 = opaque(g.s1 + (size_t)c0 * iplane);

    // keyframe windows: four register sets per map (top, mid, bottom, and the row after it, loaded a row
    // ahead); the sets rotate roles from row to row by unrolling the row loop four times -- a register copy of a
    // value that is still in flight would stall on it and undo the prefetch
    float fA[CH], fB[CH], fC[CH], fD[CH], sA[CH], sB[CH], sC[CH], sD[CH];
    {
      const unsigned ot = (unsigned)(max(y0 - 1, 0) * W + xc), om = (unsigned)(min(y0, H - 1) * W + xc);
      const unsigned ob = (unsigned)(min(y0 + 1, H - 1) * W + xc);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        fA[c] = ldf(X0, ot + c * iplane);
        fB[c] = ldf(X0, om + c * iplane);
        fC[c] = ldf(X0, ob + c * iplane);
        sA[c] = ldf(S0, ot + c * iplane);
        sB[c] = ldf(S0, om + c * iplane);
        sC[c] = ldf(S0, ob + c * iplane);
      }
    }
    // prologue: geometry of the first row, its lookups into stage 0
    RowGeom gA = geometry(y0, __ldg(g.d0 + (unsigned)(y0 * W + xc))), gB = gA;
    issue_row<TRU>(g, X1, S1, iplane, Wu, gA.tap, wstage, lane);
    cp_async_commit();
    float d0_next = (y0 + 1 < y1) ? __ldg(g.d0 + (unsigned)((y0 + 1) * W + xc)) : 0.f;

    // one row: ft/fm/fb (st/sm/sb) are the window of row y, fn/sn receive row y+2; cur is row y's geometry,
    // nxt receives row y+1's
    auto row = [&](const float (&ft)[CH], const float (&fm)[CH], const float (&fb)[CH], float (&fn)[CH],
                   const float (&st)[CH], const float (&sm)[CH], const float (&sb)[CH], float (&sn)[CH],
                   const RowGeom& cur, RowGeom& nxt, const int y) {
      float* stage = wstage + ((y - y0) & 1) * kAsyncStageFloats;
      // ---- a row ahead: geometry of row y+1, its lookups into the other stage, keyframe row y+2, inverse depth y+2
      if (y + 1 < y1) {
        nxt = geometry(y + 1, d0_next);
        issue_row<TRU>(g, X1, S1, iplane, Wu, nxt.tap, wstage + ((y + 1 - y0) & 1) * kAsyncStageFloats, lane);
        const unsigned on = (unsigned)(min(y + 2, H - 1) * W + xc);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          fn[c] = ldf(X0, on + c * iplane);
          sn[c] = ldf(S0, on + c * iplane);
        }
        if (y + 2 < y1) d0_next = __ldg(g.d0 + (unsigned)((y + 2) * W + xc));
      }
      cp_async_commit();
      cp_async_wait<1>();       // everything but the group just committed has landed: row y is in `stage`
      __syncwarp();

      // ---- consume row y
      const float* sl = stage + lane;
      const Tap& tap = cur.tap;
      const float d1w = blend_exact(sl[64 * 32], sl[65 * 32], sl[66 * 32], sl[67 * 32], tap);
      bool occ = occluded(cur.u, cur.v, cur.inv_z, d1w, H, W);
      const unsigned o = (unsigned)(y * W + xc);
      if (g.m0) occ = occ || (__ldg(g.m0 + o) == 0);
      if (g.m1) occ = occ || !(sample_mask(g.m1, tap, W) > 0.f);
      if (TRU) {
        const float s0c0 = (c0 == 0) ? sm[0] : __ldg(g.s0 + o);
        occ = occ || (s0c0 == g.s0lo) || (s0c0 == g.s0hi);
      }
      float saa = 0.f, sab = 0.f, sbb = 0.f, sar = 0.f, sbr = 0.f, sca = 0.f, scb = 0.f;
      float pmin = CUDART_INF_F, pmax = -CUDART_INF_F, sr0 = 0.f;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        // unit Sobel gradient of x0 and sigma0 (algorithms.py:1844-1865), separable form
        const float fvs = ft[k] + 2.f * fm[k] + fb[k], fvd = fb[k] - ft[k];
        const float svs = st[k] + 2.f * sm[k] + sb[k], svd = sb[k] - st[k];
        const float fSx = __shfl_down_sync(0xffffffffu, fvs, 1) - __shfl_up_sync(0xffffffffu, fvs, 1);
        const float fSy = __shfl_up_sync(0xffffffffu, fvd, 1) + 2.f * fvd + __shfl_down_sync(0xffffffffu, fvd, 1);
        const float sSx = __shfl_down_sync(0xffffffffu, svs, 1) - __shfl_up_sync(0xffffffffu, svs, 1);
        const float sSy = __shfl_up_sync(0xffffffffu, svd, 1) + 2.f * svd + __shfl_down_sync(0xffffffffu, svd, 1);
        const float fin = rsqrt_fast(fmaf(fSx, fSx, fmaf(fSy, fSy, 1e-8f)));
        const float sin_ = rsqrt_fast(fmaf(sSx, sSx, fmaf(sSy, sSy, 1e-8f)));
        const float gfx = fSx * fin, gfy = fSy * fin, gsx = sSx * sin_, gsy = sSy * sin_;

        const float* q = sl + (4 * k) * 32;
        const float fr = blend_fast(q[0], q[32], q[64], q[96], tap);
        const float* z = sl + (32 + 4 * k) * 32;
        const float sr = TRU ? blend_exact(z[0], z[32], z[64], z[96], tap) : blend_fast(z[0], z[32], z[64], z[96], tap);
        const float res = fr - fm[k];
        const float s0v = sm[k];
        const float rs = rsqrt_fast(fmaf(sr, sr, s0v * s0v));
        const float wres = res * rs;
        const float qq = wres * (s0v * (rs * rs));
        const float a = fmaf(gfx, rs, qq * gsx);
        const float bq = fmaf(gfy, rs, qq * gsy);
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
      if (TRU && c0 != 0) {
        sr0 = sample_exact(g.s1, tap, W);
        pmin = fminf(pmin, sr0);
        pmax = fmaxf(pmax, sr0);
      }
      if (!col_out) { saa = sab = sbb = sar = sbr = 0.f; }
      float ju[6], jv[6];
      warp_rows(px, cur.py, cur.d0, fx, fy, ju, jv);
      accumulate_system(S.acc, ju, jv, saa, sab, sbb, sar, sbr);
      if (TRU) {
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
      // the stage just consumed is overwritten by the lookups issued for the row after next: every lane must
      // be done reading it before any lane issues them
      __syncwarp();
    };

    for (int y = y0; y < y1; y += 4) {
      row(fA, fB, fC, fD, sA, sB, sC, sD, gA, gB, y);
      if (y + 1 < y1) row(fB, fC, fD, fA, sB, sC, sD, sA, gB, gA, y + 1);
      if (y + 2 < y1) row(fC, fD, fA, fB, sC, sD, sA, sB, gA, gB, y + 2);
      if (y + 3 < y1) row(fD, fA, fB, fC, sD, sA, sB, sC, gB, gA, y + 3);
    }
    cp_async_wait<0>();
    __syncwarp();
  }
}

}  // namespace dpft
