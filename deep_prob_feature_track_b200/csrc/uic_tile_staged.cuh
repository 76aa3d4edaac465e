// Warp tile with a shared-memory staged lookup footprint.
//
// process_tile (uic_tile.cuh) fetches the 68 bilinear taps of a pixel (x1, sigma1: 8 channels x 4 texels, plus
// the 4 texels of the live inverse depth) with scalar global loads and waits for them twice per row; that wait
// is where its time goes (profiles/r1_uic_iter_kernel_level0.txt).  But the taps of a warp row are not
// scattered: the 30 pixels of a row land on two or three consecutive SOURCE rows, ~32 consecutive texels
// each, and the next tile row needs the same source rows shifted down by one.  So every warp keeps a ring of
// kStageRows source rows x 17 maps x kStageWidth texels in shared memory, fills it with 16-byte cp.async.cg
// (L2 -> shared, no L1 allocation, no registers) a row or two AHEAD of the row that needs it, and the taps
// become LDS at immediate offsets.  Every source texel is fetched once per tile instead of up to four times.
//
// Exactness: the staged values are the same floats; the blend arithmetic is untouched.  A lane whose footprint
// is not resident (large parallax spread, a jump in the warp field, image border clamps far from the segment)
// takes the direct global loads instead -- per lane, behind one warp-uniform branch -- so the results do not
// depend on what is staged.  Requires C == CH == 8, W % 4 == 0, W >= kStageWidth and 16-byte aligned maps.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "uic_tile.cuh"

namespace dpft {

constexpr int kStageRows = 4;        // ring depth (power of two: slot = source row & 3)
constexpr int kStageWidth = 44;      // texels per staged row segment (30 output columns + margin)
constexpr int kStageMaps = 17;       // x1[0..7], sigma1[0..7], invd1
constexpr int kStageLookahead = 2;   // source rows requested ahead of the row being computed
// slot stride = 0 mod 32 banks: lanes of one warp row that sit on different source rows (same map, distinct
// columns) then never share a bank
constexpr int kStageSlotFloats = (kStageMaps * kStageWidth + 31) / 32 * 32;
constexpr int kStageWarpFloats = kStageRows * kStageSlotFloats + 32;   // + x origin of each slot (as int)

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const float* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_wait(int pending) {
  // cp.async groups complete in order; `pending` newest groups may stay in flight
  if (pending <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
  else asm volatile("cp.async.wait_group 2;" ::: "memory");
}

struct TapXY {      // Tap plus the integer texel coordinates of its north-west corner
  Tap t;
  int xi, yi;
};

__device__ __forceinline__ TapXY make_tap_xy(float u, float v, int H, int W, float rcp_half_w, float rcp_half_h) {
  const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
  const float gx = xsub(div_by(u, hw, rcp_half_w), 1.f), gy = xsub(div_by(v, hh, rcp_half_h), 1.f);
  const float ix = fminf(fmaxf(xmul(xmul(xadd(gx, 1.f), 0.5f), (float)(W - 1)), 0.f), (float)(W - 1));
  const float iy = fminf(fmaxf(xmul(xmul(xadd(gy, 1.f), 0.5f), (float)(H - 1)), 0.f), (float)(H - 1));
  const float xw = floorf(ix), yn = floorf(iy);
  float txr = xsub(ix, xw), tys = xsub(iy, yn);
  float txl = xsub(xadd(xw, 1.f), ix), tyn = xsub(xadd(yn, 1.f), iy);
  int xi = (int)xw, yi = (int)yn;
  if (xi > W - 2) { xi = W - 2; const float t = txl; txl = txr; txr = t; }
  if (yi > H - 2) { yi = H - 2; const float t = tyn; tyn = tys; tys = t; }
  xi = max(xi, 0);
  yi = max(yi, 0);
  TapXY r;
  r.t.wa = xmul(txl, tyn);
  r.t.wb = xmul(txr, tyn);
  r.t.wc = xmul(txl, tys);
  r.t.wd = xmul(txr, tys);
  r.t.o = yi * W + xi;
  r.xi = xi;
  r.yi = yi;
  return r;
}

template <bool TRU, int GW = 0, int GH = 0>
__device__ __forceinline__ void process_tile_staged(const PairView& g, const float* spose, float (*scorr)[33],
                                                    float* ring /* kStageWarpFloats of this warp */, const int seg,
                                                    const int y0, const int y1, const int lane, TileSums& S) {
  constexpr int CH = 8;
  constexpr bool FIXED = GW > 0 && GH > 0;
  constexpr int PLANE = GW * GH;
  constexpr int SW = kStageWidth, NM = kStageMaps, CPR = SW / 4;   // 16-byte chunks per staged map row
  const int H = FIXED ? GH : g.H, W = FIXED ? GW : g.W;
  const unsigned iplane = (unsigned)(H * W), Wu = (unsigned)W;
  const int x = seg * kTileCols - 1 + lane;
  const int xc = min(max(x, 0), W - 1);
  const bool col_out = lane >= 1 && lane <= kTileCols && x < W;
  const float fx = g.fx, fy = g.fy, cx = g.cx, cy = g.cy;
  const float px = xdiv(xsub((float)xc, cx), fx);
  const float rcp_fy = __frcp_rn(fy);
  const float rcp_hw = __frcp_rn(0.5f * (float)(W - 1)), rcp_hh = __frcp_rn(0.5f * (float)(H - 1));
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  int* slot_xs = reinterpret_cast<int*>(ring + kStageRows * kStageSlotFloats);

  const float* X0 = g.x0;
  const float* S0 = g.s0;
  const float* X1 = g.x1;
  const float* S1 = g.s1;
  if (!FIXED) { X0 = opaque(X0); S0 = opaque(S0); X1 = opaque(X1); S1 = opaque(S1); }

  // one source row of all 17 maps -> ring slot (row & 3), columns [xs, xs + SW)
  auto stage_row = [&](const int row, const int xs) {
    const unsigned slot_s = ring_s + (unsigned)((row & (kStageRows - 1)) * kStageSlotFloats) * 4u;
    const unsigned src_off = (unsigned)(row * W + xs);
#pragma unroll
    for (int k = 0; k < (NM * CPR + 31) / 32; ++k) {
      const int i = lane + 32 * k;
      if (i < NM * CPR) {
        const int m = i / CPR, ch = i - m * CPR;
        const float* base = (m < CH) ? X1 : (m < 2 * CH) ? S1 : g.d1;
        const unsigned plane_off = (m < 2 * CH) ? (unsigned)(m & (CH - 1)) * iplane : 0u;
        const float* src;
        asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(src) : "r"(src_off + plane_off + 4u * ch), "l"(base));
        cp_async16(slot_s + (unsigned)(m * SW + 4 * ch) * 4u, src);
      }
    }
    if (lane == 0) slot_xs[row & (kStageRows - 1)] = xs;
    stage_commit();
  };

  float ft[CH], fm[CH], st[CH], sm[CH];
  {
    const unsigned ot = (unsigned)(max(y0 - 1, 0) * W + xc), om = (unsigned)(min(y0, H - 1) * W + xc);
    if (FIXED) {
      const float *xt = X0 + ot, *xm = X0 + om, *zt = S0 + ot, *zm = S0 + om;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        ft[c] = __ldg(xt + c * PLANE);
        fm[c] = __ldg(xm + c * PLANE);
        st[c] = __ldg(zt + c * PLANE);
        sm[c] = __ldg(zm + c * PLANE);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        ft[c] = ldf(X0, ot + c * iplane);
        fm[c] = ldf(X0, om + c * iplane);
        st[c] = ldf(S0, ot + c * iplane);
        sm[c] = ldf(S0, om + c * iplane);
      }
    }
  }
  // the ring holds source rows max(base, top - 3) .. top; base = first row requested since the last restart
  int top = -0x40000000, base = 0x40000000;
  float d0_next = (y0 < y1) ? __ldg(g.d0 + (unsigned)(y0 * W + xc)) : 0.f;

  for (int y = y0; y < y1; ++y) {
    const unsigned ob = (unsigned)(min(y + 1, H - 1) * W + xc);
    float fb[CH], sb[CH];
    if (FIXED) {
      const float *xr = X0 + ob, *zr = S0 + ob;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        fb[c] = __ldg(xr + c * PLANE);
        sb[c] = __ldg(zr + c * PLANE);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        fb[c] = ldf(X0, ob + c * iplane);
        sb[c] = ldf(S0, ob + c * iplane);
      }
    }
    const unsigned o = (unsigned)(y * W + xc);
    const float d0 = d0_next;
    if (y + 1 < y1) d0_next = __ldg(g.d0 + o + Wu);
    const float py = div_by(xsub((float)y, cy), fy, rcp_fy);

    float u, v, inv_z;
    {
      const float4 ra = lds_v4(spose), rb = lds_v4(spose + 4), rc = lds_v4(spose + 8);
      const float wx = xadd(xadd(xadd(xmul(ra.x, px), xmul(ra.y, py)), ra.z), xmul(rc.y, d0));
      const float wy = xadd(xadd(xadd(xmul(ra.w, px), xmul(rb.x, py)), rb.y), xmul(rc.z, d0));
      const float wz = xadd(xadd(xadd(xmul(rb.z, px), xmul(rb.w, py)), rc.x), xmul(rc.w, d0));
      const float rz = __frcp_rn(wz);
      u = xadd(xmul(div_by(wx, wz, rz), fx), cx);
      v = xadd(xmul(div_by(wy, wz, rz), fy), cy);
      inv_z = div_by(d0, wz, rz);
    }
    const TapXY txy = make_tap_xy(u, v, H, W, rcp_hw, rcp_hh);
    const Tap& tap = txy.t;

    // ---- footprint of this warp row and the ring ----------------------------------------------------------
    const int ylo = __reduce_min_sync(0xffffffffu, col_out ? txy.yi : 0x3fffffff);
    const int yhi = __reduce_max_sync(0xffffffffu, col_out ? txy.yi + 1 : -1);
    const int xlo = __reduce_min_sync(0xffffffffu, col_out ? txy.xi : 0x3fffffff);
    int ready_top = top;
    if (yhi >= 0) {
      if (ylo > top + 1 || ylo < base - 2) {            // a jump in the warp field: restart the ring at ylo
        top = ylo - 1;
        base = ylo;
      }
      const int keep = max(max(base, top - (kStageRows - 1)), ylo);      // lowest row this warp row still needs
      const int target = min(min(yhi + kStageLookahead, keep + kStageRows - 1), H - 1);
      const int xs = min(max((xlo - (SW - 32) / 2) & ~3, 0), W - SW);
      __syncwarp();                                                      // every lane is done with the old slots
      while (top < target) stage_row(++top, xs);
      ready_top = min(yhi, top);
      stage_wait(top - ready_top);
      __syncwarp();
    }
    const int lowest = max(base, top - (kStageRows - 1));
    const int s0i = txy.yi & (kStageRows - 1), s1i = (txy.yi + 1) & (kStageRows - 1);
    const int xs0 = slot_xs[s0i], xs1 = slot_xs[s1i];
    const bool resident = txy.yi >= lowest && txy.yi + 1 <= ready_top && txy.xi >= xs0 && txy.xi + 1 < xs0 + SW &&
                          txy.xi >= xs1 && txy.xi + 1 < xs1 + SW;
    const bool any_direct = __any_sync(0xffffffffu, col_out && !resident);
    // north-west / south-west texel of map 0 in the ring (kept in range for lanes that are not resident)
    const float* a0 = ring + s0i * kStageSlotFloats + min(max(txy.xi - xs0, 0), SW - 2);
    const float* a1 = ring + s1i * kStageSlotFloats + min(max(txy.xi - xs1, 0), SW - 2);

    float d1w;
    {
      float da = a0[2 * CH * SW], db = a0[2 * CH * SW + 1];
      float dc = a1[2 * CH * SW], dd = a1[2 * CH * SW + 1];
      if (any_direct && !resident) {
        const float* q = g.d1 + tap.o;
        da = __ldg(q); db = __ldg(q + 1); dc = __ldg(q + W); dd = __ldg(q + W + 1);
      }
      d1w = blend_exact(da, db, dc, dd, tap);
    }
    bool occ = occluded(u, v, inv_z, d1w, H, W);
    if (g.m0) occ = occ || (__ldg(g.m0 + o) == 0);
    if (g.m1) occ = occ || !(sample_mask(g.m1, tap, W) > 0.f);
    if (TRU) occ = occ || (sm[0] == g.s0lo) || (sm[0] == g.s0hi);

    float saa = 0.f, sab = 0.f, sbb = 0.f, sar = 0.f, sbr = 0.f, sca = 0.f, scb = 0.f;
    float pmin = CUDART_INF_F, pmax = -CUDART_INF_F, sr0 = 0.f;
    constexpr int G = DPFT_GATHER_GROUP;
#pragma unroll
    for (int g0 = 0; g0 < CH; g0 += G) {
      float xa[G], xb[G], xc_[G], xd[G], za[G], zb[G], zc[G], zd[G];
#pragma unroll
      for (int c = 0; c < G; ++c) {
        const int kx = (g0 + c) * SW, kz = (CH + g0 + c) * SW;
        xa[c] = a0[kx]; xb[c] = a0[kx + 1]; xc_[c] = a1[kx]; xd[c] = a1[kx + 1];
        za[c] = a0[kz]; zb[c] = a0[kz + 1]; zc[c] = a1[kz]; zd[c] = a1[kz + 1];
      }
      if (any_direct && !resident) {
#pragma unroll
        for (int c = 0; c < G; ++c) {
          const unsigned ia = (unsigned)tap.o + (unsigned)(g0 + c) * iplane, ic = ia + Wu;
          ldf2(X1, ia, xa[c], xb[c]); ldf2(X1, ic, xc_[c], xd[c]);
          ldf2(S1, ia, za[c], zb[c]); ldf2(S1, ic, zc[c], zd[c]);
        }
      }
      float gfx[G], gfy[G], gsx[G], gsy[G];
#pragma unroll
      for (int c = 0; c < G; ++c) {
        const int k = g0 + c;
#ifdef DPFT_EXPERIMENT_NO_SOBEL   // timing experiment only: wrong results
        gfx[c] = ft[k]; gfy[c] = fb[k]; gsx[c] = st[k]; gsy[c] = sb[k];
        continue;
#endif
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
        const float sr = TRU ? blend_exact(za[c], zb[c], zc[c], zd[c], tap) : blend_fast(za[c], zb[c], zc[c], zd[c], tap);
        const float res = fr - fm[k];
        const float s0v = sm[k];
        const float rs = rsqrt_fast(fmaf(sr, sr, s0v * s0v));
        const float wres = res * rs;
        const float q = wres * (s0v * (rs * rs));
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
          if (k == 0) sr0 = sr;
        }
      }
    }
    if (!col_out) { saa = sab = sbb = sar = sbr = 0.f; }
    float ju[6], jv[6];
    warp_rows(px, py, d0, fx, fy, ju, jv);
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
    if (g.occ_out && col_out) {
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
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // nothing of this warp may still land in the ring
}

}  // namespace dpft
