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
// Output columns per warp row: 30 (lanes 0 and 31 are halo columns of the horizontal Sobel taps, as in the plain
// kernel) or 32 (every lane is an output column; the two halo columns' vertical Sobel sums are formed by the lanes
// in turn -- one map and one side each -- a row ahead, handed over in shared memory and applied as corrections on
// lanes 0 and 31).  32 columns walk a 160-wide row in 5 segments instead of 6.
#ifndef DPFT_STAGED_COLS
#define DPFT_STAGED_COLS 30
#endif
constexpr int kStagedCols = DPFT_STAGED_COLS;
static_assert(kStagedCols == 30 || kStagedCols == 32, "30 or 32 output columns per warp row");
#ifndef DPFT_STAGE_WIDTH
#define DPFT_STAGE_WIDTH (DPFT_STAGED_COLS == 32 ? 48 : 44)     // tuning hook; 4 * (chunks per map row)
#endif
constexpr int kHaloFloats = kStagedCols == 32 ? 128 : 0;        // [row parity][side][16 maps: vs, vs, vd, vd per pair]
constexpr int kStageWidth = DPFT_STAGE_WIDTH;      // texels per staged row segment (30 output columns + margin)
constexpr int kStageMaps = 17;       // x1[0..7], sigma1[0..7], invd1
#ifndef DPFT_STAGE_LOOKAHEAD
#define DPFT_STAGE_LOOKAHEAD 2
#endif
constexpr int kStageLookahead = DPFT_STAGE_LOOKAHEAD;   // source rows requested ahead of the row being computed
// slot stride = 0 mod 32 banks: lanes of one warp row that sit on different source rows (same map, distinct
// columns) then never share a bank
constexpr int kStageSlotFloats = (kStageMaps * kStageWidth + 31) / 32 * 32;
constexpr int kStageWarpFloats = kStageRows * kStageSlotFloats + 32;   // + x origin of each slot (as int)
// One uncertainty map per frame (SB): a staged source row holds 10 maps (x1[0..7], sigma1, invd1) instead of 17, and a
// worker's ring shrinks from 12 to 7 KB -- what lets a fourth CTA of the work-queue kernel fit an SM.  The routine
// always lays its ring out by stage_maps(SB); callers that size their areas with the 17-map constants simply leave the
// tail unused.
__host__ __device__ constexpr int stage_maps(bool sb) { return sb ? 8 + 2 : kStageMaps; }
__host__ __device__ constexpr int stage_slot_floats(bool sb) { return (stage_maps(sb) * kStageWidth + 31) / 32 * 32; }
__host__ __device__ constexpr int stage_warp_floats(bool sb) { return kStageRows * stage_slot_floats(sb) + 32; }
// lanes of a row whose footprint is not resident park their own 17 x 4 taps here (see the direct-load body)
#ifndef DPFT_OUT_LANES
#define DPFT_OUT_LANES 4
#endif
constexpr int kOutLanes = DPFT_OUT_LANES;     // tuning hook
constexpr int kOutFloats = kOutLanes * kStageMaps * 4;
// shared memory of one warp: ring | corrections of remove_tru_sigma | outlier taps | halo sums; a multiple of 128
// bytes so that every ring slot is 128-byte aligned (what the tensor-map copies of DPFT_STAGED_TMA=1 require of their
// destination).  The 27 rows of the final reduction overlay the tail of the ring once the tile is done.
constexpr int kStageAreaFloats = (kStageWarpFloats + 12 * 33 + kOutFloats + kHaloFloats + 31) / 32 * 32;
__host__ __device__ constexpr int stage_area_floats(bool sb) {
  return (stage_warp_floats(sb) + 12 * 33 + kOutFloats + kHaloFloats + 31) / 32 * 32;
}
static_assert(stage_warp_floats(true) >= 27 * 33 && stage_warp_floats(true) % 4 == 0, "the reduction rows overlay the ring");
static_assert(stage_warp_floats(false) == kStageWarpFloats && stage_area_floats(false) == kStageAreaFloats, "17-map layout");
static_assert((kStageWarpFloats + 12 * 33 + kOutFloats) % 4 == 0, "the halo sums are read as float4");
static_assert(kStageWarpFloats >= 27 * 33, "the reduction rows overlay the ring");

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const float* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const float* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// TMA variant of the ring (-DDPFT_STAGED_TMA=1): a staged source row is three tensor-map copies issued by ONE lane
// (all x1 planes, all sigma1 planes, the inverse depth; box = stage width x 1 row x channels) and completion is
// tracked by one mbarrier per ring slot instead of cp.async groups.  Parity-green, measured slower than the cp.async
// ring (92 vs 87 us per level-0 launch, profiles/r1h_sweep_tma.txt), so it is off by default.
#ifndef DPFT_STAGED_TMA
#define DPFT_STAGED_TMA 0
#endif
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned mbar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned mbar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(mbar), "r"(parity)
      : "memory");
  return ok != 0;
}
// one box {SW texels, 1 row, nc channel planes, 1 pair} of a (W, H, C, B) tensor -> shared memory, [channel][texel]
__device__ __forceinline__ void tma_row_g2s(unsigned dst, const void* tmap, int x, int y, int b, unsigned mbar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(tmap), "r"(x), "r"(y), "r"(0), "r"(b), "r"(mbar)
      : "memory");
}
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

// SB ("sigma broadcast"): sigma0 / sigma1 are ONE map per frame -- what the reference's encoder emits before it
// repeats it to C channels (algorithms.py:1425-1427, uncertainty_channel = 1 in every shipped configuration).  The
// map is staged, looked up, blended and differentiated once per pixel instead of once per channel; the results
// are those of the repeated tensor.
// AUX: object masks and / or the per-pixel debug outputs may be present (their tests are compiled out otherwise).
// NARROW: maps narrower than the ring (W < kStageWidth, W % 4 == 0 -- the 30x40 and 15x20 levels of a 120x160 pyramid):
// a staged source row is the WHOLE map row at column origin 0; the chunks of a ring row past the map row re-read its
// last chunk (never looked up), so nothing is fetched from beyond a row.
template <bool TRU, bool SB = false, int GW = 0, int GH = 0, bool AUX = true, bool NARROW = false>
__device__ __forceinline__ void process_tile_staged(const PairView& g, const float* spose, float (*scorr)[33],
                                                    float* ring /* kStageWarpFloats of this warp */,
                                                    float* outl /* kOutFloats (+ kHaloFloats) of this warp */, const int seg,
                                                    const int y0, const int y1, const int lane, TileSums& S) {
  constexpr int CH = 8;
  constexpr bool FIXED = GW > 0 && GH > 0;
  constexpr int PLANE = GW * GH;
  constexpr int SW = kStageWidth, CPR = SW / 4;   // 16-byte chunks per staged map row
  constexpr int NM = stage_maps(SB), SLOT = stage_slot_floats(SB);   // maps and floats of one staged source row
  constexpr int DMAP = NM - 1;                    // map slot of the inverse depth (sigma1: slots CH.. or the one slot CH)
  const int H = FIXED ? GH : g.H, W = FIXED ? GW : g.W;
  const unsigned iplane = (unsigned)(H * W), Wu = (unsigned)W;
  constexpr bool WIDE = kStagedCols == 32;
  const int x = WIDE ? seg * 32 + lane : seg * kTileCols - 1 + lane;
  const int xc = min(max(x, 0), W - 1);
  const bool col_out = WIDE ? (x < W) : (lane >= 1 && lane <= kTileCols && x < W);
  const float fx = g.fx, fy = g.fy, cx = g.cx, cy = g.cy;
  const float px = xdiv(xsub((float)xc, cx), fx);
  const float rcp_fy = __frcp_rn(fy);
  const float rcp_hw = __frcp_rn(0.5f * (float)(W - 1)), rcp_hh = __frcp_rn(0.5f * (float)(H - 1));
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  int* slot_xs = reinterpret_cast<int*>(ring + kStageRows * SLOT);
  // TMA variant: four mbarriers behind the slot origins; `par` holds the phase parity to wait for per slot,
  // `inflight` the slots whose copy has not been confirmed yet (both warp-uniform)
#if DPFT_STAGED_TMA
  static_assert(!NARROW, "the tensor-map ring has no narrow form");
  const unsigned mbar0 = ring_s + (unsigned)(kStageRows * SLOT + 8) * 4u;
  unsigned par = 0u, inflight = 0u;
#endif

  const float* X0 = g.x0;
  const float* S0 = g.s0;
  const float* X1 = g.x1;
  const float* S1 = g.s1;
  X0 = opaque(X0);
  S0 = opaque(S0);
  // the live frame's bases stay in registers: left to itself ptxas re-derives them from the kernel parameters in
  // every row, and the constant load it uses for that ends up sharing a scoreboard with the sixteen window loads
  // issued next to it -- the first staging address of the row then waits for all of them (a DRAM round trip in
  // the serial part of every row; 18 % of the stall samples in profiles/r1h_uic_iter_staged_kernel_level0.txt)
  X1 = opaque(X1);
  S1 = opaque(S1);
  const float* D1 = opaque(g.d1);
  const float* D0 = opaque(g.d0);

  // one source row of all 17 maps -> ring slot (row & 3), columns [xs, xs + SW).  A map row is CPR 16-byte
  // chunks and SW = 4 CPR, so chunk i = 11 m + ch of a tensor lands at float 4 i of the slot: the destination is
  // 16 lane + an immediate, and the source offsets of this lane's three chunks are fixed for the whole tile.
  static_assert(SW == 4 * CPR, "chunk i of a tensor sits at float 4 i of the slot");
#if DPFT_STAGED_TMA
  // map rows per staged source row: x1 channels in map slots 0..7, sigma1 in 8..15 (one map in slot 8 with SB),
  // the inverse depth in slot 16
  constexpr int kBulkMaps = NM;
  auto wait_slot = [&](const int sl) {
    const unsigned mb = mbar0 + 8u * (unsigned)sl;
    int spins = 0;
    while (!mbar_try_wait(mb, (par >> sl) & 1u)) {
      if (++spins > (1 << 24)) __trap();               // a copy that never lands must not hang the device
    }
    par ^= 1u << sl;
    inflight &= ~(1u << sl);
  };
  if (lane == 0) {
#pragma unroll
    for (int sl = 0; sl < kStageRows; ++sl) mbar_init(mbar0 + 8u * sl, 1u);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto stage_row = [&](const int row, const int xs) {
    const int sl = row & (kStageRows - 1);
    if ((inflight >> sl) & 1u) wait_slot(sl);          // a staged row nobody waited for (lookahead past a restart)
    const unsigned mb = mbar0 + 8u * (unsigned)sl;
    // three tensor copies by one lane: the eight x1 planes, the sigma1 planes, the inverse depth (tensor maps of
    // the level's (W, H, C, B) tensors travel as kernel parameters; box = SW texels x 1 row x all channels)
    if (lane == 0) {
      slot_xs[sl] = xs;
      mbar_expect_tx(mb, (unsigned)(kBulkMaps * SW * 4));
      const unsigned dst = ring_s + (unsigned)(sl * SLOT) * 4u;
      tma_row_g2s(dst, g.tm_x1, xs, row, g.b, mb);
      tma_row_g2s(dst + 4u * (CH * SW), g.tm_s1, xs, row, g.b, mb);
      tma_row_g2s(dst + 4u * (DMAP * SW), g.tm_d1, xs, row, g.b, mb);
    }
    inflight |= 1u << sl;
  };
#else
  unsigned soff[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = lane + 32 * k, m = i / CPR, ch = i - m * CPR;
    soff[k] = (unsigned)m * iplane + 4u * (unsigned)(NARROW ? min(ch, (W >> 2) - 1) : ch);
  }
  const int dchunk = NARROW ? 4 * min(lane, (W >> 2) - 1) : 0;   // NARROW: this lane's chunk of a one-map row
  auto stage_row = [&](const int row, const int xs) {
    const unsigned dst = ring_s + (unsigned)((row & (kStageRows - 1)) * SLOT) * 4u + 16u * (unsigned)lane;
    const unsigned row_off = (unsigned)(row * W + xs);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < 2 || lane + 64 < CH * CPR) {
        const float* sx;
        asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(sx) : "r"(row_off + soff[k]), "l"(X1));
        cp_async16(dst + 512u * k, sx);
        if (!SB) {
          const float* sz;
          asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(sz) : "r"(row_off + soff[k]), "l"(S1));
          cp_async16(dst + 512u * k + 4u * (CH * SW), sz);
        }
      }
    }
    if (lane < CPR) {
      const int dc = NARROW ? dchunk : 4 * lane;
      cp_async16(dst + 4u * (DMAP * SW), D1 + row_off + dc);
      if (SB) cp_async16(dst + 4u * (CH * SW), S1 + row_off + dc);           // the one sigma map, in map slot CH
    }
    if (lane == 0) slot_xs[row & (kStageRows - 1)] = xs;
    stage_commit();
  };
#endif

  // channel PAIRS travel together as float2 so the per-channel arithmetic issues as packed FFMA2 / FMUL2 / FADD2
  constexpr int NP = CH / 2;
  constexpr int NSP = SB ? 1 : NP;        // sigma windows: one map (its .x half) or one per channel pair
  float2 ft[NP], fm[NP], st[NSP], sm[NSP];
  auto load_pair = [&](const float* base, const unsigned idx, const int p) {
    float2 r;
    if (FIXED) {
      const float* q = base + idx;
      r.x = DPFT_LDW(q + (2 * p) * PLANE);
      r.y = DPFT_LDW(q + (2 * p + 1) * PLANE);
    } else {
      r.x = ldf(base, idx + (2 * p) * iplane);
      r.y = ldf(base, idx + (2 * p + 1) * iplane);
    }
    return r;
  };
  {
    const unsigned ot = (unsigned)(max(y0 - 1, 0) * W + xc), om = (unsigned)(min(y0, H - 1) * W + xc);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      ft[p] = load_pair(X0, ot, p);
      fm[p] = load_pair(X0, om, p);
    }
    if (SB) {
      st[0] = make_float2(DPFT_LDW(S0 + ot), 0.f);
      sm[0] = make_float2(DPFT_LDW(S0 + om), 0.f);
    } else {
#pragma unroll
      for (int p = 0; p < NSP; ++p) {
        st[p] = load_pair(S0, ot, p);
        sm[p] = load_pair(S0, om, p);
      }
    }
  }
  // 32 output columns: lane l forms the vertical Sobel sums of map l & 15 (x0 channels, then sigma0 channels) in the
  // halo column left (l < 16) or right (l >= 16) of the segment, one row AHEAD of the row loop (the hand-over through
  // shared memory then needs no synchronisation of its own).  Clamped columns reproduce the replicate padding: at
  // the image border the halo column IS the lane's own column and the correction below vanishes exactly.
  float* hbuf = outl + kOutFloats;
  const float* hsrc = nullptr;      // halo column of this lane's map, row 0
  float ht = 0.f, hm = 0.f;
  float2 msgn2 = make_float2(0.f, 0.f), mabs2 = make_float2(0.f, 0.f);
  auto halo_store = [&](const int row, const float hb) {
    // same expressions as unit_sobel's vs / vd, so a halo sum equals what the neighbouring segment's lane forms
    const int hmap = lane & 15;
    float* q = hbuf + (row & 1) * 64 + (lane >> 4) * 32 + 4 * (hmap >> 1) + (hmap & 1);
    q[0] = fmaf(hm, 2.f, ht + hb);
    q[2] = fmaf(ht, -1.f, hb);
  };
  if (WIDE) {
    const int hmap = lane & 15;
    const int hcol = (lane >> 4) ? min(seg * 32 + 32, W - 1) : max(seg * 32 - 1, 0);
    hsrc = (hmap < CH ? X0 + (unsigned)hmap * iplane : S0 + (SB ? 0u : (unsigned)(hmap - CH) * iplane)) + hcol;
    ht = __ldg(hsrc + max(y0 - 1, 0) * W);
    hm = __ldg(hsrc + min(y0, H - 1) * W);
    const float hb = __ldg(hsrc + min(y0 + 1, H - 1) * W);
    halo_store(y0, hb);
    ht = hm;
    hm = hb;
    const float sg = lane == 0 ? 1.f : lane == 31 ? -1.f : 0.f;
    msgn2 = make_float2(sg, sg);
    mabs2 = make_float2(fabsf(sg), fabsf(sg));
    __syncwarp();
  }
  // the ring holds source rows max(base, top - 3) .. top; base = first row requested since the last restart
  int top = -0x40000000, base = 0x40000000;
  float d0_next = (y0 < y1) ? DPFT_LDW(D0 + (unsigned)(y0 * W + xc)) : 0.f;

#ifdef DPFT_HOIST_PX
  // the lane's column is fixed for the whole tile: its three products with the first column of R are, too
  const float rpx0 = xmul(spose[0], px), rpx1 = xmul(spose[3], px), rpx2 = xmul(spose[6], px);
#endif
  for (int y = y0; y < y1; ++y) {
#ifdef DPFT_DEBUG_STAMPS
    const long long ck0 = clock64();
    long long ck1 = ck0, ck2 = ck0;
#endif
    const unsigned ob = (unsigned)(min(y + 1, H - 1) * W + xc);
    float2 fb[NP], sb[NSP];
#pragma unroll
    for (int p = 0; p < NP; ++p) fb[p] = load_pair(X0, ob, p);
    if (SB) {
      sb[0] = make_float2(DPFT_LDW(S0 + ob), 0.f);
    } else {
#pragma unroll
      for (int p = 0; p < NSP; ++p) sb[p] = load_pair(S0, ob, p);
    }
    float hnext = 0.f;
    if (WIDE) hnext = __ldg(hsrc + min(y + 2, H - 1) * W);
    const unsigned o = (unsigned)(y * W + xc);
    const float d0 = d0_next;
    if (y + 1 < y1) d0_next = DPFT_LDW(D0 + o + Wu);
    const float py = div_by(xsub((float)y, cy), fy, rcp_fy);

    float u, v, inv_z;
    {
      const float4 ra = lds_v4(spose), rb = lds_v4(spose + 4), rc = lds_v4(spose + 8);
#ifdef DPFT_HOIST_PX
      const float wx = xadd(xadd(xadd(rpx0, xmul(ra.y, py)), ra.z), xmul(rc.y, d0));
      const float wy = xadd(xadd(xadd(rpx1, xmul(rb.x, py)), rb.y), xmul(rc.z, d0));
      const float wz = xadd(xadd(xadd(rpx2, xmul(rb.w, py)), rc.x), xmul(rc.w, d0));
#else
      const float wx = xadd(xadd(xadd(xmul(ra.x, px), xmul(ra.y, py)), ra.z), xmul(rc.y, d0));
      const float wy = xadd(xadd(xadd(xmul(ra.w, px), xmul(rb.x, py)), rb.y), xmul(rc.z, d0));
      const float wz = xadd(xadd(xadd(xmul(rb.z, px), xmul(rb.w, py)), rc.x), xmul(rc.w, d0));
#endif
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
    const int xlo = NARROW ? 0 : __reduce_min_sync(0xffffffffu, col_out ? txy.xi : 0x3fffffff);
    int ready_top = top;
    if (yhi >= 0) {
      if (ylo > top + 1 || ylo < base - 2) {            // a jump in the warp field: restart the ring at ylo
#ifdef DPFT_DEBUG_STAMPS
        S.nrestart += 1;
#endif
        top = ylo - 1;
        base = ylo;
      }
      const int keep = max(max(base, top - (kStageRows - 1)), ylo);      // lowest row this warp row still needs
      const int target = min(min(yhi + kStageLookahead, keep + kStageRows - 1), H - 1);
      const int xs = NARROW ? 0 : min(max((xlo - (SW - 32) / 2) & ~3, 0), W - SW);
      __syncwarp();                                                      // every lane is done with the old slots
      if (WIDE) {                                                        // ... and with the halo sums of row y - 1
        halo_store(y + 1, hnext);
        ht = hm;
        hm = hnext;
      }
#pragma unroll 1
      while (top < target) {
        stage_row(++top, xs);
#ifdef DPFT_DEBUG_STAMPS
        S.nstaged += 1;
#endif
      }
      ready_top = min(yhi, top);
#ifdef DPFT_DEBUG_STAMPS
      ck1 = clock64();
#endif
#if DPFT_STAGED_TMA
      {
        // rows keep .. ready_top are read by this warp row: confirm the ones still in flight (usually the newest)
#pragma unroll 1
        for (int r = max(keep, top - (kStageRows - 1)); r <= ready_top; ++r) {
          const int sl = r & (kStageRows - 1);
          if ((inflight >> sl) & 1u) wait_slot(sl);
        }
      }
#else
      stage_wait(top - ready_top);
#endif
      __syncwarp();
#ifdef DPFT_DEBUG_STAMPS
      ck2 = clock64();
#endif
    }
    // The sixteen window loads issued at the top of the row land in scratch registers, and ptxas copies them into
    // the register PAIRS the packed arithmetic wants as early as it can -- a hundred instructions after the loads,
    // in the serial part of the row, where the copies then wait a DRAM round trip (17 % of the stall samples in the
    // first r1h capture).  Adding a zero that is only known after the ring logic turns each copy into an FADD that
    // cannot be scheduled before this point, by which time the loads have landed.
    const float late0 = (top == 0x7ffffff0) ? 1.f : 0.f;       // always 0.f (top is a row index), opaque to the compiler
#ifndef DPFT_NO_LATE0
#pragma unroll
    for (int p = 0; p < NP; ++p) { fb[p].x += late0; fb[p].y += late0; }
#pragma unroll
    for (int p = 0; p < NSP; ++p) { sb[p].x += late0; if (!SB) sb[p].y += late0; }
#endif
    const int lowest = max(base, top - (kStageRows - 1));
    const int s0i = txy.yi & (kStageRows - 1), s1i = (txy.yi + 1) & (kStageRows - 1);
    const int xs0 = slot_xs[s0i], xs1 = slot_xs[s1i];
    const bool resident = txy.yi >= lowest && txy.yi + 1 <= ready_top && txy.xi >= xs0 && txy.xi + 1 < xs0 + SW &&
                          txy.xi >= xs1 && txy.xi + 1 < xs1 + SW;
    const bool any_direct = __any_sync(0xffffffffu, col_out && !resident);
#ifdef DPFT_DEBUG_STAMPS
    S.ndirect += any_direct ? 1 : 0;
    S.nlanes += __popc(__ballot_sync(0xffffffffu, col_out && !resident));
#endif
    // north-west / south-west texel of map 0 in the ring (kept in range for lanes that are not resident)
    const float* a0 = ring + s0i * SLOT + min(max(txy.xi - xs0, 0), SW - 2);
    const float* a1 = ring + s1i * SLOT + min(max(txy.xi - xs1, 0), SW - 2);

    // The rest of the row exists twice: the copy that runs when every lane's footprint is resident has no
    // lane-divergent code at all (one basic block from the first tap to the accumulation, which is what lets the
    // scheduler interleave the channel groups), the other one adds the per-lane direct loads.
#ifdef DPFT_STAGED_ONE_BODY
    constexpr int kBodies = 1;
#else
    constexpr int kBodies = 2;
#endif
#pragma unroll
    for (int body = 0; body < kBodies; ++body) {
    if (kBodies == 2 && (body == 1) != any_direct) continue;
    const bool DIRECT = kBodies == 1 ? any_direct : (body == 1);
    // A lane whose footprint is not resident (a pixel whose depth differs from its neighbours': a hole, an object
    // edge) reads its 68 taps from global memory.  Read as they are needed they cost three dependent round trips
    // (depth, channels 0-3, channels 4-7), and the warps that meet many such pixels were the stragglers every
    // launch waited for.  So the first kOutLanes such lanes of a row copy all their taps into a small scratch area
    // with 4-byte cp.async -- no registers, everything in flight at once, one round trip -- and then read them
    // like everybody else, through a per-lane base and map stride; only the overflow takes the dependent loads.
    // Lanes outside the tile's columns never need real values (their sums are discarded): they read the ring
    // wherever their clamped offsets point.
    const float *b0 = a0, *b1 = a1;     // north / south texel row of map 0 for this lane
    int ms = SW;                        // floats between maps
    bool slow = false;                  // overflow: dependent direct loads
    if (DIRECT) {
      const bool need = col_out && !resident;
      const int rank = __popc(__ballot_sync(0xffffffffu, need) & ((1u << lane) - 1u));
      const bool parked = need && rank < kOutLanes;
      slow = need && !parked;
      if (parked) {
        float* sc = outl + rank * (kStageMaps * 4);
        const unsigned sc_s = (unsigned)__cvta_generic_to_shared(sc);
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          const float* src = (m < CH) ? X1 + (unsigned)tap.o + (unsigned)m * iplane
                           : (m < DMAP) ? S1 + (unsigned)tap.o + (SB ? 0u : (unsigned)(m - CH) * iplane)
                                        : D1 + tap.o;
          cp_async4(sc_s + 16u * m, src);
          cp_async4(sc_s + 16u * m + 4u, src + 1);
          cp_async4(sc_s + 16u * m + 8u, src + W);
          cp_async4(sc_s + 16u * m + 12u, src + W + 1);
        }
        b0 = sc;
        b1 = sc + 2;
        ms = 4;
      }
      stage_commit();
      stage_wait(0);
      __syncwarp();
    }
    float d1w;
    {
      float da = b0[DMAP * ms], db = b0[DMAP * ms + 1];
      float dc = b1[DMAP * ms], dd = b1[DMAP * ms + 1];
      if (DIRECT && slow) {
        const float* q = D1 + tap.o;
        da = __ldg(q); db = __ldg(q + 1); dc = __ldg(q + W); dd = __ldg(q + W + 1);
      }
      d1w = blend_exact(da, db, dc, dd, tap);
    }
    bool occ = occluded(u, v, inv_z, d1w, H, W);
    if (AUX && g.m0) occ = occ || (__ldg(g.m0 + o) == 0);
    if (AUX && g.m1) occ = occ || !(sample_mask(g.m1, tap, W) > 0.f);
    if (TRU) occ = occ || (sm[0].x == g.s0lo) || (sm[0].x == g.s0hi);

    const float2 zero2 = make_float2(0.f, 0.f), two2 = make_float2(2.f, 2.f), neg1 = make_float2(-1.f, -1.f);
    const float2 rt_zero2 = make_float2(late0, late0);        // (0, 0), opaque to the compiler
    const float2 wa2 = make_float2(tap.wa, tap.wa), wb2 = make_float2(tap.wb, tap.wb);
    const float2 wc2 = make_float2(tap.wc, tap.wc), wd2 = make_float2(tap.wd, tap.wd);
    float2 saa2 = zero2, sab2 = zero2, sbb2 = zero2, sar2 = zero2, sbr2 = zero2, sca2 = zero2, scb2 = zero2;
    float pmin = CUDART_INF_F, pmax = -CUDART_INF_F, sr0 = 0.f;
    auto shfl2 = [](const float2 v, const bool down) {
      float2 r;
      r.x = down ? __shfl_down_sync(0xffffffffu, v.x, 1) : __shfl_up_sync(0xffffffffu, v.x, 1);
      r.y = down ? __shfl_down_sync(0xffffffffu, v.y, 1) : __shfl_up_sync(0xffffffffu, v.y, 1);
      return r;
    };
    // unit Sobel gradient of a map pair (algorithms.py:1844-1865), separable: Sx = vs(x+1) - vs(x-1),
    // Sy = vd(x-1) + 2 vd(x) + vd(x+1), vs = t + 2 m + b, vd = b - t
    const float* hrow = hbuf + (y & 1) * 64 + (lane >> 4) * 32;
    auto unit_sobel = [&](const float2 t, const float2 m, const float2 b, float2& gx, float2& gy, const int hpair) {
      const float2 vs = __ffma2_rn(m, two2, __fadd2_rn(t, b)), vd = __ffma2_rn(t, neg1, b);
      float2 Sx = __ffma2_rn(shfl2(vs, false), neg1, shfl2(vs, true));
      float2 Sy = __ffma2_rn(vd, two2, __fadd2_rn(shfl2(vd, false), shfl2(vd, true)));
      if (WIDE) {
        // lanes 0 / 31 received their OWN sums from the shuffle that has no neighbour: swap them for the halo's
        const float4 hv = *reinterpret_cast<const float4*>(hrow + 4 * hpair);
        Sx = __ffma2_rn(__ffma2_rn(make_float2(hv.x, hv.y), neg1, vs), msgn2, Sx);
        Sy = __ffma2_rn(__ffma2_rn(vd, neg1, make_float2(hv.z, hv.w)), mabs2, Sy);
      }
      const float2 n = __ffma2_rn(Sx, Sx, __ffma2_rn(Sy, Sy, make_float2(1e-8f, 1e-8f)));
      const float2 inv = make_float2(rsqrt_fast(n.x), rsqrt_fast(n.y));
      gx = __fmul2_rn(Sx, inv);
      gy = __fmul2_rn(Sy, inv);
    };
    // one sigma map per frame: its gradient, warped value and 1 / sigma once per pixel, shared by all channels
    float2 sb_gsx = zero2, sb_gsy = zero2, sb_rs = zero2, sb_k = zero2;
    if (SB) {
      float2 gx1, gy1;
      unit_sobel(st[0], sm[0], sb[0], gx1, gy1, CH / 2);
      float za = b0[CH * ms], zb = b0[CH * ms + 1], zc = b1[CH * ms], zd = b1[CH * ms + 1];
      if (DIRECT && slow) {
        const float* q = S1 + tap.o;
        za = __ldg(q); zb = __ldg(q + 1); zc = __ldg(q + W); zd = __ldg(q + W + 1);
      }
      const float sr = TRU ? blend_exact(za, zb, zc, zd, tap) : blend_fast(za, zb, zc, zd, tap);
      const float s0v = sm[0].x;
      const float rs = rsqrt_fast(fmaf(sr, sr, s0v * s0v));
      const float k3 = s0v * (rs * rs);
      sb_gsx = make_float2(gx1.x, gx1.x); sb_gsy = make_float2(gy1.x, gy1.x);
      sb_rs = make_float2(rs, rs); sb_k = make_float2(k3, k3);
      pmin = pmax = sr0 = sr;
    }
#ifndef DPFT_GATHER_PAIRS_SB
#define DPFT_GATHER_PAIRS_SB 4
#endif
    constexpr int GP = SB ? DPFT_GATHER_PAIRS_SB : DPFT_GATHER_GROUP / 2;      // channel pairs whose lookups are in flight together
#pragma unroll
    for (int p0 = 0; p0 < NP; p0 += GP) {
      float2 xa[GP], xb[GP], xc_[GP], xd[GP], za[GP], zb[GP], zc[GP], zd[GP];
#pragma unroll
      for (int j = 0; j < GP; ++j) {
        const int kx = 2 * (p0 + j) * ms, kz = (CH + 2 * (p0 + j)) * ms;
        xa[j] = make_float2(b0[kx], b0[kx + ms]); xb[j] = make_float2(b0[kx + 1], b0[kx + ms + 1]);
        xc_[j] = make_float2(b1[kx], b1[kx + ms]); xd[j] = make_float2(b1[kx + 1], b1[kx + ms + 1]);
        if (!SB) {
          za[j] = make_float2(b0[kz], b0[kz + ms]); zb[j] = make_float2(b0[kz + 1], b0[kz + ms + 1]);
          zc[j] = make_float2(b1[kz], b1[kz + ms]); zd[j] = make_float2(b1[kz + 1], b1[kz + ms + 1]);
        }
      }
      if (DIRECT && slow) {
#pragma unroll
        for (int j = 0; j < GP; ++j) {
          const unsigned ia = (unsigned)tap.o + (unsigned)(2 * (p0 + j)) * iplane, ic = ia + Wu;
          ldf2(X1, ia, xa[j].x, xb[j].x); ldf2(X1, ic, xc_[j].x, xd[j].x);
          ldf2(X1, ia + iplane, xa[j].y, xb[j].y); ldf2(X1, ic + iplane, xc_[j].y, xd[j].y);
          if (!SB) {
            ldf2(S1, ia, za[j].x, zb[j].x); ldf2(S1, ic, zc[j].x, zd[j].x);
            ldf2(S1, ia + iplane, za[j].y, zb[j].y); ldf2(S1, ic + iplane, zc[j].y, zd[j].y);
          }
        }
      }
      float2 gfx[GP], gfy[GP], gsx[GP], gsy[GP];
#pragma unroll
      for (int j = 0; j < GP; ++j) {
        unit_sobel(ft[p0 + j], fm[p0 + j], fb[p0 + j], gfx[j], gfy[j], p0 + j);
        if (SB) { gsx[j] = sb_gsx; gsy[j] = sb_gsy; }
        else unit_sobel(st[SB ? 0 : p0 + j], sm[SB ? 0 : p0 + j], sb[SB ? 0 : p0 + j], gsx[j], gsy[j], CH / 2 + p0 + j);
      }
#pragma unroll
      for (int j = 0; j < GP; ++j) {
        const int p = p0 + j;
        const float2 fr = __ffma2_rn(xd[j], wd2, __ffma2_rn(xc_[j], wc2, __ffma2_rn(xb[j], wb2, __fmul2_rn(xa[j], wa2))));
        // residual, its uncertainty and the 2-vector d(wres)/d(u,v) (algorithms.py:1969-1972, :872)
        const float2 res = __ffma2_rn(fm[p], neg1, fr);
        float2 sr = zero2, rs, k3;
        if (SB) {
          rs = sb_rs;
          k3 = sb_k;
        } else {
          // sigma is compared for equality against its batch extremes -> every product and sum rounded on its own,
          // in blend_exact's order.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although both carry .rn
          // (it honours the scalar forms); a saturated map then loses its one-ulp dips and the WRONG pixels tie with
          // the extreme (tests/test_uic_queue_gpu.py: saturated sigma).  So the products are formed as x * w + 0 with
          // a zero only known at run time -- an FFMA2 that is exactly the rounded product (sigma and the weights are
          // non-negative, so no -0 can turn into +0) and that ptxas cannot merge into the sums.
          sr = TRU ? __fadd2_rn(__fadd2_rn(__fadd2_rn(__ffma2_rn(za[j], wa2, rt_zero2), __ffma2_rn(zb[j], wb2, rt_zero2)),
                                           __ffma2_rn(zc[j], wc2, rt_zero2)), __ffma2_rn(zd[j], wd2, rt_zero2))
                   : __ffma2_rn(zd[j], wd2, __ffma2_rn(zc[j], wc2, __ffma2_rn(zb[j], wb2, __fmul2_rn(za[j], wa2))));
          const float2 s0v = sm[SB ? 0 : p];
          const float2 ss = __ffma2_rn(sr, sr, __fmul2_rn(s0v, s0v));
          rs = make_float2(rsqrt_fast(ss.x), rsqrt_fast(ss.y));                  // 1 / sigma
          k3 = __fmul2_rn(s0v, __fmul2_rn(rs, rs));                              // sigma0 / sigma^2
        }
        const float2 wres = __fmul2_rn(res, rs);
        const float2 q = __fmul2_rn(wres, k3);                                   // res * sigma0 / sigma^3
        const float2 a = __ffma2_rn(gfx[j], rs, __fmul2_rn(q, gsx[j]));
        const float2 bq = __ffma2_rn(gfy[j], rs, __fmul2_rn(q, gsy[j]));
        const float2 wm = occ ? make_float2(1e-6f, 1e-6f) : wres;
        saa2 = __ffma2_rn(a, a, saa2);
        sab2 = __ffma2_rn(a, bq, sab2);
        sbb2 = __ffma2_rn(bq, bq, sbb2);
        sar2 = __ffma2_rn(a, wm, sar2);
        sbr2 = __ffma2_rn(bq, wm, sbr2);
        if (TRU) {
          const float2 dw = __fadd2_rn(wres, make_float2(-1e-6f, -1e-6f));
          sca2 = __ffma2_rn(a, dw, sca2);
          scb2 = __ffma2_rn(bq, dw, scb2);
          if (!SB) {
            pmin = fminf(pmin, fminf(sr.x, sr.y));
            pmax = fmaxf(pmax, fmaxf(sr.x, sr.y));
            if (p == 0) sr0 = sr.x;
          }
        }
      }
    }
    float saa = saa2.x + saa2.y, sab = sab2.x + sab2.y, sbb = sbb2.x + sbb2.y, sar = sar2.x + sar2.y, sbr = sbr2.x + sbr2.y;
    const float sca = sca2.x + sca2.y, scb = scb2.x + scb2.y;
    if (!col_out) { saa = sab = sbb = sar = sbr = 0.f; }
    float ju[6], jv[6];
    warp_rows(px, py, d0, fx, fy, ju, jv);
    accumulate_system(S.acc, ju, jv, saa, sab, sbb, sar, sbr);
    if (TRU) {
      // running extremes of the warped sigma over the WARP's pixels and what the pixels sitting on them added to
      // J^T r.  A new extreme of the warp is rare after the first rows (~ log of the pixels seen), so the
      // bookkeeping sits behind one warp-uniform branch; ties (saturated sigma maps) take it every time.
      // one vote decides the common case (three rows of four): no pixel of this row reaches the warp's running extremes,
      // nothing below would change anything (-0.6 % / -2.9 % on the one-map / C-map level-0 launch)
      if (__any_sync(0xffffffffu, col_out && (pmin <= S.vmin || pmax >= S.vmax))) {
      const float wmin = ord2f(__reduce_min_sync(0xffffffffu, f2ord(col_out ? pmin : CUDART_INF_F)));
      const float wmax = ord2f(__reduce_max_sync(0xffffffffu, f2ord(col_out ? pmax : -CUDART_INF_F)));
      const bool lo = wmin < S.vmin, hi = wmax > S.vmax;
      const float nmin = lo ? wmin : S.vmin, nmax = hi ? wmax : S.vmax;
      const bool tmin = col_out && !occ && (sr0 == nmin), tmax = col_out && !occ && (sr0 == nmax);
      const bool any_min = lo || __any_sync(0xffffffffu, tmin), any_max = hi || __any_sync(0xffffffffu, tmax);
      if (any_min || any_max) {
#ifdef DPFT_DEBUG_STAMPS
        S.ntru += 1;
#endif
        S.vmin = nmin;
        S.vmax = nmax;
        float cc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          cc[i] = 0.f;
          if (i != 4) cc[i] = fmaf(sca, ju[i], cc[i]);
          if (i != 3) cc[i] = fmaf(scb, jv[i], cc[i]);
        }
        // every lane rewrites its column (no per-lane branches): a new extreme restarts the sums
        if (any_min) {
#pragma unroll
          for (int i = 0; i < 6; ++i) scorr[i][lane] = (lo ? 0.f : scorr[i][lane]) + (tmin ? cc[i] : 0.f);
        }
        if (any_max) {
#pragma unroll
          for (int i = 0; i < 6; ++i) scorr[6 + i][lane] = (hi ? 0.f : scorr[6 + i][lane]) + (tmax ? cc[i] : 0.f);
        }
      }
      }
    }
    if (AUX && g.occ_out && col_out) {
      g.occ_out[(size_t)y * W + x] = occ ? 1 : 0;
      if (TRU) g.sr0_dbg[(size_t)y * W + x] = sr0;
    }
    }   // body
#ifdef DPFT_DEBUG_STAMPS
    {
      const long long ck3 = clock64();
      S.c_front += ck1 - ck0;
      S.c_wait += ck2 - ck1;
      S.c_body += ck3 - ck2;
    }
#endif
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      ft[p] = fm[p];
      fm[p] = fb[p];
    }
#pragma unroll
    for (int p = 0; p < NSP; ++p) {
      st[p] = sm[p];
      sm[p] = sb[p];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // nothing of this warp may still land in the ring
#if DPFT_STAGED_TMA
#pragma unroll
  for (int sl = 0; sl < kStageRows; ++sl)
    if ((inflight >> sl) & 1u) wait_slot(sl);
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int sl = 0; sl < kStageRows; ++sl) mbar_inval(mbar0 + 8u * sl);
  }
  __syncwarp();
#endif
}

}  // namespace dpft
