// U_IC forward: one Gauss-Newton iteration per launch, every launch fused end to end.
//
//   uic_iter_kernel    warp + bilinear lookup + validity mask + residual + Jacobian + J^T J / J^T r
//                      partial sums, then (last CTA of a pair) the per-pair reduction and (last CTA
//                      of the grid, or of the pair) the damped 6x6 solve and the pose update.
//
// Restates reference code/models/algorithms.py:611-723 (TrustRegionInverseWUncertainty.forward) with
// the arithmetic of oracle/ic_oracle.py; see DESIGN.md for the data flow and the byte accounting.
//
// Thread mapping: a warp owns a tile of 30 output columns x TR rows of one frame pair.  Lanes 0 and
// 31 are halo columns (clamped at the image border = replicate padding), so the horizontal Sobel
// taps come from __shfl of the neighbouring lanes and the vertical taps from a 3-row register window
// that slides down the tile: every x0 / sigma0 element is loaded once per tile (coalesced 128 B rows),
// the unit Sobel gradients are recomputed instead of stored, and nothing per-pixel is written back.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"
#include "dpft_kernels.h"
#include "dpft_records.h"
#include "uic_tile.cuh"
#include "uic_tile_staged.cuh"

#ifndef DPFT_MIN_CTAS
#define DPFT_MIN_CTAS 4   // 128-thread CTAs per SM the register allocation must allow
#endif

namespace dpft {

#ifndef DPFT_WARPS
#define DPFT_WARPS 4
#endif
constexpr int kWarps = DPFT_WARPS;   // warps per CTA of the iteration kernels
constexpr int kThreads = kWarps * 32;
constexpr int kCols = kTileCols;   // output columns per warp tile
constexpr int kMaxTileRows = 40;

// Warps per CTA of the staged kernel (a tuning hook: 2-warp CTAs make the two kinds of pairs of the balanced tiling
// differ by 1/13 of their work instead of 1/6, but measured slower).
#ifndef DPFT_STAGED_WARPS
#define DPFT_STAGED_WARPS 4
#endif
constexpr int kSW = DPFT_STAGED_WARPS;   // warps per CTA of the staged kernel
constexpr int kSThreads = kSW * 32;

// Balanced tiling of the staged kernel.  With one rectangular tiling for every pair the CTA count is a multiple of
// B and rarely matches the resident CTA slots, and tiles of whole rows per segment rarely divide evenly among the
// warps.  So (1) the pairs come in two kinds: `n_more` of them (spread evenly over the batch) get one CTA more than
// the others, and (2) the warp-rows of a pair (segment-major: segment 0 top to bottom, then segment 1, ...) are
// cut into one contiguous range per warp, every range the same length to within a row.  A range that crosses a
// segment boundary is walked as two sub-tiles (the second one re-primes its windows and its ring).
constexpr int kTabWarps = 40;
struct TileTab {
  int on;                          // 0: rectangular tiling (nseg x nrt tiles of TR rows)
  int n_more;                      // pairs of kind 1
  int ctas[2];                     // CTAs per pair of kind 0 / 1
  unsigned char seg[2][kTabWarps][2];
  short y0[2][kTabWarps][2], y1[2][kTabWarps][2];
};

struct UicIterParams {
  const float *x0, *x1, *s0, *s1, *d0, *d1, *K;
  const uint8_t *m0, *m1;
  uint8_t* occ_out;      // (B,H,W) of this iteration or nullptr
  float* sr0_dbg;        // (B,H,W) warped sigma channel 0 (only with occ_out + TRU)
  int H, W, B, C;
  int nseg, nrt, TR, ctas_per_pair;
  const float* pose;     // (B,12) pose this iteration linearises at
  float* pose_next;      // (B,12)
  float* sys_out;        // (B,27)
  float* partials;       // (B, ctas_per_pair, PS)
  double* pairrec;       // (B, PS)
  int* counters;         // [B] per pair, then [n_groups] pairs done per sigma-extreme group
  int SC;                // channels of sigma0 / sigma1 the kernel works with: C, or 1 (one map per frame)
  int SCm;               // channels of sigma0 / sigma1 in MEMORY: SC, or C when SC == 1 reads channel 0 of full tensors
  const int* mism;       // device flag of sigma_replication_kernel: 0 = every channel of sigma0 / sigma1 equals channel 0
  int rep_role;          // 0: no twin launches; 1: this launch works only if *mism == 0; 2: only if *mism != 0
  const uint32_t* s0mm;  // order-encoded min, max of sigma0 over the level tensor of every group: (n_mm_groups, 2)
  float* gmm;            // (n_groups, 4) group-wide min/max of the warped sigma of this iteration, min/max of sigma0
  int32_t* status;
  uint32_t flags;
  int kf_shared;         // keyframe-side tensors (x0, sigma0, invd0, obj_mask0) have batch size 1
  int pairwise;          // sigma extremes per pair instead of per batch (each pair its own batch of one: group == 1)
  int group, n_mm_groups; // pairs per sigma-extreme group (B: the whole call is one batch); groups of the sigma0 extremes
  const float* icp_rec;  // (B,28) sums of the point-to-plane term of this iteration, or nullptr
  float icp_w2;          // its weight squared (w_icp scales both J and r)
  TileTab tab;           // staged kernel only
#if DPFT_STAGED_TMA
  CUtensorMap tm_x1, tm_s1, tm_d1;   // (W, H, C | SC | 1, B) maps of x1 / sigma1 / invd1, box = stage width x 1 x channels x 1
#endif
};

// Final step for one pair: corrections for the batch-global sigma extremes, damping, solve, update.
template <bool TRU>
__device__ void finalize_pair(const UicIterParams& p, int b, float gmin, float gmax) {
  const double* rec = p.pairrec + (size_t)b * PS;
  double A[21], rhs[6];
#pragma unroll
  for (int i = 0; i < 21; ++i) A[i] = __ldcg(rec + i);
#pragma unroll
  for (int i = 0; i < 6; ++i) rhs[i] = __ldcg(rec + 21 + i);
  if (TRU) {
    // pixels whose warped sigma (channel 0) equals the batch-global min or max are masked
    // (algorithms.py:1976-1979): their weighted residual becomes 1e-6, i.e. subtract what they added.
    const bool at_min = ((float)__ldcg(rec + E_VMIN) == gmin);
    const bool at_max = ((float)__ldcg(rec + E_VMAX) == gmax) && (gmax != gmin);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (at_min) rhs[i] -= __ldcg(rec + E_CMIN + i);
      if (at_max) rhs[i] -= __ldcg(rec + E_CMAX + i);
    }
  }
  if (p.icp_rec) {
    // feature-metric and point-to-plane systems are simply added (algorithms.py:685-688)
#pragma unroll
    for (int i = 0; i < 21; ++i) A[i] += (double)p.icp_w2 * (double)__ldcg(p.icp_rec + (size_t)b * 28 + i);
#pragma unroll
    for (int i = 0; i < 6; ++i) rhs[i] += (double)p.icp_w2 * (double)__ldcg(p.icp_rec + (size_t)b * 28 + 21 + i);
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < 21; ++i) finite = finite && isfinite(A[i]);
#pragma unroll
  for (int i = 0; i < 6; ++i) finite = finite && isfinite(rhs[i]);
  float* sys = p.sys_out + (size_t)b * 27;
#pragma unroll
  for (int i = 0; i < 21; ++i) sys[i] = (float)A[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) sys[21 + i] = (float)rhs[i];
  double xi[6];
#ifdef DPFT_EXPERIMENT_SKIP_SOLVE
  bool ok = true;
  for (int i = 0; i < 12; ++i) p.pose_next[(size_t)b * 12 + i] = p.pose[(size_t)b * 12 + i];
#else
  const bool ok = solve_and_update(A, rhs, true, p.pose + (size_t)b * 12, p.pose_next + (size_t)b * 12, xi);
#endif
  int st = 0;
  if (!finite) st |= DPFT_ST_NONFINITE;
  if (!ok) st |= DPFT_ST_SINGULAR;
  if (st) atomicOr(p.status, st);
}

// Everything after the per-thread accumulation, shared by both iteration kernels: warp -> CTA sums
// (fp64 once lanes are combined), one partial record per CTA, the last CTA of a pair folds its records in
// a fixed order (deterministic), and the last CTA of the grid (or of the pair, when nothing couples the
// pairs) damps, solves and updates the poses.
// Optional in-kernel timeline (compile with -DDPFT_DEBUG_STAMPS): %globaltimer at a few points of the LAST launch,
// read back with dpft_debug_read_stamps.  Not part of the ABI; absent from normal builds.
#ifdef DPFT_DEBUG_STAMPS
__device__ unsigned long long g_stamps[16];
__device__ __forceinline__ void stamp(int i) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_stamps[i] = t;
}
#define DPFT_STAMP(i, cond) do { if (cond) stamp(i); } while (0)
// per-warp timeline of the staged kernel: SM id, start of the tile walk, its end, rows walked
constexpr int kTimelineWarps = 8192;
__device__ unsigned long long g_wtl[kTimelineWarps * 4];
__device__ unsigned long long g_wtl2[kTimelineWarps * 4];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#else
#define DPFT_STAMP(i, cond) do { } while (0)
#endif

template <bool TRU, int NW = kWarps>
__device__ __forceinline__ void reduce_and_finish(const UicIterParams& p, const int b, float (*redw)[33] /* this warp's [NSUM][33] */,
                                                  const float (&acc)[27], const float vmin, const float vmax,
                                                  const int n_ctas /* CTAs of THIS pair; records are spaced by p.ctas_per_pair */) {
  __shared__ double wsum[NW][NSUM + 1];
  __shared__ float wvmin[NW], wvmax[NW];
  __shared__ float s_pair_mm[2];
  __shared__ int s_flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---------------------------------------------------------------- CTA reduction
  if (TRU) {
    const float wmn = warp_min(vmin), wmx = warp_max(vmax);
    // the corrections already sit in redw[27..38][lane]
    if (vmin != wmn) {
#pragma unroll
      for (int i = 0; i < 6; ++i) redw[27 + i][lane] = 0.f;
    }
    if (vmax != wmx) {
#pragma unroll
      for (int i = 0; i < 6; ++i) redw[33 + i][lane] = 0.f;
    }
    if (lane == 0) {
      wvmin[warp] = wmn;
      wvmax[warp] = wmx;
    }
  }
#pragma unroll
  for (int e = 0; e < 27; ++e) redw[e][lane] = acc[e];
  __syncwarp();
  constexpr int NE = TRU ? NSUM : 27;
  for (int e = lane; e < NE; e += 32) {
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) s += (double)redw[e][j];
    wsum[warp][e] = s;
  }
  __syncthreads();

  DPFT_STAMP(3, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);   // CTA reduction done
  float* part = p.partials + ((size_t)b * p.ctas_per_pair + blockIdx.x) * PS;
  float cta_min = CUDART_INF_F, cta_max = -CUDART_INF_F;
  if (TRU) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      cta_min = fminf(cta_min, wvmin[w]);
      cta_max = fmaxf(cta_max, wvmax[w]);
    }
  }
  if (threadIdx.x < NE) {
    const int e = threadIdx.x;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      bool take = true;
      if (TRU && e >= 27 && e < 33) take = (wvmin[w] == cta_min);
      if (TRU && e >= 33) take = (wvmax[w] == cta_max);
      if (take) s += wsum[w][e];
    }
    part[e < 27 ? e : e + 2] = (float)s;   // corr entries live at E_CMIN.. / E_CMAX..
  }
  if (TRU && threadIdx.x == 0) {
    part[E_VMIN] = cta_min;
    part[E_VMAX] = cta_max;
  }

  // ---------------------------------------------------------------- last CTA of the pair reduces it
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_flag = (atomicAdd(p.counters + b, 1) == n_ctas - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  DPFT_STAMP(4, threadIdx.x == 0 && b == 0);                      // last CTA of pair 0 starts the pair reduction

  const float* pp = p.partials + (size_t)b * p.ctas_per_pair * PS;
  const int n = n_ctas;
  float pair_min = CUDART_INF_F, pair_max = -CUDART_INF_F;
  if (TRU) {
    float a = CUDART_INF_F, c = -CUDART_INF_F;
    for (int i = threadIdx.x; i < n; i += (NW * 32)) {
      a = fminf(a, __ldcg(pp + (size_t)i * PS + E_VMIN));
      c = fmaxf(c, __ldcg(pp + (size_t)i * PS + E_VMAX));
    }
    a = warp_min(a);
    c = warp_max(c);
    __syncthreads();   // wvmin/wvmax reuse
    if (lane == 0) {
      wvmin[warp] = a;
      wvmax[warp] = c;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      pair_min = fminf(pair_min, wvmin[w]);
      pair_max = fmaxf(pair_max, wvmax[w]);
    }
  }
  double* rec = p.pairrec + (size_t)b * PS;
  if (threadIdx.x < NE) {
    const int e = threadIdx.x;
    const int slot = e < 27 ? e : e + 2;
    // eight records in flight at a time (the loads are L2 round trips), summed in record order
    const bool corr = TRU && e >= 27;
    const int key_at = e < 33 ? E_VMIN : E_VMAX;
    const float key_want = e < 33 ? pair_min : pair_max;
    double s = 0.0;
    for (int i0 = 0; i0 < n; i0 += 8) {
      float v[8], key[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* q = pp + (size_t)min(i0 + j, n - 1) * PS;
        v[j] = __ldcg(q + slot);
        key[j] = corr ? __ldcg(q + key_at) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i0 + j < n && (!corr || key[j] == key_want)) s += (double)v[j];
    }
    rec[slot] = s;
  }
  if (TRU && threadIdx.x == 0) {
    rec[E_VMIN] = (double)pair_min;
    rec[E_VMAX] = (double)pair_max;
  }
  if (threadIdx.x == 0) p.counters[b] = 0;   // ready for the next launch
  __threadfence();
  __syncthreads();

  if (!TRU) {
    if (threadIdx.x == 0) finalize_pair<false>(p, b, 0.f, 0.f);
    return;
  }
  if (p.pairwise) {
    // every pair is its own batch: its extremes are "the batch extremes", nothing couples the pairs
    if (threadIdx.x == 0) {
      finalize_pair<true>(p, b, pair_min, pair_max);
      const uint32_t* mm = p.s0mm + (p.n_mm_groups > 1 ? 2 * b : 0);
      float* gm = p.gmm + 4 * b;
      gm[0] = pair_min; gm[1] = pair_max; gm[2] = ord2f(__ldcg(mm)); gm[3] = ord2f(__ldcg(mm + 1));
    }
    return;
  }

  // ---------------------------------------------------------------- last CTA of the group: its extremes + its solves
  // (a group is one batch of the reference: `group` consecutive pairs; usually the whole call)
  DPFT_STAMP(5, threadIdx.x == 0 && b == 0);                      // pair 0 reduced
  const int grp = b / p.group, g_lo = grp * p.group, g_hi = g_lo + p.group;
  if (threadIdx.x == 0) s_flag = (atomicAdd(p.counters + p.B + grp, 1) == p.group - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  {
    float a = CUDART_INF_F, c = -CUDART_INF_F;
    for (int i = g_lo + threadIdx.x; i < g_hi; i += (NW * 32)) {
      a = fminf(a, (float)__ldcg(p.pairrec + (size_t)i * PS + E_VMIN));
      c = fmaxf(c, (float)__ldcg(p.pairrec + (size_t)i * PS + E_VMAX));
    }
    a = warp_min(a);
    c = warp_max(c);
    if (lane == 0) {
      wvmin[warp] = a;
      wvmax[warp] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float g0 = CUDART_INF_F, g1 = -CUDART_INF_F;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        g0 = fminf(g0, wvmin[w]);
        g1 = fmaxf(g1, wvmax[w]);
      }
      s_pair_mm[0] = g0;
      s_pair_mm[1] = g1;
      const uint32_t* mm = p.s0mm + (p.n_mm_groups > 1 ? 2 * grp : 0);
      float* gm = p.gmm + 4 * grp;
      gm[0] = g0;
      gm[1] = g1;
      gm[2] = ord2f(__ldcg(mm));
      gm[3] = ord2f(__ldcg(mm + 1));
      p.counters[p.B + grp] = 0;
    }
    __syncthreads();
  }
  DPFT_STAMP(6, threadIdx.x == 0);                                // group-last CTA: extremes known
  for (int i = g_lo + threadIdx.x; i < g_hi; i += (NW * 32)) finalize_pair<true>(p, i, s_pair_mm[0], s_pair_mm[1]);
  DPFT_STAMP(7, threadIdx.x == 0);                                // all solves written
}

//
// RES: the "resident" variant for SMALL levels (the coarse end of the pyramid).  There a level of one pair is a few
// tens of KB, so every CTA first copies the pair's whole live frame (x1, sigma1, invd1) into shared memory -- with
// cp.async, BEFORE the dependency sync, i.e. under the tail of the previous iteration's launch, because nothing a
// launch writes is read by that copy -- and the 68 footprint taps of a pixel become LDS instead of L1/L2 round trips.
constexpr int kResMinCtas = 3;     // shared memory, not registers, bounds the resident CTAs: 170 registers, no spills
__device__ __forceinline__ bool cta_has_tiles(const UicIterParams& p, unsigned cta) { return (int)(cta * kWarps) < p.nseg * p.nrt; }

template <int CH, bool TRU, int GW = 0, int GH = 0, bool RES = false>
__global__ void __launch_bounds__(kThreads, RES ? kResMinCtas : DPFT_MIN_CTAS) uic_iter_kernel(const __grid_constant__ UicIterParams p) {
  extern __shared__ __align__(16) float dyn_live[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int plane = (GW > 0) ? GW * GH : p.H * p.W;
  const int wt = blockIdx.x * kWarps + warp;
  const bool warp_on = wt < p.nseg * p.nrt;
  const int seg = warp_on ? wt % p.nseg : 0;
  const int rt = warp_on ? wt / p.nseg : 0;
  const int y0 = rt * p.TR;
  const int y1 = warp_on ? min(y0 + p.TR, p.H) : y0;

  PairView g;
  const size_t po = (size_t)b * p.C * plane;
  const size_t b0 = p.kf_shared ? 0 : (size_t)b;      // one keyframe for the whole batch (kf_vo-style tracking)
  g.x0 = p.x0 + b0 * p.C * plane; g.x1 = p.x1 + po;
  g.s0 = p.s0 + b0 * p.SCm * plane; g.s1 = p.s1 + (size_t)b * p.SCm * plane;
  g.scm = RES ? p.SC : p.SCm;           // sigma1 planes of the shared-memory copy (RES) / of the tensor in memory
  g.splane = (p.SC == p.C) ? (unsigned)plane : 0u;
  g.d0 = p.d0 + b0 * plane; g.d1 = p.d1 + (size_t)b * plane;
  g.m0 = p.m0 ? p.m0 + b0 * plane : nullptr;
  g.m1 = p.m1 ? p.m1 + (size_t)b * plane : nullptr;
  g.occ_out = p.occ_out ? p.occ_out + (size_t)b * plane : nullptr;
  g.sr0_dbg = p.sr0_dbg ? p.sr0_dbg + (size_t)b * plane : nullptr;
  g.H = p.H; g.W = p.W; g.C = p.C;
  g.fx = __ldg(p.K + 4 * b); g.fy = __ldg(p.K + 4 * b + 1); g.cx = __ldg(p.K + 4 * b + 2); g.cy = __ldg(p.K + 4 * b + 3);
  g.s0lo = g.s0hi = 0.f;

  // let the next iteration's launch become resident as soon as every CTA of this one has started; nothing
  // before the dependency sync touches what the previous launch writes
  __shared__ float red[kWarps][NSUM][33];
  __shared__ __align__(16) float s_pose[12];
  DPFT_STAMP(0, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);
  // Twin launches (one uncertainty map / C maps) for full sigma tensors: only the twin that matches what
  // sigma_replication_kernel found works.  The flag is final before any iteration launch can be resident (ordinary
  // launches -- init, sigma0 extremes -- sit between that kernel and the first programmatic launch), so the twin
  // that will not work skips its copy of the live frame as well.
  const bool idle_twin = p.rep_role && ((__ldcg(p.mism) == 0) != (p.rep_role == 1));
  if (RES && !idle_twin && cta_has_tiles(p, blockIdx.x)) {
    // x1 | sigma1 | invd1 of this pair, contiguous in that order: 16-byte copies when the planes allow it
    // (SC == 1 on full tensors: the pair's channel 0 is the one map)
    const unsigned dst0 = (unsigned)__cvta_generic_to_shared(dyn_live);
    const float* src[3] = {g.x1, g.s1, g.d1};
    const int cnt[3] = {p.C * plane, p.SC * plane, plane};
    const bool v16 = (plane % 4 == 0) && ((((uintptr_t)g.x1 | (uintptr_t)g.s1 | (uintptr_t)g.d1) & 15u) == 0);
    int off = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (v16) {
        for (int i = 4 * threadIdx.x; i < cnt[r]; i += 4 * kThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 4u * (unsigned)(off + i)), "l"(src[r] + i) : "memory");
      } else {
        for (int i = threadIdx.x; i < cnt[r]; i += kThreads)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4u * (unsigned)(off + i)), "l"(src[r] + i) : "memory");
      }
      off += cnt[r];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  DPFT_STAMP(1, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);
  if (idle_twin) return;        // (after the dependency sync: the chain of programmatic dependencies stays transitive)
  if (RES) asm volatile("cp.async.wait_group 0;" ::: "memory");
  // full sigma tensors whose channels are copies of channel 0 (found on the device): read the one map, C times less
  if (p.mism && p.rep_role == 0 && __ldcg(p.mism) == 0) g.splane = 0u;
  if (threadIdx.x < 12) s_pose[threadIdx.x] = __ldcg(p.pose + (size_t)b * 12 + threadIdx.x);
  if (TRU) {
    const uint32_t* mm = p.s0mm + (p.n_mm_groups > 1 ? 2 * (b / p.group) : 0);
    g.s0lo = ord2f(__ldcg(mm));
    g.s0hi = ord2f(__ldcg(mm + 1));
#pragma unroll
    for (int i = 0; i < 12; ++i) red[warp][27 + i][lane] = 0.f;
  }
  __syncthreads();
  TileSums S;
  S.reset();
  process_tile<CH, TRU, GW, GH, RES>(g, s_pose, &red[warp][27], seg, y0, y1, lane, S, dyn_live);
  DPFT_STAMP(2, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);
  reduce_and_finish<TRU>(p, b, red[warp], S.acc, S.vmin, S.vmax, p.ctas_per_pair);
}

// Same launch structure as uic_iter_kernel, tile walked by the staged-footprint routine (uic_tile_staged.cuh):
// every warp owns kStageAreaFloats of dynamic shared memory: its ring of source rows, then the 12 correction
// rows of remove_tru_sigma; the 27 rows of the final reduction overlay the tail of the ring once the tile is done.
#ifndef DPFT_STAGED_CTAS
#define DPFT_STAGED_CTAS (12 / DPFT_STAGED_WARPS)
#endif
template <bool TRU, bool SB = false, int GW = 0, int GH = 0, bool AUX = true>
__global__ void __launch_bounds__(kSThreads, DPFT_STAGED_CTAS) uic_iter_staged_kernel(const __grid_constant__ UicIterParams p) {
  extern __shared__ __align__(128) float dyn_stage[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int plane = (GW > 0) ? GW * GH : p.H * p.W;
  int seg[2], y0[2], y1[2], n_ctas = p.ctas_per_pair;
  if (p.tab.on) {
    // kind 1 pairs are spread evenly over the batch (Bresenham)
    const int kind = (int)(((long)(b + 1) * p.tab.n_more) / p.B - ((long)b * p.tab.n_more) / p.B);
    n_ctas = p.tab.ctas[kind];
    if ((int)blockIdx.x >= n_ctas) return;
    const int w = blockIdx.x * kSW + warp;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      seg[k] = p.tab.seg[kind][w][k];
      y0[k] = p.tab.y0[kind][w][k];
      y1[k] = p.tab.y1[kind][w][k];
    }
  } else {
    const int wt = blockIdx.x * kSW + warp;
    const bool warp_on = wt < p.nseg * p.nrt;
    seg[0] = warp_on ? wt % p.nseg : 0;
    const int rt = warp_on ? wt / p.nseg : 0;
    y0[0] = rt * p.TR;
    y1[0] = warp_on ? min(y0[0] + p.TR, p.H) : y0[0];
    seg[1] = y0[1] = y1[1] = 0;
  }

  PairView g;
  const size_t po = (size_t)b * p.C * plane;
  const size_t b0 = p.kf_shared ? 0 : (size_t)b;      // one keyframe for the whole batch (kf_vo-style tracking)
  g.x0 = p.x0 + b0 * p.C * plane; g.x1 = p.x1 + po;
  g.s0 = p.s0 + b0 * p.SCm * plane; g.s1 = p.s1 + (size_t)b * p.SCm * plane;
  g.scm = p.SCm;
  g.splane = (p.SC == p.C) ? (unsigned)plane : 0u;
  g.d0 = p.d0 + b0 * plane; g.d1 = p.d1 + (size_t)b * plane;
  g.m0 = p.m0 ? p.m0 + b0 * plane : nullptr;
  g.m1 = p.m1 ? p.m1 + (size_t)b * plane : nullptr;
  g.occ_out = p.occ_out ? p.occ_out + (size_t)b * plane : nullptr;
  g.sr0_dbg = p.sr0_dbg ? p.sr0_dbg + (size_t)b * plane : nullptr;
  g.H = p.H; g.W = p.W; g.C = p.C;
  g.fx = __ldg(p.K + 4 * b); g.fy = __ldg(p.K + 4 * b + 1); g.cx = __ldg(p.K + 4 * b + 2); g.cy = __ldg(p.K + 4 * b + 3);
  g.s0lo = g.s0hi = 0.f;
  g.b = b;
#if DPFT_STAGED_TMA
  g.tm_x1 = &p.tm_x1; g.tm_s1 = &p.tm_s1; g.tm_d1 = &p.tm_d1;
#else
  g.tm_x1 = g.tm_s1 = g.tm_d1 = nullptr;
#endif

  float* area = dyn_stage + warp * kStageAreaFloats;
  float (*redw)[33] = reinterpret_cast<float (*)[33]>(area + kStageWarpFloats - 27 * 33);   // rows 27.. start at the ring's end
  __shared__ __align__(16) float s_pose[12];
  DPFT_STAMP(0, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);   // CTA resident
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  DPFT_STAMP(1, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);   // previous launch complete
  // twin launches (one uncertainty map / C maps) for full sigma tensors: only the twin that matches what
  // sigma_replication_kernel found works; the other one has waited for its predecessor (so the chain of
  // programmatic dependencies stays transitive) and leaves
  if (p.rep_role && ((__ldcg(p.mism) == 0) != (p.rep_role == 1))) return;
  if (threadIdx.x < 12) s_pose[threadIdx.x] = __ldcg(p.pose + (size_t)b * 12 + threadIdx.x);
  if (TRU) {
    const uint32_t* mm = p.s0mm + (p.n_mm_groups > 1 ? 2 * (b / p.group) : 0);
    g.s0lo = ord2f(__ldcg(mm));
    g.s0hi = ord2f(__ldcg(mm + 1));
#pragma unroll
    for (int i = 0; i < 12; ++i) redw[27 + i][lane] = 0.f;
  }
  __syncthreads();
  TileSums S;
  S.reset();
#ifdef DPFT_DEBUG_STAMPS
  const unsigned long long wt0 = gtimer();
#endif
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    const int sg = k ? seg[1] : seg[0], ya = k ? y0[1] : y0[0], yb = k ? y1[1] : y1[0];
    if (yb > ya) process_tile_staged<TRU, SB, GW, GH, AUX>(g, s_pose, redw + 27, area, area + kStageWarpFloats + 12 * 33, sg, ya, yb, lane, S);
  }
  __syncwarp();
#ifdef DPFT_DEBUG_STAMPS
  {
    const int wi = (b * (int)gridDim.x + (int)blockIdx.x) * kSW + warp;
    if (lane == 0 && wi < kTimelineWarps) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_wtl[4 * wi] = smid | ((unsigned long long)S.nrestart << 32) | ((unsigned long long)S.nlanes << 40);
      g_wtl[4 * wi + 1] = wt0;
      g_wtl[4 * wi + 2] = gtimer();
      g_wtl2[4 * wi] = (unsigned long long)S.c_front;
      g_wtl2[4 * wi + 1] = (unsigned long long)S.c_wait;
      g_wtl2[4 * wi + 2] = (unsigned long long)S.c_body;
      g_wtl2[4 * wi + 3] = (unsigned long long)S.nstaged | ((unsigned long long)S.ntru << 32);
      g_wtl[4 * wi + 3] = (unsigned long long)((y1[0] - y0[0]) + (y1[1] - y0[1])) | ((unsigned long long)(y1[1] > y0[1]) << 16) |
                          ((unsigned long long)seg[0] << 24) | ((unsigned long long)y0[0] << 32) |
                          ((unsigned long long)S.ndirect << 48);
    }
  }
#endif
  DPFT_STAMP(2, threadIdx.x == 0 && blockIdx.x == 0 && b == 0);   // tile walked
  reduce_and_finish<TRU, kSW>(p, b, redw, S.acc, S.vmin, S.vmax, n_ctas);
}

// =========================================================================== materialised-gradient path
// The unit Sobel gradients of x0 and sigma0 do not depend on the pose, so they can be formed once per level
// (sobel_unit_kernel) and read back by the three iterations.  That costs HBM traffic (4C extra floats per
// pixel and iteration) but turns the iteration into a pixel-parallel kernel with no register windows, no
// halo lanes and every load of a pixel independent of the others -- which is what the latency-bound
// fused kernel above lacks.  DESIGN.md carries the byte accounting of both.

// g = S / sqrt(Sx^2 + Sy^2 + 1e-8), S the replicate-padded Sobel response (algorithms.py:1844-1865)
__global__ void __launch_bounds__(256) sobel_unit_kernel(const float* __restrict__ img, float* __restrict__ gx,
                                                         float* __restrict__ gy, int planes, int H, int W) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const int xl = max(x - 1, 0), xr = min(x + 1, W - 1), yt = max(y - 1, 0) * W, ym = y * W, yb = min(y + 1, H - 1) * W;
  for (int pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const float* q = img + (size_t)pl * H * W;
    const float a = __ldg(q + yt + xl), b = __ldg(q + yt + x), c = __ldg(q + yt + xr);
    const float d = __ldg(q + ym + xl), f = __ldg(q + ym + xr);
    const float g = __ldg(q + yb + xl), h = __ldg(q + yb + x), i = __ldg(q + yb + xr);
    const float sx = (c - a) + 2.f * (f - d) + (i - g);
    const float sy = (g - a) + 2.f * (h - b) + (i - c);
    const float inv = rsqrtf(fmaf(sx, sx, fmaf(sy, sy, 1e-8f)));
    gx[(size_t)pl * H * W + ym + x] = sx * inv;
    gy[(size_t)pl * H * W + ym + x] = sy * inv;
  }
}

void launch_sobel_unit(const float* img, float* gx, float* gy, int planes, int H, int W, cudaStream_t stream) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8, std::min(planes, 4096));
  sobel_unit_kernel<<<grid, 256, 0, stream>>>(img, gx, gy, planes, H, W);
}

struct PxExtra {
  const float *gfx, *gfy, *gsx, *gsy;   // (B,C,H,W) unit gradients of x0 and sigma0
  int ppt;                              // pixels per thread
};

template <int CH, bool TRU>
__global__ void __launch_bounds__(kThreads, 4) uic_iter_px_kernel(const __grid_constant__ UicIterParams p, const PxExtra e) {
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, C = p.C;
  const int iplane = H * W;

  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  float cmn[6], cmx[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) cmn[i] = cmx[i] = 0.f;
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;

  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1);
  const float cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const size_t pair_off = (size_t)b * C * iplane;
  const float* d0p = p.d0 + (size_t)b * iplane;
  const float* d1p = p.d1 + (size_t)b * iplane;
  const uint8_t* m0p = p.m0 ? p.m0 + (size_t)b * iplane : nullptr;
  const uint8_t* m1p = p.m1 ? p.m1 + (size_t)b * iplane : nullptr;

  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const Pose pose = load_pose(p.pose + (size_t)b * 12);
  float s0lo = 0.f, s0hi = 0.f;
  if (TRU) {
    s0lo = ord2f(__ldcg(p.s0mm));
    s0hi = ord2f(__ldcg(p.s0mm + 1));
  }

  for (int i = 0; i < e.ppt; ++i) {
    const int pix = (blockIdx.x * e.ppt + i) * kThreads + threadIdx.x;
    if (pix >= iplane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float d0 = __ldg(d0p + pix);
    float u, v, inv_z;
    warp_pixel(pose, px, py, d0, fx, fy, cx, cy, u, v, inv_z);
    const Tap tap = make_tap(u, v, H, W);
    const float d1w = sample_exact(d1p, tap, W);
    bool occ = occluded(u, v, inv_z, d1w, H, W);
    if (m0p) occ = occ || (__ldg(m0p + pix) == 0);
    if (m1p) occ = occ || !(sample_mask(m1p, tap, W) > 0.f);
    float sr0 = 0.f;
    if (TRU) {
      const float s0c0 = __ldg(p.s0 + pair_off + pix);
      occ = occ || (s0c0 == s0lo) || (s0c0 == s0hi);
    }
    float saa = 0.f, sab = 0.f, sbb = 0.f, sar = 0.f, sbr = 0.f, sca = 0.f, scb = 0.f;
    float pmin = CUDART_INF_F, pmax = -CUDART_INF_F;
    for (int c0 = 0; c0 < C; c0 += CH) {
      const size_t o0 = pair_off + (size_t)c0 * iplane + pix;        // keyframe-side element
      const size_t o1 = pair_off + (size_t)c0 * iplane + tap.o;      // north-west texel of the live frame
      float f0[CH], s0v[CH], gfx[CH], gfy[CH], gsx[CH], gsy[CH];
      float xa[CH], xb[CH], xc[CH], xd[CH], za[CH], zb[CH], zc[CH], zd[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const size_t k0 = o0 + (size_t)c * iplane, k1 = o1 + (size_t)c * iplane;
        f0[c] = __ldg(p.x0 + k0); s0v[c] = __ldg(p.s0 + k0);
        gfx[c] = __ldg(e.gfx + k0); gfy[c] = __ldg(e.gfy + k0);
        gsx[c] = __ldg(e.gsx + k0); gsy[c] = __ldg(e.gsy + k0);
        const float* q1 = p.x1 + k1;
        const float* q2 = p.s1 + k1;
        xa[c] = __ldg(q1); xb[c] = __ldg(q1 + 1); xc[c] = __ldg(q1 + W); xd[c] = __ldg(q1 + W + 1);
        za[c] = __ldg(q2); zb[c] = __ldg(q2 + 1); zc[c] = __ldg(q2 + W); zd[c] = __ldg(q2 + W + 1);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float fr = blend_fast(xa[c], xb[c], xc[c], xd[c], tap);
        // sigma is compared for equality against its batch extremes -> mask-grade arithmetic
        const float sr = TRU ? blend_exact(za[c], zb[c], zc[c], zd[c], tap) : blend_fast(za[c], zb[c], zc[c], zd[c], tap);
        // residual, its uncertainty and the 2-vector d(wres)/d(u,v) (algorithms.py:1969-1972, :872)
        const float res = fr - f0[c];
        const float rs = rsqrtf(fmaf(sr, sr, s0v[c] * s0v[c]));   // 1 / sigma
        const float wres = res * rs;
        const float q = wres * (s0v[c] * (rs * rs));              // res * sigma0 / sigma^3
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
          if (c0 == 0 && c == 0) sr0 = sr;
        }
      }
    }
    float ju[6], jv[6];
    warp_rows(px, py, d0, fx, fy, ju, jv);
    accumulate_system(acc, ju, jv, saa, sab, sbb, sar, sbr);
    if (TRU) {
      // running extremes of the warped sigma and what their pixels added to J^T r (rare -> one uniform branch)
      const bool lo = pmin < vmin, hi = pmax > vmax;
      const float nmin = lo ? pmin : vmin, nmax = hi ? pmax : vmax;
      const bool tmin = !occ && (sr0 == nmin), tmax = !occ && (sr0 == nmax);
      if (__any_sync(__activemask(), lo || hi || tmin || tmax)) {
        vmin = nmin;
        vmax = nmax;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          float cc = 0.f;
          if (k != 4) cc = fmaf(sca, ju[k], cc);
          if (k != 3) cc = fmaf(scb, jv[k], cc);
          cmn[k] = (lo ? 0.f : cmn[k]) + (tmin ? cc : 0.f);
          cmx[k] = (hi ? 0.f : cmx[k]) + (tmax ? cc : 0.f);
        }
      }
    }
    if (p.occ_out) {
      p.occ_out[(size_t)b * iplane + pix] = occ ? 1 : 0;
      if (TRU) p.sr0_dbg[(size_t)b * iplane + pix] = sr0;
    }
  }
  __shared__ float red[kWarps][NSUM][33];
  if (TRU) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      red[warp][27 + i][lane] = cmn[i];
      red[warp][33 + i][lane] = cmx[i];
    }
  }
  reduce_and_finish<TRU>(p, b, red[threadIdx.x >> 5], acc, vmin, vmax, p.ctas_per_pair);
}

// --------------------------------------------------------------------------- small helper kernels
__global__ void init_kernel(const float* __restrict__ pose_in, float* __restrict__ pose0, int n_pose,
                            int* __restrict__ counters, int n_counters, uint32_t* __restrict__ mm, int n_levels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pose) pose0[i] = pose_in[i];
  if (i < n_counters) counters[i] = 0;
  if (i < n_levels) {
    mm[2 * i] = 0xffffffffu;   // running min
    mm[2 * i + 1] = 0u;        // running max
  }
}

// min / max of a whole tensor (torch's sigma0.min(), sigma0.max(): algorithms.py:1976-1977)
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ v, size_t n, uint32_t* __restrict__ mm) {
  __shared__ float s_lo[8], s_hi[8];
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = ((reinterpret_cast<uintptr_t>(v) & 15) == 0) ? n / 4 : 0;
  const float4* v4 = reinterpret_cast<const float4*>(v);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {       // four independent 16-byte loads in flight per thread
    const float4 a = __ldg(v4 + i), b = __ldg(v4 + i + stride), c = __ldg(v4 + i + 2 * stride), d = __ldg(v4 + i + 3 * stride);
    lo = fminf(fminf(fminf(lo, fminf(a.x, a.y)), fminf(fminf(a.z, a.w), fminf(b.x, b.y))),
               fminf(fminf(fminf(b.z, b.w), fminf(c.x, c.y)), fminf(fminf(c.z, c.w), fminf(fminf(d.x, d.y), fminf(d.z, d.w)))));
    hi = fmaxf(fmaxf(fmaxf(hi, fmaxf(a.x, a.y)), fmaxf(fmaxf(a.z, a.w), fmaxf(b.x, b.y))),
               fmaxf(fmaxf(fmaxf(b.z, b.w), fmaxf(c.x, c.y)), fmaxf(fmaxf(c.z, c.w), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w)))));
  }
  for (; i < n4; i += stride) {
    const float4 q = __ldg(v4 + i);
    lo = fminf(fminf(lo, q.x), fminf(q.y, fminf(q.z, q.w)));
    hi = fmaxf(fmaxf(hi, q.x), fmaxf(q.y, fmaxf(q.z, q.w)));
  }
  for (size_t j = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const float q = __ldg(v + j);
    lo = fminf(lo, q);
    hi = fmaxf(hi, q);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                               // one atomic pair per CTA on the two shared words
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      lo = fminf(lo, s_lo[w]);
      hi = fmaxf(hi, s_hi[w]);
    }
    atomicMin(mm, f2ord(lo));
    atomicMax(mm + 1, f2ord(hi));
  }
}

// per-pair variant: blockIdx.y = pair, extremes of v[b, :] into mm[2b], mm[2b+1]
__global__ void __launch_bounds__(256) minmax_pairs_kernel(const float* __restrict__ v, size_t per_pair,
                                                           uint32_t* __restrict__ mm) {
  const float* q = v + (size_t)blockIdx.y * per_pair;
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_pair; i += (size_t)gridDim.x * blockDim.x) {
    const float x = __ldg(q + i);
    lo = fminf(lo, x);
    hi = fmaxf(hi, x);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 2 * blockIdx.y, f2ord(lo));
    atomicMax(mm + 2 * blockIdx.y + 1, f2ord(hi));
  }
}

void launch_minmax(const float* v, size_t n, uint32_t* mm, cudaStream_t stream) {
  const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, 148 * 8);
  minmax_kernel<<<blocks, 256, 0, stream>>>(v, n, mm);
}

// debug only: OR the batch-global sigma test into the per-iteration mask
__global__ void occ_fixup_kernel(uint8_t* __restrict__ occ, const float* __restrict__ sr0, const float* __restrict__ gmm,
                                 size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (sr0[i] == gmm[0] || sr0[i] == gmm[1])) occ[i] = 1;
}

// --------------------------------------------------------------------------- host side
// What dpft_uic_options_t carries, with the defaults filled in (no environment variables, no process state).
// Are the C channels of every sigma0 / sigma1 tensor of the call copies of channel 0?  That is what the reference's
// encoder produces: ONE uncertainty map per frame, repeated to C channels before the tracker sees it
// (algorithms.py:1425-1427, uncertainty_channel = 1 in every shipped configuration).  The kernels then read the one map
// (the DPFT_SIGMA_BROADCAST tile routines with the full tensors' pair stride): same values, C times fewer loads,
// blends, Sobel responses and normalisations of sigma.  Decided on the device -- no host synchronisation, and a CUDA
// graph of the call stays valid when its buffers are refilled with other data: *mism is 0 on entry and set to 1 by any
// thread that finds a channel differing (bitwise) from channel 0.
struct RepTensors {
  const float* p[2 * DPFT_MAX_LEVELS];
  unsigned plane[2 * DPFT_MAX_LEVELS];             // elements per channel plane
  unsigned units[2 * DPFT_MAX_LEVELS];             // work units of the tensor: pairs * ceil(plane / 4)
  unsigned first[2 * DPFT_MAX_LEVELS + 1];         // prefix sums of `units`
  int n;
  // extremes on the way (sigma0 tensors of a remove_tru_sigma call): slot of the tensor's first group in `mm` (-1: none) and
  // work units per sigma-extreme group.  The check loads every channel of every sigma0 element anyway, so the per-level,
  // per-group minimum / maximum the solve needs (algorithms.py:1976-1977) cost no further pass over the tensors.
  int mm_base[2 * DPFT_MAX_LEVELS];
  unsigned upg[2 * DPFT_MAX_LEVELS];
};
// One unit = four consecutive pixels of one pair: all C channels are loaded (16 bytes each when the plane allows it)
// before anything is compared, so a thread keeps C independent loads in flight.
template <int CT>
__global__ void __launch_bounds__(256) sigma_replication_kernel(const RepTensors t, const int C, int* __restrict__ mism,
                                                                uint32_t* __restrict__ mm) {
  const unsigned total = t.first[t.n];
  unsigned diff = 0u;
  int round = 0;
  bool stop = false;                    // a difference is known: units without extremes to collect are skipped from here on
  int key = -1;                         // mm slot the running extremes belong to
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
  auto flush = [&]() {
    if (key >= 0 && lo <= hi) {
      atomicMin(mm + 2 * key, f2ord(lo));
      atomicMax(mm + 2 * key + 1, f2ord(hi));
    }
    lo = CUDART_INF_F;
    hi = -CUDART_INF_F;
  };
  // a CTA walks one contiguous share of the units (not a grid-stride loop): a thread then stays inside one tensor and one
  // sigma-extreme group for nearly all of its units, and its running extremes are flushed a handful of times
  const unsigned per_cta = (total + gridDim.x - 1u) / gridDim.x;
  const unsigned u_begin = min(total, blockIdx.x * per_cta), u_end = min(total, u_begin + per_cta);
  for (unsigned u = u_begin + threadIdx.x; u < u_end; u += 256u) {
    int ti = 0;
#pragma unroll
    for (int k = 1; k < 2 * DPFT_MAX_LEVELS; ++k) ti += (k < t.n && u >= t.first[k]) ? 1 : 0;
    const unsigned plane = t.plane[ti], cpp = (plane + 3u) / 4u;
    const unsigned w = u - t.first[ti];
    const bool ext = mm != nullptr && t.mm_base[ti] >= 0;
    if (stop && !ext) continue;
    if (ext) {
      const int nk = t.mm_base[ti] + (int)(w / t.upg[ti]);
      if (nk != key) {
        flush();
        key = nk;
      }
    }
    const unsigned pair = w / cpp, ch = w - pair * cpp;
    const float* r = t.p[ti] + (size_t)pair * C * plane + 4u * ch;
    if ((plane & 3u) == 0u && ((uintptr_t)t.p[ti] & 15u) == 0u) {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(r));
      auto fold = [&](const uint4 a) {
        const float x = __uint_as_float(a.x), y = __uint_as_float(a.y), z = __uint_as_float(a.z), q = __uint_as_float(a.w);
        lo = fminf(fminf(lo, fminf(x, y)), fminf(z, q));
        hi = fmaxf(fmaxf(hi, fmaxf(x, y)), fmaxf(z, q));
      };
      if (ext) fold(v0);
      if (CT > 0) {
        uint4 v[CT > 0 ? CT : 1];
#pragma unroll
        for (int c = 1; c < CT; ++c) v[c] = __ldg(reinterpret_cast<const uint4*>(r + (size_t)c * plane));
#pragma unroll
        for (int c = 1; c < CT; ++c) {
          diff |= (v[c].x ^ v0.x) | (v[c].y ^ v0.y) | (v[c].z ^ v0.z) | (v[c].w ^ v0.w);
          if (ext) fold(v[c]);
        }
      } else {
        for (int c = 1; c < C; ++c) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(r + (size_t)c * plane));
          diff |= (a.x ^ v0.x) | (a.y ^ v0.y) | (a.z ^ v0.z) | (a.w ^ v0.w);
          if (ext) fold(a);
        }
      }
    } else {
      const unsigned left = min(4u, plane - 4u * ch);
      for (unsigned e = 0; e < left; ++e) {
        const unsigned v0 = __float_as_uint(__ldg(r + e));
        if (ext) { lo = fminf(lo, __uint_as_float(v0)); hi = fmaxf(hi, __uint_as_float(v0)); }
        for (int c = 1; c < C; ++c) {
          const unsigned a = __float_as_uint(__ldg(r + (size_t)c * plane + e));
          diff |= a ^ v0;
          if (ext) { lo = fminf(lo, __uint_as_float(a)); hi = fmaxf(hi, __uint_as_float(a)); }
        }
      }
    }
    // somebody found a difference: without extremes to collect there is no need to go on
    if (((++round) & 7) == 0 && (diff || *(volatile int*)mism)) {
      if (!mm) break;
      stop = true;
    }
  }
  if (mm) {
    // the last flush goes through the warp when its lanes ended on the same slot (they nearly always do: a warp's units are
    // neighbours), so an address takes a few hundred atomics per launch instead of one per thread
    const int key0 = __shfl_sync(0xffffffffu, key, 0);
    if (__all_sync(0xffffffffu, key == key0)) {
      const uint32_t wlo = __reduce_min_sync(0xffffffffu, f2ord(lo)), whi = __reduce_max_sync(0xffffffffu, f2ord(hi));
      if ((threadIdx.x & 31) == 0 && key0 >= 0 && wlo <= whi) {
        atomicMin(mm + 2 * key0, wlo);
        atomicMax(mm + 2 * key0 + 1, whi);
      }
    } else {
      flush();
    }
  }
  if (diff) *mism = 1;
}

struct Tuning {
  int group = 0;                          // pairs per sigma-extreme group (queue path); 0 = B
  int tile_rows[DPFT_MAX_LEVELS] = {};    // queue path: rows per tile, 0 = chosen
  int queue_ctas = 0;
  int queue_levels = 0;                   // finest levels that run as work-queue launches; 0 = 1
  long cta_slots = 0;                     // 0 = 148 x resident CTAs
  int tiling = 0;                         // 0 dealt, 1 rectangular, 2 linear
  bool generic_geometry = false;
  bool no_resident = false;       // measurement knob: small levels on the plain (global-memory lookup) kernel
  bool no_narrow = false;         // measurement knob: narrow levels on the work queue keep the plain tile routine
  bool no_sigma_detect = false;   // do not look for full sigma tensors whose channels are copies of channel 0
  const int* mism = nullptr;      // internal: device flag of sigma_replication_kernel for this call (run_uic sets it)
  uint32_t* mm_ready = nullptr;   // internal: sigma0 extremes of every level, collected by that kernel on the way (level, group, 2)
  float* launch_ms = nullptr;
  float* queue_kernel_ms = nullptr;
  const float* icp_weight[DPFT_MAX_LEVELS] = {};   // per level: (B,1,H,W) scale of the ICP term, or nullptr (scalar w_icp)
};

static Tuning tuning_of(const dpft_uic_options_t* o) {
  Tuning t;
  if (!o || o->struct_bytes < sizeof(dpft_uic_options_t)) return t;
  t.group = o->group;
  for (int l = 0; l < DPFT_MAX_LEVELS; ++l) t.tile_rows[l] = o->tile_rows[l];
  t.queue_ctas = o->queue_ctas;
  t.queue_levels = o->queue_levels;
  t.cta_slots = o->cta_slots;
  t.tiling = o->tiling;
  t.generic_geometry = o->generic_geometry != 0;
  t.launch_ms = o->launch_ms;
  t.queue_kernel_ms = o->launch_ms ? o->queue_kernel_ms : nullptr;
  for (int l = 0; l < DPFT_MAX_LEVELS; ++l) t.icp_weight[l] = o->icp_weight[l];
  t.no_resident = o->small_levels == 1;
  t.no_narrow = o->small_levels == 2;
  t.no_sigma_detect = o->sigma_detect == 1;
  return t;
}

struct Plan {
  int nseg[DPFT_MAX_LEVELS], nrt[DPFT_MAX_LEVELS], TR[DPFT_MAX_LEVELS], ctas[DPFT_MAX_LEVELS];
  int ppt[DPFT_MAX_LEVELS], px_ctas[DPFT_MAX_LEVELS];   // materialised-gradient path: pixels per thread, CTAs per pair
  int max_ctas;
  TileTab tab[DPFT_MAX_LEVELS];   // balanced tiling of the levels the staged kernel takes
  int res_smem[DPFT_MAX_LEVELS];  // > 0: the level runs the resident variant with this many bytes of dynamic shared memory
  size_t max_plane;
  size_t off_partials, off_pairrec, off_counters, off_mm, off_gmm, off_sr0, off_grad, grad_elems;
  size_t off_vn, off_icp, off_icp_scratch, off_dmm, total;
  // persistent (single cooperative launch) path
  int p_grid, p_TR[DPFT_MAX_LEVELS], p_nrt[DPFT_MAX_LEVELS], p_rcap;
  size_t off_records, off_gext, off_clock;
};

// Rows per warp tile.  A tile is walked row by row by one warp, so the time of a launch is about
// (waves of CTAs) x (rows per tile + ~1.5 rows of window priming and reduction tail): pick the height that
// minimises that, i.e. fill whole waves of the 148 x DPFT_MIN_CTAS resident CTAs.
static int pick_tile_rows(int H, int nseg, int B, int ctas_per_sm, int warps, const Tuning& tun, double fixed_rows = 1.5) {
  long slots = 148L * ctas_per_sm;
  if (tun.cta_slots > 0) slots = tun.cta_slots;
  int best_tr = 1;
  double best = 1e30;
  for (int tr = 1; tr <= kMaxTileRows; ++tr) {
    const long nrt = (H + tr - 1) / tr;
    const long ctas = ((nrt * nseg + warps - 1) / warps) * B;
    const long waves = (ctas + slots - 1) / slots;
    const double cost = (double)waves * (tr + fixed_rows);
    if (cost < best - 1e-9) { best = cost; best_tr = tr; }
  }
  return best_tr;
}

// can this level run the staged-footprint kernel?  (16-byte cp.async rows, one 8-channel pass)
static bool staged_ok(const dpft_level_t& L, int C) {
  auto al = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  return C == 8 && L.W % 4 == 0 && L.W >= 2 * kTileCols && L.H >= kStageRows && al(L.x1) && al(L.sigma1) && al(L.invd1);
}

// Levels narrower than the staged routine's ring (the 30x40 and 15x20 levels of a 120x160 pyramid) on the work queue: the
// routine's NARROW form stages whole map rows.  No object masks (that instantiation is not built).
static bool staged_narrow_ok(const dpft_level_t& L, int C) {
  auto al = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  return C == 8 && L.W % 4 == 0 && L.W >= 8 && L.W < kStageWidth && L.H >= kStageRows && al(L.x1) && al(L.sigma1) &&
         al(L.invd1) && !L.obj_mask0 && !L.obj_mask1;
}

// Small levels: the whole live frame of a pair (x1, sigma1, invd1) in the shared memory of every CTA that works on the
// pair (uic_iter_kernel<.., RES>).  Taken when the staged kernel does not apply and at least two CTAs fit an SM.
constexpr int kResStaticSmem = (kWarps * NSUM * 33 + 16) * 4 + kWarps * (NSUM + 1) * 8 + 64;   // red, s_pose, wsum, flags
static int resident_smem(const dpft_level_t& L, int C, int SC) {
  const size_t bytes = (size_t)(C + SC + 1) * L.H * L.W * sizeof(float);
  return bytes + kResStaticSmem <= 110 * 1024 ? (int)bytes : 0;
}

// CTAs of the resident variant an SM holds (what make_plan assumes when it picks the tile height)
static int resident_ctas(int res_smem) {
  return res_smem > 0 ? std::max(1, std::min(kResMinCtas, (227 * 1024) / (res_smem + kResStaticSmem + 1024))) : 0;
}

// Linear variant of the balanced tile table (DPFT_LINEAR_TILES=1; see TileTab).  Returns false when the rectangular tiling should stay (the
// table holds at most kTabWarps warps per pair, and a warp's range may cross one segment boundary at most).
static bool make_tile_tab_linear(int H, int nseg, int B, TileTab& tab, const Tuning& tun) {
  tab = TileTab{};
  // 35/37 of the resident slots: filling every slot makes a lone launch ~3 % shorter still, but then the small
  // launches of OTHER streams (coarse levels of independent batches) find no free slot and the multi-stream
  // throughput drops by 5 % (profiles/exp9.sh)
  long slots = 148L * DPFT_STAGED_CTAS * 35 / 37;
  if (tun.cta_slots > 0) slots = tun.cta_slots;
  if (tun.tiling == 1) return false;
  const int c0 = (int)(slots / B);               // CTAs per pair of kind 0; kind 1 has one more
  const int n_more = (int)(slots - (long)c0 * B);
  if (c0 < 1 || (c0 + 1) * kSW > kTabWarps || c0 * kSW < nseg + 1) return false;
  const long R = (long)nseg * H;                 // warp-rows of a pair
  for (int kind = 0; kind < 2; ++kind) {
    const int ctas = c0 + kind, nw = ctas * kSW;
    // range ends, snapped to a segment boundary when they fall within two rows of one (no 1- or 2-row sub-tiles)
    auto bound = [&](int w) {
      long e = R * w / nw;
      const long r = e % H;
      if (r <= 2) e -= r;
      else if (r >= H - 2) e += H - r;
      return e;
    };
    tab.ctas[kind] = ctas;
    for (int w = 0; w < nw; ++w) {
      const long s = bound(w), e = bound(w + 1);
      const long sa = s / H, cut = std::min(e, (sa + 1) * H);
      if (e > (sa + 2) * H) return false;
      tab.seg[kind][w][0] = (unsigned char)sa;
      tab.y0[kind][w][0] = (short)(s - sa * H);
      tab.y1[kind][w][0] = (short)(cut - sa * H);
      tab.seg[kind][w][1] = (unsigned char)std::min<long>(sa + 1, nseg - 1);
      tab.y0[kind][w][1] = 0;
      tab.y1[kind][w][1] = (short)(e - cut);
    }
  }
  tab.n_more = n_more;
  tab.on = 1;
  return true;
}

// Balanced tile table of one level (see TileTab), the default: every segment is cut into whole row tiles (some
// segments into one tile more than the others), taller tiles first, dealt round-robin to the CTAs so that every CTA
// gets the same mix of heights.  Returns false when the rectangular tiling should stay.
static bool make_tile_tab(int H, int nseg, int B, TileTab& tab, const Tuning& tun) {
  if (tun.tiling == 2 || kSW != 4) return make_tile_tab_linear(H, nseg, B, tab, tun);
  tab = TileTab{};
  // 420 of the 444 resident slots: filling every slot makes a lone launch ~3 % shorter still, but then the small
  // launches of OTHER streams (coarse levels of independent batches) find no free slot and the multi-stream
  // throughput drops by 5 % (profiles/exp9.sh)
  long slots = 148L * DPFT_STAGED_CTAS * 35 / 37;
  if (tun.cta_slots > 0) slots = tun.cta_slots;
  if (tun.tiling == 1) return false;
  const int c0 = (int)(slots / B);               // CTAs per pair of kind 0; kind 1 has one more
  const int n_more = (int)(slots - (long)c0 * B);
  if (c0 < 1 || (c0 + 1) * kSW > kTabWarps || kSW * c0 < nseg) return false;
  for (int kind = 0; kind < 2; ++kind) {
    const int ctas = c0 + kind, tiles = kSW * ctas;
    const int base = tiles / nseg, extra = tiles - base * nseg;     // `extra` segments get base + 1 row tiles
    if (base + 1 > H) return false;
    struct T { int seg, y0, y1; } list[kTabWarps];
    int n = 0;
    // taller tiles first, then deal round-robin: every CTA gets the same mix of heights
    for (int pass = 0; pass < 2; ++pass)
      for (int sg = 0; sg < nseg; ++sg) {
        const int nt = base + (sg < extra ? 1 : 0);
        if ((nt == base) != (pass == 0)) continue;
        for (int k = 0; k < nt; ++k) list[n++] = T{sg, (int)((long)H * k / nt), (int)((long)H * (k + 1) / nt)};
      }
    tab.ctas[kind] = ctas;
    for (int i = 0; i < n; ++i) {
      const int w = (i % ctas) * kSW + i / ctas;       // CTA i % ctas, warp i / ctas
      tab.seg[kind][w][0] = (unsigned char)list[i].seg;
      tab.y0[kind][w][0] = (short)list[i].y0;
      tab.y1[kind][w][0] = (short)list[i].y1;
    }
  }
  tab.n_more = n_more;
  tab.on = 1;
  return true;
}

static Plan make_plan(const dpft_level_t* lv, int n_levels, int B, int C, uint32_t flags, bool any_occ,
                      int p_grid, const Tuning& tun, int n_groups, int n_mm_groups) {
  Plan pl{};
  pl.max_ctas = 1;
  pl.max_plane = 0;
  for (int l = 0; l < n_levels; ++l) {
    const bool staged = (flags & DPFT_STAGED_FOOTPRINT) && (flags & DPFT_FUSED_SOBEL) && staged_ok(lv[l], C);
    const int cols = staged ? kStagedCols : kCols;
    pl.nseg[l] = (lv[l].W + cols - 1) / cols;
    const int cta_warps = staged ? kSW : kWarps;
    const int SC = (flags & DPFT_SIGMA_BROADCAST) ? 1 : C;
    pl.res_smem[l] = (!staged && (flags & DPFT_FUSED_SOBEL) && !tun.no_resident) ? resident_smem(lv[l], C, SC) : 0;
    int ctas_per_sm = staged ? DPFT_STAGED_CTAS : DPFT_MIN_CTAS;
    if (pl.res_smem[l]) ctas_per_sm = resident_ctas(pl.res_smem[l]);
    // (a resident CTA pays for its copy of the live frame first: about two tile rows' worth of time, which the cost
    // model of pick_tile_rows sees as taller tiles being cheaper)
    pl.TR[l] = pick_tile_rows(lv[l].H, pl.nseg[l], B, ctas_per_sm, cta_warps, tun, pl.res_smem[l] ? 3.5 : 1.5);
    pl.nrt[l] = (lv[l].H + pl.TR[l] - 1) / pl.TR[l];
    pl.ctas[l] = (pl.nseg[l] * pl.nrt[l] + cta_warps - 1) / cta_warps;
    pl.tab[l].on = 0;
    if ((flags & DPFT_STAGED_FOOTPRINT) && (flags & DPFT_FUSED_SOBEL) && staged_ok(lv[l], C) && lv[l].H < 32768 &&
        make_tile_tab(lv[l].H, pl.nseg[l], B, pl.tab[l], tun))
      pl.ctas[l] = pl.tab[l].ctas[1];   // grid.x and the record stride
    if (pl.ctas[l] > pl.max_ctas) pl.max_ctas = pl.ctas[l];
    const size_t plane = (size_t)lv[l].H * lv[l].W;
    {
      const long want_threads = 148L * 2048 * 2;   // two full waves of resident threads
      long ppt = ((long)B * (long)plane + want_threads - 1) / want_threads;
      ppt = std::max(1L, std::min(ppt, 8L));
      pl.ppt[l] = (int)ppt;
      pl.px_ctas[l] = (int)((plane + (size_t)kThreads * ppt - 1) / ((size_t)kThreads * ppt));
      if (pl.px_ctas[l] > pl.max_ctas) pl.max_ctas = pl.px_ctas[l];
    }
    if (plane > pl.max_plane) pl.max_plane = plane;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  pl.off_partials = take((size_t)B * pl.max_ctas * PS * sizeof(float));
  pl.off_pairrec = take((size_t)B * PS * sizeof(double));
  pl.off_counters = take((size_t)(B + n_groups) * sizeof(int));
  pl.off_mm = take((size_t)2 * DPFT_MAX_LEVELS * sizeof(uint32_t) * n_mm_groups);
  pl.off_gmm = take(4 * sizeof(float) * DPFT_MAX_LEVELS * 64 * n_groups);
  pl.off_sr0 = take(((flags & DPFT_REMOVE_TRU_SIGMA) && any_occ) ? (size_t)B * pl.max_plane * sizeof(float) : 0);
  pl.grad_elems = (flags & DPFT_FUSED_SOBEL) ? 0 : (size_t)B * C * pl.max_plane;
  pl.off_grad = take(4 * pl.grad_elems * sizeof(float));
  const bool icp = flags & DPFT_COMBINE_ICP;
  pl.off_vn = take(icp ? 6 * (size_t)B * pl.max_plane * sizeof(float) : 0);
  pl.off_icp = take(icp ? (size_t)B * 28 * sizeof(float) : 0);
  {
    size_t scratch = 0;
    for (int l = 0; l < n_levels && icp; ++l) scratch = std::max(scratch, icp_scratch_bytes(B, lv[l].H, lv[l].W));
    pl.off_icp_scratch = take(scratch);
  }
  pl.off_dmm = take(2 * DPFT_MAX_LEVELS * sizeof(uint32_t));
  // persistent path: tile height per level that minimises the rows the busiest warp walks (a tile costs its
  // rows plus about one row of window priming), record slots per pair
  pl.p_grid = p_grid;
  pl.p_rcap = 1;
  if (p_grid > 0) {
    const long nw = (long)p_grid * kWarps;
    for (int l = 0; l < n_levels; ++l) {
      double best = 1e30;
      for (int tr = 1; tr <= kMaxTileRows; ++tr) {
        const long nrt = (lv[l].H + tr - 1) / tr, tiles = nrt * pl.nseg[l] * B;
        const long T = (tiles + nw - 1) / nw;
        const double cost = (double)T * (tr + 1.0);
        if (cost < best - 1e-9) { best = cost; pl.p_TR[l] = tr; pl.p_nrt[l] = (int)nrt; }
      }
      const long tpp = (long)pl.p_nrt[l] * pl.nseg[l];
      const long T = (tpp * B + nw - 1) / nw;
      pl.p_rcap = std::max(pl.p_rcap, (int)((tpp + T - 1) / T + 1));
    }
  }
  pl.off_records = take(p_grid > 0 ? (size_t)B * pl.p_rcap * PS * sizeof(float) : 0);
  pl.off_gext = take(2 * DPFT_MAX_LEVELS * 64 * sizeof(uint32_t));
  pl.off_clock = take((DPFT_MAX_LEVELS * 64 + 1) * sizeof(unsigned long long));
  pl.total = off;
  return pl;
}

static int check_args(const dpft_level_t* lv, int n_levels, int B, int C, int iters, uint32_t flags) {
  if (!lv || n_levels < 1 || n_levels > DPFT_MAX_LEVELS) return set_error(DPFT_EINVAL, "n_levels must be 1..%d", DPFT_MAX_LEVELS);
  if (B < 1 || B > 65535) return set_error(DPFT_EINVAL, "B must be 1..65535 (got %d)", B);
  if (C < 1 || iters < 0) return set_error(DPFT_EINVAL, "C must be >= 1 and iters >= 0");
  for (int l = 0; l < n_levels; ++l) {
    const dpft_level_t& L = lv[l];
    if (L.H < 2 || L.W < 2) return set_error(DPFT_EINVAL, "level %d: H and W must be >= 2", l);
    if (!L.x0 || !L.x1 || !L.sigma0 || !L.sigma1 || !L.invd0 || !L.invd1 || !L.K)
      return set_error(DPFT_EINVAL, "level %d: x0, x1, sigma0, sigma1, invd0, invd1 and K are required", l);
    if ((flags & DPFT_COMBINE_ICP) && (!L.depth0 || !L.depth1))
      return set_error(DPFT_EINVAL, "level %d: DPFT_COMBINE_ICP needs depth0 and depth1", l);
    if ((flags & DPFT_PAIRWISE_EXTREMES) && (flags & DPFT_REMOVE_TRU_SIGMA) && L.occ_out)
      return set_error(DPFT_EINVAL, "level %d: occ_out is not produced with DPFT_PAIRWISE_EXTREMES", l);
  }
  // Flag mixes no kernel serves.  The keyframe-side tensors of DPFT_SHARED_KEYFRAME have batch size 1, which only
  // the fused U_IC kernels index accordingly: the materialised-gradient pass and the ICP term would read past them.
  if ((flags & (DPFT_SHARED_KEYFRAME | DPFT_PAIRWISE_EXTREMES)) && ((flags & DPFT_COMBINE_ICP) || !(flags & DPFT_FUSED_SOBEL)))
    return set_error(DPFT_EINVAL, "DPFT_SHARED_KEYFRAME / DPFT_PAIRWISE_EXTREMES need DPFT_FUSED_SOBEL and exclude DPFT_COMBINE_ICP");
  if ((flags & DPFT_SIGMA_BROADCAST) && !(flags & DPFT_FUSED_SOBEL))
    return set_error(DPFT_EINVAL, "DPFT_SIGMA_BROADCAST needs the fused kernels (DPFT_FUSED_SOBEL)");
  return 0;
}

// the sigma-extreme group of the options: a divisor of B; groups other than the whole batch are served by the fused
// U_IC kernels only (the ICP term's normals and the per-iteration mask output know batch-global extremes only)
static int check_group(const dpft_level_t* lv, int n_levels, int B, uint32_t flags, const Tuning& tun) {
  if (tun.group < 0 || (tun.group > 0 && B % tun.group != 0))
    return set_error(DPFT_EINVAL, "options.group (%d) must divide B (%d)", tun.group, B);
  if (tun.group > 0 && tun.group != B) {
    bool any_occ = false;
    for (int l = 0; l < n_levels; ++l) any_occ = any_occ || lv[l].occ_out;
    if (!(flags & DPFT_FUSED_SOBEL) || (flags & DPFT_COMBINE_ICP) || any_occ)
      return set_error(DPFT_EINVAL, "options.group = %d needs DPFT_FUSED_SOBEL and excludes DPFT_COMBINE_ICP and occ_out", tun.group);
  }
  return 0;
}

template <int CH>
static cudaError_t launch_iter(const UicIterParams& prm, dim3 grid, bool tru, bool pdl, cudaStream_t stream,
                               const Tuning& tun, int res_smem = 0) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)res_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  // the reference's pyramid sizes (TUM 160x120 ... and 640x480 ...) run geometry-specialised instantiations
  if (CH == 8 && !tun.generic_geometry && res_smem == 0) {
#define DPFT_FIXED(w, h)                                                                     \
    if (prm.W == w && prm.H == h)                                                            \
      return tru ? cudaLaunchKernelEx(&cfg, uic_iter_kernel<8, true, w, h>, prm)             \
                 : cudaLaunchKernelEx(&cfg, uic_iter_kernel<8, false, w, h>, prm);
    DPFT_FIXED(160, 120)
    DPFT_FIXED(80, 60)
#undef DPFT_FIXED
  }
  if (res_smem > 0) {
    auto* fn = tru ? uic_iter_kernel<CH, true, 0, 0, true> : uic_iter_kernel<CH, false, 0, 0, true>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem);
    return cudaLaunchKernelEx(&cfg, fn, prm);
  }
  if (tru) return cudaLaunchKernelEx(&cfg, uic_iter_kernel<CH, true>, prm);
  return cudaLaunchKernelEx(&cfg, uic_iter_kernel<CH, false>, prm);
}

#if DPFT_STAGED_TMA
// Tensor map of a contiguous (B, C, H, W) fp32 tensor whose box is one staged row of all its channel planes.
static bool encode_row_map(CUtensorMap* tm, const float* base, int W, int H, int C, int B) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q) return false;
    fn = (encode_fn)q;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)kStageWidth, 1, (cuuint32_t)C, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

static cudaError_t launch_staged(const UicIterParams& prm_in, dim3 grid, bool tru, bool pdl, cudaStream_t stream,
                                 const Tuning& tun) {
  UicIterParams prm = prm_in;
#if DPFT_STAGED_TMA
  if (!encode_row_map(&prm.tm_x1, prm.x1, prm.W, prm.H, prm.C, prm.B) ||
      !encode_row_map(&prm.tm_s1, prm.s1, prm.W, prm.H, prm.SC, prm.B) ||
      !encode_row_map(&prm.tm_d1, prm.d1, prm.W, prm.H, 1, prm.B))
    return cudaErrorInvalidValue;
#endif
  constexpr int smem = kSW * kStageAreaFloats * (int)sizeof(float);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kSThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  // (cudaFuncSetAttribute is per device and cheap: set it on every launch rather than caching per process)
#define DPFT_STAGED_A(TRUV, SBV, w, h, AUXV)                                                                 \
  do {                                                                                                       \
    auto* fn = uic_iter_staged_kernel<TRUV, SBV, w, h, AUXV>;                                                \
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                             \
    return cudaLaunchKernelEx(&cfg, fn, prm);                                                                \
  } while (0)
  // object masks and per-pixel debug outputs are compiled out of the instantiations that run without them
#define DPFT_STAGED(TRUV, SBV, w, h)                                                                         \
  do {                                                                                                       \
    if (aux) DPFT_STAGED_A(TRUV, SBV, w, h, true);                                                           \
    DPFT_STAGED_A(TRUV, SBV, w, h, false);                                                                   \
  } while (0)
  const bool aux = prm.m0 || prm.m1 || prm.occ_out;
  const bool sb = prm.SC != prm.C;
  if (!tun.generic_geometry) {
    if (prm.W == 160 && prm.H == 120) {
      if (sb) { if (tru) DPFT_STAGED(true, true, 160, 120); else DPFT_STAGED(false, true, 160, 120); }
      if (tru) DPFT_STAGED(true, false, 160, 120); else DPFT_STAGED(false, false, 160, 120);
    }
    if (prm.W == 80 && prm.H == 60 && !sb) { if (tru) DPFT_STAGED(true, false, 80, 60); else DPFT_STAGED(false, false, 80, 60); }
  }
  if (sb) { if (tru) DPFT_STAGED(true, true, 0, 0); else DPFT_STAGED(false, true, 0, 0); }
  if (tru) DPFT_STAGED(true, false, 0, 0);
  DPFT_STAGED(false, false, 0, 0);
#undef DPFT_STAGED_A
#undef DPFT_STAGED
}

template <int CH>
static cudaError_t launch_px(const UicIterParams& prm, const PxExtra& ex, dim3 grid, bool tru, bool pdl,
                             cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (tru) return cudaLaunchKernelEx(&cfg, uic_iter_px_kernel<CH, true>, prm, ex);
  return cudaLaunchKernelEx(&cfg, uic_iter_px_kernel<CH, false>, prm, ex);
}

}  // namespace dpft

using namespace dpft;

// The single-launch path serves the plain U_IC solve; the ICP term, the per-iteration mask output and the
// materialised-gradient variant keep one launch per iteration.
static bool persistent_ok(uint32_t flags, bool any_occ) {
  return (flags & DPFT_FUSED_SOBEL) && !any_occ &&
         !(flags & (DPFT_COMBINE_ICP | DPFT_LAUNCH_PER_ITERATION | DPFT_SHARED_KEYFRAME | DPFT_PAIRWISE_EXTREMES |
                    DPFT_SIGMA_BROADCAST));
}

// co-resident CTAs of the cooperative kernel: an occupancy query per (device, instantiation), remembered (the answer
// is a property of the device and the binary; racing first calls compute the same value)
static int persistent_grid_cached(int C, bool tru) {
  constexpr int kMaxDev = 64;
  static int cache[kMaxDev][2][4] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return std::max(persistent_grid(C, tru), -1);
  const int idx = (C % 8 == 0) ? 3 : (C % 4 == 0) ? 2 : (C % 2 == 0) ? 1 : 0;
  if (!cache[dev][tru][idx]) cache[dev][tru][idx] = std::max(persistent_grid(C, tru), -1);
  return cache[dev][tru][idx];
}

// ------------------------------------------------------------------------------------------- queue path (uic_queue.cu)
// The finest level of a solve as one work-queue launch (all its iterations, per-pair dependencies).
struct QPlan {
  int nseg, TR, nrt, tpp, kind, grid;
  size_t total_items;
  size_t off_fifo, off_qctl, off_tiles_done, off_cand, off_pairs_done, off_groups_done, off_gext, off_mm, off_records,
      off_pairrec, off_tdone, off_aux, total;
};

static int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms < 1)
    sms = 148;
  return sms;
}

// Rows per tile of the queue path.  Workers are warps; an iteration of the level offers B * nseg * ceil(H / TR) tiles.
// A tile costs its rows plus ~3.5 rows of fixed work (claim, pose, window and ring priming, record, counters; measured:
// 40-row tiles beat 30-row ones by 2 % at 8 batches of 64 pairs, profiles/r2/).  With several waves of tiles per
// iteration the pairs drift apart and only the fixed work counts, so tall tiles win; with a wave or two the last tile
// of a pair IS its iteration, so the count of waves counts.
static int queue_tile_rows(int H, int nseg, int B, long workers) {
  int best_tr = 1;
  double best = 1e30;
  for (int tr = 4; tr <= std::min(H, 48); ++tr) {
    const long nrt = (H + tr - 1) / tr;
    const int tr_eff = (int)((H + nrt - 1) / nrt);          // equal tiles: ceil(H / nrt) rows
    const long tiles = nrt * nseg * B;
    const double waves = (double)tiles / (double)workers;
    const double quant = waves >= 3.0 ? waves + 0.5 : (double)((tiles + workers - 1) / workers);
    const double cost = quant * (tr_eff + 3.5);
    if (cost < best - 1e-9) { best = cost; best_tr = tr_eff; }
  }
  // Throughput regime: while an iteration still offers 2.5 waves of tiles or more, the pairs drift apart, the workers never
  // run dry, and only the fixed work per tile counts -- the tallest such tile wins, up to whole columns (measured, one-map
  // routine, profiles/r2/r2e_fine_tile_rows_probe.txt: 120x160 at 20 batches 40 / 60 / 120 rows 2636 / 2577 / 2543 us, at 12
  // batches 1601 / 1565 / 1546; 480x640 with 64 frames 48 / 120 rows 2190 / 2140 us, with 16 frames -- 1.6 waves at 60
  // rows -- 48 rows stay best; below 2.5 waves taller tiles lose, e.g. 60x80 at 12 batches).
  for (long nrt = 1; nrt <= H; ++nrt) {
    const int tr_eff = (int)((H + nrt - 1) / nrt);
    if (tr_eff <= best_tr) break;
    if ((double)(nrt * nseg * B) / (double)workers >= 2.5) return tr_eff;
  }
  return best_tr;
}

// sigma-extreme groups of a call: `group` consecutive pairs are one batch of the reference
struct Groups {
  int group, n_groups, n_mm_groups;
};
static Groups groups_of(int B, uint32_t flags, const Tuning& tun) {
  Groups g;
  g.group = tun.group > 0 ? tun.group : ((flags & DPFT_PAIRWISE_EXTREMES) ? 1 : B);
  g.n_groups = B / g.group;
  g.n_mm_groups = (flags & DPFT_SHARED_KEYFRAME) ? 1 : g.n_groups;
  return g;
}

static bool queue_wanted(int C, int iters, uint32_t flags, bool any_occ) {
  return (flags & DPFT_QUEUE) && (flags & DPFT_FUSED_SOBEL) && !(flags & DPFT_COMBINE_ICP) && !any_occ && C == 8 && iters >= 1;
}

static QPlan make_qplan(const dpft_level_t& lv, int level_index, int B, int C, int iters, uint32_t flags, const Tuning& tun) {
  QPlan q{};
  const Groups G = groups_of(B, flags, tun);
  const int sms = device_sms();
  // (full sigma tensors whose channels turn out to be copies run the one-map twin: the tile height is shared by the
  // twins, so it is chosen for the variant the flags announce)
  const long workers = (long)sms * queue_tiles_per_sm((flags & DPFT_SIGMA_BROADCAST) != 0);
  const bool staged = (flags & DPFT_STAGED_FOOTPRINT) && staged_ok(lv, C);
  q.kind = staged ? 1 : ((flags & DPFT_STAGED_FOOTPRINT) && !tun.no_narrow && staged_narrow_ok(lv, C)) ? 2 : 0;
  q.nseg = (lv.W + kCols - 1) / kCols;
  const int want = tun.tile_rows[level_index];
  const int tr = want > 0 ? std::min(want, (int)lv.H) : queue_tile_rows(lv.H, q.nseg, B, workers);
  q.nrt = (lv.H + tr - 1) / tr;
  q.TR = (lv.H + q.nrt - 1) / q.nrt;
  q.nrt = (lv.H + q.TR - 1) / q.TR;
  q.tpp = q.nseg * q.nrt;
  q.total_items = (size_t)iters * B * q.tpp;
  q.grid = tun.queue_ctas > 0 ? tun.queue_ctas : 0;      // 0: launch_queue sizes the grid per kernel variant
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  q.off_fifo = take(q.total_items * sizeof(unsigned long long));
  q.off_qctl = take(2 * sizeof(unsigned));
  q.off_tiles_done = take((size_t)B * sizeof(int));
  q.off_cand = take((size_t)B * sizeof(int));
  q.off_pairs_done = take((size_t)iters * G.n_groups * sizeof(int));
  q.off_groups_done = take((size_t)iters * sizeof(int));
  q.off_gext = take((size_t)iters * G.n_groups * 2 * sizeof(uint32_t));
  q.off_mm = take((size_t)G.n_mm_groups * 2 * sizeof(uint32_t));
  q.off_records = take((size_t)B * q.tpp * PS * sizeof(float));
  q.off_pairrec = take((size_t)B * PS * sizeof(double));
  q.off_tdone = take((size_t)(iters + 1) * sizeof(unsigned long long));
  q.off_aux = take((size_t)iters * G.n_groups * 4 * sizeof(float));
  q.total = off;
  return q;
}

// One level through the work queue.  pose_hist / sys_hist / aux_hist point at THIS level's rows; pose_in may be
// pose_hist (the coarser levels already left the level's starting pose there).  launch_ms (host, iters) optional.
static int run_queue(const dpft_level_t& L, int level_index, int B, int C, int iters, uint32_t flags, const float* pose_in,
                     float* pose_hist, float* sys_hist, float* aux_hist, int32_t* status, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream, const Tuning& tun, float* launch_ms,
                     const uint32_t* s0mm_ready /* this level's sigma0 extremes if already computed */, float* kernel_ms) {
  const QPlan q = make_qplan(L, level_index, B, C, iters, flags, tun);
  const Groups G = groups_of(B, flags, tun);
  if (workspace_bytes < q.total) return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, q.total);
  if (q.total_items >= (1ull << 32) || q.tpp >= (1 << 20) || iters >= (1 << 19) || B >= (1 << 24))
    return set_error(DPFT_EINVAL, "problem too large for the work queue's item encoding");
  char* ws = (char*)workspace;
  QueueParams prm{};
  QLevel& v = prm.L;
  v.x0 = L.x0; v.x1 = L.x1; v.s0 = L.sigma0; v.s1 = L.sigma1; v.d0 = L.invd0; v.d1 = L.invd1; v.K = L.K;
  v.m0 = L.obj_mask0; v.m1 = L.obj_mask1;
  v.H = L.H; v.W = L.W; v.nseg = q.nseg; v.TR = q.TR; v.nrt = q.nrt; v.tpp = q.tpp; v.kind = q.kind;
  prm.iters = iters; prm.B = B; prm.C = C;
  prm.SC = (flags & DPFT_SIGMA_BROADCAST) ? 1 : C;
  prm.SCm = prm.SC; prm.mism = (prm.SC == C) ? tun.mism : nullptr; prm.rep_role = 0;
  prm.group = G.group; prm.n_groups = G.n_groups; prm.n_mm_groups = G.n_mm_groups;
  prm.kf_shared = (flags & DPFT_SHARED_KEYFRAME) ? 1 : 0;
  prm.total_items = (unsigned)q.total_items;
  prm.pose_hist = pose_hist; prm.sys_hist = sys_hist;
  prm.aux = aux_hist ? aux_hist : (float*)(ws + q.off_aux);
  prm.records = (float*)(ws + q.off_records);
  prm.pairrec = (double*)(ws + q.off_pairrec);
  prm.fifo = (unsigned long long*)(ws + q.off_fifo);
  prm.qctl = (unsigned*)(ws + q.off_qctl);
  prm.tiles_done = (int*)(ws + q.off_tiles_done);
  prm.cand = (int*)(ws + q.off_cand);
  prm.pairs_done = (int*)(ws + q.off_pairs_done);
  prm.groups_done = (int*)(ws + q.off_groups_done);
  prm.gext = (uint32_t*)(ws + q.off_gext);
  uint32_t* mm = (uint32_t*)(ws + q.off_mm);
  prm.s0mm = s0mm_ready ? s0mm_ready : mm;
  prm.status = status;
  prm.t_done = launch_ms ? (unsigned long long*)(ws + q.off_tdone) : nullptr;
  const bool tru = flags & DPFT_REMOVE_TRU_SIGMA;
  if (tru && !s0mm_ready) {
    init_kernel<<<1, 256, 0, stream>>>(nullptr, nullptr, 0, nullptr, 0, mm, std::min(G.n_mm_groups, 256));
    if (G.n_mm_groups > 256) init_kernel<<<(G.n_mm_groups + 255) / 256, 256, 0, stream>>>(nullptr, nullptr, 0, nullptr, 0, mm, G.n_mm_groups);
    const float* src[1] = {L.sigma0};
    const size_t per_pair = (size_t)prm.SC * L.H * L.W;
    const size_t per_group[1] = {G.n_mm_groups > 1 ? per_pair * G.group : per_pair * (prm.kf_shared ? 1 : B)};
    const unsigned plane1[1] = {(unsigned)(L.H * L.W)};
    launch_minmax_levels(src, per_group, 1, G.n_mm_groups, mm, stream, plane1, prm.SC, prm.mism);
  }
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (kernel_ms) {
    cudaEventCreate(&ev[0]);
    cudaEventCreate(&ev[1]);
  }
  cudaError_t err = launch_queue(prm, pose_in, tru, q.grid, stream, !tun.generic_geometry, ev[0], ev[1]);
  if (err != cudaSuccess) return set_error((int)err, "work-queue launch: %s", cudaGetErrorString(err));
  if (kernel_ms) {
    err = cudaEventSynchronize(ev[1]);
    if (err == cudaSuccess) err = cudaEventElapsedTime(kernel_ms, ev[0], ev[1]);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (err != cudaSuccess) return set_error((int)err, "kernel timing: %s", cudaGetErrorString(err));
  }
  if (launch_ms) {
    unsigned long long stamps[64 + 1];
    err = cudaMemcpyAsync(stamps, prm.t_done, (iters + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
    if (err != cudaSuccess) return set_error((int)err, "stamp read-back: %s", cudaGetErrorString(err));
    for (int i = 0; i < iters; ++i) launch_ms[i] = (float)((double)(stamps[i + 1] - stamps[i]) * 1e-6);
  }
  return 0;
}

// Workspace of a call: [launch-per-iteration plan of the levels it serves | work-queue plan of the finest level | the
// flag of sigma_replication_kernel].
constexpr size_t kRepFlagBytes = 256;
// ... and behind the flag the sigma0 extremes that kernel collects on the way: (level, group, 2) order-encoded floats
static size_t rep_mm_bytes(int n_levels, int n_mm_groups) {
  return (((size_t)n_levels * n_mm_groups * 2 * sizeof(uint32_t)) + 255) & ~(size_t)255;
}
static size_t lpi_bytes(const dpft_level_t* levels, int n_levels, int B, int C, uint32_t flags, bool any_occ, const Tuning& tun) {
  if (n_levels < 1) return 0;
  const Groups G = groups_of(B, flags, tun);
  const int pg = (persistent_ok(flags, any_occ) && G.n_groups == 1) ? persistent_grid_cached(C, flags & DPFT_REMOVE_TRU_SIGMA) : 0;
  return make_plan(levels, n_levels, B, C, flags, any_occ, std::max(pg, 0), tun, G.n_groups, G.n_mm_groups).total;
}

extern "C" size_t dpft_uic_workspace_bytes_ex(const dpft_level_t* levels, int n_levels, int B, int C, int iters,
                                              uint32_t flags, const dpft_uic_options_t* opt) {
  if (check_args(levels, n_levels, B, C, iters, flags)) return 0;
  const Tuning tun = tuning_of(opt);
  if (check_group(levels, n_levels, B, flags, tun)) return 0;
  bool any_occ = false;
  for (int l = 0; l < n_levels; ++l) any_occ = any_occ || levels[l].occ_out;
  if (queue_wanted(C, iters, flags, any_occ)) {
    const int nc = n_levels - std::min(n_levels, std::max(1, tun.queue_levels));
    size_t q = 0;                                     // the queue launches run one after the other: one region
    for (int l = nc; l < n_levels; ++l) q = std::max(q, make_qplan(levels[l], l, B, C, iters, flags, tun).total);
    return lpi_bytes(levels, nc, B, C, flags, any_occ, tun) + q + kRepFlagBytes + rep_mm_bytes(n_levels, groups_of(B, flags, tun).n_mm_groups);
  }
  return lpi_bytes(levels, n_levels, B, C, flags, any_occ, tun) + kRepFlagBytes + rep_mm_bytes(n_levels, groups_of(B, flags, tun).n_mm_groups);
}

extern "C" size_t dpft_uic_workspace_bytes(const dpft_level_t* levels, int n_levels, int B, int C, int iters,
                                           uint32_t flags) {
  return dpft_uic_workspace_bytes_ex(levels, n_levels, B, C, iters, flags, nullptr);
}

// Launch-per-iteration kernels (or the single cooperative launch) over `n_levels` levels; arguments already checked.
// `n_mm_levels` >= n_levels: levels whose sigma0 extremes are computed here (the ones past n_levels belong to work-queue
// launches that follow; `mm_out` receives where they are).
static int run_lpi(const dpft_level_t* levels, int n_levels, int B, int C, int iters, uint32_t flags,
                   float w_icp, const float* pose_in, float* pose_hist, float* sys_hist, float* aux_hist,
                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream_, cudaEvent_t* ev,
                   unsigned long long* clock_host, const Tuning& tun, int n_mm_levels = 0,
                   const uint32_t** mm_out = nullptr) {
  n_mm_levels = std::max(n_mm_levels, n_levels);
  bool any_occ = false;
  for (int l = 0; l < n_levels; ++l) any_occ = any_occ || levels[l].occ_out;
  const Groups G = groups_of(B, flags, tun);
  if (ev && !clock_host) flags |= DPFT_LAUNCH_PER_ITERATION;   // event timing needs separate launches
  const bool persist = persistent_ok(flags, any_occ) && G.n_groups == 1 &&
                       persistent_grid_cached(C, flags & DPFT_REMOVE_TRU_SIGMA) > 0;
  const Plan pl = make_plan(levels, n_levels, B, C, flags, any_occ,
                            persist ? persistent_grid_cached(C, flags & DPFT_REMOVE_TRU_SIGMA) : 0, tun, G.n_groups,
                            G.n_mm_groups);
  if (workspace_bytes < pl.total) return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, pl.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  float* partials = (float*)(ws + pl.off_partials);
  double* pairrec = (double*)(ws + pl.off_pairrec);
  int* counters = (int*)(ws + pl.off_counters);
  // (sigma0 extremes: already collected by the replication check of this call, or found below)
  uint32_t* mm = tun.mm_ready ? tun.mm_ready : (uint32_t*)(ws + pl.off_mm);
  float* gmm = aux_hist ? aux_hist : (float*)(ws + pl.off_gmm);   // 4 floats per iteration and group
  float* sr0 = (float*)(ws + pl.off_sr0);
  const bool tru = flags & DPFT_REMOVE_TRU_SIGMA;
  const bool pdl = !(flags & DPFT_NO_PDL);
  const bool fused = flags & DPFT_FUSED_SOBEL;
  const bool icp = flags & DPFT_COMBINE_ICP;
  const int SC = (flags & DPFT_SIGMA_BROADCAST) ? 1 : C;     // channels of the sigma maps
  float* grad = (float*)(ws + pl.off_grad);
  float* vn = (float*)(ws + pl.off_vn);
  float* icp_rec = (float*)(ws + pl.off_icp);
  uint32_t* dmm = (uint32_t*)(ws + pl.off_dmm);

  if (mm_out) *mm_out = mm;
  {
    const int n_mm = tun.mm_ready ? 0 : n_mm_levels * G.n_mm_groups;
    const int n = std::max(std::max(B * 12, B + G.n_groups), n_mm);
    init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pose_in, pose_hist, pose_in == pose_hist ? 0 : B * 12, counters,
                                                     B + G.n_groups, mm, n_mm);
    if (icp) init_kernel<<<1, 32, 0, stream>>>(pose_in, pose_hist, 0, counters, 0, dmm, n_levels);
  }
  const bool shared_kf = flags & DPFT_SHARED_KEYFRAME, pairwise = G.group == 1;
  const int mm_per_level = G.n_mm_groups;
  if (tru && !tun.mm_ready) {
    // extremes of sigma0 per level and group, all levels in one launch
    const float* src[DPFT_MAX_LEVELS];
    size_t per_group[DPFT_MAX_LEVELS];
    unsigned planes[DPFT_MAX_LEVELS];
    for (int l = 0; l < n_mm_levels; ++l) {
      const size_t per_pair = (size_t)SC * levels[l].H * levels[l].W;
      src[l] = levels[l].sigma0;
      per_group[l] = G.n_mm_groups > 1 ? per_pair * G.group : per_pair * (shared_kf ? 1 : B);
      planes[l] = (unsigned)(levels[l].H * levels[l].W);
    }
    // (full sigma tensors found to be C copies of one map: the extremes of channel 0 are the tensor's)
    launch_minmax_levels(src, per_group, n_mm_levels, G.n_mm_groups, mm, stream, planes, SC, (fused && SC == C) ? tun.mism : nullptr);
  }
  if (persist) {
    PersistParams pp{};
    for (int l = 0; l < n_levels; ++l) {
      const dpft_level_t& L = levels[l];
      PLevel& q = pp.lv[l];
      q.x0 = L.x0; q.x1 = L.x1; q.s0 = L.sigma0; q.s1 = L.sigma1; q.d0 = L.invd0; q.d1 = L.invd1; q.K = L.K;
      q.m0 = L.obj_mask0; q.m1 = L.obj_mask1;
      q.H = L.H; q.W = L.W; q.nseg = pl.nseg[l]; q.TR = pl.p_TR[l]; q.nrt = pl.p_nrt[l]; q.tpp = q.nseg * q.nrt;
    }
    pp.n_levels = n_levels; pp.iters = iters; pp.B = B; pp.C = C; pp.rcap = pl.p_rcap;
    pp.pose_hist = pose_hist; pp.sys_hist = sys_hist; pp.aux = gmm;
    pp.records = (float*)(ws + pl.off_records); pp.pairrec = pairrec; pp.s0mm = mm;
    pp.gext = (uint32_t*)(ws + pl.off_gext); pp.status = status;
    pp.clock_out = clock_host ? (unsigned long long*)(ws + pl.off_clock) : nullptr;
    init_kernel<<<(n_levels * iters + 255) / 256, 256, 0, stream>>>(pose_in, pose_hist, 0, counters, 0, pp.gext, n_levels * iters);
    const cudaError_t err = launch_persistent(pp, pl.p_grid, tru, stream);
    if (err != cudaSuccess) return set_error((int)err, "cooperative launch: %s", cudaGetErrorString(err));
    if (clock_host) {
      const int n = n_levels * iters + 1;
      cudaError_t e2 = cudaMemcpyAsync(clock_host, pp.clock_out, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
      if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(stream);
      if (e2 != cudaSuccess) return set_error((int)e2, "clock read-back: %s", cudaGetErrorString(e2));
    }
    return 0;
  }
  int k = 0;
  for (int l = 0; l < n_levels; ++l) {
    const dpft_level_t& L = levels[l];
    const size_t plane = (size_t)L.H * L.W;
    if (icp) {
      // vertex and normal maps of the live frame; normals vanish on the batch-global depth extremes
      const size_t n = (size_t)B * plane;
      const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, 148 * 8);
      minmax_kernel<<<blocks, 256, 0, stream>>>(L.depth1, n, dmm + 2 * l);
      launch_vertex_normal(L.depth1, L.K, dmm + 2 * l, vn, vn + 3 * (size_t)B * plane, B, L.H, L.W, stream);
    }
    PxExtra ex{};
    if (!fused) {
      // unit Sobel gradients of this level's keyframe maps, once for all its iterations
      float* g = grad;
      ex.gfx = g; ex.gfy = g + pl.grad_elems; ex.gsx = g + 2 * pl.grad_elems; ex.gsy = g + 3 * pl.grad_elems;
      ex.ppt = pl.ppt[l];
      launch_sobel_unit(L.x0, g, g + pl.grad_elems, B * C, L.H, L.W, stream);
      launch_sobel_unit(L.sigma0, g + 2 * pl.grad_elems, g + 3 * pl.grad_elems, B * C, L.H, L.W, stream);
    }
#ifndef DPFT_MAX_CH
#define DPFT_MAX_CH 8
#endif
    const int CH = (C % 8 == 0 && DPFT_MAX_CH >= 8) ? 8 : (C % 4 == 0) ? 4 : (C % 2 == 0) ? 2 : 1;
    for (int it = 0; it < iters; ++it, ++k) {
      UicIterParams prm{};
      prm.x0 = L.x0; prm.x1 = L.x1; prm.s0 = L.sigma0; prm.s1 = L.sigma1;
      prm.d0 = L.invd0; prm.d1 = L.invd1; prm.K = L.K;
      prm.m0 = L.obj_mask0; prm.m1 = L.obj_mask1;
      prm.occ_out = L.occ_out ? L.occ_out + (size_t)it * B * plane : nullptr;
      prm.sr0_dbg = sr0;
      prm.H = L.H; prm.W = L.W; prm.B = B; prm.C = C; prm.SC = SC;
      prm.SCm = SC; prm.mism = (fused && SC == C) ? tun.mism : nullptr; prm.rep_role = 0;
      prm.nseg = pl.nseg[l]; prm.nrt = pl.nrt[l]; prm.TR = pl.TR[l];
      prm.ctas_per_pair = fused ? pl.ctas[l] : pl.px_ctas[l];
      prm.tab = pl.tab[l];
      prm.pose = pose_hist + (size_t)k * B * 12;
      prm.pose_next = pose_hist + (size_t)(k + 1) * B * 12;
      prm.sys_out = sys_hist + (size_t)k * B * 27;
      prm.partials = partials; prm.pairrec = pairrec; prm.counters = counters;
      prm.s0mm = mm + 2 * (size_t)l * mm_per_level; prm.gmm = gmm + 4 * (size_t)k * G.n_groups;
      prm.kf_shared = shared_kf; prm.pairwise = pairwise; prm.status = status; prm.flags = flags;
      prm.group = G.group; prm.n_mm_groups = G.n_mm_groups;
      if (icp) {
        const float* wmap = tun.icp_weight[l];      // a learned scaler's per-pixel map replaces the scalar weight
        launch_icp_term(L.depth0, L.K, vn, vn + 3 * (size_t)B * plane, prm.pose, L.obj_mask0, L.obj_mask1, icp_rec,
                        nullptr, nullptr, wmap, B, L.H, L.W, stream, ws + pl.off_icp_scratch);
        prm.icp_rec = icp_rec;
        prm.icp_w2 = wmap ? 1.f : w_icp * w_icp;
      }
      const dim3 grid(prm.ctas_per_pair, B);
      // the debug mask pass reads what this launch wrote, so keep plain stream order around it
      const bool use_pdl = pdl && !any_occ && !ev;
      cudaError_t err;
      if (ev) cudaEventRecord(ev[k], stream);
      if (!fused) {
        switch (CH) {
          case 8: err = launch_px<8>(prm, ex, grid, tru, use_pdl, stream); break;
          case 4: err = launch_px<4>(prm, ex, grid, tru, use_pdl, stream); break;
          case 2: err = launch_px<2>(prm, ex, grid, tru, use_pdl, stream); break;
          default: err = launch_px<1>(prm, ex, grid, tru, use_pdl, stream); break;
        }
      } else if ((flags & DPFT_STAGED_FOOTPRINT) && staged_ok(L, C)) {
        if (prm.mism) {
          // full sigma tensors, replication decided on the device: the one-map twin, then the C-map twin
          UicIterParams one = prm;
          one.SC = 1; one.rep_role = 1;
          err = launch_staged(one, grid, tru, use_pdl, stream, tun);
          prm.rep_role = 2;
          if (err == cudaSuccess) err = launch_staged(prm, grid, tru, use_pdl, stream, tun);
        } else {
          err = launch_staged(prm, grid, tru, use_pdl, stream, tun);
        }
      } else if (prm.mism && pl.res_smem[l] > 0 && CH == 8 && resident_ctas(resident_smem(L, C, 1)) > resident_ctas(pl.res_smem[l])) {
        // resident level, replication decided on the device: the one-map twin holds C + 2 instead of 2 C + 1 planes
        // of the live frame, then the C-map twin; one of them returns at once.  Only where the smaller copy lets more
        // CTAs share an SM (30x40: 4 against 2, 476 against 628 us per level at 20 batches); a 15x20 level would
        // only pay the extra launches (+5 us each)
        UicIterParams one = prm;
        one.SC = 1; one.rep_role = 1;
        err = launch_iter<8>(one, grid, tru, use_pdl, stream, tun, resident_smem(L, C, 1));
        prm.rep_role = 2;
        if (err == cudaSuccess) err = launch_iter<8>(prm, grid, tru, use_pdl, stream, tun, pl.res_smem[l]);
      } else
      switch (CH) {
        case 8: err = launch_iter<8>(prm, grid, tru, use_pdl, stream, tun, pl.res_smem[l]); break;
        case 4: err = launch_iter<4>(prm, grid, tru, use_pdl, stream, tun, pl.res_smem[l]); break;
        case 2: err = launch_iter<2>(prm, grid, tru, use_pdl, stream, tun, pl.res_smem[l]); break;
        default: err = launch_iter<1>(prm, grid, tru, use_pdl, stream, tun, pl.res_smem[l]); break;
      }
      if (err != cudaSuccess) return set_error((int)err, "uic_iter_kernel launch: %s", cudaGetErrorString(err));
      if (tru && prm.occ_out) {
        const size_t n = (size_t)B * plane;
        occ_fixup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(prm.occ_out, sr0, prm.gmm, n);
      }
    }
  }
  if (ev) cudaEventRecord(ev[k], stream);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_error((int)err, "launch: %s", cudaGetErrorString(err));
  return 0;
}

// launch-per-iteration levels with the device time of every iteration (events, or stamps of the cooperative launch)
static int lpi_timed(const dpft_level_t* levels, int n_levels, int B, int C, int iters, uint32_t flags, float w_icp,
                     const float* pose_in, float* pose_hist, float* sys_hist, float* aux_hist, int32_t* status,
                     void* workspace, size_t workspace_bytes, void* stream, const Tuning& tun, float* launch_ms,
                     int n_mm_levels = 0, const uint32_t** mm_out = nullptr) {
  const int n = n_levels * iters;
  bool any_occ = false;
  for (int l = 0; l < n_levels; ++l) any_occ = any_occ || levels[l].occ_out;
  if (persistent_ok(flags, any_occ) && groups_of(B, flags, tun).n_groups == 1 &&
      persistent_grid_cached(C, flags & DPFT_REMOVE_TRU_SIGMA) > 0) {
    // single cooperative launch: iteration boundaries are stamped on the device with %globaltimer
    unsigned long long stamps[DPFT_MAX_LEVELS * 64 + 1];
    const int rc = run_lpi(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                           workspace, workspace_bytes, stream, nullptr, stamps, tun, n_mm_levels, mm_out);
    for (int i = 0; i < n && rc == 0; ++i) launch_ms[i] = (float)((double)(stamps[i + 1] - stamps[i]) * 1e-6);
    return rc;
  }
  cudaEvent_t ev[DPFT_MAX_LEVELS * 64 + 1];
  for (int i = 0; i <= n; ++i) cudaEventCreate(&ev[i]);
  int rc = run_lpi(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                   workspace, workspace_bytes, stream, ev, nullptr, tun, n_mm_levels, mm_out);
  if (rc == 0) {
    const cudaError_t err = cudaStreamSynchronize((cudaStream_t)stream);
    if (err != cudaSuccess) rc = set_error((int)err, "sync: %s", cudaGetErrorString(err));
  }
  for (int i = 0; i < n && rc == 0; ++i) cudaEventElapsedTime(&launch_ms[i], ev[i], ev[i + 1]);
  for (int i = 0; i <= n; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

// Every dpft_uic_forward* entry point ends here: checks, then the levels either all through the launch-per-iteration
// kernels or -- DPFT_QUEUE on a problem that qualifies -- the coarse ones through them and the finest through the
// work queue.
static int run_uic(const dpft_level_t* levels, int n_levels, int B, int C, int iters, uint32_t flags, float w_icp,
                   const float* pose_in, float* pose_hist, float* sys_hist, float* aux_hist, int32_t* status,
                   void* workspace, size_t workspace_bytes, void* stream, const Tuning& tun_in) {
  Tuning tun = tun_in;
  tun.mism = nullptr;
  if (iters > 64) return set_error(DPFT_EINVAL, "iters must be <= 64");
  if (int e = check_args(levels, n_levels, B, C, iters, flags)) return e;
  if (int e = check_group(levels, n_levels, B, flags, tun)) return e;
  if (!pose_in || !pose_hist || !status || (iters > 0 && !sys_hist) || !workspace)
    return set_error(DPFT_EINVAL, "pose_in, pose_hist, sys_hist, status and workspace are required");
  if (tun.launch_ms && iters < 1) return set_error(DPFT_EINVAL, "timing needs at least one iteration");
  bool any_occ = false;
  for (int l = 0; l < n_levels; ++l) any_occ = any_occ || levels[l].occ_out;
  float* ms = tun.launch_ms;
  const bool use_queue = queue_wanted(C, iters, flags, any_occ);
  if ((flags & DPFT_FUSED_SOBEL) && ((flags & DPFT_LAUNCH_PER_ITERATION) || use_queue) && C > 1 && iters >= 1 &&
      !(flags & (DPFT_SIGMA_BROADCAST | DPFT_COMBINE_ICP)) && !any_occ && !tun.no_sigma_detect) {
    // where the flag lives: behind everything else the call uses
    size_t used;
    if (use_queue) {
      const int nc0 = n_levels - std::min(n_levels, std::max(1, tun.queue_levels));
      size_t q = 0;
      for (int l = nc0; l < n_levels; ++l) q = std::max(q, make_qplan(levels[l], l, B, C, iters, flags, tun).total);
      used = lpi_bytes(levels, nc0, B, C, flags, any_occ, tun) + q;
    } else {
      used = lpi_bytes(levels, n_levels, B, C, flags, any_occ, tun);
    }
    int* mism = (int*)((char*)workspace + used);
    const Groups Gd = groups_of(B, flags, tun);
    // behind the flag: the sigma0 extremes the check collects on the way, (level, group, 2) -- what launch_minmax_levels
    // would otherwise compute in a pass of its own
    const bool fuse_mm = (flags & DPFT_REMOVE_TRU_SIGMA) != 0;
    uint32_t* mm_fused = (uint32_t*)((char*)workspace + used + kRepFlagBytes);
    const size_t mm_bytes = rep_mm_bytes(n_levels, Gd.n_mm_groups);
    if (workspace_bytes < used + kRepFlagBytes + mm_bytes)
      return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, used + kRepFlagBytes + mm_bytes);
    RepTensors rt{};
    const size_t Bk = (flags & DPFT_SHARED_KEYFRAME) ? 1 : (size_t)B;
    unsigned long long units = 0;
    for (int l = 0; l < n_levels; ++l) {
      const unsigned plane = (unsigned)(levels[l].H * levels[l].W), cpp = (plane + 3) / 4;
      const float* tp[2] = {levels[l].sigma0, levels[l].sigma1};
      const size_t np[2] = {Bk, (size_t)B};
      for (int k = 0; k < 2; ++k) {
        rt.p[rt.n] = tp[k]; rt.plane[rt.n] = plane; rt.units[rt.n] = (unsigned)(np[k] * cpp);
        rt.first[rt.n] = (unsigned)units;
        // sigma0: one slot per sigma-extreme group (one for a shared keyframe or a single batch)
        rt.mm_base[rt.n] = (k == 0 && fuse_mm) ? l * Gd.n_mm_groups : -1;
        rt.upg[rt.n] = (unsigned)((Gd.n_mm_groups > 1 ? (size_t)Gd.group : np[k]) * cpp);
        units += np[k] * cpp;
        ++rt.n;
      }
    }
    rt.first[rt.n] = (unsigned)units;
    // (larger calls simply go without the check; the margin keeps the kernel's 32-bit share bounds, begin + ceil(units / CTAs),
    // from wrapping)
    if (units < (1ull << 32) - (1ull << 16)) {
      // flag <- 0, extremes <- (+inf, -inf) in their order encoding: one small launch
      const int n_mm = fuse_mm ? n_levels * Gd.n_mm_groups : 0;
      init_kernel<<<(std::max(n_mm, 1) + 255) / 256, 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, 0, mism, 1, mm_fused, n_mm);
      const unsigned gx = (unsigned)std::min<unsigned long long>((units + 255) / 256, 148ull * 16);
      if (C == 8) sigma_replication_kernel<8><<<gx, 256, 0, (cudaStream_t)stream>>>(rt, C, mism, fuse_mm ? mm_fused : nullptr);
      else sigma_replication_kernel<0><<<gx, 256, 0, (cudaStream_t)stream>>>(rt, C, mism, fuse_mm ? mm_fused : nullptr);
      tun.mism = mism;
      tun.mm_ready = fuse_mm ? mm_fused : nullptr;
    }
  }
  if (!use_queue) {
    if (ms)
      return lpi_timed(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                       workspace, workspace_bytes, stream, tun, ms);
    return run_lpi(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status, workspace,
                   workspace_bytes, stream, nullptr, nullptr, tun);
  }
  const int nc = n_levels - std::min(n_levels, std::max(1, tun.queue_levels));   // coarse levels
  const Groups G = groups_of(B, flags, tun);
  const size_t coarse_bytes = lpi_bytes(levels, nc, B, C, flags, any_occ, tun);
  size_t qbytes = 0;
  for (int l = nc; l < n_levels; ++l) qbytes = std::max(qbytes, make_qplan(levels[l], l, B, C, iters, flags, tun).total);
  if (workspace_bytes < coarse_bytes + qbytes)
    return set_error(DPFT_ENOSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, coarse_bytes + qbytes);
  // the coarse levels' launch computes the sigma0 extremes of EVERY level in its one reduction launch (unless the
  // replication check already collected them)
  const uint32_t* mm_all = tun.mm_ready;
  if (nc > 0) {
    const int rc = ms ? lpi_timed(levels, nc, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                                  workspace, coarse_bytes, stream, tun, ms, n_levels, &mm_all)
                      : run_lpi(levels, nc, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                                workspace, coarse_bytes, stream, nullptr, nullptr, tun, n_levels, &mm_all);
    if (rc) return rc;
  }
  for (int l = nc; l < n_levels; ++l) {
    const size_t k0 = (size_t)l * iters;
    float* ph = pose_hist + k0 * B * 12;
    const bool tru = flags & DPFT_REMOVE_TRU_SIGMA;
    const int rc = run_queue(levels[l], l, B, C, iters, flags, l > 0 ? ph : pose_in, ph, sys_hist + k0 * B * 27,
                             aux_hist ? aux_hist + k0 * G.n_groups * 4 : nullptr, status, (char*)workspace + coarse_bytes,
                             qbytes, (cudaStream_t)stream, tun, ms ? ms + k0 : nullptr,
                             (tru && mm_all) ? mm_all + 2 * (size_t)l * G.n_mm_groups : nullptr,
                             tun.queue_kernel_ms ? tun.queue_kernel_ms + (l - nc) : nullptr);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int dpft_uic_forward_ex(const dpft_level_t* levels, int n_levels, int B, int C, int iters, uint32_t flags,
                                   float w_icp, const float* pose_in, float* pose_hist, float* sys_hist,
                                   float* aux_hist, int32_t* status, void* workspace, size_t workspace_bytes,
                                   void* stream, const dpft_uic_options_t* opt) {
  return run_uic(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                 workspace, workspace_bytes, stream, tuning_of(opt));
}

extern "C" int dpft_uic_forward(const dpft_level_t* levels, int n_levels, int B, int C, int iters, uint32_t flags,
                                float w_icp, const float* pose_in, float* pose_hist, float* sys_hist,
                                float* aux_hist, int32_t* status, void* workspace, size_t workspace_bytes,
                                void* stream) {
  return dpft_uic_forward_ex(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status,
                             workspace, workspace_bytes, stream, nullptr);
}

extern "C" int dpft_uic_forward_timed(const dpft_level_t* levels, int n_levels, int B, int C, int iters,
                                      uint32_t flags, float w_icp, const float* pose_in, float* pose_hist,
                                      float* sys_hist, float* aux_hist, int32_t* status, void* workspace,
                                      size_t workspace_bytes, void* stream, float* launch_ms) {
  if (!launch_ms) return set_error(DPFT_EINVAL, "launch_ms is required");
  Tuning tun;
  tun.launch_ms = launch_ms;
  return run_uic(levels, n_levels, B, C, iters, flags, w_icp, pose_in, pose_hist, sys_hist, aux_hist, status, workspace,
                 workspace_bytes, stream, tun);
}

// Rows per work-queue tile the planner picks for a level of H x W with B pairs in the call (test hook, not part of the ABI;
// `workers` <= 0: what the device -- or, without one, a B200 -- holds of the one-map routine).
extern "C" int dpft_debug_queue_tile_rows(int H, int W, int B, long workers) {
  if (H < 1 || W < 1 || B < 1) return 0;
  if (workers <= 0) workers = (long)device_sms() * queue_tiles_per_sm(true);
  return queue_tile_rows(H, (W + kCols - 1) / kCols, B, workers);
}

// Host-side view of the balanced tile table of one level (test hook, not part of the ABI): returns 1 and fills
// `tiles` ([2 kinds][max_warps][2 sub-tiles][seg, y0, y1]), `ctas` (CTAs per pair of kind 0 / 1), `n_more` (pairs of
// kind 1), `nseg` and `warps_per_cta` when a table is used for (H, W, B), 0 when the rectangular tiling stays.
extern "C" int dpft_debug_tile_table(int H, int W, int B, int linear, int max_warps, int* tiles, int* ctas,
                                     int* n_more, int* nseg, int* warps_per_cta) {
  TileTab tab;
  const int ns = (W + kStagedCols - 1) / kStagedCols;
  const Tuning tun;
  const bool on = linear ? make_tile_tab_linear(H, ns, B, tab, tun) : make_tile_tab(H, ns, B, tab, tun);
  *nseg = ns;
  *warps_per_cta = kSW;
  if (!on) return 0;
  ctas[0] = tab.ctas[0];
  ctas[1] = tab.ctas[1];
  *n_more = tab.n_more;
  for (int kind = 0; kind < 2; ++kind)
    for (int w = 0; w < max_warps && w < kTabWarps; ++w)
      for (int k = 0; k < 2; ++k) {
        int* t = tiles + ((kind * max_warps + w) * 2 + k) * 3;
        t[0] = tab.seg[kind][w][k];
        t[1] = tab.y0[kind][w][k];
        t[2] = tab.y1[kind][w][k];
      }
  return 1;
}

#ifdef DPFT_DEBUG_STAMPS
extern "C" int dpft_debug_read_phases(unsigned long long* host, int n_warps) {
  return (int)cudaMemcpyFromSymbol(host, dpft::g_wtl2, sizeof(unsigned long long) * 4 * (size_t)n_warps);
}
extern "C" int dpft_debug_read_timeline(unsigned long long* host, int n_warps) {
  return (int)cudaMemcpyFromSymbol(host, dpft::g_wtl, sizeof(unsigned long long) * 4 * (size_t)n_warps);
}
extern "C" int dpft_debug_read_stamps(unsigned long long* host16) {
  return (int)cudaMemcpyFromSymbol(host16, dpft::g_stamps, sizeof(unsigned long long) * 16);
}
#endif
