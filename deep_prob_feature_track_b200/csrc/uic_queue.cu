// The finest pyramid level of the U_IC solve as ONE launch driven by a work queue: all Gauss-Newton iterations of all
// frame pairs of the call, with the dependencies of the algorithm expressed per PAIR instead of per launch.
//
// Why.  With one launch per iteration (uic_forward.cu) every iteration ends in a drain: the launch is as long as its
// slowest warp, then a chain of hand-offs (CTA sums -> pair sums -> batch extremes -> the solves) runs on a nearly idle
// machine, then the next launch starts.  The in-kernel timeline of round 1 put a third of a finest-level launch into
// that drain (profiles/r1h_timeline_*.txt).  But the algorithm only orders the iterations of ONE pair (iteration k+1
// of a pair needs the pose iteration k gave it); pairs do not wait for each other -- with one exception, the
// batch-global sigma extremes of remove_tru_sigma (algorithms.py:1976-1979), handled below.
//
// How.  A unit of work is a warp tile (30 columns x TR rows of one pair at one iteration), walked by the same tile
// routines as the launch-per-iteration kernels (uic_tile.cuh / uic_tile_staged.cuh).  Warps are workers: a worker
// takes the next slot of a FIFO in global memory (one atomicAdd), waits until the slot holds an item, walks the tile
// and writes one record.  The worker that completes the LAST tile of a pair-iteration folds the pair's records in tile
// order (fp64, deterministic whoever ran which tile), damps and solves the 6x6 system, writes the pose of iteration
// k+1 and appends that iteration's tiles to the FIFO.  Nothing waits for a launch boundary; pairs drift apart and the
// drain of one is hidden behind the tiles of the others.  It pays when the call holds more work than one wave of
// tiles: several independent batches ("groups") per call, or pairs that nothing couples.  Small calls keep the coarse
// levels on the launch-per-iteration kernels: their iterations are latency chains (measured, profiles/r2/) that a whole
// grid walks faster in lockstep than pair by pair.  Large calls give the second- and third-finest level a work-queue
// launch of their own (algorithms.default_queue_levels; levels narrower than the staged routine's ring run its narrow
// form, tile routine kind 2), and tiles grow to whole columns once an iteration offers 2.5 waves of them
// (queue_tile_rows, uic_forward.cu).
//
// Batch-global sigma extremes without a barrier.  A pixel is masked when its warped sigma equals the minimum or the
// maximum over the whole batch (group).  The mask only moves J^T r, and every record carries what its extreme pixels
// added (uic_forward.cu: finalize_pair), so a pair only has to know whether ITS extreme is the group's.  When a pair
// is folded it merges its extremes into the group's running extremes with atomicMin / atomicMax: if the running value
// was already beyond the pair's, the pair can never be the group's extreme (the running value only moves outwards)
// and it is solved at once.  Otherwise it is a candidate: its sums are parked and the LAST pair of the group to be
// folded -- which sees the final extremes -- solves the candidates.  Typically two pairs of 64 wait; with saturated
// sigma maps (ties) all of them do, which degrades to the barrier the reference has, never to a wrong mask.
//
// Progress.  Slots are claimed in order and an item is only ever waited for by a worker that holds no other item, so
// every claimed item is being walked by a resident warp and the item a waiting worker needs is produced by them: no
// co-residency guarantee (cooperative launch) is needed.  A wait that exceeds kWaitLimitNs traps instead of hanging.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_host.h"
#include "dpft_kernels.h"
#include "dpft_records.h"
#include "uic_reduce.cuh"
#include "uic_tile.cuh"
#include "uic_tile_staged.cuh"

namespace dpft {

constexpr int kQW = 4;                 // warps (workers) per CTA
#ifndef DPFT_QW_SB
#define DPFT_QW_SB 4                   // ... of the one-map staged routine (tuning hook)
#endif
__host__ __device__ constexpr int q_warps(bool sb, int kind) { return (sb && kind >= 1) ? DPFT_QW_SB : kQW; }
// Resident CTAs per SM: 3 (168 registers, 12 workers).  The one-map routine's 10-map ring would let a fourth CTA fit
// by shared memory, and DPFT_Q_CTAS_SB=4 / DPFT_QW_SB / DPFT_Q_MAXNREG_SB build such variants -- all measured SLOWER
// (profiles/r2/r2c_occupancy.txt): the register file is 16 K per scheduler, so between 12 and 16 workers per SM there is
// nothing (144 registers still hold 3 warps per scheduler), and at 128 registers the row body spills ~30 accesses per
// row whose lines no longer fit the 28-60 KB of L1 that 16 workers' shared memory leaves (3.60 against 2.66 ms).
#ifndef DPFT_Q_CTAS_SB
#define DPFT_Q_CTAS_SB 3
#endif
constexpr int kQCtasPerSm = 3;
__host__ __device__ constexpr int q_ctas_per_sm(bool sb, int kind) { return (sb && kind >= 1) ? DPFT_Q_CTAS_SB : kQCtasPerSm; }
// per-worker shared memory: the staged routine's area (ring | corrections | outlier taps | halo sums), then the pose
__host__ __device__ constexpr int q_area_floats(bool sb) { return stage_area_floats(sb) + 32; }
constexpr unsigned long long kWaitLimitNs = 4000000000ull;   // a worker that waits this long for an item traps
constexpr unsigned long long kItemValid = 1ull << 63;

static_assert(stage_warp_floats(true) - 27 * 33 >= 2 * PS, "the fold's fp64 scratch sits in front of the reduction rows");
static_assert(DPFT_QW_SB * q_area_floats(true) * 4 * DPFT_Q_CTAS_SB + 1024 * DPFT_Q_CTAS_SB <= 227 * 1024, "one-map CTAs per SM by shared memory");

__device__ __forceinline__ unsigned long long q_encode(int k, int b, int t) {
  return kItemValid | ((unsigned long long)k << 44) | ((unsigned long long)b << 20) | (unsigned long long)t;
}
__device__ __forceinline__ int q_item_k(unsigned long long it) { return (int)((it >> 44) & 0x7ffffu); }
__device__ __forceinline__ int q_item_b(unsigned long long it) { return (int)((it >> 20) & 0xffffffu); }
__device__ __forceinline__ int q_item_t(unsigned long long it) { return (int)(it & 0xfffffu); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* q) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(q) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* q, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(q), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* q, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(q), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long q_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Optional phase stamps of pair 0 (compile with -DDPFT_QUEUE_STAMPS; read back with dpft_debug_queue_stamps): per
// iteration k the %globaltimer at [0] first tile dequeued, [1] walk start / [2] walk end / [3] record flushed /
// [4] counted of the pair's LAST tile, [5] folded, [6] solved or parked, [7] next tiles pushed, [8] group resolved.
#ifdef DPFT_QUEUE_STAMPS
__device__ unsigned long long g_qstamps[512 * 16];
#define DPFT_QSTAMP(k, b, i) do { if ((b) == 0 && (threadIdx.x & 31) == 0 && (k) < 512) g_qstamps[(k) * 16 + (i)] = q_now_ns(); } while (0)
#define DPFT_QSTAMP_MIN(k, b, i) do { if ((b) == 0 && (threadIdx.x & 31) == 0 && (k) < 512) atomicMin(&g_qstamps[(k) * 16 + (i)], q_now_ns()); } while (0)
#else
#define DPFT_QSTAMP(k, b, i) do { } while (0)
#define DPFT_QSTAMP_MIN(k, b, i) do { } while (0)
#endif

// lane 0: wait until the slot holds an item (the producer may still be folding the pair that feeds it)
__device__ __forceinline__ unsigned long long q_wait_item(const unsigned long long* slot) {
  unsigned long long v = ld_acquire_u64(slot);
  if (v) return v;
  const unsigned long long t0 = q_now_ns();
  unsigned ns = 64;
  while (true) {
    __nanosleep(ns);
    v = ld_acquire_u64(slot);
    if (v) return v;
    if (ns < 1024) ns *= 2;
    if (q_now_ns() - t0 > kWaitLimitNs) __trap();   // a lost item must not hang the device
  }
}

// Damp, solve, update one pair (algorithms.py:2017-2054, 2094-2103) from its 39 folded sums.
template <bool TRU, bool GLOBAL>
__device__ __forceinline__ void q_finalize(const QueueParams& p, const int k, const int b, const double* rec,
                                           const bool at_min, const bool at_max) {
  auto ld = [&](int i) { return GLOBAL ? __ldcg(rec + i) : rec[i]; };
  double A[21], rhs[6];
#pragma unroll
  for (int i = 0; i < 21; ++i) A[i] = ld(i);
#pragma unroll
  for (int i = 0; i < 6; ++i) rhs[i] = ld(21 + i);
  if (TRU) {
    // pixels whose warped sigma (channel 0) sits on the group's extreme are masked: their weighted residual becomes
    // 1e-6, i.e. subtract what they added beyond that
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (at_min) rhs[i] -= ld(E_CMIN + i);
      if (at_max) rhs[i] -= ld(E_CMAX + i);
    }
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < 21; ++i) finite = finite && isfinite(A[i]);
#pragma unroll
  for (int i = 0; i < 6; ++i) finite = finite && isfinite(rhs[i]);
  float* sys = p.sys_hist + ((size_t)k * p.B + b) * 27;
#pragma unroll
  for (int i = 0; i < 21; ++i) sys[i] = (float)A[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) sys[21 + i] = (float)rhs[i];
  double xi[6];
  const bool ok = solve_and_update(A, rhs, true, p.pose_hist + ((size_t)k * p.B + b) * 12,
                                   p.pose_hist + ((size_t)(k + 1) * p.B + b) * 12, xi);
  int st = 0;
  if (!finite) st |= DPFT_ST_NONFINITE;
  if (!ok) st |= DPFT_ST_SINGULAR;
  if (st) atomicOr(p.status, st);
}

// Append the tiles of iteration k+1 of pair b (its pose has been written by this warp's lane 0).
__device__ __forceinline__ void q_push_next(const QueueParams& p, const int k, const int b, const int lane) {
  if (k + 1 >= p.iters) return;
  const int n = p.L.tpp;
  __threadfence();                       // the pose (lane 0) and everything before it, ahead of the items
  __syncwarp();
  unsigned pos = 0;
  if (lane == 0) pos = atomicAdd(p.qctl + 1, (unsigned)n);
  pos = __shfl_sync(0xffffffffu, pos, 0);
  // (the fence above orders the pose before the items; a release per store would fence 32 more times)
  for (int i = lane; i < n; i += 32) st_relaxed_u64(p.fifo + pos + i, q_encode(k + 1, b, i));
}

// One item: walk the tile, write its record, count it.  Returns true for the worker that completed the pair-iteration.
// The level lives in the kernel parameters (constant bank): nothing of it occupies a register across the row loop.
template <bool TRU, bool SB, bool AUX, int GW, int GH, int KIND>
__device__ __forceinline__ bool q_walk_item(const QueueParams& p, float* area, const unsigned long long item) {
  const int lane = threadIdx.x & 31;
  constexpr int WF = stage_warp_floats(SB && KIND >= 1);
  float (*redw)[33] = reinterpret_cast<float (*)[33]>(area + WF - 27 * 33);   // rows 27.. follow the ring
  float* spose = area + stage_area_floats(SB && KIND >= 1);
  const int k = q_item_k(item), b = q_item_b(item), t = q_item_t(item);
  const int B = p.B, C = p.C;
  const QLevel& L = p.L;
  const int plane = L.H * L.W;

  PairView g;
  {
    const size_t b0 = p.kf_shared ? 0 : (size_t)b;      // one keyframe for every pair (kf_vo-style tracking)
    g.x0 = L.x0 + b0 * C * plane; g.x1 = L.x1 + (size_t)b * C * plane;
    g.s0 = L.s0 + b0 * p.SCm * plane; g.s1 = L.s1 + (size_t)b * p.SCm * plane;
    g.scm = p.SCm;
    g.splane = (p.SC == C) ? (unsigned)plane : 0u;
    g.d0 = L.d0 + b0 * plane; g.d1 = L.d1 + (size_t)b * plane;
    g.m0 = (AUX && L.m0) ? L.m0 + b0 * plane : nullptr;
    g.m1 = (AUX && L.m1) ? L.m1 + (size_t)b * plane : nullptr;
    g.occ_out = nullptr; g.sr0_dbg = nullptr;
    g.H = L.H; g.W = L.W; g.C = C;
    g.fx = __ldg(L.K + 4 * b); g.fy = __ldg(L.K + 4 * b + 1); g.cx = __ldg(L.K + 4 * b + 2); g.cy = __ldg(L.K + 4 * b + 3);
    g.s0lo = g.s0hi = 0.f;
    g.b = b;
    g.tm_x1 = g.tm_s1 = g.tm_d1 = nullptr;
    if (TRU) {
      const uint32_t* mm = p.s0mm + 2 * (p.n_mm_groups > 1 ? b / p.group : 0);
      g.s0lo = ord2f(__ldcg(mm));
      g.s0hi = ord2f(__ldcg(mm + 1));
    }
  }
  DPFT_QSTAMP_MIN(k, b, 0);
  DPFT_QSTAMP(k, b, 1);
  __syncwarp();
  if (lane < 12) spose[lane] = __ldcg(p.pose_hist + ((size_t)k * B + b) * 12 + lane);
  if (TRU) {
#pragma unroll
    for (int i = 0; i < 12; ++i) redw[27 + i][lane] = 0.f;
  }
  __syncwarp();

  TileSums S;
  S.reset();
  {
    const int seg = t % L.nseg, rt = t / L.nseg;
    const int y0 = rt * L.TR, y1 = min(y0 + L.TR, L.H);
    float* outl = area + WF + 12 * 33;
    if (KIND >= 1)      // 2: the staged routine's form for maps narrower than its ring
      process_tile_staged<TRU, SB, GW, GH, AUX, KIND == 2>(g, spose, redw + 27, area, outl, seg, y0, y1, lane, S);
    else
      process_tile<8, TRU>(g, spose, redw + 27, seg, y0, y1, lane, S);
  }
  __syncwarp();
  DPFT_QSTAMP(k, b, 2);
  flush_warp<TRU>(S, redw, p.records + ((size_t)b * L.tpp + t) * PS, lane);
  DPFT_QSTAMP(k, b, 3);

  __threadfence();
  __syncwarp();
  int last = 0;
  if (lane == 0) last = (atomicAdd(p.tiles_done + b, 1) == L.tpp - 1);
  DPFT_QSTAMP(k, b, 4);
  return __shfl_sync(0xffffffffu, last, 0) != 0;
}

// The worker that completed the last tile of (iteration k, pair b): fold, solve (or park), append the next tiles.
template <bool TRU>
__device__ __noinline__ void q_finish_pair(const QueueParams& p, float* area, const int k, const int b) {
  const int lane = threadIdx.x & 31;
  constexpr int NE = TRU ? NSUM : 27;
  double* sdbl = reinterpret_cast<double*>(area);        // PS doubles, free between two tile walks
  const int B = p.B;
  const QLevel& L = p.L;
  const int grp = b / p.group;
  __threadfence();

  // ---------------------------------------------------------------- fold the pair's records in tile order
  const float* recs = p.records + (size_t)b * L.tpp * PS;
  const int n = L.tpp;
  // Lane e sums entry e of the records (and entry e + 32: the six corrections of the maximum and one of the minimum
  // live past 32), sixteen loads in flight per lane, always in record order.  A correction only counts when its
  // record sits on the pair's extreme, which is itself folded from the records first.
  float pair_min = CUDART_INF_F, pair_max = -CUDART_INF_F;
  if (TRU) {
    for (int i = lane; i < n; i += 32) {
      pair_min = fminf(pair_min, __ldcg(recs + (size_t)i * PS + E_VMIN));
      pair_max = fmaxf(pair_max, __ldcg(recs + (size_t)i * PS + E_VMAX));
    }
    pair_min = warp_min(pair_min);
    pair_max = warp_max(pair_max);
  }
  {
    const int e0 = lane, e1 = lane + 32;
    const bool two = e1 < NE;
    const int sl0 = e0 < 27 ? e0 : e0 + 2, sl1 = two ? e1 + 2 : sl0;
    const bool corr0 = TRU && e0 >= 27;
    const int key0_at = E_VMIN;                       // entries 27..31 belong to the minimum
    const int key1_at = e1 < 33 ? E_VMIN : E_VMAX;    // entry 32 is the last of the minimum's, 33..38 the maximum's
    const float want1 = e1 < 33 ? pair_min : pair_max;
    double s0 = 0.0, s1 = 0.0;
    for (int i0 = 0; i0 < n; i0 += 8) {
      float v0[8], v1[8], k0[8], k1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* q = recs + (size_t)min(i0 + j, n - 1) * PS;
        v0[j] = __ldcg(q + sl0);
        k0[j] = corr0 ? __ldcg(q + key0_at) : 0.f;
        v1[j] = two ? __ldcg(q + sl1) : 0.f;
        k1[j] = (TRU && two) ? __ldcg(q + key1_at) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (i0 + j < n) {
          if (!corr0 || k0[j] == pair_min) s0 += (double)v0[j];
          if (two && k1[j] == want1) s1 += (double)v1[j];
        }
      }
    }
    if (e0 < NE) sdbl[sl0] = s0;
    if (two) sdbl[sl1] = s1;
  }
  DPFT_QSTAMP(k, b, 5);
  float s0lo = 0.f, s0hi = 0.f;
  if (lane == 0) {
    sdbl[E_VMIN] = (double)pair_min;
    sdbl[E_VMAX] = (double)pair_max;
    p.tiles_done[b] = 0;               // ready for the pair's next iteration
    if (TRU && p.aux) {
      const uint32_t* mm = p.s0mm + 2 * (p.n_mm_groups > 1 ? grp : 0);
      s0lo = ord2f(__ldcg(mm));
      s0hi = ord2f(__ldcg(mm + 1));
    }
  }
  __syncwarp();

  if (!TRU || p.group == 1) {
    // nothing couples the pairs (a group of one is its own batch: its extremes are the batch extremes)
    if (lane == 0) {
      q_finalize<TRU, false>(p, k, b, sdbl, TRU, TRU && (pair_max != pair_min));
      if (TRU && p.aux) {
        float* a = p.aux + ((size_t)k * p.n_groups + grp) * 4;
        a[0] = pair_min; a[1] = pair_max; a[2] = s0lo; a[3] = s0hi;
      }
    }
    q_push_next(p, k, b, lane);
    if (p.t_done && lane == 0) {
      if (atomicAdd(p.groups_done + k, 1) == B - 1) p.t_done[k + 1] = q_now_ns();
    }
    return;
  }

  // ---------------------------------------------------------------- group extremes: solve now, or park as a candidate
  uint32_t* gx = p.gext + 2 * ((size_t)k * p.n_groups + grp);
  int cand = 0;
  if (lane == 0) {
    const uint32_t emin = f2ord(pair_min), emax = f2ord(pair_max);
    const uint32_t old_min = atomicMin(gx, emin), old_max = atomicMax(gx + 1, emax);
    cand = (emin <= old_min) || (emax >= old_max);
  }
  cand = __shfl_sync(0xffffffffu, cand, 0);
  if (cand) {
    double* rec = p.pairrec + (size_t)b * PS;
    for (int i = lane; i < PS; i += 32) rec[i] = sdbl[i];
    if (lane == 0) p.cand[b] = k + 1;
  } else {
    if (lane == 0) q_finalize<TRU, false>(p, k, b, sdbl, false, false);
    DPFT_QSTAMP(k, b, 6);
    q_push_next(p, k, b, lane);
    DPFT_QSTAMP(k, b, 7);
  }
  __threadfence();
  __syncwarp();
  int group_done = 0;
  if (lane == 0) group_done = (atomicAdd(p.pairs_done + (size_t)k * p.n_groups + grp, 1) == p.group - 1);
  group_done = __shfl_sync(0xffffffffu, group_done, 0);
  if (!group_done) return;
  __threadfence();

  // ---------------------------------------------------------------- last pair of the group: the extremes are final
  const float gmin = ord2f(__ldcg(gx)), gmax = ord2f(__ldcg(gx + 1));
  if (lane == 0 && p.aux) {
    float* a = p.aux + ((size_t)k * p.n_groups + grp) * 4;
    a[0] = gmin; a[1] = gmax; a[2] = s0lo; a[3] = s0hi;
  }
  const int b_lo = grp * p.group, b_hi = b_lo + p.group;
  for (int base = b_lo; base < b_hi; base += 32) {
    const int bb = base + lane;
    const bool mine = bb < b_hi && __ldcg(p.cand + bb) == k + 1;
    if (mine) {
      const double* rec = p.pairrec + (size_t)bb * PS;
      const bool at_min = ((float)__ldcg(rec + E_VMIN) == gmin);
      const bool at_max = ((float)__ldcg(rec + E_VMAX) == gmax) && (gmax != gmin);
      q_finalize<TRU, true>(p, k, bb, rec, at_min, at_max);
    }
    unsigned m = __ballot_sync(0xffffffffu, mine);
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      q_push_next(p, k, base + j, lane);
    }
  }
  DPFT_QSTAMP(k, 0, 8);
  if (p.t_done && lane == 0) {
    if (atomicAdd(p.groups_done + k, 1) == p.n_groups - 1) p.t_done[k + 1] = q_now_ns();
  }
}

template <bool TRU, bool SB, bool AUX, int GW, int GH, int KIND>
#ifdef DPFT_Q_MAXNREG_SB
__global__ void __maxnreg__((SB && KIND >= 1) ? DPFT_Q_MAXNREG_SB : 168) uic_queue_kernel(
#else
__global__ void __launch_bounds__(q_warps(SB, KIND) * 32, q_ctas_per_sm(SB, KIND)) uic_queue_kernel(
#endif
    const __grid_constant__ QueueParams p) {
  extern __shared__ __align__(128) float q_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* area = q_dyn + warp * q_area_floats(SB && KIND >= 1);
  // Twin launches for full sigma tensors whose channels may be copies of channel 0 (QueueParams::mism): only the twin
  // whose tile routine matches what sigma_replication_kernel found does the work, the other leaves the queue alone.
  if (p.rep_role && ((__ldcg(p.mism) == 0) != (p.rep_role == 1))) return;
  // Workers stay until the queue is exhausted (one CTA per resident slot).  Workers that leave after a few items, so
  // that retiring CTAs let the kernels of other streams in between, were measured and bought nothing
  // (profiles/r2/stream_probe_items.txt).
  while (true) {
    unsigned slot = 0;
    if (lane == 0) slot = atomicAdd(p.qctl, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= p.total_items) break;
    unsigned long long item = 0;
    if (lane == 0) {
      item = q_wait_item(p.fifo + slot);
      if (p.t_done && slot == 0) p.t_done[0] = q_now_ns();
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    if (q_walk_item<TRU, SB, AUX, GW, GH, KIND>(p, area, item)) q_finish_pair<TRU>(p, area, q_item_k(item), q_item_b(item));
  }
}

// head / tail, counters, running extremes, candidate marks, the FIFO with the first iteration of every pair in it
__global__ void __launch_bounds__(256) queue_init_kernel(const QueueParams p, const float* __restrict__ pose_in) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned n0 = (unsigned)p.B * (unsigned)p.L.tpp;
  if (i < p.total_items) {
    unsigned long long v = 0;
    if (i < n0) v = q_encode(0, (int)(i / p.L.tpp), (int)(i % p.L.tpp));
    p.fifo[i] = v;
  }
  if (i == 0) {
    p.qctl[0] = 0u;
    p.qctl[1] = n0;
  }
  if (i < (size_t)p.B * 12 && pose_in != p.pose_hist) p.pose_hist[i] = pose_in[i];
  if (i < (size_t)p.B) {
    p.tiles_done[i] = 0;
    p.cand[i] = 0;
  }
  if (i < (size_t)p.iters * p.n_groups) {
    p.pairs_done[i] = 0;
    p.gext[2 * i] = 0xffffffffu;
    p.gext[2 * i + 1] = 0u;
  }
  if (i < (size_t)p.iters) p.groups_done[i] = 0;
}

// extremes of sigma0 per (level, group): blockIdx.z = level, blockIdx.y = group
struct MmLevels {
  const float* v[DPFT_MAX_LEVELS];
  size_t per_group[DPFT_MAX_LEVELS];     // elements of one group's slice
  unsigned plane[DPFT_MAX_LEVELS];       // elements of one channel plane (0: unknown, never the channel-0 walk)
  const int* mism;                       // device flag of sigma_replication_kernel (nullptr: none); 0 = the C channels of
  int C;                                 //   a pair are copies of channel 0, whose extremes are the tensor's
};
__global__ void __launch_bounds__(256) minmax_levels_kernel(const MmLevels q, uint32_t* __restrict__ mm, const int n_groups) {
  __shared__ float s_lo[8], s_hi[8];
  const int l = blockIdx.z, grp = blockIdx.y;
  const size_t n = q.per_group[l];
  const float* v = q.v[l] + (size_t)grp * n;
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const unsigned plane = q.plane[l];
  if (q.mism && q.C > 1 && plane >= 4u && (plane & 3u) == 0u && n % ((size_t)q.C * plane) == 0 &&
      (reinterpret_cast<uintptr_t>(v) & 15) == 0 && __ldcg(q.mism) == 0) {
    // replicated channels: walk channel 0 of every pair of the slice, a C-th of the bytes
    const unsigned p4 = plane / 4u;
    const size_t n4 = n / q.C / 4u;                  // 16-byte chunks of the channel-0 planes
    const size_t pair_stride4 = (size_t)q.C * p4;
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const size_t pair = i / p4;
      const float4 a = __ldg(v4 + pair * pair_stride4 + (i - pair * p4));
      lo = fminf(fminf(lo, a.x), fminf(a.y, fminf(a.z, a.w)));
      hi = fmaxf(fmaxf(hi, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
    }
  } else {
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
    const float4 a = __ldg(v4 + i);
    lo = fminf(fminf(lo, a.x), fminf(a.y, fminf(a.z, a.w)));
    hi = fmaxf(fmaxf(hi, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
  }
  for (size_t j = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const float a = __ldg(v + j);
    lo = fminf(lo, a);
    hi = fmaxf(hi, a);
  }
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      lo = fminf(lo, s_lo[w]);
      hi = fmaxf(hi, s_hi[w]);
    }
    uint32_t* out = mm + 2 * ((size_t)l * n_groups + grp);   // (level, group)
    atomicMin(out, f2ord(lo));
    atomicMax(out + 1, f2ord(hi));
  }
}

void launch_minmax_levels(const float* const* v, const size_t* per_group, int n_levels, int n_groups, uint32_t* mm,
                          cudaStream_t stream, const unsigned* plane, int C, const int* mism) {
  MmLevels q{};
  size_t widest = 0;
  for (int l = 0; l < n_levels; ++l) {
    q.v[l] = v[l];
    q.per_group[l] = per_group[l];
    q.plane[l] = (plane && mism) ? plane[l] : 0u;
    widest = std::max(widest, per_group[l]);
  }
  q.mism = (plane && C > 1) ? mism : nullptr;
  q.C = C;
  const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>((widest / 4 + 255) / 256, (148 * 8) / std::max(1, n_groups) + 1));
  minmax_levels_kernel<<<dim3(bx, n_groups, n_levels), 256, 0, stream>>>(q, mm, n_groups);
}

template <bool TRU, bool SB, bool AUX, int GW, int GH, int KIND>
static cudaError_t launch_q(const QueueParams& prm, int grid, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1) {
  constexpr int smem = q_warps(SB, KIND) * q_area_floats(SB && KIND >= 1) * (int)sizeof(float);
  auto* fn = uic_queue_kernel<TRU, SB, AUX, GW, GH, KIND>;
  // (cudaFuncSetAttribute is per device and cheap: set it on every launch rather than caching per process)
  cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return err;
#ifdef DPFT_Q_CARVEOUT_MAX
  // (tuning hook: by default the driver picks the carve-out, which may hold fewer CTAs than the launch bounds name)
  err = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess) return err;
#endif
  if (ev0) cudaEventRecord(ev0, stream);
  fn<<<grid, q_warps(SB, KIND) * 32, smem, stream>>>(prm);
  if (ev1) cudaEventRecord(ev1, stream);
  return cudaGetLastError();
}

int queue_tiles_per_sm(bool one_map) { return q_ctas_per_sm(one_map, 1) * q_warps(one_map, 1); }

static cudaError_t launch_queue_variant(const QueueParams& prm, bool tru, int grid, cudaStream_t stream, bool allow_fixed_geometry,
                                        cudaEvent_t ev0, cudaEvent_t ev1);

// `prm.s0mm` must already hold the level's sigma0 extremes (launch_minmax_levels); pose_in may be prm.pose_hist.
cudaError_t launch_queue(const QueueParams& prm, const float* pose_in, bool tru, int grid, cudaStream_t stream,
                         bool allow_fixed_geometry, cudaEvent_t ev0, cudaEvent_t ev1) {
  {
    const size_t n = std::max<size_t>(std::max<size_t>(prm.total_items, (size_t)prm.B * 12), (size_t)prm.iters * prm.n_groups);
    queue_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(prm, pose_in);
  }
  const bool staged = prm.L.kind >= 1;
  if (staged && prm.mism && prm.SC == prm.C && prm.rep_role == 0) {
    // full sigma tensors, replication decided on the device: the one-map twin, then the C-map twin (one of them
    // returns at once); the events bracket both
    QueueParams one = prm, full = prm;
    one.SC = 1; one.rep_role = 1;
    full.rep_role = 2;
    cudaError_t err = launch_queue_variant(one, tru, grid, stream, allow_fixed_geometry, ev0, nullptr);
    if (err != cudaSuccess) return err;
    return launch_queue_variant(full, tru, grid, stream, allow_fixed_geometry, nullptr, ev1);
  }
  return launch_queue_variant(prm, tru, grid, stream, allow_fixed_geometry, ev0, ev1);
}

// one launch of the kernel instantiation that serves prm (prm.SC == 1: one uncertainty map per frame)
static cudaError_t launch_queue_variant(const QueueParams& prm, bool tru, int grid, cudaStream_t stream, bool allow_fixed_geometry,
                                 cudaEvent_t ev0, cudaEvent_t ev1) {
  const bool sb = prm.SC != prm.C;
  const bool aux = prm.L.m0 || prm.L.m1;
  const bool staged = prm.L.kind >= 1;
  if (grid <= 0) {     // what the device holds of THIS variant (the one-map routine fits a fourth CTA per SM)
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
      sms = 148;
    const int qw = q_warps(sb, staged ? 1 : 0);
    const long want = ((long)prm.B * prm.L.tpp + qw - 1) / qw;
    grid = (int)std::max<long>(1, std::min<long>((long)sms * q_ctas_per_sm(sb, staged ? 1 : 0), want));
  }
#define DPFT_Q(TRUV, SBV, AUXV, w, h, KINDV) launch_q<TRUV, SBV, AUXV, w, h, KINDV>(prm, grid, stream, ev0, ev1)
  if (!staged) {   // levels the staged routine does not take (narrow, unaligned): the plain tile routine
    if (tru) return aux ? DPFT_Q(true, false, true, 0, 0, 0) : DPFT_Q(true, false, false, 0, 0, 0);
    return aux ? DPFT_Q(false, false, true, 0, 0, 0) : DPFT_Q(false, false, false, 0, 0, 0);
  }
  if (prm.L.kind == 2) {   // narrow maps (make_qplan only plans them without object masks)
    if (tru) return sb ? DPFT_Q(true, true, false, 0, 0, 2) : DPFT_Q(true, false, false, 0, 0, 2);
    return sb ? DPFT_Q(false, true, false, 0, 0, 2) : DPFT_Q(false, false, false, 0, 0, 2);
  }
  // the reference's TUM level 0 (160x120) runs a geometry-specialised tile routine
  if (allow_fixed_geometry && !sb && !aux && prm.L.W == 160 && prm.L.H == 120)
    return tru ? DPFT_Q(true, false, false, 160, 120, 1) : DPFT_Q(false, false, false, 160, 120, 1);
  // (the one-map routine specialised for 160x120 measured 4 % SLOWER than its generic instantiation: 2760 against 2654 us
  // for 20 batches, profiles/r2/r2b_sigma_detect_fixedgeo.txt -- not instantiated)
  if (tru) {
    if (sb) return aux ? DPFT_Q(true, true, true, 0, 0, 1) : DPFT_Q(true, true, false, 0, 0, 1);
    return aux ? DPFT_Q(true, false, true, 0, 0, 1) : DPFT_Q(true, false, false, 0, 0, 1);
  }
  if (sb) return aux ? DPFT_Q(false, true, true, 0, 0, 1) : DPFT_Q(false, true, false, 0, 0, 1);
  return aux ? DPFT_Q(false, false, true, 0, 0, 1) : DPFT_Q(false, false, false, 0, 0, 1);
#undef DPFT_Q
}

}  // namespace dpft

#ifdef DPFT_QUEUE_STAMPS
// debug builds only: reset (fill with ~0) or read the phase stamps of pair 0
extern "C" int dpft_debug_queue_stamps(unsigned long long* host, int n_it, int reset) {
  if (reset) {
    static unsigned long long ones[512 * 16];
    for (auto& v : ones) v = ~0ull;
    return (int)cudaMemcpyToSymbol(dpft::g_qstamps, ones, sizeof(ones));
  }
  return (int)cudaMemcpyFromSymbol(host, dpft::g_qstamps, sizeof(unsigned long long) * 16 * (size_t)n_it);
}
#endif
