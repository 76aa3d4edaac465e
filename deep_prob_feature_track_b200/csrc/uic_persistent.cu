// U_IC forward as ONE cooperative launch: every level and every Gauss-Newton iteration inside a persistent
// grid (one CTA slot per SM x occupancy), grid-wide barriers where the algorithm has a real dependency.
//
// Why: with one launch per iteration the coarse pyramid levels (9 of the 12 iterations) are dominated by
// fixed costs -- launch, drain, three fence/atomic hand-offs in the reduction tail -- and every level pays
// wave quantisation.  Here a warp owns a contiguous range of warp tiles (pair-major order) for the whole
// launch, writes one partial record per pair it touched, and after a grid barrier CTA b folds the records of
// pair b in a fixed order (deterministic), the batch-global sigma extremes are exchanged through one
// atomicMin/atomicMax and a second barrier (only with remove_tru_sigma -- that coupling is the reference's,
// algorithms.py:1976-1979), the pair's system is damped and solved in fp64, and a last barrier publishes the
// poses for the next iteration.
#include <cooperative_groups.h>
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

namespace cg = cooperative_groups;

namespace dpft {

#ifndef DPFT_MIN_CTAS
#define DPFT_MIN_CTAS 4
#endif

#ifndef DPFT_WARPS
#define DPFT_WARPS 4
#endif
// the plan (make_plan in uic_forward.cu) sizes the record slots with DPFT_WARPS warps per CTA: one constant for both
constexpr int kPW = DPFT_WARPS, kPT = kPW * 32;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int CH, bool TRU>
__global__ void __launch_bounds__(kPT, DPFT_MIN_CTAS) uic_persistent_kernel(const PersistParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float red[kPW][NSUM][33];
  __shared__ float s_mn[kPW], s_mx[kPW];
  __shared__ __align__(16) float s_pose[kPW][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kPW + warp, nw = gridDim.x * kPW;
  const int B = p.B, C = p.C;
  constexpr int NE = TRU ? NSUM : 27;

  int k = 0;
  for (int l = 0; l < p.n_levels; ++l) {
    const PLevel& L = p.lv[l];
    const int plane = L.H * L.W;
    const int total_tiles = L.tpp * B;
    const int T = (total_tiles + nw - 1) / nw;   // tiles per warp
    float s0lo = 0.f, s0hi = 0.f;
    if (TRU) {
      s0lo = ord2f(__ldcg(p.s0mm + 2 * l));
      s0hi = ord2f(__ldcg(p.s0mm + 2 * l + 1));
    }
    for (int it = 0; it < p.iters; ++it, ++k) {
      if (p.clock_out && blockIdx.x == 0 && threadIdx.x == 0) p.clock_out[k] = global_ns();
      const float* pose_k = p.pose_hist + (size_t)k * B * 12;
      // ------------------------------------------------------------ this warp's tiles
      {
        const int t0 = min(gw * T, total_tiles), t1 = min(t0 + T, total_tiles);
        int cur_b = -1;
        TileSums S;
        PairView g;
        for (int t = t0; t < t1; ++t) {
          const int b = t / L.tpp, tl = t - b * L.tpp;
          if (b != cur_b) {
            if (cur_b >= 0)
              flush_warp<TRU>(S, red[warp], p.records + ((size_t)cur_b * p.rcap + (gw - (cur_b * L.tpp) / T)) * PS, lane);
            cur_b = b;
            S.reset();
            if (TRU) {
#pragma unroll
              for (int i = 0; i < 12; ++i) red[warp][27 + i][lane] = 0.f;
            }
            const size_t po = (size_t)b * C * plane;
            g.x0 = L.x0 + po; g.x1 = L.x1 + po; g.s0 = L.s0 + po; g.s1 = L.s1 + po;
            g.d0 = L.d0 + (size_t)b * plane; g.d1 = L.d1 + (size_t)b * plane;
            g.m0 = L.m0 ? L.m0 + (size_t)b * plane : nullptr;
            g.m1 = L.m1 ? L.m1 + (size_t)b * plane : nullptr;
            g.occ_out = nullptr; g.sr0_dbg = nullptr;
            g.H = L.H; g.W = L.W; g.C = C; g.splane = (unsigned)plane;
            g.fx = __ldg(L.K + 4 * b); g.fy = __ldg(L.K + 4 * b + 1); g.cx = __ldg(L.K + 4 * b + 2); g.cy = __ldg(L.K + 4 * b + 3);
            g.s0lo = s0lo; g.s0hi = s0hi;
            __syncwarp();
            if (lane < 12) s_pose[warp][lane] = __ldcg(pose_k + (size_t)b * 12 + lane);
            __syncwarp();
          }
          const int seg = tl % L.nseg, rt = tl / L.nseg;
          const int y0 = rt * L.TR, y1 = min(y0 + L.TR, L.H);
          process_tile<CH, TRU>(g, s_pose[warp], &red[warp][27], seg, y0, y1, lane, S);
        }
        if (cur_b >= 0)
          flush_warp<TRU>(S, red[warp], p.records + ((size_t)cur_b * p.rcap + (gw - (cur_b * L.tpp) / T)) * PS, lane);
      }
      grid.sync();
      // ------------------------------------------------------------ CTA b folds the records of pair b
      for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int w_first = (b * L.tpp) / T, w_last = ((b + 1) * L.tpp - 1) / T;
        const int n = w_last - w_first + 1;
        const float* recs = p.records + (size_t)b * p.rcap * PS;
        float pmin = CUDART_INF_F, pmax = -CUDART_INF_F;
        if (TRU) {
          float a = CUDART_INF_F, c = -CUDART_INF_F;
          for (int i = threadIdx.x; i < n; i += kPT) {
            a = fminf(a, __ldcg(recs + (size_t)i * PS + E_VMIN));
            c = fmaxf(c, __ldcg(recs + (size_t)i * PS + E_VMAX));
          }
          a = warp_min(a);
          c = warp_max(c);
          __syncthreads();
          if (lane == 0) {
            s_mn[warp] = a;
            s_mx[warp] = c;
          }
          __syncthreads();
#pragma unroll
          for (int w = 0; w < kPW; ++w) {
            pmin = fminf(pmin, s_mn[w]);
            pmax = fmaxf(pmax, s_mx[w]);
          }
        }
        double* rec = p.pairrec + (size_t)b * PS;
        if (threadIdx.x < NE) {
          const int e = threadIdx.x, slot = e < 27 ? e : e + 2;
          double s = 0.0;
          for (int i = 0; i < n; ++i) {
            const float* q = recs + (size_t)i * PS;
            bool take = true;
            if (TRU && e >= 27 && e < 33) take = (__ldcg(q + E_VMIN) == pmin);
            if (TRU && e >= 33) take = (__ldcg(q + E_VMAX) == pmax);
            if (take) s += (double)__ldcg(q + slot);
          }
          rec[slot] = s;
        }
        if (TRU && threadIdx.x == 0) {
          rec[E_VMIN] = (double)pmin;
          rec[E_VMAX] = (double)pmax;
          atomicMin(p.gext + 2 * k, f2ord(pmin));
          atomicMax(p.gext + 2 * k + 1, f2ord(pmax));
        }
      }
      if (TRU) grid.sync();
      else __syncthreads();
      // ------------------------------------------------------------ damp, solve, update (one thread per pair)
      if (threadIdx.x == 0) {
        float gmin = 0.f, gmax = 0.f;
        if (TRU) {
          gmin = ord2f(__ldcg(p.gext + 2 * k));
          gmax = ord2f(__ldcg(p.gext + 2 * k + 1));
        }
        for (int b = blockIdx.x; b < B; b += gridDim.x) {
          const double* rec = p.pairrec + (size_t)b * PS;
          double A[21], rhs[6];
#pragma unroll
          for (int i = 0; i < 21; ++i) A[i] = __ldcg(rec + i);
#pragma unroll
          for (int i = 0; i < 6; ++i) rhs[i] = __ldcg(rec + 21 + i);
          if (TRU) {
            const bool at_min = ((float)__ldcg(rec + E_VMIN) == gmin);
            const bool at_max = ((float)__ldcg(rec + E_VMAX) == gmax) && (gmax != gmin);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              if (at_min) rhs[i] -= __ldcg(rec + E_CMIN + i);
              if (at_max) rhs[i] -= __ldcg(rec + E_CMAX + i);
            }
          }
          bool finite = true;
#pragma unroll
          for (int i = 0; i < 21; ++i) finite = finite && isfinite(A[i]);
#pragma unroll
          for (int i = 0; i < 6; ++i) finite = finite && isfinite(rhs[i]);
          float* sys = p.sys_hist + ((size_t)k * B + b) * 27;
#pragma unroll
          for (int i = 0; i < 21; ++i) sys[i] = (float)A[i];
#pragma unroll
          for (int i = 0; i < 6; ++i) sys[21 + i] = (float)rhs[i];
          double xi[6];
          const bool ok = solve_and_update(A, rhs, true, pose_k + (size_t)b * 12,
                                           p.pose_hist + ((size_t)(k + 1) * B + b) * 12, xi);
          int st = 0;
          if (!finite) st |= DPFT_ST_NONFINITE;
          if (!ok) st |= DPFT_ST_SINGULAR;
          if (st) atomicOr(p.status, st);
          if (b == 0 && p.aux) {
            p.aux[4 * k] = gmin; p.aux[4 * k + 1] = gmax; p.aux[4 * k + 2] = s0lo; p.aux[4 * k + 3] = s0hi;
          }
        }
      }
      grid.sync();
    }
  }
  if (p.clock_out && blockIdx.x == 0 && threadIdx.x == 0) p.clock_out[k] = global_ns();
}

template <int CH, bool TRU>
static cudaError_t launch_one(const PersistParams& prm, int grid, cudaStream_t stream) {
  void* args[] = {(void*)&prm};
  return cudaLaunchCooperativeKernel((void*)uic_persistent_kernel<CH, TRU>, dim3(grid), dim3(kPT), args, 0, stream);
}

template <int CH, bool TRU>
static int resident_ctas() {
  int nb = 0, dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, uic_persistent_kernel<CH, TRU>, kPT, 0);
  return nb * sms;
}

int persistent_grid(int C, bool tru) {
  const int CH = (C % 8 == 0) ? 8 : (C % 4 == 0) ? 4 : (C % 2 == 0) ? 2 : 1;
  switch (CH) {
    case 8: return tru ? resident_ctas<8, true>() : resident_ctas<8, false>();
    case 4: return tru ? resident_ctas<4, true>() : resident_ctas<4, false>();
    case 2: return tru ? resident_ctas<2, true>() : resident_ctas<2, false>();
    default: return tru ? resident_ctas<1, true>() : resident_ctas<1, false>();
  }
}

cudaError_t launch_persistent(const PersistParams& prm, int grid, bool tru, cudaStream_t stream) {
  const int C = prm.C;
  const int CH = (C % 8 == 0) ? 8 : (C % 4 == 0) ? 4 : (C % 2 == 0) ? 2 : 1;
  switch (CH) {
    case 8: return tru ? launch_one<8, true>(prm, grid, stream) : launch_one<8, false>(prm, grid, stream);
    case 4: return tru ? launch_one<4, true>(prm, grid, stream) : launch_one<4, false>(prm, grid, stream);
    case 2: return tru ? launch_one<2, true>(prm, grid, stream) : launch_one<2, false>(prm, grid, stream);
    default: return tru ? launch_one<1, true>(prm, grid, stream) : launch_one<1, false>(prm, grid, stream);
  }
}

}  // namespace dpft
