// Launchers shared between translation units of libdpft.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpft {

// gx, gy <- unit Sobel gradient of `planes` images of H x W (reference algorithms.py:1844-1865)
void launch_sobel_unit(const float* img, float* gx, float* gy, int planes, int H, int W, cudaStream_t stream);

// adjoint of launch_sobel_unit: g_img (accumulated) <- d/d(gx), d/d(gy)
void launch_sobel_unit_bwd(const float* img, const float* ggx, const float* ggy, float* g_img, int planes, int H, int W,
                           cudaStream_t stream);

// V (B,3,H,W) = [x,y,1] depth and N (B,3,H,W) its unit normals (algorithms.py:2148-2171); dmm = order-encoded
// min / max of the whole depth tensor
void launch_vertex_normal(const float* depth, const float* K, const uint32_t* dmm, float* V, float* N, int B, int H,
                          int W, cudaStream_t stream);

// Point-to-plane sums of one iteration into rec (B,28); zeroes rec first (algorithms.py:916-973).  wmap: optional
// (B,H,W) per-pixel scale applied to residual and Jacobian (a learned ScaleNet's output)
// scratch (optional, icp_scratch_bytes): per-CTA sums and per-pair counters -- with it the sums are folded in CTA order by
// the pair's last CTA (bitwise reproducible); without it they are added with float atomics
size_t icp_scratch_bytes(int B, int H, int W);
void launch_icp_term(const float* depth0, const float* K, const float* V1, const float* N1, const float* pose,
                     const uint8_t* m0, const uint8_t* m1, float* rec, uint8_t* occ_out, float* r_out, const float* wmap,
                     int B, int H, int W, cudaStream_t stream, void* scratch = nullptr);

// Pose gradient of the point-to-plane term of one iteration, accumulated into gpose (B,12)
void launch_icp_bwd(const float* depth0, const float* K, const float* V1, const float* N1, const float* pose,
                    const float* mlam, const uint8_t* m0, const uint8_t* m1, float* gpose, float w2, int B, int H, int W,
                    cudaStream_t stream);
void launch_minmax(const float* v, size_t n, uint32_t* mm, cudaStream_t stream);

// ---- persistent (single cooperative launch) U_IC forward, uic_persistent.cu
struct PLevel {
  const float *x0, *x1, *s0, *s1, *d0, *d1, *K;
  const uint8_t *m0, *m1;
  int H, W, nseg, nrt, TR, tpp;   // tpp = warp tiles per pair
};
struct PersistParams {
  PLevel lv[8];
  int n_levels, iters, B, C, rcap;   // rcap = record slots per pair
  float *pose_hist, *sys_hist, *aux;
  float* records;                    // (B, rcap, PS)
  double* pairrec;                   // (B, PS)
  const uint32_t* s0mm;              // per level: order-encoded min, max of sigma0
  uint32_t* gext;                    // per iteration: order-encoded min, max of the warped sigma (pre-initialised)
  int32_t* status;
  unsigned long long* clock_out;     // optional (n_it + 1) %globaltimer stamps at the iteration boundaries
};
int persistent_grid(int C, bool tru);   // CTAs that can be co-resident on the current device
cudaError_t launch_persistent(const PersistParams& prm, int grid, bool tru, cudaStream_t stream);

// ---- work-queue launch of one pyramid level (uic_queue.cu): per-pair dependencies, any number of sigma-extreme groups
struct QLevel {
  const float *x0, *x1, *s0, *s1, *d0, *d1, *K;
  const uint8_t *m0, *m1;
  int H, W, nseg, TR, nrt, tpp;      // TR rows per tile, tpp = tiles per pair
  int kind;                          // tile routine: 0 plain, 1 staged footprint, 2 staged, maps narrower than the ring
};
struct QueueParams {
  QLevel L;
  int iters, B, C, SC;
  int SCm;                           // channels of sigma0 / sigma1 in memory (C or 1); SC == 1 with SCm == C: read channel 0
  const int* mism;                   // device flag written by sigma_replication_kernel (0: every channel of sigma0 and
  int rep_role;                      //  sigma1 equals channel 0); role 1 runs only if it is 0, role 2 only if not, 0 always
  int group, n_groups;               // pairs per sigma-extreme group (the reference's batch), groups in this call
  int n_mm_groups;                   // groups of the sigma0 extremes (1 with a shared keyframe)
  int kf_shared;
  unsigned total_items;
  float *pose_hist, *sys_hist, *aux; // rows of THIS level: (iters + 1, B, 12), (iters, B, 27), (iters, n_groups, 4) or nullptr
  float* records;                    // (B, tpp, PS)
  double* pairrec;                   // (B, PS) parked sums of the candidates
  unsigned long long* fifo;          // (total_items)
  unsigned* qctl;                    // [0] head (claims), [1] tail (reservations)
  int *tiles_done, *cand;            // (B)
  int *pairs_done, *groups_done;     // (iters, n_groups), (iters)
  uint32_t* gext;                    // (iters, n_groups, 2) running extremes of the warped sigma, order-encoded
  const uint32_t* s0mm;              // (n_mm_groups, 2) extremes of this level's sigma0, order-encoded
  int32_t* status;
  unsigned long long* t_done;        // optional (iters + 1) %globaltimer stamps: start, then every iteration complete
};
int queue_tiles_per_sm(bool one_map);   // resident workers (warps) per SM of the staged routine (one sigma map: a 4th CTA)
// grid <= 0: as many CTAs as the device holds of the variant that runs
// ev0 / ev1 (optional): events recorded on `stream` right before / after the work-queue kernel itself
cudaError_t launch_queue(const QueueParams& prm, const float* pose_in, bool tru, int grid, cudaStream_t stream,
                         bool allow_fixed_geometry, cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);
// order-encoded [min, max] of v[l][g * per_group[l] .. (g + 1) * per_group[l]) into mm[(l * n_groups + g) * 2 ..]
// (atomicMin / atomicMax: mm must hold 0xffffffff, 0 on entry)
// plane / C / mism (optional): elements per channel plane of every level, channels per pair and the device flag of
// sigma_replication_kernel -- when the flag says the channels are copies, only channel 0 of every pair is read
void launch_minmax_levels(const float* const* v, const size_t* per_group, int n_levels, int n_groups, uint32_t* mm,
                          cudaStream_t stream, const unsigned* plane = nullptr, int C = 1, const int* mism = nullptr);

}  // namespace dpft
