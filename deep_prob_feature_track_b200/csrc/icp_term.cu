// Point-to-plane ICP term of the U_IC tracker (reference code/models/algorithms.py:916-997, 2148-2171,
// geometry.py:1129-1136), added to the feature-metric normal equations when combine_icp is on.
//
//   vertex_normal_kernel  once per level: V1 = [x, y, 1] depth1 and N1 = normalised cross product of the
//                         (un-normalised, replicate-padded) Sobel responses of V1; zero where depth1 sits
//                         on its batch-global extremes.
//   icp_term_kernel       per iteration, one thread per pixel: P = R V0 + t, project, bilinear lookup of
//                         V1 / N1, r = N1.(P - V1), J = -[ (N1^T R) x V0 , -(N1^T R) ] / sigma_icp,
//                         accumulates sum J J^T and sum J r per pair.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>

#include "dpft.h"
#include "dpft_device.cuh"
#include "dpft_kernels.h"

namespace dpft {

__global__ void __launch_bounds__(256) vertex_normal_kernel(const float* __restrict__ depth, const float* __restrict__ K,
                                                            const uint32_t* __restrict__ dmm, float* __restrict__ V,
                                                            float* __restrict__ N, int B, int H, int W) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= W || y >= H) return;
  const float fx = __ldg(K + 4 * b), fy = __ldg(K + 4 * b + 1), cx = __ldg(K + 4 * b + 2), cy = __ldg(K + 4 * b + 3);
  const float* d = depth + (size_t)b * H * W;
  const int xs[3] = {max(x - 1, 0), x, min(x + 1, W - 1)};
  const int ys[3] = {max(y - 1, 0), y, min(y + 1, H - 1)};
  float v[3][3][3];   // [row][col][component]
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float z = __ldg(d + ys[r] * W + xs[c]);
      const float px = xdiv(xsub((float)xs[c], cx), fx), py = xdiv(xsub((float)ys[r], cy), fy);
      v[r][c][0] = px * z;
      v[r][c][1] = py * z;
      v[r][c][2] = z;
    }
  float sx[3], sy[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    sx[k] = (v[0][2][k] - v[0][0][k]) + 2.f * (v[1][2][k] - v[1][0][k]) + (v[2][2][k] - v[2][0][k]);
    sy[k] = (v[2][0][k] - v[0][0][k]) + 2.f * (v[2][1][k] - v[0][1][k]) + (v[2][2][k] - v[0][2][k]);
  }
  float n[3] = {sx[1] * sy[2] - sx[2] * sy[1], sx[2] * sy[0] - sx[0] * sy[2], sx[0] * sy[1] - sx[1] * sy[0]};
  const float mag = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-8f;
  const float z = v[1][1][2];
  const bool bad = (z == ord2f(dmm[0])) || (z == ord2f(dmm[1]));
  const size_t plane = (size_t)H * W, o = (size_t)b * 3 * plane + (size_t)y * W + x;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    V[o + k * plane] = v[1][1][k];
    N[o + k * plane] = bad ? 0.f : n[k] / mag;
  }
}

struct IcpParams {
  const float *depth0, *K, *V1, *N1, *pose;
  const uint8_t *m0, *m1;
  float* rec;         // (B,28): 21 + 6 sums (atomically accumulated; zero before the launch)
  uint8_t* occ_out;   // optional (B,H,W)
  float* r_out;       // optional (B,H,W) weighted-by-sigma residual (1e-6 where masked)
  const float* wmap;  // optional (B,H,W) per-pixel scale of residual and Jacobian (learned ScaleNet, alg:677-682)
  float* part;        // optional (B, gridDim.x, 28) per-CTA sums: the pair's last CTA folds them in CTA order into rec
  int* done;          //   (B) CTAs of the pair finished (zero before the launch, zero again after it)
  int H, W, B, ppt;
};

__global__ void __launch_bounds__(128, 4) icp_term_kernel(const IcpParams p) {
  __shared__ float s_red[4][27];
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, plane = H * W;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const Pose pose = load_pose(p.pose + (size_t)b * 12);
  const float* V1 = p.V1 + (size_t)b * 3 * plane;
  const float* N1 = p.N1 + (size_t)b * 3 * plane;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  for (int i = 0; i < p.ppt; ++i) {
    const int pix = (blockIdx.x * p.ppt + i) * 128 + threadIdx.x;
    if (pix >= plane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float z = __ldg(p.depth0 + (size_t)b * plane + pix);
    const float v0[3] = {xmul(px, z), xmul(py, z), z};
    float P[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      P[k] = xadd(xadd(xadd(xmul(pose.r[3 * k], v0[0]), xmul(pose.r[3 * k + 1], v0[1])), xmul(pose.r[3 * k + 2], v0[2])),
                  pose.t[k]);
    const float u = xadd(xmul(xdiv(P[0], P[2]), fx), cx);
    const float v = xadd(xmul(xdiv(P[1], P[2]), fy), cy);
    const bool inview = (u > 0.f) && (u < (float)(W - 1)) && (v > 0.f) && (v < (float)(H - 1));
    const Tap tap = make_tap(u, v, H, W);
    float v1[3], n1[3], diff[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v1[k] = sample_exact(V1 + (size_t)k * plane, tap, W);
      n1[k] = sample_exact(N1 + (size_t)k * plane, tap, W);
      diff[k] = xsub(P[k], v1[k]);
    }
    const float dist = sqrtf(xadd(xadd(xmul(diff[0], diff[0]), xmul(diff[1], diff[1])), xmul(diff[2], diff[2])));
    bool occ = !inview || (dist > 0.1f);
    if (p.m0) occ = occ || (__ldg(p.m0 + (size_t)b * plane + pix) == 0);
    if (p.m1) occ = occ || !(sample_mask(p.m1 + (size_t)b * plane, tap, W) > 0.f);
    float r = n1[0] * diff[0] + n1[1] * diff[1] + n1[2] * diff[2];
    float nr[3];   // N1^T R
#pragma unroll
    for (int j = 0; j < 3; ++j) nr[j] = n1[0] * pose.r[j] + n1[1] * pose.r[3 + j] + n1[2] * pose.r[6 + j];
    // TUM stereo noise model projected on the rotated normal (algorithms.py:975-997)
    const float s_lat = z / 525.f * 5.5f, s_z = z * z * 0.4f / (525.f * 1.2f);
    const float sig = sqrtf((nr[0] * s_lat) * (nr[0] * s_lat) + (nr[1] * s_lat) * (nr[1] * s_lat) +
                            (nr[2] * s_z) * (nr[2] * s_z) + 1e-8f);
    const float inv = 1.f / (sig + 1e-8f);
    r *= inv;
    float J[6];
    J[0] = -(nr[1] * v0[2] - nr[2] * v0[1]) * inv;
    J[1] = -(nr[2] * v0[0] - nr[0] * v0[2]) * inv;
    J[2] = -(nr[0] * v0[1] - nr[1] * v0[0]) * inv;
    J[3] = nr[0] * inv;
    J[4] = nr[1] * inv;
    J[5] = nr[2] * inv;
    float rm = occ ? 1e-6f : r;
    if (p.wmap) {   // w scales r and J alike: sum (w J)(w J)^T, sum (w J)(w r)
      const float w = __ldg(p.wmap + (size_t)b * plane + pix);
      rm *= w;
#pragma unroll
      for (int a = 0; a < 6; ++a) J[a] *= w;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
      for (int c = a; c < 6; ++c) acc[tri(a, c)] = fmaf(J[a], J[c], acc[tri(a, c)]);
      acc[21 + a] = fmaf(J[a], rm, acc[21 + a]);
    }
    if (p.occ_out) p.occ_out[(size_t)b * plane + pix] = occ ? 1 : 0;
    if (p.r_out) p.r_out[(size_t)b * plane + pix] = occ ? 1e-6f : r;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 27; ++i) {
    float s = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[warp][i] = s;
  }
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < 27) s = s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x];
  if (!p.part) {      // callers that only want the per-pixel outputs: the order of the sums does not matter to them
    if (threadIdx.x < 27) atomicAdd(p.rec + (size_t)b * 28 + threadIdx.x, s);
    return;
  }
  // the solver's sums: fixed-order fold by the pair's last CTA (bitwise the same whatever the CTAs' timing)
  __shared__ int s_last;
  float* mine = p.part + ((size_t)b * gridDim.x + blockIdx.x) * 28;
  if (threadIdx.x < 27) mine[threadIdx.x] = s;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(p.done + b, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 27) {
    double acc = 0.0;
    const float* q = p.part + (size_t)b * gridDim.x * 28 + threadIdx.x;
    for (unsigned c0 = 0; c0 < gridDim.x; c0 += 8) {       // eight loads in flight, added in CTA order
      float v[8];
#pragma unroll
      for (unsigned j = 0; j < 8; ++j) v[j] = (c0 + j < gridDim.x) ? __ldcg(q + (size_t)(c0 + j) * 28) : 0.f;
#pragma unroll
      for (unsigned j = 0; j < 8; ++j) acc += (double)v[j];
    }
    p.rec[(size_t)b * 28 + threadIdx.x] = (float)acc;
  }
  if (threadIdx.x == 0) p.done[b] = 0;
}

// Reverse mode of icp_term_kernel for one iteration: only the pose receives a gradient (the depth maps are
// inputs without gradient and the features do not enter this term).  With J = Jn * inv, Jn = [V0 x nr, nr],
// nr = N1r^T R, r = (N1r . (P - V1r)) * inv, the upstream is  dL/dJ = w^2 (M J + rm lambda),  dL/dr = w^2 J.lambda
// (zero where masked), exactly as for the feature term (uic_backward.cu).
struct IcpBwdParams {
  const float *depth0, *K, *V1, *N1, *pose, *mlam;
  const uint8_t *m0, *m1;
  float* gpose;     // (B,12) accumulated
  float w2;
  int H, W, B, ppt;
};

__global__ void __launch_bounds__(128, 4) icp_bwd_kernel(const IcpBwdParams p) {
  __shared__ float s_red[4][12];
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, plane = H * W;
  const float fx = __ldg(p.K + 4 * b), fy = __ldg(p.K + 4 * b + 1), cx = __ldg(p.K + 4 * b + 2), cy = __ldg(p.K + 4 * b + 3);
  const Pose pose = load_pose(p.pose + (size_t)b * 12);
  const float* V1 = p.V1 + (size_t)b * 3 * plane;
  const float* N1 = p.N1 + (size_t)b * 3 * plane;
  float M[21], lam[6];
#pragma unroll
  for (int i = 0; i < 21; ++i) M[i] = __ldg(p.mlam + (size_t)b * 27 + i);
#pragma unroll
  for (int i = 0; i < 6; ++i) lam[i] = __ldg(p.mlam + (size_t)b * 27 + 21 + i);
  float gR[9], gt[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) gR[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) gt[i] = 0.f;
  for (int it = 0; it < p.ppt; ++it) {
    const int pix = (blockIdx.x * p.ppt + it) * 128 + threadIdx.x;
    if (pix >= plane) break;
    const int y = pix / W, x = pix - y * W;
    const float px = xdiv(xsub((float)x, cx), fx), py = xdiv(xsub((float)y, cy), fy);
    const float z = __ldg(p.depth0 + (size_t)b * plane + pix);
    const float v0[3] = {xmul(px, z), xmul(py, z), z};
    float P[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      P[k] = xadd(xadd(xadd(xmul(pose.r[3 * k], v0[0]), xmul(pose.r[3 * k + 1], v0[1])), xmul(pose.r[3 * k + 2], v0[2])),
                  pose.t[k]);
    const float u = xadd(xmul(xdiv(P[0], P[2]), fx), cx);
    const float v = xadd(xmul(xdiv(P[1], P[2]), fy), cy);
    const bool inview = (u > 0.f) && (u < (float)(W - 1)) && (v > 0.f) && (v < (float)(H - 1));
    const Tap tap = make_tap(u, v, H, W);
    float v1[3], n1[3], diff[3], dv1x[3], dv1y[3], dn1x[3], dn1y[3];
    const float tyn = tap.wa + tap.wb, tys = tap.wc + tap.wd, txl = tap.wa + tap.wc, txr = tap.wb + tap.wd;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* q = V1 + (size_t)k * plane + tap.o;
      const float a = __ldg(q), bb = __ldg(q + 1), c = __ldg(q + W), d = __ldg(q + W + 1);
      v1[k] = blend_exact(a, bb, c, d, tap);
      dv1x[k] = fmaf(bb - a, tyn, (d - c) * tys);
      dv1y[k] = fmaf(c - a, txl, (d - bb) * txr);
      const float* qn = N1 + (size_t)k * plane + tap.o;
      const float na = __ldg(qn), nb = __ldg(qn + 1), nc = __ldg(qn + W), nd = __ldg(qn + W + 1);
      n1[k] = blend_exact(na, nb, nc, nd, tap);
      dn1x[k] = fmaf(nb - na, tyn, (nd - nc) * tys);
      dn1y[k] = fmaf(nc - na, txl, (nd - nb) * txr);
      diff[k] = xsub(P[k], v1[k]);
    }
    const float dist = sqrtf(xadd(xadd(xmul(diff[0], diff[0]), xmul(diff[1], diff[1])), xmul(diff[2], diff[2])));
    bool occ = !inview || (dist > 0.1f);
    if (p.m0) occ = occ || (__ldg(p.m0 + (size_t)b * plane + pix) == 0);
    if (p.m1) occ = occ || !(sample_mask(p.m1 + (size_t)b * plane, tap, W) > 0.f);
    // forward values
    const float r0 = n1[0] * diff[0] + n1[1] * diff[1] + n1[2] * diff[2];
    float nr[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) nr[j] = n1[0] * pose.r[j] + n1[1] * pose.r[3 + j] + n1[2] * pose.r[6 + j];
    const float s_lat = z / 525.f * 5.5f, s_z = z * z * 0.4f / (525.f * 1.2f);
    const float sd2[3] = {s_lat * s_lat, s_lat * s_lat, s_z * s_z};
    const float qsum = nr[0] * nr[0] * sd2[0] + nr[1] * nr[1] * sd2[1] + nr[2] * nr[2] * sd2[2];
    const float sig = sqrtf(qsum + 1e-8f);
    const float inv = 1.f / (sig + 1e-8f);
    const float Jn[6] = {v0[1] * nr[2] - v0[2] * nr[1], v0[2] * nr[0] - v0[0] * nr[2], v0[0] * nr[1] - v0[1] * nr[0],
                         nr[0], nr[1], nr[2]};
    float J[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) J[k] = Jn[k] * inv;
    const float r = r0 * inv;
    const float rm = occ ? 1e-6f : r;
    // upstream
    float gJ[6], gr = 0.f;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      float sacc = rm * lam[a];
#pragma unroll
      for (int c = 0; c < 6; ++c) sacc = fmaf(M[a <= c ? tri(a, c) : tri(c, a)], J[c], sacc);
      gJ[a] = p.w2 * sacc;
      gr = fmaf(J[a], lam[a], gr);
    }
    gr = occ ? 0.f : p.w2 * gr;
    // J = Jn inv, r = r0 inv
    float ginv = gr * r0, gJn[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      ginv = fmaf(gJ[k], Jn[k], ginv);
      gJn[k] = gJ[k] * inv;
    }
    const float gr0 = gr * inv;
    // Jn = [V0 x nr, nr]
    float gnr[3] = {gJn[3] + (gJn[1] * v0[2] - gJn[2] * v0[1]), gJn[4] + (gJn[2] * v0[0] - gJn[0] * v0[2]),
                    gJn[5] + (gJn[0] * v0[1] - gJn[1] * v0[0])};
    // inv = 1 / (sig + 1e-8), sig = sqrt(sum (nr sd)^2 + 1e-8)
    const float gq = -ginv * inv * inv / (2.f * sig);
#pragma unroll
    for (int k = 0; k < 3; ++k) gnr[k] = fmaf(gq, 2.f * nr[k] * sd2[k], gnr[k]);
    // r0 = N1r . diff ; nr = N1r^T R ; diff = P - V1r
    float gn1[3], gP[3], gv1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gn1[i] = gr0 * diff[i] + pose.r[3 * i] * gnr[0] + pose.r[3 * i + 1] * gnr[1] + pose.r[3 * i + 2] * gnr[2];
      gP[i] = gr0 * n1[i];
      gv1[i] = -gP[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) gR[3 * i + j] = fmaf(n1[i], gnr[j], gR[3 * i + j]);
    }
    // sampled maps -> sample position (zero on / outside the clip range, as grid_sampler)
    const float ixu = xmul(xmul(xadd(xsub(xdiv(u, 0.5f * (float)(W - 1)), 1.f), 1.f), 0.5f), (float)(W - 1));
    const float iyu = xmul(xmul(xadd(xsub(xdiv(v, 0.5f * (float)(H - 1)), 1.f), 1.f), 0.5f), (float)(H - 1));
    float gu = 0.f, gv = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gu = fmaf(gv1[k], dv1x[k], fmaf(gn1[k], dn1x[k], gu));
      gv = fmaf(gv1[k], dv1y[k], fmaf(gn1[k], dn1y[k], gv));
    }
    if (!(ixu > 0.f && ixu < (float)(W - 1))) gu = 0.f;
    if (!(iyu > 0.f && iyu < (float)(H - 1))) gv = 0.f;
    const float iz = 1.f / P[2];
    gP[0] = fmaf(gu, fx * iz, gP[0]);
    gP[1] = fmaf(gv, fy * iz, gP[1]);
    gP[2] -= (gu * fx * P[0] + gv * fy * P[1]) * iz * iz;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) gR[3 * i + j] = fmaf(gP[i], v0[j], gR[3 * i + j]);
      gt[i] += gP[i];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float sacc = i < 9 ? gR[i] : gt[i - 9];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    if (lane == 0) s_red[warp][i] = sacc;
  }
  __syncthreads();
  if (threadIdx.x < 12)
    atomicAdd(p.gpose + (size_t)b * 12 + threadIdx.x,
              s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
}

void launch_icp_bwd(const float* depth0, const float* K, const float* V1, const float* N1, const float* pose,
                    const float* mlam, const uint8_t* m0, const uint8_t* m1, float* gpose, float w2, int B, int H, int W,
                    cudaStream_t stream) {
  IcpBwdParams p{};
  p.depth0 = depth0; p.K = K; p.V1 = V1; p.N1 = N1; p.pose = pose; p.mlam = mlam; p.m0 = m0; p.m1 = m1;
  p.gpose = gpose; p.w2 = w2; p.H = H; p.W = W; p.B = B;
  const long plane = (long)H * W;
  const long want_threads = 148L * 2048 * 2;
  long ppt = ((long)B * plane + want_threads - 1) / want_threads;
  p.ppt = (int)std::max(1L, std::min(ppt, 8L));
  const dim3 grid((unsigned)((plane + 128L * p.ppt - 1) / (128L * p.ppt)), B);
  icp_bwd_kernel<<<grid, 128, 0, stream>>>(p);
}

void launch_vertex_normal(const float* depth, const float* K, const uint32_t* dmm, float* V, float* N, int B, int H,
                          int W, cudaStream_t stream) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8, B);
  vertex_normal_kernel<<<grid, 256, 0, stream>>>(depth, K, dmm, V, N, B, H, W);
}

#ifndef DPFT_ICP_WANT_THREADS
#define DPFT_ICP_WANT_THREADS (148L * 2048 * 2)      // tuning hook: resident threads the launch is sized for
#endif
static int icp_pixels_per_thread(int B, long plane) {
  const long want_threads = DPFT_ICP_WANT_THREADS;
  const long ppt = ((long)B * plane + want_threads - 1) / want_threads;
  return (int)std::max(1L, std::min(ppt, 8L));
}
static unsigned icp_ctas_per_pair(int B, long plane) {
  const long per_cta = 128L * icp_pixels_per_thread(B, plane);
  return (unsigned)((plane + per_cta - 1) / per_cta);
}

size_t icp_scratch_bytes(int B, int H, int W) {
  const size_t part = (size_t)B * icp_ctas_per_pair(B, (long)H * W) * 28 * sizeof(float);
  return ((part + 255) & ~(size_t)255) + (size_t)B * sizeof(int);
}

void launch_icp_term(const float* depth0, const float* K, const float* V1, const float* N1, const float* pose,
                     const uint8_t* m0, const uint8_t* m1, float* rec, uint8_t* occ_out, float* r_out, const float* wmap,
                     int B, int H, int W, cudaStream_t stream, void* scratch) {
  IcpParams p{};
  p.depth0 = depth0; p.K = K; p.V1 = V1; p.N1 = N1; p.pose = pose; p.m0 = m0; p.m1 = m1;
  p.rec = rec; p.occ_out = occ_out; p.r_out = r_out; p.wmap = wmap; p.H = H; p.W = W; p.B = B;
  const long plane = (long)H * W;
  p.ppt = icp_pixels_per_thread(B, plane);
  const dim3 grid(icp_ctas_per_pair(B, plane), B);
  if (scratch) {
    const size_t part = ((size_t)B * grid.x * 28 * sizeof(float) + 255) & ~(size_t)255;
    p.part = (float*)scratch;
    p.done = (int*)((char*)scratch + part);
    cudaMemsetAsync(p.done, 0, (size_t)B * sizeof(int), stream);
  } else {
    cudaMemsetAsync(rec, 0, (size_t)B * 28 * sizeof(float), stream);
  }
  icp_term_kernel<<<grid, 128, 0, stream>>>(p);
}

}  // namespace dpft
