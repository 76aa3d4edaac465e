// Adjoint of one damped Gauss-Newton pose update (fp64, one thread per pair):
//   xi = Hd^-1 rhs,  dR = exp(-xi_w) (Rodrigues, no small-angle guard),  dt = -dR xi_v,
//   R' = R dR,  t' = R dt + t                       (reference algorithms.py:2035-2054, geometry.py:105-185)
// Given d/d(R',t') it returns xi, lambda = Hd^-1 d/d(xi) (so d/d(rhs) = lambda, d/d(Hd) = -lambda xi^T) and
// ADDS d/d(R,t) through the compose into gpose_k.  Used by the U_IC backward and by the IC tracker's update.
#pragma once
#include "dpft_device.cuh"

namespace dpft {

__device__ inline void solve_update_adjoint(const double* Hd /*21, upper triangle, damping included*/, const double* rhs,
                                            const float* pose_k /*12*/, const float* gpose_next /*12*/, double* xi,
                                            double* lam, double* gpose_k /*12, accumulated*/) {
  // Cholesky with reciprocal pivots (one rsqrt per column, no division): this runs serially, one thread per pair
  double L[6][6], invd[6];
  for (int j = 0; j < 6; ++j) {
    double s = Hd[tri(j, j)];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    const double inv = rsqrt(s);
    invd[j] = inv;
    for (int i = j + 1; i < 6; ++i) {
      double v = Hd[tri(j, i)];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v * inv;
    }
  }
  auto chol_solve = [&](const double* r, double* out) {
    double z[6];
    for (int i = 0; i < 6; ++i) {
      double v = r[i];
      for (int k = 0; k < i; ++k) v -= L[i][k] * z[k];
      z[i] = v * invd[i];
    }
    for (int i = 5; i >= 0; --i) {
      double v = z[i];
      for (int k = i + 1; k < 6; ++k) v -= L[k][i] * out[k];
      out[i] = v * invd[i];
    }
  };
  chol_solve(rhs, xi);
  // forward: w = -xi_w, theta = |w|, k = w/theta, dR = I + K s + K^2 c1, dt = -dR xi_v
  const double w[3] = {-xi[0], -xi[1], -xi[2]};
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double ith = 1.0 / th;
  const double kv[3] = {w[0] * ith, w[1] * ith, w[2] * ith};
  double s, c;
  sincos(th, &s, &c);
  const double c1 = 1.0 - c;
  const double Kx[9] = {0, -kv[2], kv[1], kv[2], 0, -kv[0], -kv[1], kv[0], 0};
  double K2[9], dR[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double kk = 0;
      for (int m = 0; m < 3; ++m) kk += Kx[3 * i + m] * Kx[3 * m + j];
      K2[3 * i + j] = kk;
      dR[3 * i + j] = (i == j ? 1.0 : 0.0) + Kx[3 * i + j] * s + kk * c1;
    }
  double dt[3];
  for (int i = 0; i < 3; ++i) dt[i] = -(dR[3 * i] * xi[3] + dR[3 * i + 1] * xi[4] + dR[3 * i + 2] * xi[5]);
  double R[9], GR[9], Gt[3];
  for (int i = 0; i < 9; ++i) {
    R[i] = (double)pose_k[i];
    GR[i] = (double)gpose_next[i];
  }
  for (int i = 0; i < 3; ++i) Gt[i] = (double)gpose_next[9 + i];
  // R' = R dR, t' = R dt + t
  double gdR[9], gdt[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0, bb = 0;
      for (int m = 0; m < 3; ++m) {
        a += GR[3 * i + m] * dR[3 * j + m];      // GR dR^T
        bb += R[3 * m + i] * GR[3 * m + j];      // R^T GR
      }
      gpose_k[3 * i + j] += a + Gt[i] * dt[j];
      gdR[3 * i + j] = bb;
    }
  for (int i = 0; i < 3; ++i) {
    gdt[i] = R[i] * Gt[0] + R[3 + i] * Gt[1] + R[6 + i] * Gt[2];
    gpose_k[9 + i] += Gt[i];
  }
  // dt = -dR xi_v
  double gxi[6];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) gdR[3 * i + j] -= gdt[i] * xi[3 + j];
  for (int j = 0; j < 3; ++j) gxi[3 + j] = -(dR[j] * gdt[0] + dR[3 + j] * gdt[1] + dR[6 + j] * gdt[2]);
  // Rodrigues
  double g_s = 0, g_c1 = 0, gK[9];
  for (int i = 0; i < 9; ++i) {
    g_s += gdR[i] * Kx[i];
    g_c1 += gdR[i] * K2[i];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0;
      for (int m = 0; m < 3; ++m) a += gdR[3 * i + m] * Kx[3 * j + m] + Kx[3 * m + i] * gdR[3 * m + j];   // G K^T + K^T G
      gK[3 * i + j] = gdR[3 * i + j] * s + c1 * a;
    }
  const double gk[3] = {gK[7] - gK[5], gK[2] - gK[6], gK[3] - gK[1]};
  double g_th = g_s * c + g_c1 * s;
  const double kdotw = gk[0] * w[0] + gk[1] * w[1] + gk[2] * w[2];
  g_th -= kdotw * ith * ith;
  for (int i = 0; i < 3; ++i) gxi[i] = -(gk[i] * ith + g_th * w[i] * ith);
  chol_solve(gxi, lam);
}

}  // namespace dpft
