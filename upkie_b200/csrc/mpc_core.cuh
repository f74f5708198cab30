// SPDX-License-Identifier: Apache-2.0
//
// mpc_core.cuh -- per-robot arithmetic of the batched MPC balancer kernel.
//
// Replaces MPCBalancer.step (upkie/controllers/mpc_balancer.py:237-312) whose QP
// the reference builds with qpmpc (condensed, dense N x N Hessian) and solves
// with ProxQP. The problem is a linear-quadratic tracking problem over the
// wheeled-inverted-pendulum model with box-bounded inputs; here it is solved in
// its un-condensed form by a primal-dual active-set iteration whose
// equality-constrained sub-problems are solved EXACTLY by a Riccati sweep
// (O(N) per sweep, numerically benign in fp32 where the condensed Hessian's
// conditioning ~1e6 is not). The minimiser of a strictly convex QP is unique, so
// the result coincides with ProxQP's up to its eps_abs = 1e-3 tolerance.
//
//   cost   1/2 [ w_u sum_{k<N} u_k^2 + w_x sum_{k<N} |x_k - r_k|^2 + w_T |x_N - r_N|^2 ]
//   s.t.   x_{k+1} = A x_k + B u_k,  |u_k| <= a_max,
//   r_k = [p0 + k T v, 0, v, 0]      (get_target_states, mpc_balancer.py:18-37)
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define UPKIE_MPC_HD __host__ __device__ __forceinline__
#else
#define UPKIE_MPC_HD inline
#endif

namespace upkie_b200 {

template <typename T>
struct MpcParams {
  // A = [[1,0,Ts,0],[0,ch,0,sho],[0,0,1,0],[0,osh,0,ch]], B = [b0,b1,b2,b3]
  T Ts, ch, sho, osh;
  T b0, b1, b2, b3;
  T w_u, w_x, w_T;
  T a_max, v_max, fall_pitch;
  int N, max_iterations;
};

template <typename T>
UPKIE_MPC_HD void mpc_mulA(const MpcParams<T>& M, const T x[4], T y[4]) {
  y[0] = x[0] + M.Ts * x[2];
  y[1] = M.ch * x[1] + M.sho * x[3];
  y[2] = x[2];
  y[3] = M.osh * x[1] + M.ch * x[3];
}
template <typename T>
UPKIE_MPC_HD void mpc_mulAT(const MpcParams<T>& M, const T y[4], T x[4]) {
  x[0] = y[0];
  x[1] = M.ch * y[1] + M.osh * y[3];
  x[2] = M.Ts * y[0] + y[2];
  x[3] = M.sho * y[1] + M.ch * y[3];
}

// symmetric 4x4 packed: 00 01 02 03 11 12 13 22 23 33
UPKIE_MPC_HD constexpr int S4(int i, int j) { return i <= j ? (i * (7 - i)) / 2 + j : (j * (7 - j)) / 2 + i; }

// Scratch accessor: element (k, c) of the per-step gains of this robot.
// Layout [5 N][stride] so that consecutive robots are consecutive in memory.
template <typename T>
struct MpcScratch {
  T* base;
  int stride;
  UPKIE_MPC_HD T& at(int k, int c) const { return base[size_t(k * 5 + c) * stride]; }
};

// ---- free-tail tables ---------------------------------------------------------------------------------------------
// The matrix part of the Riccati recursion (P_k, s_k = P_{k+1} B, q_uu) depends on the ACTIVE SET only, not on the
// robot's state or target: behind the last bounded step of the horizon ("free tail", where saturation rarely reaches:
// the bounds bite in the first steps of a plan) it is the same for every robot and every tick. It is computed once per
// handle in double precision - row k: s_k (4), A^T s_k (4), 1 / q_uu,k, pad, P_{k+1} (10) - and the sweep over the free
// tail only propagates the affine part p_k (~25 instructions per step instead of ~150); the explicit recursion starts
// at the last bounded step from the tabulated P.
constexpr int kMpcTabRow = 20;
constexpr int kMpcTabS = 0, kMpcTabATs = 4, kMpcTabInv = 8, kMpcTabP = 10;

// index of the highest set bit, -1 for 0
UPKIE_MPC_HD int mpc_last_bound(uint64_t m) {
#if defined(__CUDA_ARCH__)
  return 63 - __clzll((long long)m);
#else
  int k = -1;
  while (m) { ++k; m >>= 1; }
  return k;
#endif
}

// host: fill tab[N * kMpcTabRow] (double arithmetic whatever T is)
template <typename T>
inline void mpc_build_tables(const MpcParams<double>& M, T* tab) {
  const int N = M.N;
  const double B[4] = {M.b0, M.b1, M.b2, M.b3};
  double P[4][4] = {{0}};
  for (int i = 0; i < 4; ++i) P[i][i] = M.w_T;
  const double A[4][4] = {{1, 0, M.Ts, 0}, {0, M.ch, 0, M.sho}, {0, 0, 1, 0}, {0, M.osh, 0, M.ch}};
  for (int k = N - 1; k >= 0; --k) {
    T* row = tab + size_t(k) * kMpcTabRow;
    double s[4], ATs[4], quu = M.w_u;
    for (int i = 0; i < 4; ++i) { s[i] = 0; for (int j = 0; j < 4; ++j) s[i] += P[i][j] * B[j]; }
    for (int i = 0; i < 4; ++i) quu += B[i] * s[i];
    for (int i = 0; i < 4; ++i) { ATs[i] = 0; for (int j = 0; j < 4; ++j) ATs[i] += A[j][i] * s[j]; }
    for (int i = 0; i < 4; ++i) { row[kMpcTabS + i] = T(s[i]); row[kMpcTabATs + i] = T(ATs[i]); }
    row[kMpcTabInv] = T(1.0 / quu);
    row[kMpcTabInv + 1] = T(0);
    for (int i = 0; i < 4; ++i)
      for (int j = i; j < 4; ++j) row[kMpcTabP + S4(i, j)] = T(P[i][j]);
    double PA[4][4], Z[4][4];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) { PA[i][j] = 0; for (int l = 0; l < 4; ++l) PA[i][j] += P[i][l] * A[l][j]; }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) { Z[i][j] = 0; for (int l = 0; l < 4; ++l) Z[i][j] += A[l][i] * PA[l][j]; }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) P[i][j] = Z[i][j] - ATs[i] * ATs[j] / quu + (i == j ? M.w_x : 0.0);
  }
}

// One Riccati sweep for the active set (up, lo): bit k set = u_k fixed at
// +a_max / -a_max. Stores beta_k = B^T p_{k+1} for every step and s_k = B^T P_{k+1} (4 values) for the steps at or
// before the last bounded one (behind it s_k is the tabulated one).
template <typename T>
UPKIE_MPC_HD void mpc_backward(const MpcParams<T>& M, const T* tab, T p0, T v, uint64_t up, uint64_t lo,
                               const MpcScratch<T>& sc) {
  const int N = M.N;
  T P[10], p[4];
  {
    const T rN[4] = {p0 + T(N) * M.Ts * v, T(0), v, T(0)};
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = -M.w_T * rN[i];
  }
  const T B[4] = {M.b0, M.b1, M.b2, M.b3};
  const int kmax = mpc_last_bound(up | lo);
  // free tail: every later step is free, the matrices are tabulated, only the affine part moves
  // the row of the NEXT step is fetched while this one is computed (the loads do not depend on the recursion)
  T rw[5];
  {
    const T* row = tab + size_t(N - 1) * kMpcTabRow;
#pragma unroll
    for (int i = 0; i < 4; ++i) rw[i] = row[kMpcTabATs + i];
    rw[4] = row[kMpcTabInv];
  }
  for (int k = N - 1; k > kmax; --k) {
    const T ats[4] = {rw[0], rw[1], rw[2], rw[3]};
    const T inv = rw[4];
    if (k > 0) {
      const T* row = tab + size_t(k - 1) * kMpcTabRow;
#pragma unroll
      for (int i = 0; i < 4; ++i) rw[i] = row[kMpcTabATs + i];
      rw[4] = row[kMpcTabInv];
    }
    const T beta = B[0] * p[0] + B[1] * p[1] + B[2] * p[2] + B[3] * p[3];
    sc.at(k, 4) = beta;
    T ATp[4];
    mpc_mulAT(M, p, ATp);
    const T rk0 = p0 + T(k) * M.Ts * v;
    const T g = beta * inv;
    p[0] = -M.w_x * rk0 + ATp[0] - ats[0] * g;
    p[1] = ATp[1] - ats[1] * g;
    p[2] = -M.w_x * v + ATp[2] - ats[2] * g;
    p[3] = ATp[3] - ats[3] * g;
  }
  if (kmax < 0) return;
  {
    const T* row = tab + size_t(kmax) * kMpcTabRow;
#pragma unroll
    for (int i = 0; i < 10; ++i) P[i] = row[kMpcTabP + i];
  }
  for (int k = kmax; k >= 0; --k) {
    // s = P B, beta = B . p
    T s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = P[S4(i, 0)] * B[0] + P[S4(i, 1)] * B[1] + P[S4(i, 2)] * B[2] + P[S4(i, 3)] * B[3];
    const T beta = B[0] * p[0] + B[1] * p[1] + B[2] * p[2] + B[3] * p[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) sc.at(k, i) = s[i];
    sc.at(k, 4) = beta;
    const bool is_up = (up >> k) & 1ull, is_lo = (lo >> k) & 1ull;
    // Z = A^T P A (symmetric): columns of P A, then A^T
    T PA[4][4];  // PA[:, j] = P * A[:, j]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      PA[i][0] = P[S4(i, 0)];
      PA[i][1] = M.ch * P[S4(i, 1)] + M.osh * P[S4(i, 3)];
      PA[i][2] = M.Ts * P[S4(i, 0)] + P[S4(i, 2)];
      PA[i][3] = M.sho * P[S4(i, 1)] + M.ch * P[S4(i, 3)];
    }
    T Z[10];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const T col[4] = {PA[0][j], PA[1][j], PA[2][j], PA[3][j]};
      T r[4];
      mpc_mulAT(M, col, r);
#pragma unroll
      for (int i = 0; i <= j; ++i) Z[S4(i, j)] = r[i];
    }
    T ATp[4], ATs[4];
    mpc_mulAT(M, p, ATp);
    mpc_mulAT(M, s, ATs);  // = (B^T P A)^T = Q_ux^T
    const T rk[4] = {p0 + T(k) * M.Ts * v, T(0), v, T(0)};
    if (is_up || is_lo) {
      const T ub = is_up ? M.a_max : -M.a_max;
#pragma unroll
      for (int i = 0; i < 10; ++i) P[i] = Z[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = -M.w_x * rk[i] + ATp[i] + ATs[i] * ub;
    } else {
      const T quu = M.w_u + (B[0] * s[0] + B[1] * s[1] + B[2] * s[2] + B[3] * s[3]);
      const T inv = T(1) / quu;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) P[S4(i, j)] = Z[S4(i, j)] - ATs[i] * ATs[j] * inv;
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = -M.w_x * rk[i] + ATp[i] - ATs[i] * (beta * inv);
    }
    P[S4(0, 0)] += M.w_x; P[S4(1, 1)] += M.w_x; P[S4(2, 2)] += M.w_x; P[S4(3, 3)] += M.w_x;
  }
}

// Forward rollout with the stored gains; returns the updated active set and
// writes the inputs into the scratch (slot 4 of each step is re-used for u_k).
template <typename T>
UPKIE_MPC_HD void mpc_forward(const MpcParams<T>& M, const T* tab, const T x0[4], uint64_t up, uint64_t lo,
                              const MpcScratch<T>& sc, uint64_t& new_up, uint64_t& new_lo, T& u0) {
  const int N = M.N;
  const T B[4] = {M.b0, M.b1, M.b2, M.b3};
  T x[4] = {x0[0], x0[1], x0[2], x0[3]};
  new_up = 0;
  new_lo = 0;
  const int kmax = mpc_last_bound(up | lo);
  for (int k = 0; k <= kmax; ++k) {  // up to the last bounded step: the gains this robot's sweep stored
    T Ax[4];
    mpc_mulA(M, x, Ax);
    const T beta = sc.at(k, 4);
    const T s[4] = {sc.at(k, 0), sc.at(k, 1), sc.at(k, 2), sc.at(k, 3)};
    const bool is_up = (up >> k) & 1ull, is_lo = (lo >> k) & 1ull;
    T u;
    if (is_up || is_lo) {
      u = is_up ? M.a_max : -M.a_max;
      // multiplier sign: gradient of the cost-to-go w.r.t. u_k at the bound
      const T g = M.w_u * u + s[0] * (Ax[0] + B[0] * u) + s[1] * (Ax[1] + B[1] * u) + s[2] * (Ax[2] + B[2] * u) +
                  s[3] * (Ax[3] + B[3] * u) + beta;
      if (is_up && !(g > T(0))) new_up |= (1ull << k);
      if (is_lo && !(g < T(0))) new_lo |= (1ull << k);
    } else {
      const T quu = M.w_u + (B[0] * s[0] + B[1] * s[1] + B[2] * s[2] + B[3] * s[3]);
      u = -(s[0] * Ax[0] + s[1] * Ax[1] + s[2] * Ax[2] + s[3] * Ax[3] + beta) / quu;
      if (u > M.a_max) { new_up |= (1ull << k); }
      else if (u < -M.a_max) { new_lo |= (1ull << k); }
    }
    sc.at(k, 4) = u;
    if (k == 0) u0 = u;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = Ax[i] + B[i] * u;
  }
  // free tail: tabulated gains; the row and the affine term of the NEXT step are fetched while this one is computed
  if (kmax + 1 < N) {
    T rw[6];
    {
      const T* row = tab + size_t(kmax + 1) * kMpcTabRow;
#pragma unroll
      for (int i = 0; i < 4; ++i) rw[i] = row[kMpcTabS + i];
      rw[4] = row[kMpcTabInv];
      rw[5] = sc.at(kmax + 1, 4);
    }
    for (int k = kmax + 1; k < N; ++k) {
      const T s[4] = {rw[0], rw[1], rw[2], rw[3]};
      const T inv = rw[4], beta = rw[5];
      if (k + 1 < N) {
        const T* row = tab + size_t(k + 1) * kMpcTabRow;
#pragma unroll
        for (int i = 0; i < 4; ++i) rw[i] = row[kMpcTabS + i];
        rw[4] = row[kMpcTabInv];
        rw[5] = sc.at(k + 1, 4);
      }
      T Ax[4];
      mpc_mulA(M, x, Ax);
      const T u = -(s[0] * Ax[0] + s[1] * Ax[1] + s[2] * Ax[2] + s[3] * Ax[3] + beta) * inv;
      if (u > M.a_max) { new_up |= (1ull << k); }
      else if (u < -M.a_max) { new_lo |= (1ull << k); }
      sc.at(k, 4) = u;
      if (k == 0) u0 = u;
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = Ax[i] + B[i] * u;
    }
  }
}

// Full solve. Returns true when the active set reached a fixed point (optimal).
// On exit the plan sits in scratch slot 4 (clipped to the bounds).
template <typename T>
UPKIE_MPC_HD bool mpc_solve(const MpcParams<T>& M, const T* tab, const T x0[4], T v_target, const MpcScratch<T>& sc,
                            T& u0, uint64_t& up, uint64_t& lo) {
  bool converged = false;
  for (int it = 0; it < M.max_iterations; ++it) {
    mpc_backward(M, tab, x0[0], v_target, up, lo, sc);
    uint64_t nu, nl;
    mpc_forward(M, tab, x0, up, lo, sc, nu, nl, u0);
    if (nu == up && nl == lo) { converged = true; break; }
    up = nu;
    lo = nl;
  }
  return converged;
}

// MPCBalancer.step post-processing (mpc_balancer.py:260,295-311)
template <typename T>
UPKIE_MPC_HD T mpc_command_update(const MpcParams<T>& M, T v_cmd, T pitch, bool floor_contact, bool found, T u0, T dt) {
  const bool fallen = fabs(pitch) > M.fall_pitch;
  if (fallen || !floor_contact) {
    const T alpha = dt / T(0.1);  // low_pass_filter(cutoff_period=0.1), filters.py:63-80
    return v_cmd + alpha * (T(0) - v_cmd);
  }
  if (!found) return v_cmd;  // plan.is_empty: re-send previous ground velocity
  T v = v_cmd + u0 * dt / T(2);
  if (v < -M.v_max) v = -M.v_max;  // clamp_abs, clamp.py:33-39
  if (v > M.v_max) v = M.v_max;
  return v;
}

}  // namespace upkie_b200
