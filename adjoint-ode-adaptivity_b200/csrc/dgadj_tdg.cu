// dgadj_tdg.cu -- DG-in-time ODE march and reverse-time DG adjoint with the per-element
// adjoint-weighted indicator: matlab/dg_march.m:1-80 and matlab/adj_march.m:1-122, batched
// over the initial value on a shared mesh (one thread per trajectory, elements in sequence,
// the whole element system in registers).
//
// Everything that does not depend on the trajectory is computed once per mesh by the host
// (fem_setup.m:1-41 operators, the polyfit/polyval interpolation of dg_march.m:47-49 /
// adj_march.m:75-79 as matrices -- quirk C-5 -- including the mirrored quadrature interval
// of adj_march.m:72,78 -- quirk C-3) and passed as per-element constant blocks:
//   march   block: A[Np*Np] | Iq[nq*Np] | Phi[nq*Np] | w[nq] | hk | np_k
//   adjoint block: A0[Na*Na] | f1[Na] | A2[Na*Na] | Ix[Na*Npp] | Iq[nq*Npp] | Phi[nq*Na] | w[nq] | hk | na_k | last_{k-1}
// Mixed orders over the mesh (Ns(k), matlab/MAIN.m:21,141): Np / Na / nq are the
// mesh maxima and every element's block is padded to them by the host -- identity rows in the
// system matrices, zero rows/columns in the interpolation matrices, zero quadrature weights --
// so the padded unknowns stay exactly 0 and the leading np_k x np_k system is eliminated with
// the same pivots and the same arithmetic as an unpadded solve.  np_k / na_k = the element's
// own node counts, last_{k-1} = index of the last primal node of the element before it.
// The device does what depends on the state: sin/cos at the quadrature points, the element
// residual and Jacobian, Newton's iteration with the reference's stopping rule
// (||dU||_2 <= tol, at most maxit+1 iterations: dg_march.m:36,44), the dense solves
// (Gaussian elimination with partial pivoting, MATLAB's `\`), the indicator dot product.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "dgadj_internal.h"

namespace dgadj {

// x / h from the correctly rounded reciprocal r = RN(1/h): q = RN(x r) is within an ulp of the quotient, the
// residual x - h q is exact in an fma, and RN(q + residual r) is the correctly rounded quotient (Markstein's
// theorem): the bits of the division without a second division sequence.
__device__ __forceinline__ double div_rn_q(double x, double h, double r) {
  const double q = x * r;
  return fma(fma(-h, q, x), r, q);
}

// Gaussian elimination with partial pivoting.  The pivots' reciprocals of the elimination serve the back
// substitution as well (b[r] = s / A[r][r] through div_rn_q: same bits, half the division sequences -- they are a
// fifth of a Newton iteration of the 2 x 2 systems of N = 1).
template <int N>
__device__ __forceinline__ void solve_dense(double (&A)[N][N], double (&b)[N]) {
  double inv[N];
#pragma unroll
  for (int c = 0; c < N; ++c) {
    // partial pivoting: bring the largest |A[r][c]|, r >= c, to row c (branch-free row swaps)
#pragma unroll
    for (int r = c + 1; r < N; ++r) {
      const bool sw = fabs(A[r][c]) > fabs(A[c][c]);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double t = A[c][j];
        A[c][j] = sw ? A[r][j] : t;
        A[r][j] = sw ? t : A[r][j];
      }
      const double tb = b[c];
      b[c] = sw ? b[r] : tb;
      b[r] = sw ? tb : b[r];
    }
    inv[c] = 1.0 / A[c][c];
#pragma unroll
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] * inv[c];
#pragma unroll
      for (int j = c + 1; j < N; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
      b[r] = fma(-f, b[c], b[r]);
    }
  }
#pragma unroll
  for (int r = N - 1; r >= 0; --r) {
    double s = b[r];
#pragma unroll
    for (int j = r + 1; j < N; ++j) s = fma(-A[r][j], b[j], s);
    b[r] = div_rn_q(s, A[r][r], inv[r]);
  }
}

// forward march: y[b][k][:] = element solution, its[b][k] = Newton iterations
template <int NP>
__global__ void tdg_march_kernel(long long B, int Ks, int nq, int linear, double tol, int maxit,
                                 const double* __restrict__ ec, const double* __restrict__ y0,
                                 double* __restrict__ y, int* __restrict__ its) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int blk = NP * NP + 2 * nq * NP + nq + 2;
  double uR = y0[b];
  for (int k = 0; k < Ks; ++k) {
    const double* A = ec + (size_t)k * blk;
    const double* Iq = A + NP * NP;
    const double* Phi = Iq + nq * NP;
    const double* w = Phi + nq * NP;
    const double hk2 = 0.5 * w[nq];
    const int npk = (int)w[nq + 1];
    double U[NP];
    int it = 0;
    if (linear) {  // dg_march.m:11-25: one solve A U = F, F(1) = uR_prev
      double M[NP][NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        U[i] = (i == 0) ? uR : 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) M[i][j] = A[i * NP + j];
      }
      solve_dense<NP>(M, U);
      it = 1;
    } else {  // dg_march.m:27-77
#pragma unroll
      for (int i = 0; i < NP; ++i) U[i] = (i < npk) ? uR : 0.0;
      double err = 1.0;
      while (it <= maxit && err > tol) {
        double Mt[NP], J[NP][NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          Mt[i] = 0.0;
#pragma unroll
          for (int j = 0; j < NP; ++j) J[i][j] = 0.0;
        }
        for (int q = 0; q < nq; ++q) {
          double ur = 0.0, ph[NP];
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            ur = fma(Iq[q * NP + i], U[i], ur);
            ph[i] = Phi[q * NP + i];
          }
          double sn, cs;
          sincos(ur, &sn, &cs);
          const double ws = w[q] * sn, wc = w[q] * cs;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            Mt[i] = fma(ph[i], ws, Mt[i]);
            const double pw = ph[i] * wc;
#pragma unroll
            for (int j = 0; j < NP; ++j) J[i][j] = fma(pw, ph[j], J[i][j]);
          }
        }
        double R[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          double r = fma(hk2, Mt[i], (i == 0) ? uR : 0.0);   // M_tilde + F
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            r = fma(A[i * NP + j], U[j], r);                  // + A*U_old
            J[i][j] = fma(hk2, J[i][j], A[i * NP + j]);       // dRdU = A + dMtdU
          }
          R[i] = r;
        }
        solve_dense<NP>(J, R);                                // delta_u = dRdU \ R
        double e2 = 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          U[i] -= R[i];
          e2 = fma(R[i], R[i], e2);
        }
        err = sqrt(e2);
        ++it;
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (i == npk - 1) uR = U[i];
      y[((size_t)b * Ks + k) * NP + i] = U[i];
    }
    if (its) its[(size_t)b * Ks + k] = it;
  }
}
// Small batches (nonlinear branch): a GROUP of G lanes per trajectory (G = 16: two trajectories per warp).  A
// thread-per-trajectory launch of a few thousand trajectories leaves one warp per SM.  A group's lanes take the
// 30 N + 1 quadrature points of an element (dg_march.m:29) in turns, the partial sums meet in a butterfly of
// shuffles -- every lane of the group ends with the same bits -- and each lane repeats the small dense solve.  Same
// stopping rule per trajectory (a group that has converged idles through the iterations its warp-mate still
// needs); the quadrature sums are associated differently from the sequential kernel (rounding-level differences in
// U, same Newton iteration counts).
// MEASURED (B = 4096, Ks = 31, N = 1; ncu launch list): G = 4: 249 us, 8: 164 us, 16: 134 us (129 us with the
// stopping rule decided without the square root), 32: 169 us (thread
// form: ~1 ms).  With 32 lanes everything but the sin / cos is replicated 32 times and the fp64 pipe (16 lanes per
// cycle, whatever they compute) is the bound; below 16 the march is bound by the latency of one Newton iteration
// (~3000 cycles: two reciprocal sequences and the substitution of the solve, the sin / cos pair, the butterfly, the
// square root of the stopping rule), 85 of which follow one another along a trajectory.
template <int NP, int G>
__global__ void tdg_march_group_kernel(long long B, int Ks, int nq, double tol, int maxit,
                                       const double* __restrict__ ec, const double* __restrict__ y0,
                                       double* __restrict__ y, int* __restrict__ its) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long bb = gt / G;
  const int l = (int)(gt % G);
  const bool act = bb < B;              // (idle groups of the last warp compute on the last trajectory, write nothing)
  const long long b = act ? bb : B - 1;
  const int blk = NP * NP + 2 * nq * NP + nq + 2;
  double uR = y0[b];
  for (int k = 0; k < Ks; ++k) {
    const double* A = ec + (size_t)k * blk;
    const double* Iq = A + NP * NP;
    const double* Phi = Iq + nq * NP;
    const double* w = Phi + nq * NP;
    const double hk2 = 0.5 * w[nq];
    const int npk = (int)w[nq + 1];
    double U[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) U[i] = (i < npk) ? uR : 0.0;
    int it = 0;
    // the stopping rule `err > tol`, err = sqrt(sum R^2) (dg_march.m:66-68), decided without the square root
    // wherever the sum is clear of tol^2 (sqrt is monotone and correctly rounded: a sum beyond tol^2 (1 +- 1e-12)
    // cannot round across tol); the square root only inside that band -- the same decisions (measured: 3 % of the
    // kernel time)
    const double tol2 = tol * tol, t2hi = tol2 * (1.0 + 1e-12), t2lo = tol2 * (1.0 - 1e-12);
    bool above = 1.0 > tol;
    while (__any_sync(0xffffffffu, it <= maxit && above)) {   // (warp-uniform trip count: the shuffles below)
      const bool go = it <= maxit && above;                  // this trajectory's own stopping rule
      double Mt[NP], J[NP][NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        Mt[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) J[i][j] = 0.0;
      }
#pragma unroll 4   // (the sin / cos evaluations of a lane's points side by side: one after the other they are the
                   //  longest dependent chain of an iteration)
      for (int q0 = 0; q0 < nq; q0 += G) {
        const int q = q0 + l;
        if (q < nq) {
          double ur = 0.0, ph[NP];
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            ur = fma(Iq[q * NP + i], U[i], ur);
            ph[i] = Phi[q * NP + i];
          }
          double sn, cs;
          sincos(ur, &sn, &cs);
          const double ws = w[q] * sn, wc = w[q] * cs;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            Mt[i] = fma(ph[i], ws, Mt[i]);
            const double pw = ph[i] * wc;
#pragma unroll
            for (int j = i; j < NP; ++j) J[i][j] = fma(pw, ph[j], J[i][j]);   // symmetric: upper triangle
          }
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          Mt[i] += __shfl_xor_sync(0xffffffffu, Mt[i], o);
#pragma unroll
          for (int j = i; j < NP; ++j) J[i][j] += __shfl_xor_sync(0xffffffffu, J[i][j], o);
        }
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) J[i][j] = J[j][i];
      }
      double R[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        double r = fma(hk2, Mt[i], (i == 0) ? uR : 0.0);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          r = fma(A[i * NP + j], U[j], r);
          J[i][j] = fma(hk2, J[i][j], A[i * NP + j]);
        }
        R[i] = r;
      }
      solve_dense<NP>(J, R);
      double e2 = 0.0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        U[i] = go ? U[i] - R[i] : U[i];
        e2 = fma(R[i], R[i], e2);
      }
      bool ab = e2 > t2hi;
      if (!ab && !(e2 < t2lo)) ab = sqrt(e2) > tol;   // (NaN lands here: sqrt(NaN) > tol is false, as in the plain rule)
      above = go ? ab : above;
      it += go ? 1 : 0;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (i == npk - 1) uR = U[i];
      if (act && l == 0) y[((size_t)b * Ks + k) * NP + i] = U[i];
    }
    if (its && act && l == 0) its[(size_t)b * Ks + k] = it;
  }
}

// reverse march: v[b][k][:], err[b][k]
template <int NPP>
__global__ void tdg_adjoint_kernel(long long B, int Ks, int nq, int linear, double y0_hard,
                                   const double* __restrict__ y0_arr,
                                   const double* __restrict__ ec, const double* __restrict__ y,
                                   double* __restrict__ v, double* __restrict__ err) {
  constexpr int NA = NPP + 1;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int blk = NA * NA + NA + NA * NA + NA * NPP + nq * NPP + nq * NA + nq + 3;
  double vL = 0.0;
  for (int k = Ks - 1; k >= 0; --k) {
    const double* A0 = ec + (size_t)k * blk;
    const double* f1 = A0 + NA * NA;
    const double* A2 = f1 + NA;
    const double* Ix = A2 + NA * NA;
    const double* Iq = Ix + NA * NPP;
    const double* Phi = Iq + nq * NPP;
    const double* w = Phi + nq * NA;
    const double hk2 = 0.5 * w[nq];
    const int nak = (int)w[nq + 1], lastprev = (int)w[nq + 2];
    double Uk[NPP];
#pragma unroll
    for (int i = 0; i < NPP; ++i) Uk[i] = y[((size_t)b * Ks + k) * NPP + i];
    double Mt[NA], M[NA][NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      Mt[i] = 0.0;
#pragma unroll
      for (int j = 0; j < NA; ++j) M[i][j] = 0.0;
    }
    if (!linear) {
      for (int q = 0; q < nq; ++q) {  // adj_march.m:78-82,104-105 (mirrored points: quirk C-3)
        double ur = 0.0, ph[NA];
#pragma unroll
        for (int i = 0; i < NPP; ++i) ur = fma(Iq[q * NPP + i], Uk[i], ur);
#pragma unroll
        for (int i = 0; i < NA; ++i) ph[i] = Phi[q * NA + i];
        double sn, cs;
        sincos(ur, &sn, &cs);
        const double ws = w[q] * sn, wc = w[q] * cs;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          Mt[i] = fma(ph[i], ws, Mt[i]);
          const double pw = ph[i] * wc;
#pragma unroll
          for (int j = 0; j < NA; ++j) M[i][j] = fma(pw, ph[j], M[i][j]);
        }
      }
    }
    double F[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      F[i] = f1[i] - ((i == nak - 1) ? vL : 0.0);          // F = M_k*1; F(end) -= vL_prev
#pragma unroll
      for (int j = 0; j < NA; ++j) M[i][j] = fma(-hk2, M[i][j], A0[i * NA + j]);   // A = A0 - M_v
    }
    solve_dense<NA>(M, F);                                  // v_k = A \ F
    vL = F[0];
    // indicator: err_k = v_k' * ( -A2*uh_k - M_tilde + F0 ),  F0(1) = y0 (k == 1) or y1{k-1}(end)
    double uh[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NPP; ++j) s = fma(Ix[i * NPP + j], Uk[j], s);
      uh[i] = s;
    }
    const double f0 = (k == 0) ? (y0_arr ? y0_arr[b] : y0_hard) : y[((size_t)b * Ks + (k - 1)) * NPP + lastprev];
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = fma(-hk2, Mt[i], (i == 0) ? f0 : 0.0);
#pragma unroll
      for (int j = 0; j < NA; ++j) s = fma(-A2[i * NA + j], uh[j], s);
      e = fma(F[i], s, e);
      if (v) v[((size_t)b * Ks + k) * NA + i] = F[i];
    }
    if (err) err[(size_t)b * Ks + k] = e;
  }
}

// solve_dense in two parts: everything that does not touch the right-hand side (pivot search, row swaps of
// the matrix, the multipliers, the reciprocals of the pivots) and the part that does.  lu_apply performs on b
// exactly the operations solve_dense performs on it, in the same order, so solve_dense(A, b) and
// lu_factor(A, ..) + lu_apply(A, .., b) give the same bits (tests/lu_split_check.c restates both on the host and
// compares them on random systems with row swaps).
template <int N>
__device__ __forceinline__ void lu_factor(double (&A)[N][N], unsigned long long& swmask, double (&rcp)[N]) {
  swmask = 0ull;   // (bit c N + r: up to 41 at N = 7)
#pragma unroll
  for (int c = 0; c < N; ++c) {
#pragma unroll
    for (int r = c + 1; r < N; ++r) {
      const bool sw = fabs(A[r][c]) > fabs(A[c][c]);
      // (the multipliers already stored in columns < c travel with their rows: solve_dense has applied them to b
      //  before this swap, so lu_apply replays swaps and eliminations in the original order, column by column)
#pragma unroll
      for (int j = c; j < N; ++j) {
        const double t = A[c][j];
        A[c][j] = sw ? A[r][j] : t;
        A[r][j] = sw ? t : A[r][j];
      }
      swmask |= sw ? (1ull << (c * N + r)) : 0ull;
    }
    rcp[c] = 1.0 / A[c][c];
#pragma unroll
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] * rcp[c];
#pragma unroll
      for (int j = c + 1; j < N; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
      A[r][c] = f;   // kept in the eliminated position
    }
  }
}
template <int N>
__device__ __forceinline__ void lu_apply(const double (&A)[N][N], unsigned long long swmask, const double (&rcp)[N], double (&b)[N]) {
#pragma unroll
  for (int c = 0; c < N; ++c) {
#pragma unroll
    for (int r = c + 1; r < N; ++r) {
      const bool sw = (swmask >> (c * N + r)) & 1ull;
      const double tb = b[c];
      b[c] = sw ? b[r] : tb;
      b[r] = sw ? tb : b[r];
    }
#pragma unroll
    for (int r = c + 1; r < N; ++r) b[r] = fma(-A[r][c], b[c], b[r]);
  }
#pragma unroll
  for (int r = N - 1; r >= 0; --r) {
    double s = b[r];
#pragma unroll
    for (int j = r + 1; j < N; ++j) s = fma(-A[r][j], b[j], s);
    b[r] = div_rn_q(s, A[r][r], rcp[r]);
  }
}

// Small batches: one WARP per trajectory, one LANE per element.  Everything expensive in an element of the adjoint
// march depends on the PRIMAL only -- the nq sin / cos evaluations, the mass-type matrices, the factorisation of the
// element matrix, the weights of the indicator -- and is done for 32 elements at once; what is sequential is the
// right-hand side (the inflow value vL of the element behind) and the substitution, done by the lane that owns the
// element while the others wait, vL handed on by a shuffle.  Same arithmetic per element as tdg_adjoint_kernel, in
// the same order: bit-identical outputs.
template <int NPP>
__global__ void tdg_adjoint_warp_kernel(long long B, int Ks, int nq, int linear, double y0_hard,
                                        const double* __restrict__ y0_arr, const double* __restrict__ ec,
                                        const double* __restrict__ y, double* __restrict__ v, double* __restrict__ err) {
  constexpr int NA = NPP + 1;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;   // whole warps leave together
  const int blk = NA * NA + NA + NA * NA + NA * NPP + nq * NPP + nq * NA + nq + 3;
  double vL = 0.0;
  for (int base = ((Ks - 1) >> 5) << 5; base >= 0; base -= 32) {
    const int k = base + lane;
    const bool act = k < Ks;
    const int kk = act ? k : Ks - 1;   // idle lanes repeat the last element (results unused)
    const double* A0 = ec + (size_t)kk * blk;
    const double* f1 = A0 + NA * NA;
    const double* A2 = f1 + NA;
    const double* Ix = A2 + NA * NA;
    const double* Iq = Ix + NA * NPP;
    const double* Phi = Iq + nq * NPP;
    const double* w = Phi + nq * NA;
    const double hk2 = 0.5 * w[nq];
    const int nak = (int)w[nq + 1], lastprev = (int)w[nq + 2];
    double Uk[NPP];
#pragma unroll
    for (int i = 0; i < NPP; ++i) Uk[i] = y[((size_t)b * Ks + kk) * NPP + i];
    double Mt[NA], M[NA][NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      Mt[i] = 0.0;
#pragma unroll
      for (int j = 0; j < NA; ++j) M[i][j] = 0.0;
    }
    if (!linear) {
      for (int q = 0; q < nq; ++q) {
        double ur = 0.0, ph[NA];
#pragma unroll
        for (int i = 0; i < NPP; ++i) ur = fma(Iq[q * NPP + i], Uk[i], ur);
#pragma unroll
        for (int i = 0; i < NA; ++i) ph[i] = Phi[q * NA + i];
        double sn, cs;
        sincos(ur, &sn, &cs);
        const double ws = w[q] * sn, wc = w[q] * cs;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          Mt[i] = fma(ph[i], ws, Mt[i]);
          const double pw = ph[i] * wc;
#pragma unroll
          for (int j = 0; j < NA; ++j) M[i][j] = fma(pw, ph[j], M[i][j]);
        }
      }
    }
    double F0[NA], Sw[NA], rcp[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      F0[i] = f1[i];
#pragma unroll
      for (int j = 0; j < NA; ++j) M[i][j] = fma(-hk2, M[i][j], A0[i * NA + j]);
    }
    unsigned long long swmask;
    lu_factor<NA>(M, swmask, rcp);
    double uh[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NPP; ++j) s = fma(Ix[i * NPP + j], Uk[j], s);
      uh[i] = s;
    }
    const double f0 = (kk == 0) ? (y0_arr ? y0_arr[b] : y0_hard) : y[((size_t)b * Ks + (kk - 1)) * NPP + lastprev];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = fma(-hk2, Mt[i], (i == 0) ? f0 : 0.0);
#pragma unroll
      for (int j = 0; j < NA; ++j) s = fma(-A2[i * NA + j], uh[j], s);
      Sw[i] = s;
    }
    // ---- the sequential part: elements base+31 .. base, each by its lane
    const int top = (Ks - 1 - base) < 31 ? (Ks - 1 - base) : 31;
    for (int src = top; src >= 0; --src) {
      double v0 = 0.0;
      if (lane == src) {
        double F[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) F[i] = F0[i] - ((i == nak - 1) ? vL : 0.0);
        lu_apply<NA>(M, swmask, rcp, F);
        v0 = F[0];
        double e = 0.0;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          e = fma(F[i], Sw[i], e);
          if (v) v[((size_t)b * Ks + k) * NA + i] = F[i];
        }
        if (err) err[(size_t)b * Ks + k] = e;
      }
      vL = __shfl_sync(0xffffffffu, v0, src);
    }
  }
}

// Radau-reconstructed adjoint, matlab/adj_rec.m:18-71 (linear branch; the reference's nonlinear
// branch is unfinished and returns zeros, handled by the host): adjoint solved at the PRIMAL
// order, interpolated to the element's Radau points, extended by the inflow value and
// re-interpolated at the nodes of order N+1, where the indicator is evaluated.
//   block: A0[NP*NP] | f1[NP] | R[NP*NP] | H[NA*NA] | A2[NA*NA] | Ix[NA*NP] | np_k | last_{k-1}
// (NA = NP + 1; same padding rules as above).  v[b][k][:] = [v at the Radau points; v at t_{k+1}].
template <int NP>
__global__ void tdg_adjrec_kernel(long long B, int Ks, double y0_hard, const double* __restrict__ y0_arr,
                                  const double* __restrict__ ec,
                                  const double* __restrict__ y, double* __restrict__ v,
                                  double* __restrict__ err) {
  constexpr int NA = NP + 1;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  constexpr int blk = NP * NP + NP + NP * NP + NA * NA + NA * NA + NA * NP + 2;
  double vL = 0.0;
  for (int k = Ks - 1; k >= 0; --k) {
    const double* A0 = ec + (size_t)k * blk;
    const double* f1 = A0 + NP * NP;
    const double* R = f1 + NP;
    const double* H = R + NP * NP;
    const double* A2 = H + NA * NA;
    const double* Ix = A2 + NA * NA;
    const int npk = (int)Ix[NA * NP], lastprev = (int)Ix[NA * NP + 1];
    double M[NP][NP], F[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      F[i] = f1[i] - ((i == npk - 1) ? vL : 0.0);           // adj_rec.m:31
#pragma unroll
      for (int j = 0; j < NP; ++j) M[i][j] = A0[i * NP + j];
    }
    solve_dense<NP>(M, F);                                   // :40  v_s = A \ F
    double vrec[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = (i == npk) ? vL : 0.0;                      // :44  v_rec(end+1) = vL_prev
      if (i < NP) {
#pragma unroll
        for (int j = 0; j < NP; ++j) s = fma(R[i * NP + j], F[j], s);
      }
      vrec[i] = s;
    }
    double Uk[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) Uk[i] = y[((size_t)b * Ks + k) * NP + i];
    const double f0 = (k == 0) ? (y0_arr ? y0_arr[b] : y0_hard) : y[((size_t)b * Ks + (k - 1)) * NP + lastprev];
    double uh[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NP; ++j) s = fma(Ix[i * NP + j], Uk[j], s);
      uh[i] = s;
    }
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double vh = 0.0, r = (i == 0) ? f0 : 0.0;
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        vh = fma(H[i * NA + j], vrec[j], vh);                // :65
        r = fma(-A2[i * NA + j], uh[j], r);                  // :66
      }
      e = fma(vh, r, e);
      if (v) v[((size_t)b * Ks + k) * NA + i] = vrec[i];
    }
    if (err) err[(size_t)b * Ks + k] = e;
    vL = vrec[0];                                            // :69
  }
}

// matlab/err_contribution.m:1-50 (unused by the reference, MAIN.m:50): err_i = int over element i of
// a(t) (u_h - u_h')(t) dt with the exact adjoint a(t) = e^{1-t} - 1 of a' = -a - 1, a(1) = 0 (:23-25),
// plus u(1) - 1 on the first element (:42-43).  u_h is the polyfit/polyval interpolant (:10-14), so
// the integral is linear in the element's nodal values: err[b][k] = cvec_k . y[b][k][:], the weight
// vectors built by the host (Gauss quadrature in place of MATLAB's adaptive `integral`).
__global__ void tdg_errcon_kernel(long long B, int Ks, int Np, const double* __restrict__ cvec,
                                  const double* __restrict__ y, double* __restrict__ err) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * Ks) return;
  const int k = (int)(t % Ks);
  const double* yk = y + (size_t)t * Np;
  double e = 0.0;
  for (int i = 0; i < Np; ++i) e = fma(cvec[(size_t)k * Np + i], yk[i], e);
  if (k == 0) e += yk[0] - 1.0;
  err[t] = e;
}

// The constant blocks of a call on the device: found among the handle's last few (same length, same bytes), or
// uploaded into the least recently used slot.  A hit costs a memcmp of a few KB; a miss the pageable upload (staged by
// the driver before cudaMemcpyAsync returns, so the caller may release its buffer; stream-ordered after the kernels
// that may still read the slot).
static int tdg_consts(dgadj_handle* h, const double* host, size_t n, cudaStream_t st, const double** dev_out) {
  const size_t need = n * sizeof(double);
  dgadj_handle::TdgConstSlot* lru = &h->tdg_cache[0];
  for (auto& sl : h->tdg_cache) {
    if (sl.dev && sl.n == n && memcmp(sl.host, host, need) == 0) {
      sl.stamp = ++h->tdg_clock;
      *dev_out = sl.dev;
      return DGADJ_OK;
    }
    if (sl.stamp < lru->stamp) lru = &sl;
  }
  if (n > lru->cap) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(lru->dev);
    free(lru->host);
    lru->dev = nullptr;
    lru->host = nullptr;
    lru->cap = lru->n = 0;
    lru->host = (double*)malloc(need);
    if (!lru->host) return fail(h, DGADJ_ERR_NOMEM, "host copy of the time-DG constant blocks (%zu B)", need);
    CUDA_TRY(h, cudaMalloc((void**)&lru->dev, need));
    lru->cap = n;
  }
  lru->n = 0;   // (not a valid entry until the upload is queued)
  memcpy(lru->host, host, need);
  CUDA_TRY(h, cudaMemcpyAsync(lru->dev, host, need, cudaMemcpyHostToDevice, st));
  lru->n = n;
  lru->stamp = ++h->tdg_clock;
  *dev_out = lru->dev;
  return DGADJ_OK;
}

// Lane groups / warps per trajectory below this batch size: the thread-per-trajectory kernels need
// ~150k trajectories to fill the device.  dgadj_set_tuning(block_threads = 1 / 32) forces either form.
constexpr long long TDG_WARP_BATCH = 16384;
static int tdg_launch_march(dgadj_handle* h, long long B, int Ks, int Np, int nq, int linear, double tol, int maxit,
                            const double* ec_dev, const double* y0_dev, double* y_dev, int* its_dev, cudaStream_t st) {
  const bool warp = !linear && (h->tune_block == 32 || (h->tune_block != 1 && B <= TDG_WARP_BATCH));
  if (warp) {
    const int block = 128;
    constexpr int G = 16;   // lanes per trajectory
#define DGADJ_TDG_W(n) case n: tdg_march_group_kernel<n, G><<<(unsigned)((B * G + block - 1) / block), block, 0, st>>>(B, Ks, nq, tol, maxit, ec_dev, y0_dev, y_dev, its_dev); break;
    switch (Np) { DGADJ_TDG_W(2) DGADJ_TDG_W(3) DGADJ_TDG_W(4) DGADJ_TDG_W(5) DGADJ_TDG_W(6) }
#undef DGADJ_TDG_W
  } else {
    const int block = 128;
    const unsigned grid = (unsigned)((B + block - 1) / block);
#define DGADJ_TDG_M(n) case n: tdg_march_kernel<n><<<grid, block, 0, st>>>(B, Ks, nq, linear, tol, maxit, ec_dev, y0_dev, y_dev, its_dev); break;
    switch (Np) { DGADJ_TDG_M(2) DGADJ_TDG_M(3) DGADJ_TDG_M(4) DGADJ_TDG_M(5) DGADJ_TDG_M(6) }
#undef DGADJ_TDG_M
  }
  CUDA_TRY(h, cudaGetLastError());
  return DGADJ_OK;
}

static int tdg_launch_adjoint(dgadj_handle* h, long long B, int Ks, int Npp, int nq, int linear, double y0_hard,
                              const double* y0_dev, const double* ec_dev, const double* y_dev, double* v_dev,
                              double* err_dev, cudaStream_t st) {
  const bool warp = h->tune_block == 32 || (h->tune_block != 1 && B <= TDG_WARP_BATCH);
  if (warp) {   // one warp per trajectory, one lane per element
    const int block = 128;
    const unsigned grid = (unsigned)((B * 32 + block - 1) / block);
#define DGADJ_TDG_AW(n) case n: tdg_adjoint_warp_kernel<n><<<grid, block, 0, st>>>(B, Ks, nq, linear, y0_hard, y0_dev, ec_dev, y_dev, v_dev, err_dev); break;
    switch (Npp) { DGADJ_TDG_AW(2) DGADJ_TDG_AW(3) DGADJ_TDG_AW(4) DGADJ_TDG_AW(5) DGADJ_TDG_AW(6) }
#undef DGADJ_TDG_AW
  } else {
    // (forced thread form on a small batch: 32 threads per CTA spread the trajectories over four times as many SMs)
    const int block = (B <= TDG_WARP_BATCH) ? 32 : 128;
    const unsigned grid = (unsigned)((B + block - 1) / block);
#define DGADJ_TDG_A(n) case n: tdg_adjoint_kernel<n><<<grid, block, 0, st>>>(B, Ks, nq, linear, y0_hard, y0_dev, ec_dev, y_dev, v_dev, err_dev); break;
    switch (Npp) { DGADJ_TDG_A(2) DGADJ_TDG_A(3) DGADJ_TDG_A(4) DGADJ_TDG_A(5) DGADJ_TDG_A(6) }
#undef DGADJ_TDG_A
  }
  CUDA_TRY(h, cudaGetLastError());
  return DGADJ_OK;
}

}  // namespace dgadj

extern "C" int dgadj_tdg_march(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np, int32_t nq, int32_t linear,
                               double tol, int32_t maxit, const double* elem_consts_host,
                               const double* y0_dev, double* y_dev, int32_t* its_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || Ks <= 0 || nq < 0 || !elem_consts_host || !y0_dev || !y_dev) return fail(h, DGADJ_ERR_INVALID, "bad tdg_march arguments");
  if (Np < 2 || Np > 6) return fail(h, DGADJ_ERR_UNSUPPORTED, "time-DG march supports 1 <= N <= 5 (Np = %d)", Np);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t blk = (size_t)Np * Np + 2 * (size_t)nq * Np + nq + 2;
  const double* ec = nullptr;
  int rc = tdg_consts(h, elem_consts_host, blk * Ks, st, &ec);
  if (rc) return rc;
  rc = tdg_launch_march(h, B, Ks, Np, nq, linear, tol, maxit, ec, y0_dev, y_dev, its_dev, st);
  if (rc) return rc;
  h->launches++;
  return DGADJ_OK;
}

extern "C" int dgadj_tdg_adjoint(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np_primal, int32_t nq,
                                 int32_t linear, double y0_hard, const double* y0_dev,
                                 const double* elem_consts_host, const double* y_dev, double* v_dev,
                                 double* err_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || Ks <= 0 || nq < 0 || !elem_consts_host || !y_dev) return fail(h, DGADJ_ERR_INVALID, "bad tdg_adjoint arguments");
  if (Np_primal < 2 || Np_primal > 6) return fail(h, DGADJ_ERR_UNSUPPORTED, "time-DG adjoint supports primal 1 <= N <= 5");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Na = Np_primal + 1;
  const size_t blk = Na * Na + Na + Na * Na + Na * Np_primal + (size_t)nq * Np_primal + (size_t)nq * Na + nq + 3;
  const double* ec = nullptr;
  int rc = tdg_consts(h, elem_consts_host, blk * Ks, st, &ec);
  if (rc) return rc;
  rc = tdg_launch_adjoint(h, B, Ks, Np_primal, nq, linear, y0_hard, y0_dev, ec, y_dev, v_dev, err_dev, st);
  if (rc) return rc;
  h->launches++;
  return DGADJ_OK;
}

extern "C" int dgadj_tdg_adjoint_rec(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np_primal, double y0_hard,
                                     const double* y0_dev, const double* elem_consts_host, const double* y_dev,
                                     double* v_dev, double* err_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || Ks <= 0 || !elem_consts_host || !y_dev) return fail(h, DGADJ_ERR_INVALID, "bad tdg_adjoint_rec arguments");
  // utils/Globals1D.m:37-42 tabulates Radau points up to m = 5  =>  N <= 4
  if (Np_primal < 2 || Np_primal > 5) return fail(h, DGADJ_ERR_UNSUPPORTED, "adj_rec supports 1 <= N <= 4 (Radau tables of Globals1D.m)");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Np = Np_primal, Na = Np + 1;
  const size_t blk = Np * Np + Np + Np * Np + Na * Na + Na * Na + Na * Np + 2;
  const double* ec = nullptr;
  int rc = tdg_consts(h, elem_consts_host, blk * Ks, st, &ec);
  if (rc) return rc;
  const int block = 128;
  const unsigned grid = (unsigned)((B + block - 1) / block);
#define DGADJ_TDG_R(n) case n: tdg_adjrec_kernel<n><<<grid, block, 0, st>>>(B, Ks, y0_hard, y0_dev, ec, y_dev, v_dev, err_dev); break;
  switch (Np_primal) { DGADJ_TDG_R(2) DGADJ_TDG_R(3) DGADJ_TDG_R(4) DGADJ_TDG_R(5) }
#undef DGADJ_TDG_R
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return DGADJ_OK;
}

extern "C" int dgadj_tdg_err_contribution(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np,
                                          const double* cvec_host, const double* y_dev, double* err_dev,
                                          void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || Ks <= 0 || Np < 2 || !cvec_host || !y_dev || !err_dev) return fail(h, DGADJ_ERR_INVALID, "bad tdg_err_contribution arguments");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const double* ec = nullptr;
  int rc = tdg_consts(h, cvec_host, (size_t)Ks * Np, st, &ec);
  if (rc) return rc;
  const int block = 128;
  const long long n = (long long)B * Ks;
  tdg_errcon_kernel<<<(unsigned)((n + block - 1) / block), block, 0, st>>>(B, Ks, Np, ec, y_dev, err_dev);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return DGADJ_OK;
}

// =======================================================================================================
// Device-resident adaptive refinement loop of matlab/MAIN.m:29-166, batched over the
// initial values on a shared mesh with the batch-mean indicator of python/Main_variable_params.py:340-341:
//   per iteration  build the element blocks of the current mesh -> dg_march (MAIN.m:32) -> adj_march at order
//                  n+1 with the indicator (MAIN.m:34) -> batch mean of |err| (MAIN.m:51) -> argmax element,
//                  lowest index on ties, midpoint insertion (MAIN.m:137-141)
// all enqueued on the caller's stream by ONE call: the mesh lives on the device, the number of elements of
// iteration `it` is Ks0 + it whatever gets refined, so no launch shape depends on a result and nothing is
// read back inside the loop.
// Element blocks: every entry of a block is affine in the element width h (fem_setup.m:27-39 operators are
// those of the reference element; M_k = h/2 M, hk = +-h; the polyfit/polyval interpolation matrices of
// dg_march.m:47-49 / adj_march.m:75-79 are invariant under the affine map of an element), so the host hands
// over two template blocks T0, T1 per kind and the device forms block_k = T0 + h_k T1.
// =======================================================================================================
namespace dgadj {

__global__ void tdg_build_blocks_kernel(int Ks, const double* __restrict__ times, int blk_m, const double* __restrict__ m0,
                                        const double* __restrict__ m1, double* __restrict__ out_m, int blk_a,
                                        const double* __restrict__ a0, const double* __restrict__ a1,
                                        double* __restrict__ out_a) {
  const int k = blockIdx.x;
  if (k >= Ks) return;
  const double h = times[k + 1] - times[k];
  for (int i = threadIdx.x; i < blk_m; i += blockDim.x) out_m[(size_t)k * blk_m + i] = fma(h, m1[i], m0[i]);
  for (int i = threadIdx.x; i < blk_a; i += blockDim.x) {
    double v = fma(h, a1[i], a0[i]);
    if (i == blk_a - 1 && k == 0) v = 0.0;   // last_{k-1}: there is no element before the first
    out_a[(size_t)k * blk_a + i] = v;
  }
}

// batch mean of |err[b][k]| per element (block k; fixed-order tree: the same bits for every launch), the mean
// terminal value, the largest Newton count, and the status of the solve:
//   stats[0] = mean_b y[b][Ks-1][Np-1];  istats[0] = max its;  istats[1] = #elements-solves that hit maxit
//   without converging;  istats[2] = #non-finite values in err
__global__ void tdg_loop_reduce_kernel(long long B, int Ks, int Np, int maxit, const double* __restrict__ err,
                                       const double* __restrict__ y, const int* __restrict__ its,
                                       double* __restrict__ mean_err, double* __restrict__ stats, int* __restrict__ istats) {
  __shared__ double sm[256];
  __shared__ int smi[3][256];
  const int k = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  int mx = 0, nc = 0, bad = 0;
  if (k < Ks) {
    for (long long b = tid; b < B; b += 256) {
      const double e = err[(size_t)b * Ks + k];
      acc += fabs(e);
      bad += !isfinite(e);
      const int it = its[(size_t)b * Ks + k];
      mx = max(mx, it);
      nc += (it > maxit);
    }
  } else {   // block Ks: the terminal value of the primal
    for (long long b = tid; b < B; b += 256) acc += y[((size_t)b * Ks + (Ks - 1)) * Np + (Np - 1)];
  }
  sm[tid] = acc;
  smi[0][tid] = mx;
  smi[1][tid] = nc;
  smi[2][tid] = bad;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      sm[tid] += sm[tid + o];
      smi[0][tid] = max(smi[0][tid], smi[0][tid + o]);
      smi[1][tid] += smi[1][tid + o];
      smi[2][tid] += smi[2][tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (k < Ks) {
      mean_err[k] = sm[0] / (double)B;
      atomicMax(&istats[0], smi[0][0]);
      atomicAdd(&istats[1], smi[1][0]);
      atomicAdd(&istats[2], smi[2][0]);
    } else {
      stats[0] = sm[0] / (double)B;
    }
  }
}

// argmax element (lowest index on ties: find(abs(err)==max(abs(err))) of MAIN.m:137 taken as its first
// entry, np.argmax of Main_finite_difference.py:337) and midpoint insertion (MAIN.m:138-141 /
// Main_finite_difference.py:338-341) into the next row of the mesh history
__global__ void refine_shared_kernel(int K, const double* __restrict__ ind, const double* __restrict__ times,
                                     double* __restrict__ times_next, int* __restrict__ ref_idx, double* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int best = 0;
  double bv = ind[0], tot = 0.0;
  for (int k = 0; k < K; ++k) {
    const double v = ind[k];
    tot += v;
    if (v > bv) {
      bv = v;
      best = k;
    }
  }
  *ref_idx = best;
  if (total) *total = tot;
  if (times_next) {
    for (int j = 0; j <= best; ++j) times_next[j] = times[j];
    times_next[best + 1] = (times[best] + times[best + 1]) / 2.0;
    for (int j = best + 1; j <= K; ++j) times_next[j + 1] = times[j];
  }
}

}  // namespace dgadj

extern "C" int dgadj_tdg_adapt_loop(dgadj_handle* h, const dgadj_tdg_loop_args* a, const double* y0_dev,
                                    double* times_hist_dev, double* err_hist_dev, int32_t* ref_idx_dev,
                                    double* stats_dev, int32_t* istats_dev, double* y_last_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (!a || a->B <= 0 || a->iters < 0 || a->Ks0 < 1 || !a->times0_host || !a->march_T0_host || !a->march_T1_host ||
      !a->adj_T0_host || !a->adj_T1_host || !y0_dev || !times_hist_dev || !err_hist_dev || !ref_idx_dev)
    return fail(h, DGADJ_ERR_INVALID, "bad tdg_adapt_loop arguments");
  const int Np = a->Np, Na = Np + 1;
  if (Np < 2 || Np > 6) return fail(h, DGADJ_ERR_UNSUPPORTED, "time-DG loop supports 1 <= N <= 5 (Np = %d)", Np);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nqm = a->nq_march, nqa = a->nq_adj;
  const size_t blk_m = (size_t)Np * Np + 2 * (size_t)nqm * Np + nqm + 2;
  const size_t blk_a = (size_t)Na * Na + Na + (size_t)Na * Na + (size_t)Na * Np + (size_t)nqa * Np + (size_t)nqa * Na + nqa + 3;
  const int Kmax = a->Ks0 + a->iters;          // elements of the last solve
  const int W = Kmax + 2;                      // row stride of the mesh history
  const long long B = a->B;
  // scratch: templates | blocks of the current mesh | y | its | err | mean
  const size_t n_tmpl = 2 * blk_m + 2 * blk_a;
  const size_t n_d = n_tmpl + (size_t)Kmax * (blk_m + blk_a) + (size_t)B * Kmax * Np + (size_t)B * Kmax + (size_t)Kmax + 8;
  const size_t need = n_d * sizeof(double) + (size_t)B * Kmax * sizeof(int) + 64;
  if (need > h->tdg_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->tdg_scratch);
    h->tdg_scratch = nullptr;
    h->tdg_bytes = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->tdg_scratch, need));
    h->tdg_bytes = need;
  }
  double* d = h->tdg_scratch;
  double *m0 = d, *m1 = m0 + blk_m, *a0 = m1 + blk_m, *a1 = a0 + blk_a;
  double* bm = a1 + blk_a;
  double* ba = bm + (size_t)Kmax * blk_m;
  double* y = ba + (size_t)Kmax * blk_a;
  double* err = y + (size_t)B * Kmax * Np;
  double* mean = err + (size_t)B * Kmax;
  int* its = (int*)(mean + Kmax + 8);
  CUDA_TRY(h, cudaMemcpyAsync(m0, a->march_T0_host, blk_m * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(m1, a->march_T1_host, blk_m * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(a0, a->adj_T0_host, blk_a * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(a1, a->adj_T1_host, blk_a * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(times_hist_dev, a->times0_host, (size_t)(a->Ks0 + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  if (istats_dev) CUDA_TRY(h, cudaMemsetAsync(istats_dev, 0, (size_t)(a->iters + 1) * 3 * sizeof(int32_t), st));
  int rc = DGADJ_OK;
  for (int it = 0; it <= a->iters && rc == DGADJ_OK; ++it) {
    const int Ks = a->Ks0 + it;
    const double* times = times_hist_dev + (size_t)it * W;
    tdg_build_blocks_kernel<<<Ks, 128, 0, st>>>(Ks, times, (int)blk_m, m0, m1, bm, (int)blk_a, a0, a1, ba);
    const bool last = (it == a->iters);
    double* yo = (last && y_last_dev) ? y_last_dev : y;
    rc = tdg_launch_march(h, B, Ks, Np, nqm, a->linear, a->tol, a->maxit, bm, y0_dev, yo, its, st);
    if (rc) break;
    rc = tdg_launch_adjoint(h, B, Ks, Np, nqa, a->linear, a->y0_hard, a->y0_per_trajectory ? y0_dev : nullptr, ba, yo, nullptr, err, st);
    if (rc) break;
    double* mrow = err_hist_dev + (size_t)it * Kmax;
    int* ist = istats_dev ? istats_dev + (size_t)it * 3 : (int*)(mean + Kmax);   // (scratch when not wanted)
    if (!istats_dev) CUDA_TRY(h, cudaMemsetAsync(ist, 0, 3 * sizeof(int), st));
    tdg_loop_reduce_kernel<<<Ks + 1, 256, 0, st>>>(B, Ks, Np, a->maxit, err, yo, its, mrow,
                                                    stats_dev ? stats_dev + (size_t)it * 2 : mean + Kmax + 4, ist);
    refine_shared_kernel<<<1, 32, 0, st>>>(Ks, mrow, times, last ? nullptr : times_hist_dev + (size_t)(it + 1) * W,
                                            ref_idx_dev + it, stats_dev ? stats_dev + (size_t)it * 2 + 1 : nullptr);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 5;
  }
  return rc;
}

// =======================================================================================================
// Per-trajectory meshes (build plan step 8 "second"): every trajectory refines ITS OWN mesh --
// the reference's single-trajectory loop (matlab/MAIN.m:29-166) run for a whole batch at once, no batch rule.
// One warp per trajectory: the lanes build the element block T0 + h T1 of the trajectory's current element in
// shared memory, share the quadrature points (as tdg_march_group_kernel), and repeat the small solves.
// Meshes times[B][W] (W = Ks0 + iters + 2) live on the device; one call enqueues all iterations.
// =======================================================================================================
namespace dgadj {

template <int NP>
__global__ void tdg_march_pt_kernel(long long B, int Ks, int W, int nq, int linear, double tol, int maxit,
                                    const double* __restrict__ T0, const double* __restrict__ T1,
                                    const double* __restrict__ times, const double* __restrict__ y0,
                                    double* __restrict__ y, int ystride, int* __restrict__ its) {
  extern __shared__ double tdg_sm[];
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (b >= B) return;
  const int blk = NP * NP + 2 * nq * NP + nq + 2;
  double* sb = tdg_sm + (size_t)wl * blk;
  const double* tb = times + (size_t)b * W;
  double uR = y0[b];
  for (int k = 0; k < Ks; ++k) {
    const double h = tb[k + 1] - tb[k];
    __syncwarp();
    for (int i = lane; i < blk; i += 32) sb[i] = fma(h, T1[i], T0[i]);
    __syncwarp();
    const double* A = sb;
    const double* Iq = A + NP * NP;
    const double* Phi = Iq + nq * NP;
    const double* w = Phi + nq * NP;
    const double hk2 = 0.5 * w[nq];
    double U[NP];
    int it = 0;
    if (linear) {
      double M[NP][NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        U[i] = (i == 0) ? uR : 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) M[i][j] = A[i * NP + j];
      }
      solve_dense<NP>(M, U);
      it = 1;
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) U[i] = uR;
      double err = 1.0;
      while (it <= maxit && err > tol) {
        double Mt[NP], J[NP][NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          Mt[i] = 0.0;
#pragma unroll
          for (int j = 0; j < NP; ++j) J[i][j] = 0.0;
        }
        for (int q = lane; q < nq; q += 32) {
          double ur = 0.0, ph[NP];
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            ur = fma(Iq[q * NP + i], U[i], ur);
            ph[i] = Phi[q * NP + i];
          }
          double sn, cs;
          sincos(ur, &sn, &cs);
          const double ws = w[q] * sn, wc = w[q] * cs;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            Mt[i] = fma(ph[i], ws, Mt[i]);
            const double pw = ph[i] * wc;
#pragma unroll
            for (int j = i; j < NP; ++j) J[i][j] = fma(pw, ph[j], J[i][j]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            Mt[i] += __shfl_xor_sync(0xffffffffu, Mt[i], o);
#pragma unroll
            for (int j = i; j < NP; ++j) J[i][j] += __shfl_xor_sync(0xffffffffu, J[i][j], o);
          }
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
#pragma unroll
          for (int j = 0; j < i; ++j) J[i][j] = J[j][i];
        }
        double R[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          double r = fma(hk2, Mt[i], (i == 0) ? uR : 0.0);
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            r = fma(A[i * NP + j], U[j], r);
            J[i][j] = fma(hk2, J[i][j], A[i * NP + j]);
          }
          R[i] = r;
        }
        solve_dense<NP>(J, R);
        double e2 = 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          U[i] -= R[i];
          e2 = fma(R[i], R[i], e2);
        }
        err = sqrt(e2);
        ++it;
      }
    }
    uR = U[NP - 1];
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) y[((size_t)b * ystride + k) * NP + i] = U[i];
      if (its) its[(size_t)b * ystride + k] = it;
    }
  }
}

template <int NPP>
__global__ void tdg_adjoint_pt_kernel(long long B, int Ks, int W, int nq, int linear, const double* __restrict__ y0,
                                      const double* __restrict__ T0, const double* __restrict__ T1,
                                      const double* __restrict__ times, const double* __restrict__ y, int ystride,
                                      double* __restrict__ err) {
  constexpr int NA = NPP + 1;
  extern __shared__ double tdg_sm[];
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (b >= B) return;
  const int blk = NA * NA + NA + NA * NA + NA * NPP + nq * NPP + nq * NA + nq + 3;
  double* sb = tdg_sm + (size_t)wl * blk;
  const double* tb = times + (size_t)b * W;
  double vL = 0.0;
  for (int k = Ks - 1; k >= 0; --k) {
    const double h = tb[k + 1] - tb[k];
    __syncwarp();
    for (int i = lane; i < blk; i += 32) sb[i] = fma(h, T1[i], T0[i]);
    __syncwarp();
    const double* A0 = sb;
    const double* f1 = A0 + NA * NA;
    const double* A2 = f1 + NA;
    const double* Ix = A2 + NA * NA;
    const double* Iq = Ix + NA * NPP;
    const double* Phi = Iq + nq * NPP;
    const double* w = Phi + nq * NA;
    const double hk2 = 0.5 * w[nq];
    double Uk[NPP];
#pragma unroll
    for (int i = 0; i < NPP; ++i) Uk[i] = y[((size_t)b * ystride + k) * NPP + i];
    double Mt[NA], M[NA][NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      Mt[i] = 0.0;
#pragma unroll
      for (int j = 0; j < NA; ++j) M[i][j] = 0.0;
    }
    if (!linear) {
      for (int q = lane; q < nq; q += 32) {
        double ur = 0.0, ph[NA];
#pragma unroll
        for (int i = 0; i < NPP; ++i) ur = fma(Iq[q * NPP + i], Uk[i], ur);
#pragma unroll
        for (int i = 0; i < NA; ++i) ph[i] = Phi[q * NA + i];
        double sn, cs;
        sincos(ur, &sn, &cs);
        const double ws = w[q] * sn, wc = w[q] * cs;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          Mt[i] = fma(ph[i], ws, Mt[i]);
          const double pw = ph[i] * wc;
#pragma unroll
          for (int j = i; j < NA; ++j) M[i][j] = fma(pw, ph[j], M[i][j]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          Mt[i] += __shfl_xor_sync(0xffffffffu, Mt[i], o);
#pragma unroll
          for (int j = i; j < NA; ++j) M[i][j] += __shfl_xor_sync(0xffffffffu, M[i][j], o);
        }
      }
#pragma unroll
      for (int i = 0; i < NA; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) M[i][j] = M[j][i];
      }
    }
    double F[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      F[i] = f1[i] - ((i == NA - 1) ? vL : 0.0);
#pragma unroll
      for (int j = 0; j < NA; ++j) M[i][j] = fma(-hk2, M[i][j], A0[i * NA + j]);
    }
    solve_dense<NA>(M, F);
    vL = F[0];
    double uh[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NPP; ++j) s = fma(Ix[i * NPP + j], Uk[j], s);
      uh[i] = s;
    }
    const double f0 = (k == 0) ? y0[b] : y[((size_t)b * ystride + (k - 1)) * NPP + (NPP - 1)];
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = fma(-hk2, Mt[i], (i == 0) ? f0 : 0.0);
#pragma unroll
      for (int j = 0; j < NA; ++j) s = fma(-A2[i * NA + j], uh[j], s);
      e = fma(F[i], s, e);
    }
    if (lane == 0) err[(size_t)b * ystride + k] = e;
  }
}

// each trajectory: argmax |err| (lowest index on ties, MAIN.m:137), midpoint insertion (MAIN.m:138-141)
__global__ void tdg_refine_pt_kernel(long long B, int Ks, int W, int estride, const double* __restrict__ err,
                                     double* __restrict__ times, int* __restrict__ ref_hist, int hstride, int it,
                                     double* __restrict__ tot_hist, bool insert) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* e = err + (size_t)b * estride;
  int best = 0;
  double bv = fabs(e[0]), tot = 0.0;
  for (int k = 0; k < Ks; ++k) {
    const double v = fabs(e[k]);
    tot += v;
    if (v > bv) {
      bv = v;
      best = k;
    }
  }
  ref_hist[(size_t)b * hstride + it] = best;
  if (tot_hist) tot_hist[(size_t)b * hstride + it] = tot;
  if (insert) {
    double* t = times + (size_t)b * W;
    const double mid = (t[best] + t[best + 1]) / 2.0;
    for (int j = Ks; j > best; --j) t[j + 1] = t[j];
    t[best + 1] = mid;
  }
}

}  // namespace dgadj

extern "C" int dgadj_tdg_adapt_loop_pt(dgadj_handle* h, const dgadj_tdg_loop_args* a, const double* y0_dev,
                                       double* times_dev, int32_t* ref_hist_dev, double* tot_hist_dev,
                                       double* y_last_dev, int32_t* its_last_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (!a || a->B <= 0 || a->iters < 0 || a->Ks0 < 1 || !a->march_T0_host || !a->march_T1_host || !a->adj_T0_host ||
      !a->adj_T1_host || !y0_dev || !times_dev || !ref_hist_dev)
    return fail(h, DGADJ_ERR_INVALID, "bad tdg_adapt_loop_pt arguments");
  const int Np = a->Np, Na = Np + 1;
  if (Np < 2 || Np > 6) return fail(h, DGADJ_ERR_UNSUPPORTED, "time-DG loop supports 1 <= N <= 5 (Np = %d)", Np);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nqm = a->nq_march, nqa = a->nq_adj;
  const size_t blk_m = (size_t)Np * Np + 2 * (size_t)nqm * Np + nqm + 2;
  const size_t blk_a = (size_t)Na * Na + Na + (size_t)Na * Na + (size_t)Na * Np + (size_t)nqa * Np + (size_t)nqa * Na + nqa + 3;
  const int Kmax = a->Ks0 + a->iters, W = Kmax + 2;
  const long long B = a->B;
  const size_t n_d = 2 * blk_m + 2 * blk_a + (size_t)B * Kmax * Np + (size_t)B * Kmax + 8;
  const size_t need = n_d * sizeof(double) + (size_t)B * Kmax * sizeof(int) + 64;
  if (need > h->tdg_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->tdg_scratch);
    h->tdg_scratch = nullptr;
    h->tdg_bytes = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->tdg_scratch, need));
    h->tdg_bytes = need;
  }
  double* d = h->tdg_scratch;
  double *m0 = d, *m1 = m0 + blk_m, *a0 = m1 + blk_m, *a1 = a0 + blk_a;
  double* y = a1 + blk_a;
  double* err = y + (size_t)B * Kmax * Np;
  int* its = (int*)(err + (size_t)B * Kmax + 8);
  CUDA_TRY(h, cudaMemcpyAsync(m0, a->march_T0_host, blk_m * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(m1, a->march_T1_host, blk_m * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(a0, a->adj_T0_host, blk_a * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(a1, a->adj_T1_host, blk_a * sizeof(double), cudaMemcpyHostToDevice, st));
  const int wpb = 4, block = 32 * wpb;
  const unsigned grid = (unsigned)((B + wpb - 1) / wpb);
  const size_t sm_m = wpb * blk_m * sizeof(double), sm_a = wpb * blk_a * sizeof(double);
  if (sm_m > 200 * 1024 || sm_a > 200 * 1024) return fail(h, DGADJ_ERR_UNSUPPORTED, "element blocks too large for shared memory");
  for (int it = 0; it <= a->iters; ++it) {
    const int Ks = a->Ks0 + it;
    const bool last = (it == a->iters);
    double* yo = (last && y_last_dev) ? y_last_dev : y;
    int* io = (last && its_last_dev) ? its_last_dev : its;
#define DGADJ_PT_M(n)                                                                                                      \
  case n:                                                                                                                 \
    CUDA_TRY(h, cudaFuncSetAttribute(tdg_march_pt_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_m));     \
    tdg_march_pt_kernel<n><<<grid, block, sm_m, st>>>(B, Ks, W, nqm, a->linear, a->tol, a->maxit, m0, m1, times_dev, y0_dev, yo, Kmax, io); \
    break;
    switch (Np) { DGADJ_PT_M(2) DGADJ_PT_M(3) DGADJ_PT_M(4) DGADJ_PT_M(5) DGADJ_PT_M(6) }
#undef DGADJ_PT_M
#define DGADJ_PT_A(n)                                                                                                      \
  case n:                                                                                                                 \
    CUDA_TRY(h, cudaFuncSetAttribute(tdg_adjoint_pt_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));   \
    tdg_adjoint_pt_kernel<n><<<grid, block, sm_a, st>>>(B, Ks, W, nqa, a->linear, y0_dev, a0, a1, times_dev, yo, Kmax, err);   \
    break;
    switch (Np) { DGADJ_PT_A(2) DGADJ_PT_A(3) DGADJ_PT_A(4) DGADJ_PT_A(5) DGADJ_PT_A(6) }
#undef DGADJ_PT_A
    tdg_refine_pt_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(B, Ks, W, Kmax, err, times_dev, ref_hist_dev, a->iters + 1,
                                                                      it, tot_hist_dev, !last);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 3;
  }
  return DGADJ_OK;
}
