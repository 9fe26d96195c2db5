/* The dense solve of the time-DG kernels (csrc/dgadj_tdg.cu), restated on the host to check two claims bit for bit:
 *  (1) solve_dense with the pivots' reciprocals reused for the back substitution (b[r] = s / A[r][r] formed as
 *      q + (s - A q) r, r = RN(1 / A[r][r])) equals the plain elimination with IEEE divisions;
 *  (2) lu_factor + lu_apply (everything that does not touch the right-hand side done first: the lane-per-element
 *      adjoint kernel) performs on b exactly the operations of the one-piece solve, row swaps included.
 * Random systems of size 2..7, with and without diagonal dominance (so that pivoting happens).              */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define NMAX 7
static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double ur(void) { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0) - 0.5; }
static double div_rn(double x, double h, double r) { const double q = x * r; return fma(fma(-h, q, x), r, q); }

static void solve_plain(int N, double A[NMAX][NMAX], double* b) {   /* the elimination with IEEE divisions */
  for (int c = 0; c < N; ++c) {
    for (int r = c + 1; r < N; ++r) {
      const int sw = fabs(A[r][c]) > fabs(A[c][c]);
      for (int j = 0; j < N; ++j) { const double t = A[c][j]; A[c][j] = sw ? A[r][j] : t; A[r][j] = sw ? t : A[r][j]; }
      const double tb = b[c]; b[c] = sw ? b[r] : tb; b[r] = sw ? tb : b[r];
    }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] * inv;
      for (int j = c + 1; j < N; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
      b[r] = fma(-f, b[c], b[r]);
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double t = b[r];
    for (int j = r + 1; j < N; ++j) t = fma(-A[r][j], b[j], t);
    b[r] = t / A[r][r];
  }
}
static void solve_dense(int N, double A[NMAX][NMAX], double* b) {   /* dgadj_tdg.cu: solve_dense */
  double inv[NMAX];
  for (int c = 0; c < N; ++c) {
    for (int r = c + 1; r < N; ++r) {
      const int sw = fabs(A[r][c]) > fabs(A[c][c]);
      for (int j = 0; j < N; ++j) { const double t = A[c][j]; A[c][j] = sw ? A[r][j] : t; A[r][j] = sw ? t : A[r][j]; }
      const double tb = b[c]; b[c] = sw ? b[r] : tb; b[r] = sw ? tb : b[r];
    }
    inv[c] = 1.0 / A[c][c];
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] * inv[c];
      for (int j = c + 1; j < N; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
      b[r] = fma(-f, b[c], b[r]);
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double t = b[r];
    for (int j = r + 1; j < N; ++j) t = fma(-A[r][j], b[j], t);
    b[r] = div_rn(t, A[r][r], inv[r]);
  }
}
static void lu_factor(int N, double A[NMAX][NMAX], uint64_t* swmask, double* rcp) {   /* dgadj_tdg.cu: lu_factor */
  *swmask = 0;
  for (int c = 0; c < N; ++c) {
    for (int r = c + 1; r < N; ++r) {
      const int sw = fabs(A[r][c]) > fabs(A[c][c]);
      for (int j = c; j < N; ++j) { const double t = A[c][j]; A[c][j] = sw ? A[r][j] : t; A[r][j] = sw ? t : A[r][j]; }
      *swmask |= sw ? ((uint64_t)1 << (c * N + r)) : 0;
    }
    rcp[c] = 1.0 / A[c][c];
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] * rcp[c];
      for (int j = c + 1; j < N; ++j) A[r][j] = fma(-f, A[c][j], A[r][j]);
      A[r][c] = f;
    }
  }
}
static void lu_apply(int N, double A[NMAX][NMAX], uint64_t swmask, const double* rcp, double* b) {   /* lu_apply */
  for (int c = 0; c < N; ++c) {
    for (int r = c + 1; r < N; ++r) {
      const int sw = (int)((swmask >> (c * N + r)) & 1u);
      const double tb = b[c]; b[c] = sw ? b[r] : tb; b[r] = sw ? tb : b[r];
    }
    for (int r = c + 1; r < N; ++r) b[r] = fma(-A[r][c], b[c], b[r]);
  }
  for (int r = N - 1; r >= 0; --r) {
    double t = b[r];
    for (int j = r + 1; j < N; ++j) t = fma(-A[r][j], b[j], t);
    b[r] = div_rn(t, A[r][r], rcp[r]);
  }
}

int main(void) {
  long bad1 = 0, bad2 = 0, swaps = 0, n = 0;
  for (int trial = 0; trial < 200000; ++trial) {
    const int N = 2 + (int)(rnd() % 6);
    double A[NMAX][NMAX], A1[NMAX][NMAX], A2[NMAX][NMAX], A3[NMAX][NMAX], b[NMAX], b1[NMAX], b2[NMAX], b3[NMAX], rcp[NMAX];
    const int dominant = (int)(rnd() & 1);
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < N; ++j) A[i][j] = ur() + ((dominant && i == j) ? 3.0 : 0.0);
      b[i] = ur() * 10.0;
    }
    memcpy(A1, A, sizeof(A)); memcpy(A2, A, sizeof(A)); memcpy(A3, A, sizeof(A));
    memcpy(b1, b, sizeof(b)); memcpy(b2, b, sizeof(b)); memcpy(b3, b, sizeof(b));
    solve_plain(N, A1, b1);
    solve_dense(N, A2, b2);
    uint64_t sw;
    lu_factor(N, A3, &sw, rcp);
    lu_apply(N, A3, sw, rcp, b3);
    swaps += sw != 0;
    for (int i = 0; i < N; ++i) {
      if (memcmp(&b1[i], &b2[i], sizeof(double)) && !(b1[i] == 0.0 && b2[i] == 0.0)) ++bad1;
      if (memcmp(&b2[i], &b3[i], sizeof(double))) ++bad2;
    }
    ++n;
  }
  printf("%ld systems (%ld with row swaps): %ld differences to the IEEE divisions, %ld between the one-piece and the split solve\n",
         n, swaps, bad1, bad2);
  return (bad1 || bad2 || swaps < n / 10) ? 1 : 0;
}
