// dgadj_burgers.cu -- inviscid Burgers DG forward march with the reference's SlopeLimitN
// (utils/SlopeLimitN.m:9-32, SlopeLimitLin.m:10-18, minmod.m:6-12) fused after every LSERK4
// stage, writing the forward checkpoints an adjoint needs: per-step states, per-stage limiter
// flags and wave speeds (BASELINE config 3).  The reference ships the limiter but no Burgers
// right-hand side; the RHS is build-specified (oracle/burgers.py):
//     f = u^2/2, C = max|u| over the mesh,  flux = nx (f^- - f^+)/2 - C/2 (u^- - u^+),
//     rhs = -rx o (Dr f) + LIFT (Fscale o flux).
// One CTA marches one trajectory, one thread owns one element (nodal state in registers).
// Dr f runs through the even/odd blocks of the advection kernels (dgadj_kernels.cuh).  Two
// exchanges per stage: traces + the CTA-wide max|u| before the RHS, cell averages before the
// limiter.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "dgadj_internal.h"

// resident CTAs per SM the 256-thread kernels are compiled for (measured on config 3, N = 4:
// forward 4.90 / 4.77 / 5.22 / 5.46e10 updates/s at 1 / 2 / 3 / 4, adjoint 2.17 / 3.21 / 2.75 / 2.69e10)
#ifndef BG_MINB_FWD
#define BG_MINB_FWD(NP) ((NP) <= 5 ? 4 : ((NP) <= 7 ? 2 : 1))
#endif
#ifndef BG_MINB_ADJ
#define BG_MINB_ADJ(NP) ((NP) <= 5 ? 2 : 1)
#endif

namespace dgadj {

struct BurgersArgs {
  long long B;
  int K, S, periodic, limit;   // limit: 0 none, 1 SlopeLimitN (detect, then limit), 2 SlopeLimit1 (every cell)
  double tvbM;                 // M of minmodB (utils/minmodB.m:6-11) in the slope minmod; 0 = plain minmod
  double eps0;                 // detection threshold of SlopeLimitN.m:13 (1e-8); -1 in SlopeLimit1 mode: every cell
  double dt;
  const double* dt_arr;
  const double* rxk;   // [K]
  const double* fs0;   // [K]
  const double* fs1;   // [K]
  const double* xc;    // [NP][K]  x - x0 (SlopeLimitLin.m:11-12)
  double xcn[MAXNP];   // x - x0 of an element in units of h/2 (= the reference nodes r): the map is affine
  const double* hk;    // [K]      x(Np,k) - x(1,k)
  const double* u0;    // [B][NP][K]
  double* uT;          // [B][NP][K]
  double* hist;        // [B][S+1][NP][K] or null
  unsigned short* lim;   // [B][S][K] bits 0-4: limited after stage s; bits 5+2s..6+2s: winning
                         // minmod argument (0 = slope zero, 1 = ux, 2 = (v+ - v)/h, 3 = (v - v-)/h)
  unsigned char* lim0;   // [B][K] same code for the limiter pass on the initial state (bit 0, bits 1-2)
  int* amax;             // [B][S][5] flat index i*K+k of max|u| per stage, +1, negated if u < 0
  double* maxvel;      // [B][S][5] or null
  // adjoint sweep
  const double* jw;      // [NP][K] weights of J = sum jw o u(T)
  double* lam0;          // [B][NP][K] dJ/du0
  double* Jout;          // [B]
  double* stage_scratch; // [grid][5][NP+2][blockDim] stage input states + neighbour traces
  double Dr[MAXNP * MAXNP];  // nodal Dr (adjoint volume term)
  double LIFT[MAXNP * 2];
  StageOps so;
  double aw[MAXNP];    // cell average weights  V(1,1)*invV(1,:)
  double sl[MAXNP];    // slope weights         Dr(1,:)*V(:,1:2)*invV(1:2,:)
  double rka[5], rkb[5];
};

// utils/minmod.m:7-11: s = sum(sign(v))/3; |s| == 1 -> s * min|v|, else 0.  |s| == 1 exactly when
// the three arguments are all > 0 or all < 0, and then s * min|v| is min(v) resp. max(v) with
// the same bits (a product with +-1 is exact): no division, no sign arithmetic.  The two calls
// of SlopeLimitN.m:21-22 share their last two arguments (v - v_{k-1}, v_{k+1} - v).
struct MinmodBC {
  bool pos, neg;   // b, c both > 0 / both < 0
  double t;        // min(b, c) if pos, max(b, c) otherwise: the bound the first argument competes with
};
__device__ __forceinline__ MinmodBC minmod_bc(double b, double c) {
  MinmodBC r;
  r.pos = (b > 0.0) & (c > 0.0);
  r.neg = (b < 0.0) & (c < 0.0);
  r.t = ((b < c) == r.pos) ? b : c;
  return r;
}
__device__ __forceinline__ double minmod3(double a, const MinmodBC& q) {
  const double r = ((a < q.t) == q.pos) ? a : q.t;   // min(a, t) if pos, max(a, t) if neg
  return ((q.pos & (a > 0.0)) | (q.neg & (a < 0.0))) ? r : 0.0;
}

// minmod of three with the index (1..3) of the winning argument (first smallest |v|, as
// MATLAB's min picks it); 0 when the result is zero
__device__ __forceinline__ double minmod3b(double a, double b, double c, int* br) {
  const bool pos = (a > 0.0) & (b > 0.0) & (c > 0.0), neg = (a < 0.0) & (b < 0.0) & (c < 0.0);
  *br = 0;
  if (!(pos | neg)) return 0.0;
  const double fa = fabs(a), fb = fabs(b), fc = fabs(c);
  double m = fa;
  int w = 1;
  if (fb < m) { m = fb; w = 2; }
  if (fc < m) { m = fc; w = 3; }
  *br = w;
  return pos ? m : -m;
}

// warp-wide max of values that are >= 0 or exactly -1.0 (the filler of idle lanes): for those the
// bit pattern is ordered like the value (high word as a signed int, then the low word), so two
// integer warp reductions (redux.sync) replace a five-step shuffle tree of double compares
__device__ __forceinline__ double warp_max_nonneg(double m) {
  const int hi = __double2hiint(m);
  const int mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned lo = (hi == mh) ? (unsigned)__double2loint(m) : 0u;
  const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
  return __hiloint2double(mh, (int)ml);
}

// One LSERK4 stage update of an element, the step size folded into the thread's constants:
//   c1 = -rx dt/4, k0 = -Fscale(1,k) dt/8, k1 = +Fscale(2,k) dt/8.
// With F = u^2 (= 2f), C = max|u|:  dt/2 Fscale_f flux_f = k_f (u^- - u^+)(u^- + u^+ +- 2C), and in
// the even/odd basis  dt rhs_i = E_i + O_i,  dt rhs_{N-i} = E_i - O_i,  dt rhs_mid = 2 E_mid  with
//   E_i = c1 (DE Fo)_i + LS_i (G0 + G1),   O_i = c1 (DO Fe)_i + LA_i (G0 - G1).
// Shared by the forward march and the adjoint's recomputation, so both see the same bits.
struct BgCoef {
  double c1, k0, k1;
};
template <int NP>
__device__ __forceinline__ void burgers_stage_update(const BurgersArgs& p, const BgCoef& c, double (&u)[NP],
                                                     double (&res)[NP], double uL, double uR, double maxvel,
                                                     double rka, double rkb) {
  constexpr int HE = (NP + 1) / 2, HO = NP / 2;
  const double twoC = maxvel + maxvel;
  const double G0 = (c.k0 * (u[0] - uL)) * ((u[0] + uL) + twoC);
  const double G1 = (c.k1 * (u[NP - 1] - uR)) * ((u[NP - 1] + uR) - twoC);
  const double ge = G0 + G1, go = G0 - G1;
  double fe[HE], fo[HO > 0 ? HO : 1];
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    const double ti = c.c1 * u[i], tj = c.c1 * u[NP - 1 - i];
    const double a = ti * u[i];
    fe[i] = fma(tj, u[NP - 1 - i], a);
    fo[i] = fma(-tj, u[NP - 1 - i], a);
  }
  if (NP & 1) fe[NP / 2] = (c.c1 * u[NP / 2]) * u[NP / 2];
  double E[HE], O[HO > 0 ? HO : 1];
#pragma unroll
  for (int i = 0; i < HE; ++i) {
    const double2 l2 = p.so.LS2[i / 2];
    double acc = ((i & 1) ? l2.y : l2.x) * ge;
#pragma unroll
    for (int j = 0; j < HO; ++j) {
      const double2 c2 = p.so.DE2[i * HP + j / 2];
      acc = fma((j & 1) ? c2.y : c2.x, fo[j], acc);
    }
    E[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < HO; ++i) {
    const double2 l2 = p.so.LA2[i / 2];
    double acc = ((i & 1) ? l2.y : l2.x) * go;
#pragma unroll
    for (int j = 0; j < HE; ++j) {
      const double2 c2 = p.so.DO2[i * HP + j / 2];
      acc = fma((j & 1) ? c2.y : c2.x, fe[j], acc);
    }
    O[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    res[i] = fma(rka, res[i], E[i] + O[i]);
    res[NP - 1 - i] = fma(rka, res[NP - 1 - i], E[i] - O[i]);
  }
  if (NP & 1) res[NP / 2] = fma(rka, res[NP / 2], E[NP / 2] + E[NP / 2]);
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = fma(rkb, res[i], u[i]);
}

// (Measured alternative: two trajectories interleaved per thread -- shared barriers / reductions /
// address arithmetic, twice the independent chains, 128 registers, 2 CTAs per SM -- ran at the same
// 7.9e10 updates/s as this one-trajectory form at 4 CTAs per SM: not kept.)
// Shared-memory exchanges alternate between two buffers (`par`): a buffer is rewritten two
// exchanges later, and the barrier of the exchange in between orders that write after every
// read of the earlier one -- one barrier per exchange.
template <int NP, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? BG_MINB_FWD(NP) : 1) burgers_kernel(const __grid_constant__ BurgersArgs p) {
  // traces; trL doubles as the cell-average buffer.  Slots MAXT / MAXT+1 are the ghosts of a
  // non-periodic mesh: the end elements publish their own values there and point nbL / nbR at them
  // (ghost state = own trace; ghost average = own average, SlopeLimitN.m:18)
  __shared__ double trL[2][MAXT + 2], trR[2][MAXT + 2];
  __shared__ double wmax[2][32];
  __shared__ int cand[2];
  const int tid = threadIdx.x, K = p.K;
  const int lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
  const bool in = tid < K;
  const int k = in ? tid : 0;
  const bool gfirst = in && tid == 0 && !p.periodic, glast = in && tid == K - 1 && !p.periodic;
  const int nbL = in ? (tid == 0 ? (p.periodic ? K - 1 : MAXT) : tid - 1) : tid;
  const int nbR = in ? (tid == K - 1 ? (p.periodic ? 0 : MAXT + 1) : tid + 1) : tid;
  const double rx = in ? p.rxk[k] : 0.0, fs0 = in ? p.fs0[k] : 0.0, fs1 = in ? p.fs1[k] : 0.0;
  const double h = in ? p.hk[k] : 1.0;
  const double twoh = 2.0 / h;
  if (tid < 2) cand[tid] = 0x7fffffff;
  int* pend = nullptr;   // where the argmax still being voted on goes (uniform over the CTA)
  int par = 0;   // exchange buffer in use
  int vs = 0;    // cand[] slot of the current stage's vote (alternates per stage: the slot is read and
                 // reset one barrier after the vote, while the next stage already votes into the other)
  __syncthreads();

  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    const double dt = p.dt_arr ? p.dt_arr[b] : p.dt;
    const BgCoef cf = {-rx * dt / 4.0, -fs0 * dt / 8.0, fs1 * dt / 8.0};
    double u[NP], res[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      u[i] = in ? p.u0[((size_t)b * NP + i) * K + k] : 0.0;
      res[i] = 0.0;
    }
    // limiter applied to the element's current state; returns flag | branch << 1
    auto limiter = [&]() -> int {
      double v = p.aw[0] * u[0];
#pragma unroll
      for (int i = 1; i < NP; ++i) v = fma(p.aw[i], u[i], v);
      trL[par][tid] = v;
      if (gfirst) trL[par][MAXT] = v;       // quirk C-16: ghost averages copy the end cells
      if (glast) trL[par][MAXT + 1] = v;
      __syncthreads();
      const double vm = trL[par][nbL], vp = trL[par][nbR];
      par ^= 1;
      if (!in) return 0;
      const double ue1 = u[0], ue2 = u[NP - 1];
      const MinmodBC q = minmod_bc(v - vm, vp - v);
      const double ve1 = v - minmod3(v - ue1, q);
      const double ve2 = v + minmod3(ue2 - v, q);
      if (!(fabs(ve1 - ue1) > p.eps0 || fabs(ve2 - ue2) > p.eps0)) return 0;
      double d = 0.0;
#pragma unroll
      for (int i = 0; i < NP; ++i) d = fma(p.sl[i], u[i], d);
      const double ux = twoh * d;
      int br = 1;
      double slope = ux;   // minmodB: a slope below M h^2 passes (recorded as argument 1 winning)
      if (!(p.tvbM > 0.0 && fabs(ux) <= p.tvbM * (h * h))) slope = minmod3b(ux, (vp - v) / h, (v - vm) / h, &br);
      const double sh = slope * (0.5 * h);   // x - x0 = (h/2) r: no geometry loads on this path
#pragma unroll
      for (int i = 0; i < NP; ++i) u[i] = fma(p.xcn[i], sh, v);
      return 1 | (br << 1);
    };
    auto flush_argmax = [&]() {   // after a barrier that follows the vote
      if (pend) {
        if (tid == 0) {
          const int c = cand[vs ^ 1];
          *pend = (c & 1) ? -((c >> 1) + 1) : ((c >> 1) + 1);
          cand[vs ^ 1] = 0x7fffffff;
        }
        pend = nullptr;
      }
    };
    {
      const int c0 = p.limit ? limiter() : 0;
      if (p.lim0 && in) p.lim0[(size_t)b * K + k] = (unsigned char)c0;
    }
    double* hist = (p.hist && in) ? p.hist + ((size_t)b * (p.S + 1) * NP) * K + k : nullptr;
    if (hist) {
#pragma unroll
      for (int i = 0; i < NP; ++i) hist[(size_t)i * K] = u[i];
    }
    int* amax_b = p.amax ? p.amax + (size_t)b * p.S * 5 : nullptr;
    double* maxvel_b = (p.maxvel && tid == 0) ? p.maxvel + (size_t)b * p.S * 5 : nullptr;
    unsigned short* lim_b = (p.lim && in) ? p.lim + (size_t)b * p.S * K + k : nullptr;
    for (int n = 0; n < p.S; ++n) {
      unsigned fl = 0u;
#pragma unroll 1
      for (int s = 0; s < 5; ++s) {
        // ---- exchange 1: traces and the mesh-wide max|u| (the value first: max is exact in any
        // order; where it sits is settled afterwards by the few threads that hold it)
        double m = -1.0;
        if (in) {
          m = fabs(u[0]);
#pragma unroll
          for (int i = 1; i < NP; ++i) {
            const double t = fabs(u[i]);
            m = (t > m) ? t : m;
          }
        }
        const double mw = warp_max_nonneg(m);
        trL[par][tid] = u[0];
        trR[par][tid] = u[NP - 1];
        if (gfirst) trR[par][MAXT] = u[0];            // ghost state = own trace
        if (glast) trL[par][MAXT + 1] = u[NP - 1];
        if (lane == 0) wmax[par][wid] = mw;
        __syncthreads();
        flush_argmax();   // the stage before: every candidate has voted by now
        const double maxvel = warp_max_nonneg((lane < nw) ? wmax[par][lane] : -1.0);
        if (amax_b) {
          // first occurrence in row-major order (i*K + k) of max|u|, and the sign of u there
          if (m == maxvel) {   // (idle lanes hold -1.0: never equal)
            int i0 = 0;
            double ui = u[0];
#pragma unroll
            for (int i = NP - 1; i >= 0; --i) {
              const bool hit = fabs(u[i]) == maxvel;
              i0 = hit ? i : i0;
              ui = hit ? u[i] : ui;
            }
            atomicMin(&cand[vs], ((i0 * K + k) << 1) | (ui < 0.0 ? 1 : 0));
          }
          pend = amax_b + n * 5 + s;
          vs ^= 1;
        }
        const double uL = trR[par][nbL], uR = trL[par][nbR];
        par ^= 1;
        if (maxvel_b) maxvel_b[n * 5 + s] = maxvel;
        burgers_stage_update<NP>(p, cf, u, res, uL, uR, maxvel, p.rka[s], p.rkb[s]);
        // ---- exchange 2: cell averages, limiter
        if (p.limit) {
          const int c = limiter();
          fl |= (unsigned)(c & 1) << s | (unsigned)(c >> 1) << (5 + 2 * s);
        }
      }
      if (lim_b) lim_b[(size_t)n * K] = (unsigned short)fl;
      if (hist) {
        double* hn = hist + (size_t)(n + 1) * NP * K;
#pragma unroll
        for (int i = 0; i < NP; ++i) hn[(size_t)i * K] = u[i];
      }
    }
    if (p.uT && in) {
#pragma unroll
      for (int i = 0; i < NP; ++i) p.uT[((size_t)b * NP + i) * K + k] = u[i];
    }
    __syncthreads();
    flush_argmax();   // the last stage's
  }
}


// ---------------------------------------------------------------------------------------
// Discrete adjoint of the limited Burgers march (oracle/burgers.py: burgers_adjoint).  Per step,
// backwards: the five stage input states are recomputed from the checkpoint u^n with the
// *recorded* limiter decisions and wave speeds (so they are the forward run's states, bit for
// bit) and parked in a per-CTA scratch (L2 resident); then the stages are transposed in
// reverse: limiter^T (frozen flags / minmod branches), lk += rkb lu, lu += dt (dR/du)^T lk
// including the rank-one term through C = max|u|, lk *= rka.
// ---------------------------------------------------------------------------------------
// The stage scratch (5 stage input states + neighbour traces per element) lives in dynamic shared
// memory when the CTA is at most 256 threads (SS_SMEM; 5 (NP+2) 256 doubles = 70 KB at N = 4), in a
// per-CTA global buffer (L2 resident) otherwise.
template <int NP, int MAXT, bool SS_SMEM>
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? BG_MINB_ADJ(NP) : 1) burgers_adjoint_kernel(const __grid_constant__ BurgersArgs p) {
  extern __shared__ double ss_smem[];
  __shared__ double exA[2][MAXT], exB[2][MAXT];
  constexpr int NWMAX = MAXT / 32;
  __shared__ __align__(16) double wsum[2][NWMAX];   // per-warp partials; slots of absent warps stay 0
  __shared__ double sm_mv[2][5];   // the step's recorded wave speeds / argmax (by step parity: fetched
  __shared__ int sm_am[2][5];      // once per step, visible after the step's first exchange barrier);
  __shared__ int sm_ak[2][5];      // sm_am = node index i of the argmax, negative-coded sign; sm_ak = its element
  const int tid = threadIdx.x, K = p.K, BD = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = (BD + 31) >> 5;
  const bool in = tid < K;
  const int k = in ? tid : 0;
  const int nbL = in ? (tid == 0 ? K - 1 : tid - 1) : tid;
  const int nbR = in ? (tid == K - 1 ? 0 : tid + 1) : tid;
  const bool first = in && tid == 0, last = in && tid == K - 1;
  const double rx = in ? p.rxk[k] : 0.0, fs0 = in ? p.fs0[k] : 0.0, fs1 = in ? p.fs1[k] : 0.0;
  const double h = in ? p.hk[k] : 1.0;
  const double twoh = 2.0 / h;
  double* const ss_g = SS_SMEM ? nullptr : p.stage_scratch + (size_t)blockIdx.x * 5 * (NP + 2) * BD + tid;
  // slot (stage s, row i) of this thread's scratch column
  auto ss = [&](int s, int i) -> double& {
    if constexpr (SS_SMEM) return ss_smem[(s * (NP + 2) + i) * MAXT + tid];
    else return ss_g[(size_t)(s * (NP + 2) + i) * BD];
  };
  int par = 0;   // exchange buffer in use (see the note above burgers_kernel)
  if (tid < 2 * NWMAX) wsum[tid / NWMAX][tid % NWMAX] = 0.0;
  __syncthreads();
  // deterministic CTA sum: shuffle tree, then the warps' partials in order
  auto warps_sum = [&]() -> double {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NWMAX; ++w) t += wsum[par][w];
    return t;
  };
  auto block_sum = [&](double v) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wsum[par][wid] = v;
    __syncthreads();
    const double t = warps_sum();
    par ^= 1;
    return t;
  };

  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    const double dt = p.dt_arr ? p.dt_arr[b] : p.dt;
    const BgCoef cf = {-rx * dt / 4.0, -fs0 * dt / 8.0, fs1 * dt / 8.0};
    const double* hist_b = p.hist + (size_t)b * (p.S + 1) * NP * K + k;
    double lu[NP], lk[NP];
    {
      double jp = 0.0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        lu[i] = in ? p.jw[(size_t)i * K + k] : 0.0;
        lk[i] = 0.0;
        const double uT = in ? hist_b[((size_t)p.S * NP + i) * K] : 0.0;
        jp = fma(lu[i], uT, jp);
      }
      const double J = block_sum(jp);
      if (p.Jout && tid == 0) p.Jout[b] = J;
    }
    // limiter^T with a frozen decision code (flag | branch << 1)
    auto limiter_T = [&](int code) {
      const int flag = code & 1, br = code >> 1;
      double a = 0.0, c = 0.0, ch = 0.0;
      if (code) {   // rare (a limited cell): keeps the sums and the division off the common path
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          a += lu[i];
          c = fma(p.xcn[i], lu[i], c);
        }
        c *= 0.5 * h;
        if (br >= 2) ch = c / h;
      }
      const double tr = (br == 2) ? ch : 0.0, tl = (br == 3) ? -ch : 0.0;
      exA[par][tid] = tr;   // goes to cell k+1
      exB[par][tid] = tl;   // goes to cell k-1
      __syncthreads();
      double fromL = exA[par][nbL], fromR = exB[par][nbR];
      par ^= 1;
      if (!p.periodic) {   // end cells see a copied ghost average: the term comes back to the cell
        if (first) fromL = 0.0;
        if (last) fromR = 0.0;
        if (last) fromL += tr;
        if (first) fromR += tl;
      }
      const double lv = (flag ? a : 0.0) + ((br == 2) ? -ch : ((br == 3) ? ch : 0.0)) + fromL + fromR;
      if (code || lv != 0.0) {
        const double cs = (br == 1) ? twoh * c : 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) lu[i] = (flag ? 0.0 : lu[i]) + p.aw[i] * lv + p.sl[i] * cs;
      }
    };

    const double* maxvel_b = p.maxvel + (size_t)b * p.S * 5;
    const int* amax_b = p.amax + (size_t)b * p.S * 5;
    for (int n = p.S - 1; n >= 0; --n) {
      const unsigned code = in ? (unsigned)p.lim[((size_t)b * p.S + n) * K + k] : 0u;
      const int np_ = n & 1;
      const double* hist_n = hist_b + (size_t)n * NP * K;
      if (tid < 5) {   // the recorded argmax decoded once per step: node, element, sign
        sm_mv[np_][tid] = maxvel_b[n * 5 + tid];
        const int am = amax_b[n * 5 + tid];
        const int flat = (am < 0 ? -am : am) - 1;
        sm_ak[np_][tid] = flat % K;
        sm_am[np_][tid] = (am < 0) ? -(flat / K) - 1 : flat / K;
      }
      // ---- recompute the stage input states of step n
      {
        double u[NP], res[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          u[i] = in ? hist_n[i * K] : 0.0;
          res[i] = 0.0;
        }
#pragma unroll 1
        for (int s = 0; s < 5; ++s) {
          exA[par][tid] = u[0];
          exB[par][tid] = u[NP - 1];
          __syncthreads();
          double uL = exB[par][nbL], uR = exA[par][nbR];
          par ^= 1;
          if (!p.periodic) {
            if (first) uL = u[0];
            if (last) uR = u[NP - 1];
          }
#pragma unroll
          for (int i = 0; i < NP; ++i) ss(s, i) = u[i];
          ss(s, NP) = uL;
          ss(s, NP + 1) = uR;
          if (s == 4) break;   // the state after the last stage is u^{n+1}: not needed
          burgers_stage_update<NP>(p, cf, u, res, uL, uR, sm_mv[np_][s], p.rka[s], p.rkb[s]);
          // limiter with the recorded decision
          const int flag = (code >> s) & 1, br = (code >> (5 + 2 * s)) & 3;
          double v = p.aw[0] * u[0];
#pragma unroll
          for (int i = 1; i < NP; ++i) v = fma(p.aw[i], u[i], v);
          exA[par][tid] = v;
          __syncthreads();
          double vm = exA[par][nbL], vp = exA[par][nbR];
          par ^= 1;
          if (!p.periodic) {
            if (first) vm = v;
            if (last) vp = v;
          }
          if (flag) {
            double d = 0.0;
#pragma unroll
            for (int i = 0; i < NP; ++i) d = fma(p.sl[i], u[i], d);
            const double slope = (br == 1) ? twoh * d : ((br == 2) ? (vp - v) / h : ((br == 3) ? (v - vm) / h : 0.0));
            const double sh = slope * (0.5 * h);
#pragma unroll
            for (int i = 0; i < NP; ++i) u[i] = fma(p.xcn[i], sh, v);
          }
        }
      }
      // ---- transpose the stages in reverse
#pragma unroll 1
      for (int s = 4; s >= 0; --s) {
        limiter_T((int)(((code >> s) & 1u) | (((code >> (5 + 2 * s)) & 3u) << 1)));
        const double rka = p.rka[s], rkb = p.rkb[s];
#pragma unroll
        for (int i = 0; i < NP; ++i) lk[i] = fma(rkb, lu[i], lk[i]);
        double us[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) us[i] = ss(s, i);
        const double uL = ss(s, NP), uR = ss(s, NP + 1);
        const double mv = sm_mv[np_][s];
        double G0 = 0.0, G1 = 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          G0 = fma(p.LIFT[i * 2], lk[i], G0);
          G1 = fma(p.LIFT[i * 2 + 1], lk[i], G1);
        }
        G0 *= fs0;
        G1 *= fs1;
        const double d0m = (-us[0] / 2.0 - mv / 2.0) * G0, d0p = (uL / 2.0 + mv / 2.0) * G0;
        const double d1m = (us[NP - 1] / 2.0 - mv / 2.0) * G1, d1p = (-uR / 2.0 + mv / 2.0) * G1;
        double gam = in ? G0 * (-(us[0] - uL) / 2.0) + G1 * (-(us[NP - 1] - uR) / 2.0) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gam += __shfl_xor_sync(0xffffffffu, gam, o);
        exA[par][tid] = d0p;   // belongs to the left neighbour's last node
        exB[par][tid] = d1p;   // belongs to the right neighbour's first node
        if (lane == 0) wsum[par][wid] = gam;
        __syncthreads();
        double toN = exA[par][nbR], to0 = exB[par][nbL];
        gam = warps_sum();
        par ^= 1;
        if (!p.periodic) {   // ghost = own trace
          if (last) toN = d1p;
          if (first) to0 = d0p;
        }
        double out[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < NP; ++i) acc = fma(p.Dr[i * NP + j], -rx * lk[i], acc);
          out[j] = us[j] * acc;
        }
        out[0] += d0m + to0;
        out[NP - 1] += d1m + toN;
        if (tid == sm_ak[np_][s]) {   // the element that held max|u|: the rank-one term of C
          const int am = sm_am[np_][s];
          const int ii = am < 0 ? -am - 1 : am;
          const double add = (am < 0) ? -gam : gam;
#pragma unroll
          for (int q = 0; q < NP; ++q) out[q] += (q == ii) ? add : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          lu[i] = fma(dt, out[i], lu[i]);
          lk[i] *= rka;
        }
      }
    }
    limiter_T(in ? (int)p.lim0[(size_t)b * K + k] : 0);
    if (p.lam0 && in) {
#pragma unroll
      for (int i = 0; i < NP; ++i) p.lam0[((size_t)b * NP + i) * K + k] = lu[i];
    }
  }
}

template <int NP>
static cudaError_t burgers_adjoint_launch(int grid, int block, cudaStream_t st, const BurgersArgs& a) {
  if (block <= 256) {
    const int smem = 5 * (NP + 2) * 256 * (int)sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(burgers_adjoint_kernel<NP, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    burgers_adjoint_kernel<NP, 256, true><<<grid, block, smem, st>>>(a);
  } else {
    burgers_adjoint_kernel<NP, 1024, false><<<grid, block, 0, st>>>(a);
  }
  return cudaGetLastError();
}

template <int NP>
static cudaError_t burgers_launch(int grid, int block, cudaStream_t st, const BurgersArgs& a) {
  if (block <= 256)
    burgers_kernel<NP, 256><<<grid, block, 0, st>>>(a);
  else
    burgers_kernel<NP, 1024><<<grid, block, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace dgadj

int dgadj_set_functional_weights_level0(dgadj_handle* h, const double* jw_c);

static int burgers_setup(dgadj_handle* h, BurgersArgs& a, int64_t B, int32_t S, double dt, const double* dt_dev,
                         int limit, const double* invV_host, const double* V_host, const double* x_host,
                         cudaStream_t st) {
  const int Np = h->Np, K = h->K;
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.K = K;
  a.S = S;
  a.periodic = (h->cfg.bc == DGADJ_BC_PERIODIC);
  a.limit = limit;
  a.tvbM = 0.0;
  a.eps0 = (limit == 2) ? -1.0 : 1.0e-8;
  a.dt = dt;
  a.dt_arr = dt_dev;
  a.rxk = h->d_mesh[0][0];
  a.fs0 = h->d_mesh[0][1];
  a.fs1 = h->d_mesh[0][2];
  a.so = h->base_ops[0];
  for (int s = 0; s < 5; ++s) {
    a.rka[s] = h->cops.rka[s];
    a.rkb[s] = h->cops.rkb[s];
  }
  for (int i = 0; i < Np * Np; ++i) a.Dr[i] = h->Dr_nodal[i];
  for (int i = 0; i < Np * 2; ++i) a.LIFT[i] = h->LIFT_nodal[i];
  // aw = V(1,1)*invV(1,:) (SlopeLimitN.m:9);  sl = Dr(1,:)*V(:,1:2)*invV(1:2,:) (SlopeLimitN.m:27,
  // SlopeLimitLin.m:16);  xc = x - x0, h = x(Np,:) - x(1,:) (SlopeLimitLin.m:10-12)
  std::vector<double> xc((size_t)Np * K, 0.0), hk(K, 1.0);
  if (limit) {
    for (int i = 0; i < Np; ++i) a.aw[i] = V_host[0] * invV_host[i];
    const double* Dr = h->Dr_nodal;
    for (int i = 0; i < Np; ++i) {
      double sum = 0.0;
      for (int j = 0; j < Np; ++j)
        sum += Dr[j] * (V_host[(size_t)j * Np + 0] * invV_host[0 * Np + i] + V_host[(size_t)j * Np + 1] * invV_host[1 * Np + i]);
      a.sl[i] = sum;
    }
    for (int k = 0; k < K; ++k) {
      const double hh = x_host[(size_t)(Np - 1) * K + k] - x_host[k];
      const double x0 = x_host[k] + hh / 2;
      hk[k] = hh;
      for (int i = 0; i < Np; ++i) xc[(size_t)i * K + k] = x_host[(size_t)i * K + k] - x0;
      if (k == 0)
        for (int i = 0; i < Np; ++i) a.xcn[i] = xc[(size_t)i * K] / (hh / 2);
    }
  }
  // (xc is kept in the argument struct for reference; no kernel reads it: the limited path reconstructs
  //  through the affine map, xcn)
  const size_t need = ((size_t)Np * K + K) * sizeof(double);
  if (need > h->bg_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->bg_scratch);
    h->bg_scratch = nullptr;
    h->bg_bytes = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->bg_scratch, need));
    h->bg_bytes = need;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->bg_scratch, xc.data(), (size_t)Np * K * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->bg_scratch + (size_t)Np * K, hk.data(), K * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  a.xc = h->bg_scratch;
  a.hk = h->bg_scratch + (size_t)Np * K;
  return DGADJ_OK;
}

extern "C" int dgadj_burgers_forward(dgadj_handle* h, int64_t B, int32_t S, double dt, const double* dt_dev,
                                     int32_t limit, double tvb_M, const double* invV_host, const double* V_host,
                                     const double* x_host, const double* u0_dev, double* uT_dev,
                                     double* hist_dev, uint16_t* lim_dev, uint8_t* lim0_dev, int32_t* amax_dev,
                                     double* maxvel_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || S < 0 || !u0_dev) return fail(h, DGADJ_ERR_INVALID, "bad burgers arguments");
  if (!h->ops_set) return fail(h, DGADJ_ERR_STATE, "dgadj_set_operators has not been called");
  if (limit && (!invV_host || !V_host || !x_host)) return fail(h, DGADJ_ERR_INVALID, "the limiter needs V, invV and x");
  if (h->nstages != 5) return fail(h, DGADJ_ERR_UNSUPPORTED, "the Burgers march is LSERK4 only");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Np = h->Np, K = h->K;
  BurgersArgs a;
  if (limit < 0 || limit > 2 || !(tvb_M >= 0.0)) return fail(h, DGADJ_ERR_INVALID, "limit must be 0, 1 or 2 and tvb_M >= 0");
  int rc = burgers_setup(h, a, B, S, dt, dt_dev, limit, invV_host, V_host, x_host, st);
  if (rc) return rc;
  a.tvbM = tvb_M;
  a.u0 = u0_dev;
  a.uT = uT_dev;
  a.hist = hist_dev;
  a.lim = lim_dev;
  a.lim0 = lim0_dev;
  a.amax = amax_dev;
  a.maxvel = maxvel_dev;
  const int block = (K + 31) / 32 * 32;
  const int per_sm = std::max(1, std::min(16, (block <= 256 ? 2048 : 1024) / block));
  const int grid = (int)std::min<int64_t>(B, (int64_t)h->sm_count * per_sm);
  cudaError_t e = cudaSuccess;
#define DGADJ_BG(n) case n: e = burgers_launch<n>(grid, block, st, a); break;
  switch (Np) {
    DGADJ_BG(2) DGADJ_BG(3) DGADJ_BG(4) DGADJ_BG(5) DGADJ_BG(6) DGADJ_BG(7) DGADJ_BG(8) DGADJ_BG(9)
    default: return fail(h, DGADJ_ERR_UNSUPPORTED, "Burgers march supports 1 <= N <= 8");
  }
#undef DGADJ_BG
  if (e != cudaSuccess) return fail(h, DGADJ_ERR_CUDA, "burgers kernel launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  return DGADJ_OK;
}

extern "C" int dgadj_burgers_adjoint(dgadj_handle* h, int64_t B, int32_t S, double dt, const double* dt_dev,
                                     const double* invV_host, const double* V_host, const double* x_host,
                                     const double* jw_host, const double* hist_dev, const uint16_t* lim_dev,
                                     const uint8_t* lim0_dev, const int32_t* amax_dev, const double* maxvel_dev,
                                     double* lam0_dev, double* J_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  // (a march of S = 0 steps has no per-step records: lim / amax / maxvel may be NULL then)
  if (B <= 0 || S < 0 || !jw_host || !hist_dev || !lim0_dev || (S > 0 && (!lim_dev || !amax_dev || !maxvel_dev)))
    return fail(h, DGADJ_ERR_INVALID, "bad burgers_adjoint arguments (the forward checkpoints are all required)");
  if (!h->ops_set) return fail(h, DGADJ_ERR_STATE, "dgadj_set_operators has not been called");
  if (h->nstages != 5) return fail(h, DGADJ_ERR_UNSUPPORTED, "the Burgers march is LSERK4 only");
  if (!invV_host || !V_host || !x_host) return fail(h, DGADJ_ERR_INVALID, "the limiter needs V, invV and x");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Np = h->Np, K = h->K;
  BurgersArgs a;
  int rc = burgers_setup(h, a, B, S, dt, dt_dev, 1, invV_host, V_host, x_host, st);
  if (rc) return rc;
  rc = dgadj_set_functional_weights_level0(h, jw_host);
  if (rc) return rc;
  a.jw = h->d_jwc;
  a.hist = const_cast<double*>(hist_dev);
  a.lim = const_cast<uint16_t*>(lim_dev);
  a.lim0 = const_cast<uint8_t*>(lim0_dev);
  a.amax = const_cast<int32_t*>(amax_dev);
  a.maxvel = const_cast<double*>(maxvel_dev);
  a.lam0 = lam0_dev;
  a.Jout = J_dev;
  const int block = (K + 31) / 32 * 32;
  const int per_sm = std::max(1, std::min(16, (block <= 256 ? 2048 : 1024) / block));
  const int grid = (int)std::min<int64_t>(B, (int64_t)h->sm_count * per_sm);
  // stage scratch: shared memory for CTAs of at most 256 threads, a global per-CTA buffer beyond
  const size_t need = block <= 256 ? 0 : (size_t)grid * 5 * (Np + 2) * block * sizeof(double);
  if (need > h->bgs_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->bgs_scratch);
    h->bgs_scratch = nullptr;
    h->bgs_bytes = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->bgs_scratch, need));
    h->bgs_bytes = need;
  }
  a.stage_scratch = h->bgs_scratch;
  cudaError_t e = cudaSuccess;
#define DGADJ_BGA(n) case n: e = burgers_adjoint_launch<n>(grid, block, st, a); break;
  switch (Np) {
    DGADJ_BGA(2) DGADJ_BGA(3) DGADJ_BGA(4) DGADJ_BGA(5) DGADJ_BGA(6) DGADJ_BGA(7) DGADJ_BGA(8) DGADJ_BGA(9)
    default: return fail(h, DGADJ_ERR_UNSUPPORTED, "Burgers adjoint supports 1 <= N <= 8");
  }
#undef DGADJ_BGA
  if (e != cudaSuccess) return fail(h, DGADJ_ERR_CUDA, "burgers adjoint launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  return DGADJ_OK;
}
