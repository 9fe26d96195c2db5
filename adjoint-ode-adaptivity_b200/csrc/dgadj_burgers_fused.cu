// dgadj_burgers_fused.cu -- BASELINE config 3 as one persistent kernel: inviscid Burgers DG march with the
// reference's SlopeLimitN fused after every LSERK4 stage (utils/SlopeLimitN.m:9-32, SlopeLimitLin.m:10-18,
// minmod.m:6-12, minmodB.m:6-11), forward checkpointing into a per-CTA ring, the reverse-time discrete
// adjoint on the frozen limiter / minmod / argmax branches, and the per-element adjoint-weighted error
// indicator in the same pass (oracle/burgers.py: burgers_fwd_adj_indicator; conventions of
// matlab/adj_march.m:103-117 and errEst, python/Main_finite_difference.py:79-94; adjoint one order higher,
// matlab/MAIN.m:34).  The Burgers right-hand side is build-specified.
//
// Mapping: one CTA marches one trajectory forward and then immediately backward; a thread owns EPT
// adjacent elements (nodal values in registers).  HBM is touched for: the initial state, one state tile
// u^n per step written into the CTA's ring slot (forward phase, coalesced) and streamed back with
// bulk-TMA + mbarrier one step ahead (adjoint phase), and the outputs -- live checkpoints are
// #CTAs x S x state, never B x S.  Nothing but states is checkpointed: the adjoint phase takes every step
// again from u^n with the very same stage routine (same bits, hence the same limiter flags, minmod
// branches, wave speeds and argmax), keeps the five stage states in shared memory and transposes them.
// Indicator mode (IND): the step that is re-taken and transposed is the ENRICHED one (order N+1) from the
// prolonged state P u^n, rho^n = P u^{n+1} - Phi_f(P u^n), eta_k += lam_f^{n+1}_k . rho^n_k.
//
// Exchanges: neighbour traces / cell averages / transposed face terms cross threads through a
// double-buffered pair of shared arrays and one mbarrier per CTA, split into arrive (as soon as the own
// values are published) and wait (when the neighbours' are needed) with the volume terms in between; a
// trajectory that fits one warp synchronises with __syncwarp alone.  The mesh-wide max|u| of the
// Lax-Friedrichs flux and the adjoint's mesh-wide sum ride on the same exchanges.
#define DGADJ_DEVICE_CODE 1
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "dgadj_internal.h"

// resident CTAs per SM the kernels are compiled for (register budget = 65536 / (threads x CTAs))
// Measured at config 3 (N = 4, K = 256, two elements per thread, 128 threads): the march with the adjoint in its
// own space fits 128 registers with 80 B of spills and gains 8 % from the fourth CTA (4.71 -> 5.08e10
// updates/s); the enriched step of indicator mode spills 336 B there and loses 10 % (4.09 -> 3.68e10).
#ifndef BG_MINB128_IND
#define BG_MINB128_IND 3
#endif
#ifndef BG_MINB128_PLAIN
#define BG_MINB128_PLAIN 4
#endif
#define BG_MINB(BD, IND) ((BD) <= 128 ? ((IND) ? BG_MINB128_IND : BG_MINB128_PLAIN) : 1)

namespace dgadj {

// The operators the stage loops read.  One copy per RK stage, indexed by the stage counter: the copies are
// identical, but a loop-invariant constant would be hoisted out of the stage loop into uniform registers --
// 36 + 12 + 18 doubles against a file of 63 -- and spilled from there through local memory (seen in the SASS
// as R2UR / LDL pairs); indexed by the counter they stay in the loop as constant-bank operands (the same
// measure as in the advection kernel, dgadj_kernels.cuh).
// MEASURED on config 3 (B200; N = 4, K = 256, B = 16384, S = 2373; updates/s, indicator mode / plain):
//   first version (stage 4's state in registers, separate landing tile, mbarrier everywhere)   3.79e10 / 4.46e10
//   + state tile lands in the idle stage slot, hardware barrier for exchanges without overlap  4.09e10 / 4.71e10
//   + fourth resident CTA where it fits 128 registers (plain mode only)                           --    / 5.08e10
//   + all five stage states in shared memory (the conditional register copy of stage 4 cost ~50 moves per
//     stage), volume terms folded into the residual / adjoint before the wait, per-stage operator copies (no
//     hoisting into uniform registers: 336 B of spills -> 0)                                    4.38e10 / 5.29e10
//   split exchanges (BG_SPLIT 1) vs a plain hardware barrier at the arrive point (0): 4.399 / 4.415e10 -- no
//   difference; nor does the suspend-time hint of the mbarrier wait matter (BG_SUSPEND_NS 0 vs 2000).
//   Four elements per thread (64-thread CTAs): 2.25e10 (half the warps: latency bound); one per thread: 2.4e10.
//   + no division sequences in the limited cells' path (div_rn: the quotient's bits from the host-rounded
//     reciprocal; 2/h tabulated): the warps that hold limited cells are the last to reach the next exchange,
//     so their divergent path is the CTA's critical path -- 36 % of the warp-stages enter it      5.03e10 / 5.96e10
//   one divergent region for all of a thread's cells (detection first, reconstructions side by side)
//     instead of one per cell: 4.79e10 / 5.67e10 -- slower, not kept.
//   + the transposed volume term through the even / odd blocks of the forward one (half the multiply-adds and
//     constant loads of the dense transpose; the exact transpose of the forward term as computed)  5.10e10 / 6.22e10
//   the exchange of the limiter's transpose split (arrive, the stage-state loads, wait) instead of one hardware
//     barrier: 4.98e10 / 6.06e10 (the state's longer live range at the 168-register cap costs the indicator mode
//     what the plain mode gains); split in plain mode only: 5.03e10 / 5.85e10 -- not kept.
// ncu of the final form (profiles/r2_burgers_fused_ncu.json): fp64 pipe 50.7 % (47.3 % before the last two steps),
// issue slots 54 % busy, 44 % of the issued instructions on the fp64 pipe (179.7 of 405.0 per update); per issued
// instruction 1.41 cycles of fixed-latency waits, 0.58 on the exchange mbarrier, 0.32 on the hardware barrier --
// 12 warps per SM (3 CTAs: the five stage states of a trajectory take 61 KB of shared memory) leave the
// schedulers without an eligible warp 45 % of the time.
#ifndef BG_SPLIT
#define BG_SPLIT 1
#endif
#ifndef BG_SUSPEND_NS
#define BG_SUSPEND_NS 2000u
#endif
#ifndef BG_STAGE_CONSTS
#define BG_STAGE_CONSTS 1
#endif
#ifndef BG_TRANSPOSED_EO
#define BG_TRANSPOSED_EO 1   // the transposed volume term through the even / odd blocks of the forward one
#endif
struct BgStageOps {
  StageOps so;                 // even/odd blocks of the nodal Dr / LIFT (forward volume and lift terms)
  double Dr[MAXNP * MAXNP];    // nodal Dr (transposed volume term)
  double LIFT[MAXNP * 2];
  double aw[MAXNP];            // cell average weights  V(1,1)*invV(1,:)                 (SlopeLimitN.m:9)
};
struct BgLevel {               // one polynomial space: primal (order N) or enriched (order N+1)
  BgStageOps st[BG_STAGE_CONSTS ? 5 : 1];
  double sl[MAXNP];            // slope weights         Dr(1,:)*V(:,1:2)*invV(1:2,:)     (SlopeLimitLin.m:16)
  double xcn[MAXNP];           // (x - x0)/(h/2) = the reference nodes                   (SlopeLimitLin.m:11-12)
  const double *rxk, *fs0, *fs1;   // [K]
  __host__ __device__ const BgStageOps& ops(int s) const { return st[BG_STAGE_CONSTS ? s : 0]; }
};

struct BgFusedArgs {
  long long B;
  int K, S, periodic, limit;
  double tvbM, eps0, dt;
  const double* dt_arr;
  const double* hk;            // [3][K]: h = x(Np,k) - x(1,k);  1/h and 2/h as the host's IEEE divisions round them
  BgLevel lv[2];
  double P[MAXNP * MAXNP];     // nodal prolongation [NPF][NP]
  const double* jw[2];         // [NPX][K] weights of J = sum jw o u(T) in both spaces
  const double* u0;            // [B][NP][K]
  double* uT;                  // [B][NP][K] or null
  double* J;                   // [B] or null
  double* lam0;                // [B][NPX][K] or null
  double* eta;                 // [B][K] (IND) or null
  int* nlim;                   // [B][2] limiter activations (cell, stage): the march; the steps taken again by the adjoint phase
  unsigned* status;            // [B] bit 0: non-finite state, or null
  double* ring;                // [grid][S][NP][EPT][BD]
  double rka[5], rkb[5];
};

// ------------------------------------------------------------------------------------------------------
// minmod without divisions (see dgadj_burgers.cu; same arithmetic, bit-identical decisions)
// ------------------------------------------------------------------------------------------------------
struct MmBC {
  bool pos, neg;
  double t;
};
__device__ __forceinline__ MmBC mm_bc(double b, double c) {
  MmBC r;
  r.pos = (b > 0.0) & (c > 0.0);
  r.neg = (b < 0.0) & (c < 0.0);
  r.t = ((b < c) == r.pos) ? b : c;
  return r;
}
__device__ __forceinline__ double mm3(double a, const MmBC& q) {
  const double r = ((a < q.t) == q.pos) ? a : q.t;
  return ((q.pos & (a > 0.0)) | (q.neg & (a < 0.0))) ? r : 0.0;
}
__device__ __forceinline__ double mm3b(double a, double b, double c, int* br) {
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
// x / h from the correctly rounded reciprocal r = RN(1/h): q = RN(x r) is within an ulp of the quotient, the
// residual x - h q is exact in an fma, and RN(q + residual r) is the correctly rounded quotient (Markstein's
// theorem) -- the bits of a division without the 30-instruction division sequence in the limited cells' path.
// (tests/div_rn_check.c: random pairs against the division on the host, bit for bit.)  Not for non-finite x: the march's
// status word reports those.
__device__ __forceinline__ double div_rn(double x, double h, double r) {
  const double q = x * r;
  return fma(fma(-h, q, x), r, q);
}
__device__ __forceinline__ double warp_max_nn(double m) {   // values >= 0 or exactly -1.0
  const int hi = __double2hiint(m);
  const int mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned lo = (hi == mh) ? (unsigned)__double2loint(m) : 0u;
  const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
  return __hiloint2double(mh, (int)ml);
}

// mbarrier wait that lets the hardware suspend the warp for up to `ns` before the test returns: fewer spin
// iterations (each one costs issue slots of the scheduler the partner warps run on)
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity, uint32_t ns) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(ns)
        : "memory");
  } while (!done);
}

// ------------------------------------------------------------------------------------------------------
// per-thread context
// ------------------------------------------------------------------------------------------------------
template <int BD>
struct BgCtx {
  static constexpr int NW = BD / 32;
  static constexpr int BDP = BD + 2;   // + the two ghost slots of a non-periodic mesh
  int tid, lane, wid, nbL, nbR, par, vs;
  bool in, gfirst, glast;   // owns elements; owns the first / last element of a NON-periodic mesh
  uint64_t* bar;
  uint32_t phase;
  double *exA, *exB;        // [2][BDP] "left-edge" / "right-edge" values
  double* wred;             // [2][8] per-warp partials
  int* cand;                // [2] argmax vote
  double* mv;               // [5] wave speeds of the step being transposed
  int* am;                  // [5] argmax of the step being transposed ((flat << 1) | negative)

  // BG_SPLIT = 1: split exchange (mbarrier arrive, work, wait); 0: a plain hardware barrier at the arrive point
  __device__ __forceinline__ void arrive() {
#if BG_SPLIT
    __syncwarp();
    if (NW > 1 && lane == 0) mbar_arrive(bar);
#else
    sync();
#endif
  }
  __device__ __forceinline__ void wait() {
#if BG_SPLIT
    if (NW > 1) {
      mbar_wait_suspend(bar, phase, BG_SUSPEND_NS);
      phase ^= 1u;
    }
#endif
  }
  // an exchange with nothing to do between publishing and reading: the hardware barrier parks the warp
  // instead of spinning on the mbarrier (spin iterations cost issue slots the other warps can use)
  __device__ __forceinline__ void sync() {
    if (NW > 1) __syncthreads();
    else __syncwarp();
  }
  __device__ __forceinline__ double* A() const { return exA + par * BDP; }
  __device__ __forceinline__ double* Bb() const { return exB + par * BDP; }
  __device__ __forceinline__ double* W() const { return wred + par * 8; }
};

struct BgCoef {
  double c1, k0, k1;   // -rx dt/4, -Fscale(1,k) dt/8, +Fscale(2,k) dt/8
};

// ------------------------------------------------------------------------------------------------------
// One LSERK4 stage of the thread's EPT elements: max|u| + trace exchange, right-hand side
// (dgadj_burgers.cu: burgers_stage_update, volume terms hoisted in front of the wait), RK update, cell
// average exchange, SlopeLimitN.  code[e] = flag | branch << 1 of the limiter pass.
// RECORD: thread 0 leaves the stage's wave speed in cx.mv[s] and the argmax vote is taken (cx.am[s] is
// written one exchange later by flush_vote).
// ------------------------------------------------------------------------------------------------------
template <int BD>
__device__ __forceinline__ void bg_flush_vote(BgCtx<BD>& cx, int& pend) {
  if (pend >= 0) {   // uniform
    if (cx.tid == 0) {
      cx.am[pend] = cx.cand[cx.vs ^ 1];
      cx.cand[cx.vs ^ 1] = 0x7fffffff;
    }
    pend = -1;
  }
}

template <int NPX, int EPT, int BD>
__device__ __forceinline__ void bg_limiter(const BgFusedArgs& p, const BgLevel& L, int s, BgCtx<BD>& cx, int k0,
                                           double (&u)[EPT][NPX], int (&code)[EPT]) {
  const BgStageOps& Ls = L.ops(s);
  double v[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    double a = Ls.aw[0] * u[e][0];
#pragma unroll
    for (int i = 1; i < NPX; ++i) a = fma(Ls.aw[i], u[e][i], a);
    v[e] = a;
  }
  double* eA = cx.A();
  double* eB = cx.Bb();
  eA[cx.tid] = v[0];
  eB[cx.tid] = v[EPT - 1];
  if (cx.gfirst) eB[BD] = v[0];             // quirk C-16: ghost averages copy the end cells (SlopeLimitN.m:18)
  if (cx.glast) eA[BD + 1] = v[EPT - 1];
  cx.sync();
  const double vmL = eB[cx.nbL], vpR = eA[cx.nbR];
  cx.par ^= 1;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    code[e] = 0;
    if (!cx.in) continue;
    const double vm = (e == 0) ? vmL : v[e - 1], vp = (e == EPT - 1) ? vpR : v[e + 1];
    const double ue1 = u[e][0], ue2 = u[e][NPX - 1];
    const MmBC q = mm_bc(v[e] - vm, vp - v[e]);
    const double ve1 = v[e] - mm3(v[e] - ue1, q);
    const double ve2 = v[e] + mm3(ue2 - v[e], q);
    if (!(fabs(ve1 - ue1) > p.eps0 || fabs(ve2 - ue2) > p.eps0)) continue;
    const double h = __ldg(p.hk + k0 + e), rh = __ldg(p.hk + p.K + k0 + e), th = __ldg(p.hk + 2 * p.K + k0 + e);
    double d = 0.0;
#pragma unroll
    for (int i = 0; i < NPX; ++i) d = fma(L.sl[i], u[e][i], d);
    const double ux = th * d;   // (2/h) d
    int br = 1;
    double slope = ux;   // minmodB: a slope below M h^2 passes (recorded as argument 1 winning)
    if (!(p.tvbM > 0.0 && fabs(ux) <= p.tvbM * (h * h)))
      slope = mm3b(ux, div_rn(vp - v[e], h, rh), div_rn(v[e] - vm, h, rh), &br);
    const double sh = slope * (0.5 * h);
#pragma unroll
    for (int i = 0; i < NPX; ++i) u[e][i] = fma(L.xcn[i], sh, v[e]);
    code[e] = 1 | (br << 1);
  }
}

template <int NPX, int EPT, int BD, bool RECORD>
__device__ __forceinline__ void bg_stage(const BgFusedArgs& p, const BgLevel& L, BgCtx<BD>& cx, int k0,
                                         const BgCoef (&cf)[EPT], double (&u)[EPT][NPX], double (&res)[EPT][NPX],
                                         int s, int (&code)[EPT], int& pend) {
  constexpr int HE = (NPX + 1) / 2, HO = NPX / 2;
  constexpr int NW = BD / 32;
  const BgStageOps& Ls = L.ops(s);
  // ---- exchange 1: traces and the mesh-wide max|u| (value first; where it sits is voted afterwards)
  // (independent chains, one per element and node parity: a single running maximum would be a dependent chain
  //  of EPT*NPX compare-select pairs at the head of every stage)
  double m = -1.0;
  if (cx.in) {
    double mc[EPT][2];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      mc[e][0] = fabs(u[e][0]);
      mc[e][1] = (NPX > 1) ? fabs(u[e][1]) : 0.0;
#pragma unroll
      for (int i = 2; i < NPX; ++i) {
        const double t = fabs(u[e][i]);
        mc[e][i & 1] = (t > mc[e][i & 1]) ? t : mc[e][i & 1];
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double t = (mc[e][1] > mc[e][0]) ? mc[e][1] : mc[e][0];
      m = (t > m) ? t : m;
    }
  }
  const double mw = warp_max_nn(m);
  double* eA = cx.A();
  double* eB = cx.Bb();
  eA[cx.tid] = u[0][0];
  eB[cx.tid] = u[EPT - 1][NPX - 1];
  if (cx.gfirst) eB[BD] = u[0][0];                  // ghost state = own trace (zero jump)
  if (cx.glast) eA[BD + 1] = u[EPT - 1][NPX - 1];
  double* wr = cx.W();
  if (NW > 1 && cx.lane == 0) wr[cx.wid] = mw;
  cx.arrive();
  // volume terms (no neighbour data), folded into the RK residual at once: with F = u^2,
  //   E_i = c1 (DE Fo)_i, O_i = c1 (DO Fe)_i;  dt rhs_i = E_i + O_i, dt rhs_{N-i} = E_i - O_i, dt rhs_mid = 2 E_mid
  const double rka = p.rka[s], rkb = p.rkb[s];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    double fe[HE], fo[HO > 0 ? HO : 1];
#pragma unroll
    for (int i = 0; i < NPX / 2; ++i) {
      const double ti = cf[e].c1 * u[e][i], tj = cf[e].c1 * u[e][NPX - 1 - i];
      const double a = ti * u[e][i];
      fe[i] = fma(tj, u[e][NPX - 1 - i], a);
      fo[i] = fma(-tj, u[e][NPX - 1 - i], a);
    }
    if (NPX & 1) fe[NPX / 2] = (cf[e].c1 * u[e][NPX / 2]) * u[e][NPX / 2];
#pragma unroll
    for (int i = 0; i < HE; ++i) {
      double E = 0.0;
#pragma unroll
      for (int j = 0; j < HO; ++j) {
        const double2 c2 = Ls.so.DE2[i * HP + j / 2];
        E = fma((j & 1) ? c2.y : c2.x, fo[j], E);
      }
      if (i < HO) {
        double O = 0.0;
#pragma unroll
        for (int j = 0; j < HE; ++j) {
          const double2 c2 = Ls.so.DO2[i * HP + j / 2];
          O = fma((j & 1) ? c2.y : c2.x, fe[j], O);
        }
        res[e][i] = fma(rka, res[e][i], E + O);
        res[e][NPX - 1 - i] = fma(rka, res[e][NPX - 1 - i], E - O);
      } else {
        res[e][i] = fma(rka, res[e][i], E + E);   // the middle node of an odd node count
      }
    }
  }
  cx.wait();
  bg_flush_vote<BD>(cx, pend);   // the stage before: every candidate has voted by now
  double maxvel = mw;
  if (NW > 1) {
    maxvel = wr[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      const double t = wr[w];
      maxvel = (t > maxvel) ? t : maxvel;
    }
  }
  const double uLL = eB[cx.nbL], uRR = eA[cx.nbR];
  cx.par ^= 1;
  if (RECORD) {
    // first occurrence in row-major order (i*K + k) of max|u|, and the sign of u there
    if (m == maxvel) {   // (idle lanes hold -1.0: never equal)
      int best = 0x7fffffff;
#pragma unroll
      for (int e = EPT - 1; e >= 0; --e) {
#pragma unroll
        for (int i = NPX - 1; i >= 0; --i) {
          const int key = ((i * p.K + k0 + e) << 1) | (u[e][i] < 0.0 ? 1 : 0);
          if (fabs(u[e][i]) == maxvel && key < best) best = key;
        }
      }
      atomicMin(&cx.cand[cx.vs], best);
    }
    pend = s;
    cx.vs ^= 1;
    if (cx.tid == 0) cx.mv[s] = maxvel;
  }
  // ---- surface terms (the lift of the Lax-Friedrichs flux), state update
  const double twoC = maxvel + maxvel;
  double G0[EPT], G1[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {   // (all fluxes first: they read the neighbours' states of this stage)
    const double uL = (e == 0) ? uLL : u[e - 1][NPX - 1];
    const double uR = (e == EPT - 1) ? uRR : u[e + 1][0];
    G0[e] = (cf[e].k0 * (u[e][0] - uL)) * ((u[e][0] + uL) + twoC);
    G1[e] = (cf[e].k1 * (u[e][NPX - 1] - uR)) * ((u[e][NPX - 1] + uR) - twoC);
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const double ge = G0[e] + G1[e], go = G0[e] - G1[e];
#pragma unroll
    for (int i = 0; i < HE; ++i) {
      const double2 l2 = Ls.so.LS2[i / 2];
      const double Es = ((i & 1) ? l2.y : l2.x) * ge;
      if (i < HO) {
        const double2 a2 = Ls.so.LA2[i / 2];
        const double Os = ((i & 1) ? a2.y : a2.x) * go;
        res[e][i] += Es + Os;
        res[e][NPX - 1 - i] += Es - Os;
      } else {
        res[e][i] += Es + Es;
      }
    }
#pragma unroll
    for (int i = 0; i < NPX; ++i) u[e][i] = fma(rkb, res[e][i], u[e][i]);
  }
  // ---- exchange 2: cell averages, limiter
  if (p.limit) {
    bg_limiter<NPX, EPT, BD>(p, L, s, cx, k0, u, code);
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) code[e] = 0;
  }
}

// deterministic CTA sum of one value per thread: shuffle tree, then the warps' partials in order
template <int BD>
__device__ __forceinline__ double bg_block_sum(BgCtx<BD>& cx, double v) {
  constexpr int NW = BD / 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (NW == 1) return v;
  double* wr = cx.W();
  if (cx.lane == 0) wr[cx.wid] = v;
  cx.sync();
  double t = wr[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) t += wr[w];
  cx.par ^= 1;
  return t;
}

// transpose of the limiter on the recorded decisions (oracle/burgers.py: limiter_T)
template <int NPX, int EPT, int BD>
__device__ __forceinline__ void bg_limiter_T(const BgFusedArgs& p, const BgLevel& L, int s, BgCtx<BD>& cx, int k0,
                                             const int (&code)[EPT], double (&lu)[EPT][NPX]) {
  const BgStageOps& Ls = L.ops(s);
  double a[EPT], c[EPT], ch[EPT], tr[EPT], tl[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    a[e] = c[e] = ch[e] = 0.0;
    const int br = code[e] >> 1;
    if (code[e]) {   // rare (a limited cell): keeps the sums and the division off the common path
      const double h = __ldg(p.hk + k0 + e);
#pragma unroll
      for (int i = 0; i < NPX; ++i) {
        a[e] += lu[e][i];
        c[e] = fma(L.xcn[i], lu[e][i], c[e]);
      }
      c[e] *= 0.5 * h;
      if (br >= 2) ch[e] = div_rn(c[e], h, __ldg(p.hk + p.K + k0 + e));
    }
    tr[e] = (br == 2) ? ch[e] : 0.0;    // goes to cell k+1
    tl[e] = (br == 3) ? -ch[e] : 0.0;   // goes to cell k-1
  }
  double* eA = cx.A();
  double* eB = cx.Bb();
  eA[cx.tid] = tl[0];         // for the left neighbour's last cell
  eB[cx.tid] = tr[EPT - 1];   // for the right neighbour's first cell
  cx.sync();
  double inL = eB[cx.nbL], inR = eA[cx.nbR];
  cx.par ^= 1;
  // end cells of a non-periodic mesh see a copied ghost average: the term comes back to the cell
  if (cx.gfirst) inL = tl[0];
  if (cx.glast) inR = tr[EPT - 1];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int flag = code[e] & 1, br = code[e] >> 1;
    const double fromL = (e == 0) ? inL : tr[e - 1], fromR = (e == EPT - 1) ? inR : tl[e + 1];
    const double lv = (flag ? a[e] : 0.0) + ((br == 2) ? -ch[e] : ((br == 3) ? ch[e] : 0.0)) + fromL + fromR;
    if (code[e] || lv != 0.0) {
      double cs = 0.0;
      if (br == 1) cs = __ldg(p.hk + 2 * p.K + k0 + e) * c[e];
#pragma unroll
      for (int i = 0; i < NPX; ++i) lu[e][i] = (flag ? 0.0 : lu[e][i]) + Ls.aw[i] * lv + L.sl[i] * cs;
    }
  }
}

// P u as u_1 + P (u - u_1) (the rows of P sum to 1): a cell the limiter has flattened (all nodal values
// bitwise equal) stays exactly flat in the enriched space, so the location of max|u| inside it follows the
// lowest-index rule instead of rounding noise (oracle/burgers.py: prolong)
template <int NP, int NPX>
__device__ __forceinline__ void bg_prolong(const double* __restrict__ P, const double (&u)[NP], double (&x)[NPX]) {
  double d[NP];
#pragma unroll
  for (int j = 1; j < NP; ++j) d[j] = u[j] - u[0];
#pragma unroll
  for (int i = 0; i < NPX; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 1; j < NP; ++j) acc = fma(P[i * NP + j], d[j], acc);
    x[i] = u[0] + acc;
  }
}

__host__ __device__ constexpr size_t bg_fused_smem(int NP, int NPX, int EPT, int BD) {
  // 16 B mbarriers | exA, exB [2][BD+2] | wred [2][8] | mv [5] (+pad) | cand [2], am [5] (+pad) |
  // stage input states ss [5][NPX][EPT][BD]; the checkpoint tile [NP][EPT][BD] of the next step lands in ss[3]
  // while that slot is idle (between the transposes of stage 3 and the next step's stage 3)
  return 16 + sizeof(double) * ((size_t)4 * (BD + 2) + 16 + 6 + 4 + (size_t)5 * NPX * EPT * BD);
}

// ------------------------------------------------------------------------------------------------------
// the kernel.  IND = false: adjoint of the coarse march itself (lam0 = dJ/du0, [NP][K]); IND = true: the
// enriched adjoint + indicator (lam0 = lam_f^0, [NP+1][K]; eta[K]).
// ------------------------------------------------------------------------------------------------------
template <int NP, int EPT, int BD, bool IND>
__global__ void __launch_bounds__(BD, BG_MINB(BD, IND)) burgers_fused_kernel(const __grid_constant__ BgFusedArgs p) {
  constexpr int NPX = IND ? NP + 1 : NP;
  constexpr int LX = IND ? 1 : 0;
  using Ctx = BgCtx<BD>;
  constexpr int BDP = Ctx::BDP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sm = reinterpret_cast<double*>(smem_raw + 16);
  Ctx cx;
  cx.exA = sm;
  cx.exB = sm + 2 * BDP;
  cx.wred = sm + 4 * BDP;
  cx.mv = cx.wred + 16;
  cx.cand = reinterpret_cast<int*>(cx.mv + 6);
  cx.am = cx.cand + 2;
  double* ss = cx.mv + 6 + 4;                                   // [5][NPX][EPT][BD]
  double* land = ss + (size_t)3 * NPX * EPT * BD;               // [NP][EPT][BD] inside ss[3]
  constexpr size_t tile = (size_t)NP * EPT * BD;
  constexpr uint32_t tile_bytes = (uint32_t)(tile * sizeof(double));
  constexpr size_t sstride = (size_t)NPX * EPT * BD;

  const int tid = threadIdx.x, K = p.K;
  const int KT = K / EPT;
  cx.tid = tid;
  cx.lane = tid & 31;
  cx.wid = tid >> 5;
  cx.par = 0;
  cx.vs = 0;
  cx.in = tid < KT;
  const int k0 = cx.in ? tid * EPT : 0;
  cx.gfirst = cx.in && tid == 0 && !p.periodic;
  cx.glast = cx.in && tid == KT - 1 && !p.periodic;
  cx.nbL = cx.in ? (tid == 0 ? (p.periodic ? KT - 1 : BD) : tid - 1) : tid;
  cx.nbR = cx.in ? (tid == KT - 1 ? (p.periodic ? 0 : BD + 1) : tid + 1) : tid;
  // neighbour threads in the periodic sense (the ends of a non-periodic mesh override what they read there)
  const int nbLp = cx.in ? (tid == 0 ? KT - 1 : tid - 1) : tid, nbRp = cx.in ? (tid == KT - 1 ? 0 : tid + 1) : tid;
  cx.bar = &mbar[1];
  cx.phase = 0u;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);                      // TMA landing tile
    mbar_init(&mbar[1], (uint32_t)Ctx::NW);      // exchanges: one arrival per warp
    fence_mbar_init();
    cx.cand[0] = cx.cand[1] = 0x7fffffff;
  }
  fence_proxy_async();
  __syncthreads();
  uint32_t land_phase = 0u;
  const BgLevel& L0 = p.lv[0];
  const BgLevel& LXv = p.lv[LX];

#pragma unroll 1
  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    const double dt = p.dt_arr ? p.dt_arr[b] : p.dt;
    double* ck = p.ring + (size_t)blockIdx.x * (size_t)p.S * tile;
    int pend = -1;
    int nlim = 0;
    double u[EPT][NP];
    // ------------------------------------------------------------------ forward phase (coarse march)
    {
      BgCoef cf[EPT];
      double res[EPT][NP];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const double rx = cx.in ? L0.rxk[k0 + e] : 0.0, f0 = cx.in ? L0.fs0[k0 + e] : 0.0, f1 = cx.in ? L0.fs1[k0 + e] : 0.0;
        cf[e] = {-rx * dt / 4.0, -f0 * dt / 8.0, f1 * dt / 8.0};
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          u[e][i] = cx.in ? p.u0[((size_t)b * NP + i) * K + k0 + e] : 0.0;
          res[e][i] = 0.0;
        }
      }
      int code[EPT];
      if (p.limit) bg_limiter<NP, EPT, BD>(p, L0, 0, cx, k0, u, code);   // the pass on the initial state
#pragma unroll 1
      for (int n = 0; n < p.S; ++n) {
        double* dst = ck + (size_t)n * tile + tid;   // u^n -> ring (coalesced)
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NP; ++i) dst[(size_t)(i * EPT + e) * BD] = u[e][i];
        }
#pragma unroll 1
        for (int s = 0; s < 5; ++s) {
          bg_stage<NP, EPT, BD, false>(p, L0, cx, k0, cf, u, res, s, code, pend);
#pragma unroll
          for (int e = 0; e < EPT; ++e) nlim += code[e] & 1;
        }
      }
    }
    // ------------------------------------------------------------------ terminal state, J
    bool bad = false;
    if (cx.in) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          bad |= !isfinite(u[e][i]);
          if (p.uT) p.uT[((size_t)b * NP + i) * K + k0 + e] = u[e][i];
        }
      }
    }
    {
      double jp = 0.0;
      if (cx.in) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NP; ++i) jp = fma(p.jw[0][(size_t)i * K + k0 + e], u[e][i], jp);
        }
      }
      const double Jt = bg_block_sum<BD>(cx, jp);
      if (p.J && tid == 0) p.J[b] = Jt;
      if (p.nlim) {
        const double nl = bg_block_sum<BD>(cx, (double)nlim);
        if (tid == 0) p.nlim[2 * b] = (int)nl;
      }
      if (p.status) {
        const int anybad = __syncthreads_or(bad ? 1 : 0);
        if (tid == 0) p.status[b] = anybad ? 1u : 0u;
      }
    }
    // ------------------------------------------------------------------ adjoint phase
    // the state tiles were written through the generic proxy; TMA reads them through the async proxy
    __threadfence();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0 && p.S >= 1) {
      mbar_expect_tx(&mbar[0], tile_bytes);
      tma_bulk_g2s(land, ck + (size_t)(p.S - 1) * tile, tile_bytes, &mbar[0]);
    }
    double lu[EPT][NPX], eta[EPT];
    int nlim2 = 0;   // limiter activations of the steps taken again (the enriched ones in indicator mode)
    // P u^{n+1} of the step being transposed sits in ss[0] (left there by the step done before it; P u^S for
    // n = S-1): a thread's own column, so no exchange is involved
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      eta[e] = 0.0;
#pragma unroll
      for (int i = 0; i < NPX; ++i) lu[e][i] = cx.in ? p.jw[LX][(size_t)i * K + k0 + e] : 0.0;
      if (IND) {
        double px[NPX];
        bg_prolong<NP, NPX>(p.P, u[e], px);
#pragma unroll
        for (int i = 0; i < NPX; ++i) ss[(size_t)(i * EPT + e) * BD + tid] = px[i];
      }
    }
#pragma unroll 1
    for (int n = p.S - 1; n >= 0; --n) {
      double x[EPT][NPX], res[EPT][NPX];
      BgCoef cf[EPT];
      int codes[EPT];   // the step's limiter decisions, 3 bits per stage
      if (IND) {
        // eta_k += lam_f^{n+1}_k . rho^n_k with rho^n = P u^{n+1} - Phi_f(P u^n), in two parts: the first
        // now (P u^{n+1} is about to be overwritten), the second once the step has been taken again
        const double* s0 = ss + tid;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NPX; ++i) eta[e] = fma(lu[e][i], s0[(size_t)(i * EPT + e) * BD], eta[e]);
        }
      }
      mbar_wait(&mbar[0], land_phase & 1u);
      ++land_phase;
      {
        const double* src = land + tid;
        double un[EPT][NP];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NP; ++i) un[e][i] = src[(size_t)(i * EPT + e) * BD];
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          if (IND) {
            bg_prolong<NP, NPX>(p.P, un[e], x[e]);
          } else {
#pragma unroll
            for (int i = 0; i < NPX; ++i) x[e][i] = un[e][i < NP ? i : 0];
          }
          const double rx = cx.in ? LXv.rxk[k0 + e] : 0.0, f0 = cx.in ? LXv.fs0[k0 + e] : 0.0, f1 = cx.in ? LXv.fs1[k0 + e] : 0.0;
          cf[e] = {-rx * dt / 4.0, -f0 * dt / 8.0, f1 * dt / 8.0};
          codes[e] = 0;
#pragma unroll
          for (int i = 0; i < NPX; ++i) res[e][i] = 0.0;
        }
      }
      // ---- the step again (same routine, same bits), the five stage input states kept in shared memory
#pragma unroll 1
      for (int s = 0; s < 5; ++s) {
        {
          double* d = ss + (size_t)s * sstride + tid;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
#pragma unroll
            for (int i = 0; i < NPX; ++i) d[(size_t)(i * EPT + e) * BD] = x[e][i];
          }
        }
        int code[EPT];
        bg_stage<NPX, EPT, BD, true>(p, LXv, cx, k0, cf, x, res, s, code, pend);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          codes[e] |= code[e] << (3 * s);
          nlim2 += code[e] & 1;
        }
      }
      if (IND) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NPX; ++i) eta[e] = fma(-lu[e][i], x[e][i], eta[e]);
        }
      }
      // ---- transpose the stages in reverse
      double lk[EPT][NPX];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
#pragma unroll
        for (int i = 0; i < NPX; ++i) lk[e][i] = 0.0;
      }
#pragma unroll 1
      for (int s = 4; s >= 0; --s) {
        int code[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) code[e] = (codes[e] >> (3 * s)) & 7;
        bg_limiter_T<NPX, EPT, BD>(p, LXv, s, cx, k0, code, lu);
        bg_flush_vote<BD>(cx, pend);   // (after an exchange that follows the last stage's vote)
        if (s == 2 && tid == 0 && n >= 1) {
          // every thread is past the transpose of stage 3, the last reader of ss[3]: the state tile of step
          // n-1 lands there while stages 2..0 are transposed
          fence_proxy_async();
          mbar_expect_tx(&mbar[0], tile_bytes);
          tma_bulk_g2s(land, ck + (size_t)(n - 1) * tile, tile_bytes, &mbar[0]);
        }
        const double rka = p.rka[s], rkb = p.rkb[s];
        double us[EPT][NPX];
        double uLL, uRR;
        {
          // the stage's input state, and the neighbours' traces straight from their columns (every column was
          // written before an exchange of the stage that produced it)
          const double* d = ss + (size_t)s * sstride;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
#pragma unroll
            for (int i = 0; i < NPX; ++i) us[e][i] = d[(size_t)(i * EPT + e) * BD + tid];
          }
          uLL = d[(size_t)((NPX - 1) * EPT + (EPT - 1)) * BD + nbLp];
          uRR = d[(size_t)(0 * EPT + 0) * BD + nbRp];
        }
        if (cx.gfirst) uLL = us[0][0];                 // ghost = own trace
        if (cx.glast) uRR = us[EPT - 1][NPX - 1];
        const double mv = cx.mv[s];
        const BgStageOps& LXs = LXv.ops(s);
        double d0m[EPT], d0p[EPT], d1m[EPT], d1p[EPT];
        double gam = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NPX; ++i) lk[e][i] = fma(rkb, lu[e][i], lk[e][i]);
          double g0 = 0.0, g1 = 0.0;
#pragma unroll
          for (int i = 0; i < NPX; ++i) {
            g0 = fma(LXs.LIFT[i * 2], lk[e][i], g0);
            g1 = fma(LXs.LIFT[i * 2 + 1], lk[e][i], g1);
          }
          const double f0 = cx.in ? __ldg(LXv.fs0 + k0 + e) : 0.0, f1 = cx.in ? __ldg(LXv.fs1 + k0 + e) : 0.0;
          const double G0 = g0 * f0, G1 = g1 * f1;
          const double uL = (e == 0) ? uLL : us[e - 1][NPX - 1];
          const double uR = (e == EPT - 1) ? uRR : us[e + 1][0];
          d0m[e] = (-us[e][0] / 2.0 - mv / 2.0) * G0;
          d0p[e] = (uL / 2.0 + mv / 2.0) * G0;
          d1m[e] = (us[e][NPX - 1] / 2.0 - mv / 2.0) * G1;
          d1p[e] = (-uR / 2.0 + mv / 2.0) * G1;
          if (cx.in) gam += G0 * (-(us[e][0] - uL) / 2.0) + G1 * (-(us[e][NPX - 1] - uR) / 2.0);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gam += __shfl_xor_sync(0xffffffffu, gam, o);
        double* eA = cx.A();
        double* eB = cx.Bb();
        double* wr = cx.W();
        eA[tid] = d0p[0];         // belongs to the left neighbour's last node
        eB[tid] = d1p[EPT - 1];   // belongs to the right neighbour's first node
        if (Ctx::NW > 1 && cx.lane == 0) wr[cx.wid] = gam;
        cx.arrive();
        // volume part (no neighbour data), folded into lu at once: lu_j += dt us_j sum_i Dr_ij (-rx lk_i);
        // then the residual's own scaling lk *= rka (everything that reads lk of this stage is done)
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#if BG_TRANSPOSED_EO
          // the transpose of the forward volume term AS COMPUTED there (even/odd blocks): with w = (-rx dt / 2) lk,
          //   we_i = w_i + w_{N-i} (2 w_mid), wo_i = w_i - w_{N-i};  A_j = sum_i DE_ij we_i,  B_j = sum_i DO_ij wo_i;
          //   lu_j += us_j (A_j + B_j),  lu_{N-j} += us_{N-j} (B_j - A_j),  lu_mid += us_mid B_mid
          // -- half the multiply-adds and constant loads of the dense transpose
          constexpr int HE = (NPX + 1) / 2, HO = NPX / 2;
          const double mh = cx.in ? -0.5 * __ldg(LXv.rxk + k0 + e) * dt : 0.0;
          double we[HE], wo[HO > 0 ? HO : 1];
#pragma unroll
          for (int i = 0; i < HO; ++i) {
            const double a = mh * lk[e][i], b = mh * lk[e][NPX - 1 - i];
            we[i] = a + b;
            wo[i] = a - b;
          }
          if (NPX & 1) we[HO] = (mh + mh) * lk[e][HO];
#pragma unroll
          for (int j = 0; j < HE; ++j) {
            double Bj = 0.0;
#pragma unroll
            for (int i = 0; i < HO; ++i) {
              const double2 c2 = LXs.so.DO2[i * HP + j / 2];
              Bj = fma((j & 1) ? c2.y : c2.x, wo[i], Bj);
            }
            if (j < HO) {
              double Aj = 0.0;
#pragma unroll
              for (int i = 0; i < HE; ++i) {
                const double2 c2 = LXs.so.DE2[i * HP + j / 2];
                Aj = fma((j & 1) ? c2.y : c2.x, we[i], Aj);
              }
              lu[e][j] = fma(us[e][j], Aj + Bj, lu[e][j]);
              lu[e][NPX - 1 - j] = fma(us[e][NPX - 1 - j], Bj - Aj, lu[e][NPX - 1 - j]);
            } else {
              lu[e][j] = fma(us[e][j], Bj, lu[e][j]);   // the middle node of an odd node count
            }
          }
#else
          const double mrx = cx.in ? -__ldg(LXv.rxk + k0 + e) * dt : 0.0;
          double w[NPX];
#pragma unroll
          for (int i = 0; i < NPX; ++i) w[i] = mrx * lk[e][i];
#pragma unroll
          for (int j = 0; j < NPX; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < NPX; ++i) acc = fma(LXs.Dr[i * NPX + j], w[i], acc);
            lu[e][j] = fma(us[e][j], acc, lu[e][j]);
          }
#endif
#pragma unroll
          for (int i = 0; i < NPX; ++i) lk[e][i] *= rka;
        }
        cx.wait();
        double inN = eA[nbRp];   // right neighbour's d0p
        double in0 = eB[nbLp];   // left neighbour's d1p
        if (Ctx::NW > 1) {
          gam = wr[0];
#pragma unroll
          for (int w = 1; w < Ctx::NW; ++w) gam += wr[w];
        }
        cx.par ^= 1;
        if (cx.glast) inN = d1p[EPT - 1];   // ghost = own trace
        if (cx.gfirst) in0 = d0p[0];
        const int amv = cx.am[s];
        const int aflat = amv >> 1;
        const int ai = aflat / K, ak = aflat - ai * K;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          lu[e][0] = fma(dt, d0m[e] + ((e == 0) ? in0 : d1p[e - 1]), lu[e][0]);
          lu[e][NPX - 1] = fma(dt, d1m[e] + ((e == EPT - 1) ? inN : d0p[e + 1]), lu[e][NPX - 1]);
          if (cx.in && k0 + e == ak) {   // the element that held max|u|: the rank-one term of C
            const double add = (amv & 1) ? -gam : gam;
#pragma unroll
            for (int q = 0; q < NPX; ++q) lu[e][q] = fma(dt, (q == ai) ? add : 0.0, lu[e][q]);
          }
        }
      }
    }
    // ------------------------------------------------------------------ outputs
    if constexpr (!IND) if (p.limit) {
      // the limiter pass on the initial state: its decisions again, then its transpose
      double v0[EPT][NP];
      int code[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
#pragma unroll
        for (int i = 0; i < NP; ++i) v0[e][i] = cx.in ? p.u0[((size_t)b * NP + i) * K + k0 + e] : 0.0;
      }
      bg_limiter<NP, EPT, BD>(p, L0, 0, cx, k0, v0, code);
      // (NPX == NP here)
      bg_limiter_T<NPX, EPT, BD>(p, LXv, 0, cx, k0, code, lu);
    }
    if (p.nlim) {
      const double nl = bg_block_sum<BD>(cx, (double)nlim2);
      if (tid == 0) p.nlim[2 * b + 1] = (int)nl;
    }
    if (cx.in) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        if (p.eta && IND) p.eta[(size_t)b * K + k0 + e] = eta[e];
        if (p.lam0) {
#pragma unroll
          for (int i = 0; i < NPX; ++i) p.lam0[((size_t)b * NPX + i) * K + k0 + e] = lu[e][i];
        }
      }
    }
    __syncthreads();   // shared memory (stage states, landing tile, exchanges) free before the next trajectory
    bg_flush_vote<BD>(cx, pend);
    __syncthreads();
  }
}

template <int NP, int EPT, int BD, bool IND>
static cudaError_t bgf_launch_one(int grid, size_t smem, cudaStream_t st, const BgFusedArgs& a) {
  auto kern = burgers_fused_kernel<NP, EPT, BD, IND>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, BD, smem, st>>>(a);
  return cudaGetLastError();
}

template <int NP, int EPT, bool IND>
static cudaError_t bgf_launch_bd(int block, int grid, size_t smem, cudaStream_t st, const BgFusedArgs& a) {
  switch (block) {
    case 32: return bgf_launch_one<NP, EPT, 32, IND>(grid, smem, st, a);
    case 64: return bgf_launch_one<NP, EPT, 64, IND>(grid, smem, st, a);
    case 128: return bgf_launch_one<NP, EPT, 128, IND>(grid, smem, st, a);
    case 256: return bgf_launch_one<NP, EPT, 256, IND>(grid, smem, st, a);
    default: return cudaErrorInvalidConfiguration;
  }
}

#ifdef DGADJ_BGF_NP
// one translation unit per primal order (-DDGADJ_BGF_NP=2..9, built in parallel): the kernels of that order
#define DGADJ_CAT2(a, b) a##b
#define DGADJ_CAT(a, b) DGADJ_CAT2(a, b)
cudaError_t DGADJ_CAT(bgf_launch_np, DGADJ_BGF_NP)(int ept, int block, int grid, size_t smem, bool ind, cudaStream_t st,
                                                   const BgFusedArgs& a) {
  constexpr int NP = DGADJ_BGF_NP;
  if (ind) {
    if constexpr (NP + 1 <= MAXNP) {
      switch (ept) {
        case 1: return bgf_launch_bd<NP, 1, true>(block, grid, smem, st, a);
        case 2: return bgf_launch_bd<NP, 2, true>(block, grid, smem, st, a);
        case 4: return bgf_launch_bd<NP, 4, true>(block, grid, smem, st, a);
      }
    }
    return cudaErrorInvalidConfiguration;
  }
  switch (ept) {
    case 1: return bgf_launch_bd<NP, 1, false>(block, grid, smem, st, a);
    case 2: return bgf_launch_bd<NP, 2, false>(block, grid, smem, st, a);
    case 4: return bgf_launch_bd<NP, 4, false>(block, grid, smem, st, a);
  }
  return cudaErrorInvalidConfiguration;
}
#else
#define DGADJ_DECL_BGF(n) cudaError_t bgf_launch_np##n(int, int, int, size_t, bool, cudaStream_t, const BgFusedArgs&);
DGADJ_DECL_BGF(2) DGADJ_DECL_BGF(3) DGADJ_DECL_BGF(4) DGADJ_DECL_BGF(5) DGADJ_DECL_BGF(6) DGADJ_DECL_BGF(7) DGADJ_DECL_BGF(8) DGADJ_DECL_BGF(9)
#endif

}  // namespace dgadj

#ifndef DGADJ_BGF_NP
// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static void bgf_limiter_weights(int Np, const double* Dr, const double* V, const double* invV, const double* x, int K,
                                BgLevel* L) {
  // aw = V(1,1)*invV(1,:) (SlopeLimitN.m:9);  sl = Dr(1,:)*V(:,1:2)*invV(1:2,:) (SlopeLimitN.m:27,
  // SlopeLimitLin.m:16);  xcn = (x - x0)/(h/2) of the first element (the map is affine: the reference nodes)
  for (int i = 0; i < Np; ++i) L->st[0].aw[i] = V[0] * invV[i];
  for (int i = 0; i < Np; ++i) {
    double sum = 0.0;
    for (int j = 0; j < Np; ++j)
      sum += Dr[j] * (V[(size_t)j * Np + 0] * invV[0 * Np + i] + V[(size_t)j * Np + 1] * invV[1 * Np + i]);
    L->sl[i] = sum;
  }
  const double hh = x[(size_t)(Np - 1) * K] - x[0];
  const double x0 = x[0] + hh / 2;
  for (int i = 0; i < Np; ++i) L->xcn[i] = (x[(size_t)i * K] - x0) / (hh / 2);
}

// device buffer of the handle, grown on demand (stream-ordered use: one stream at a time per handle)
static int bgf_ensure(dgadj_handle* h, double** buf, size_t* have, size_t need_bytes) {
  if (need_bytes <= *have) return DGADJ_OK;
  CUDA_TRY(h, cudaDeviceSynchronize());
  cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  CUDA_TRY(h, cudaMalloc((void**)buf, need_bytes));
  *have = need_bytes;
  return DGADJ_OK;
}

extern "C" int dgadj_burgers_plan(dgadj_handle* h, int64_t B, int32_t indicator, int32_t* ept, int32_t* block,
                                  int32_t* grid, int64_t* smem_bytes, int64_t* ring_bytes_per_step) {
  if (!h || B <= 0) return DGADJ_ERR_INVALID;
  const int Np = h->Np, K = h->K, NpX = indicator ? Np + 1 : Np;
  if (indicator && NpX > MAXNP) return fail(h, DGADJ_ERR_UNSUPPORTED, "the indicator needs N <= %d", MAXNP - 2);
  int e = h->tune_ept ? h->tune_ept : ((K % 2 == 0) ? 2 : 1);
  while (K % e) e >>= 1;
  while (K / e > 256 && e < 4 && K % (e * 2) == 0) e *= 2;
  const int KT = K / e;
  if (KT > 256) return fail(h, DGADJ_ERR_UNSUPPORTED, "K = %d does not fit one CTA of the fused Burgers kernel", K);
  int bd = 32;
  while (bd < KT) bd *= 2;
  const size_t smem = bg_fused_smem(Np, NpX, e, bd);
  if (smem > 227 * 1024)
    return fail(h, DGADJ_ERR_UNSUPPORTED, "the stage states of one trajectory (%zu B) do not fit shared memory: use "
                "dgadj_burgers_forward / dgadj_burgers_adjoint", smem);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, std::min(32, 2048 / bd)));
  {   // registers: what __launch_bounds__(bd, BG_MINB) lets a thread use
    const int minb = BG_MINB(bd, indicator ? 1 : 0);
    const int regs = std::min(255, (65536 / (bd * minb)) / 8 * 8);
    per_sm = std::min(per_sm, std::max(minb, 65536 / (bd * regs)));
  }
  int g = h->tune_grid ? h->tune_grid : h->sm_count * per_sm;
  g = (int)std::max<int64_t>(1, std::min<int64_t>(g, B));
  if (ept) *ept = e;
  if (block) *block = bd;
  if (grid) *grid = g;
  if (smem_bytes) *smem_bytes = (int64_t)smem;
  if (ring_bytes_per_step) *ring_bytes_per_step = (int64_t)((size_t)g * Np * e * bd * sizeof(double));
  return DGADJ_OK;
}

extern "C" int dgadj_burgers_fwd_adj(dgadj_handle* h, const dgadj_burgers_args* a, const double* u0_dev, double* uT_dev,
                                     double* J_dev, double* lam0_dev, double* eta_dev, int32_t* nlim_dev,
                                     uint32_t* status_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (!a || a->B <= 0 || a->S < 0 || !u0_dev) return fail(h, DGADJ_ERR_INVALID, "bad burgers_fwd_adj arguments");
  if (!h->ops_set) return fail(h, DGADJ_ERR_STATE, "dgadj_set_operators has not been called");
  if (h->nstages != 5) return fail(h, DGADJ_ERR_UNSUPPORTED, "the Burgers march is LSERK4 only");
  if (a->limit < 0 || a->limit > 2 || !(a->tvb_M >= 0.0)) return fail(h, DGADJ_ERR_INVALID, "limit must be 0, 1 or 2 and tvb_M >= 0");
  if (!a->invV_host || !a->V_host || !a->x_host || !a->jw_host) return fail(h, DGADJ_ERR_INVALID, "V, invV, x and jw of the primal space are required");
  const bool ind = a->indicator != 0;
  if (ind && (!h->enr_set || !a->invVF_host || !a->VF_host || !a->xF_host || !a->jwF_host))
    return fail(h, DGADJ_ERR_STATE, "the indicator needs dgadj_set_enriched and VF, invVF, xF, jwF of the enriched space");
  if (!ind && eta_dev) return fail(h, DGADJ_ERR_INVALID, "eta is produced in indicator mode only");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Np = h->Np, K = h->K, NpF = h->NpF;
  int ept = 0, block = 0, grid = 0;
  int64_t smem = 0, ring_step = 0;
  int rc = dgadj_burgers_plan(h, a->B, ind, &ept, &block, &grid, &smem, &ring_step);
  if (rc) return rc;

  BgFusedArgs k;
  memset(&k, 0, sizeof(k));
  k.B = a->B;
  k.K = K;
  k.S = a->S;
  k.periodic = (h->cfg.bc == DGADJ_BC_PERIODIC);
  k.limit = a->limit;
  k.tvbM = a->tvb_M;
  k.eps0 = (a->limit == 2) ? -1.0 : 1.0e-8;   // SlopeLimit1 mode: every cell (utils/SlopeLimit1.m:10-22)
  k.dt = a->dt;
  k.dt_arr = a->dt_dev;
  for (int s = 0; s < 5; ++s) {
    k.rka[s] = h->cops.rka[s];
    k.rkb[s] = h->cops.rkb[s];
  }
  for (int lv = 0; lv < (ind ? 2 : 1); ++lv) {
    BgLevel& L = k.lv[lv];
    const int n = lv ? NpF : Np;
    const double* Dr = lv ? h->Dr_nodal_f : h->Dr_nodal;
    const double* LIFT = lv ? h->LIFT_nodal_f : h->LIFT_nodal;
    BgStageOps& S0 = L.st[0];
    S0.so = h->base_ops[lv];
    for (int i = 0; i < n * n; ++i) S0.Dr[i] = Dr[i];
    for (int i = 0; i < n * 2; ++i) S0.LIFT[i] = LIFT[i];
    bgf_limiter_weights(n, Dr, lv ? a->VF_host : a->V_host, lv ? a->invVF_host : a->invV_host, lv ? a->xF_host : a->x_host, K, &L);
    for (int q = 1; q < (BG_STAGE_CONSTS ? 5 : 1); ++q) L.st[q] = S0;
    L.rxk = h->d_mesh[lv][0];
    L.fs0 = h->d_mesh[lv][1];
    L.fs1 = h->d_mesh[lv][2];
  }
  if (ind)
    for (int i = 0; i < NpF * Np; ++i) k.P[i] = h->P_host[i];
  // per-call constants: element widths and functional weights (pageable host memory: the copies are staged
  // before cudaMemcpyAsync returns, and ordered on the stream)
  const size_t nconst = (size_t)3 * K + (size_t)Np * K + (size_t)NpF * K;
  rc = bgf_ensure(h, &h->bgf_consts, &h->bgf_consts_bytes, nconst * sizeof(double));
  if (rc) return rc;
  {
    std::vector<double> hk((size_t)3 * K);
    for (int e = 0; e < K; ++e) {
      const double w = a->x_host[(size_t)(Np - 1) * K + e] - a->x_host[e];
      hk[e] = w;
      hk[K + e] = 1.0 / w;
      hk[2 * K + e] = 2.0 / w;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->bgf_consts, hk.data(), (size_t)3 * K * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(h->bgf_consts + 3 * K, a->jw_host, (size_t)Np * K * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ind)
      CUDA_TRY(h, cudaMemcpyAsync(h->bgf_consts + 3 * K + (size_t)Np * K, a->jwF_host, (size_t)NpF * K * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  k.hk = h->bgf_consts;
  k.jw[0] = h->bgf_consts + 3 * K;
  k.jw[1] = h->bgf_consts + 3 * K + (size_t)Np * K;
  // the per-CTA state ring
  const size_t need = (size_t)ring_step * (size_t)std::max(a->S, 1);
  if (need > h->ring_bytes) {
    if (h->ring) {
      CUDA_TRY(h, cudaDeviceSynchronize());
      cudaFree(h->ring);
      h->ring = nullptr;
      h->ring_bytes = 0;
    }
    const cudaError_t e = cudaMalloc((void**)&h->ring, need);
    if (e != cudaSuccess) {
      cudaGetLastError();
      h->ring = nullptr;
      return fail(h, DGADJ_ERR_NOMEM, "state ring of %.1f GB (%d CTAs x %d steps) does not fit on the device", need * 1e-9, grid, a->S);
    }
    h->ring_bytes = need;
  }
  k.ring = h->ring;
  k.u0 = u0_dev;
  k.uT = uT_dev;
  k.J = J_dev;
  k.lam0 = lam0_dev;
  k.eta = eta_dev;
  k.nlim = nlim_dev;
  k.status = status_dev;
  cudaError_t e = cudaSuccess;
#define DGADJ_BGF(n) case n: e = bgf_launch_np##n(ept, block, grid, (size_t)smem, ind, st, k); break;
  switch (Np) {
    DGADJ_BGF(2) DGADJ_BGF(3) DGADJ_BGF(4) DGADJ_BGF(5) DGADJ_BGF(6) DGADJ_BGF(7) DGADJ_BGF(8) DGADJ_BGF(9)
    default: return fail(h, DGADJ_ERR_UNSUPPORTED, "the fused Burgers march supports 1 <= N <= 8");
  }
#undef DGADJ_BGF
  if (e != cudaSuccess) return fail(h, DGADJ_ERR_CUDA, "fused burgers kernel launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  return DGADJ_OK;
}
#endif  // !DGADJ_BGF_NP
