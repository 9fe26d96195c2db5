// dgadj_kernels.cuh -- persistent sm_100a kernels for the batched 1-D nodal-DG march.
//
// Mapping: one thread owns one element of one trajectory for the whole march; its Np nodal
// values, the RK residual and its metric terms live in registers.  A CTA owns `tpc`
// trajectories (tpc*K <= blockDim) and loops over trajectory groups (persistent grid).
// Dr / LIFT / P / Mref / RK coefficients travel as a __grid_constant__ kernel parameter, i.e.
// they sit in constant bank 0 and feed DFMA as constant-bank operands (all operator loops
// are fully unrolled on the template order).  Neighbour traces cross threads through a
// double-buffered shared-memory pair and a split arrive/wait mbarrier per stage.
// HBM is touched for: the initial state, the final state, one coalesced checkpoint tile
// per step (forward phase, STG) which the adjoint phase streams back with bulk-TMA
// (cp.async.bulk + mbarrier, double buffered), and the outputs.
//
// Reference computations replaced (see include/dgadj.h): utils/AdvecRHS1D.m:8-19, the
// LSERK4 loop of utils/One_code.mlx, matlab/adj_march.m:67-118 (conventions), errEst of
// python/Main_finite_difference.py:79-94 (conventions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgadj {

constexpr int MAXNP = 10;      // enriched space of N = 8
constexpr int MAXSTAGES = 5;
constexpr int MAXBD = 1024;

constexpr int HM = 5;  // max half dimension of the even/odd blocks ((MAXNP+1)/2)

// Operator blocks of one space in the even/odd basis (see EO<> in the device section):
//   even-out = DE * odd-in  (+ LS * (g0+g1)),   odd-out = DO * even-in  (+ LA * (g0-g1))
// Stored as double2 pairs (row stride HP pairs) so that one 128-bit uniform constant load
// (LDCU.128) feeds two DFMA per element: entry (i, j) is DE2[i*HP + j/2].{x,y}[j & 1].
constexpr int HP = 3;  // pairs per row ((HM+1)/2)
struct alignas(16) StageOps {
  double2 DE2[HM * HP];  // i < HE, j < HO
  double2 DO2[HM * HP];  // i < HO, j < HE
  double2 LS2[HP];
  double2 LA2[HP];
};
struct ProlongOps {
  double PE[HM * HM];  // even block of T_f P T_c^-1 : [i*HM + j], i < HE_f, j < HE_c
  double PO[HM * HM];  // odd block                  : [i*HM + j], i < HO_f, j < HO_c
};
// Stage scaling.  The low-storage recurrences  r_s = rka_s r_{s-1} + R_s  (forward) and
// w_s = rka_{s+1} w_{s+1} + bm_s mu_s  (adjoint) are carried as r = sig_s r~, w = sga_s w~ with
// sig_0 = 1, sig_s = rka_s sig_{s-1} and sga_last = 1, sga_s = rka_{s+1} sga_{s+1}: the rka
// multiply disappears (r~_s = r~_{s-1} + R_s / sig_s) and the factors 1/sig_s, sga_s are folded
// into the per-stage copies of the operator blocks, which exist anyway (see fwd_step).
struct alignas(16) ConstOps {
  StageOps st[2][MAXSTAGES];      // [level][stage]: forward blocks, scaled by 1/sig_s
  StageOps sta[MAXSTAGES];        // enriched level, adjoint sweep: blocks scaled by sga_s
  double bsig[MAXSTAGES];         // rkb_s * sig_s   (state update  z += bsig_s m r~)
  double bsga[MAXSTAGES];         // rkb_s / sga_s   (adjoint update w~ += bsga_s m mu)
  ProlongOps pr[2];               // identical copies (indexed by step parity)
  double Mref[2][MAXNP * MAXNP];  // nodal reference mass matrices inv(V V') (J = int u^2)
  double P[MAXNP * MAXNP];        // nodal prolongation [NPF][NP], row stride NP
  double rka[MAXSTAGES], rkb[MAXSTAGES], rkc[MAXSTAGES];
};

struct MarchParams {
  long long B;
  int K, S, tpc, ngroups, nstages, bc, inflow, func;
  int warp_local;         // K/EPT divides 32: trajectories never straddle warps
  double alpha, a, dt, t0;
  const double* a_arr;
  const double* dt_arr;
  const double* rxk[2];   // [level][K]   rx(1,k)
  const double* fs0[2];   // [level][K]   Fscale(1,k)
  const double* fs1[2];   // [level][K]   Fscale(2,k)
  const double* jw_c;     // [NP][K]
  const double* jw_f;     // [NPF][K]
  const double* uin_table;
  const double* u0;
  double* uT;             // forward: written
  const double* uT_in;    // adjoint-only: terminal primal
  double* hist;           // [B][S+1][NP][K] or null
  double* ckpt;           // tiles [slot][S][NPF][EPT][BD]
  int ckpt_by_block;      // 1: slot = blockIdx.x (fused ring); 0: slot = group
  double* J;
  double* lam0;
  double* eta;
};

struct KArgs {
  MarchParams p;
  alignas(16) ConstOps c;
};

enum { BC_INFLOW = 0, BC_PERIODIC = 1 };
enum { INFLOW_ZERO = 0, INFLOW_SIN_AT = 1, INFLOW_SIN_AAT = 2, INFLOW_TABLE = 3 };
enum { FUNC_LINEAR = 0, FUNC_INT_U2 = 1 };
enum { VAR_FWD = 0, VAR_FWD_RESID = 1, VAR_ADJ = 2, VAR_FUSED = 3 };

// shared memory: 16-byte mbarrier header, then doubles tr[4][BD] coef[6*EPT][BD] big[nbig*EPT][BD]
__host__ __device__ constexpr size_t march_smem_bytes(int NP, int EPT, int BD, int variant) {
  const int nbig = (variant == VAR_FWD) ? 0 : NP + 1;
  return 16 + sizeof(double) * (size_t)(4 + (6 + nbig) * EPT) * (size_t)BD;
}

#if defined(__CUDACC__) && defined(DGADJ_DEVICE_CODE)
// ---------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + bulk TMA (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// per-thread context.  A thread owns EPT adjacent elements of one trajectory.  Everything
// that is not RK state lives in shared memory or in the constant bank.
// ---------------------------------------------------------------------------------------
enum { CX_FIRST = 1, CX_LAST = 2, CX_PERIODIC = 4 };
struct Ctx {
  int tid, BD, nbL, nbR, par;
  int flags;  // bit0 owns the first element, bit1 owns the last element, bit2 periodic
  uint64_t* trbar;    // mbarrier of the trace exchange (one arrival per warp)
  uint32_t trphase;   // its phase parity
};

#ifndef DGADJ_SPLIT_BARRIER
#define DGADJ_SPLIT_BARRIER 1
#endif
// Trace exchange synchronisation.  Split form: a warp *arrives* as soon as its traces are in
// shared memory and *waits* only when it needs its neighbours' -- the volume terms (80 % of a
// stage) sit in between, so warps rarely block.  DGADJ_SPLIT_BARRIER=0 keeps a plain
// __syncthreads() at the arrive point (for A/B measurements).
// warp_local (a kernel parameter, hence uniform): every trajectory lives inside one warp, so
// __syncwarp alone orders the exchange and the warps of a CTA never wait for one another.
__device__ __forceinline__ void trace_arrive(Ctx& cx, int warp_local) {
#if DGADJ_SPLIT_BARRIER
  __syncwarp();
  if (!warp_local && (cx.tid & 31) == 0) mbar_arrive(cx.trbar);
#else
  __syncthreads();
#endif
}
__device__ __forceinline__ void trace_wait(Ctx& cx, int warp_local) {
#if DGADJ_SPLIT_BARRIER
  if (!warp_local) {
    mbar_wait(cx.trbar, cx.trphase);
    cx.trphase ^= 1u;
  }
#endif
}

static __device__ __noinline__ double inflow_value(const MarchParams& p, long long b, double time, int n, int s,
                                            double rkc) {
  const double a = p.a_arr ? p.a_arr[b] : p.a;
  const double dt = p.dt_arr ? p.dt_arr[b] : p.dt;
  const double t = time + rkc * dt;
  switch (p.inflow) {
    case INFLOW_SIN_AT: return -sin(a * t);
    case INFLOW_SIN_AAT: return -sin(a * a * t);
    case INFLOW_TABLE: return p.uin_table[n * p.nstages + s];
    default: return 0.0;
  }
}

// Symmetric / antisymmetric ("even/odd") nodal representation of one element.  LGL nodes are
// mirror-symmetric, so Dr is centro-antisymmetric and LIFT / P are centro-symmetric; in
//     ze[i] = u[i] + u[N-i] (i < Np/2),  ze[mid] = u[mid] (Np odd),   zo[i] = u[i] - u[N-i]
// Dr maps odd -> even and even -> odd (two half-size blocks DE, DO), LIFT maps g0+g1 -> even
// and g0-g1 -> odd, P maps even -> even and odd -> odd.  The whole march runs in this
// basis: Np^2/2 instead of Np^2 DFMA per mat-vec.  The host builds the blocks from the
// caller's nodal Dr / LIFT / P (dgadj_api.cu: dgadj_build_const_ops) and rejects operators
// that do not have the symmetry.
template <int NPX>
struct EO {
  static constexpr int HE = (NPX + 1) / 2;  // even (symmetric) dimension, holds the mid node
  static constexpr int HO = NPX / 2;        // odd (antisymmetric) dimension
};

// One element's state in the even/odd basis (row order everywhere: e[0..HE), o[0..HO)).
template <int NPX>
struct EOVec {
  double e[EO<NPX>::HE];
  double o[EO<NPX>::HO];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < EO<NPX>::HE; ++i) e[i] = 0.0;
#pragma unroll
    for (int i = 0; i < EO<NPX>::HO; ++i) o[i] = 0.0;
  }
  // rows <-> a strided column (shared memory park / landing tile / checkpoint tile)
  __device__ __forceinline__ void load(const double* col, size_t stride) {
#pragma unroll
    for (int i = 0; i < EO<NPX>::HE; ++i) e[i] = col[(size_t)i * stride];
#pragma unroll
    for (int i = 0; i < EO<NPX>::HO; ++i) o[i] = col[(size_t)(EO<NPX>::HE + i) * stride];
  }
  __device__ __forceinline__ void store(double* col, size_t stride) const {
#pragma unroll
    for (int i = 0; i < EO<NPX>::HE; ++i) col[(size_t)i * stride] = e[i];
#pragma unroll
    for (int i = 0; i < EO<NPX>::HO; ++i) col[(size_t)(EO<NPX>::HE + i) * stride] = o[i];
  }
  // z = T u
  __device__ __forceinline__ void from_nodal(const double (&u)[NPX]) {
#pragma unroll
    for (int i = 0; i < NPX / 2; ++i) {
      e[i] = u[i] + u[NPX - 1 - i];
      o[i] = u[i] - u[NPX - 1 - i];
    }
    if (NPX & 1) e[NPX / 2] = u[NPX / 2];
  }
  // u = T^-1 z  (half = 0.5)  or  lam = T^T mu (half = 1.0)
  __device__ __forceinline__ void to_nodal(double (&u)[NPX], double half) const {
#pragma unroll
    for (int i = 0; i < NPX / 2; ++i) {
      u[i] = half * (e[i] + o[i]);
      u[NPX - 1 - i] = half * (e[i] - o[i]);
    }
    if (NPX & 1) u[NPX / 2] = e[NPX / 2];
  }
};

// One RK stage in the scaled-residual even/odd form, for the EPT elements of a thread (each
// operator constant is fetched once and used EPT times).  With m = -a*rx*dt (constant per
// element and trajectory) and r = resu/m the reference update (utils/AdvecRHS1D.m:11,19 +
// the mlx loop)   resu = rka*resu + dt*rhsu ;  u = u + rkb*resu   becomes
//     re = rka*re + DE zo + LS (g0+g1) ;  ro = rka*ro + DO ze + LA (g0-g1) ;  z += (rkb*m) r
// (the rka multiply is absorbed by the stage scaling described at ConstOps)
//     g0 = (u[0]-uL)*q0, g1 = (u[N]-uR)*q1,  q_f = dt*Fscale_f*c_f/m
template <int NPX, int EPT>
__device__ __forceinline__ void fwd_stage_volume(const StageOps& so, const EOVec<NPX> (&z)[EPT],
                                                 EOVec<NPX> (&r)[EPT]) {
  constexpr int HE = EO<NPX>::HE, HO = EO<NPX>::HO;
#pragma unroll
  for (int i = 0; i < HE; ++i) {
    double acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) acc[e] = r[e].e[i];
#pragma unroll
    for (int jp = 0; jp < (HO + 1) / 2; ++jp) {
      const double2 c2 = so.DE2[i * HP + jp];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        acc[e] = fma(c2.x, z[e].o[2 * jp], acc[e]);
        if (2 * jp + 1 < HO) acc[e] = fma(c2.y, z[e].o[2 * jp + 1 < HO ? 2 * jp + 1 : 0], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) r[e].e[i] = acc[e];
  }
#pragma unroll
  for (int i = 0; i < HO; ++i) {
    double acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) acc[e] = r[e].o[i];
#pragma unroll
    for (int jp = 0; jp < (HE + 1) / 2; ++jp) {
      const double2 c2 = so.DO2[i * HP + jp];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        acc[e] = fma(c2.x, z[e].e[2 * jp], acc[e]);
        if (2 * jp + 1 < HE) acc[e] = fma(c2.y, z[e].e[2 * jp + 1 < HE ? 2 * jp + 1 : 0], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) r[e].o[i] = acc[e];
  }
}
template <int NPX, int EPT>
__device__ __forceinline__ void fwd_stage_surface(const StageOps& so, EOVec<NPX> (&z)[EPT], EOVec<NPX> (&r)[EPT],
                                                  const double (&g0)[EPT], const double (&g1)[EPT],
                                                  const double (&bm)[EPT]) {
  constexpr int HE = EO<NPX>::HE, HO = EO<NPX>::HO;
  double ge[EPT], go[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    ge[e] = g0[e] + g1[e];
    go[e] = g0[e] - g1[e];
  }
#pragma unroll
  for (int ip = 0; ip < (HE + 1) / 2; ++ip) {
    const double2 c2 = so.LS2[ip];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      r[e].e[2 * ip] = fma(c2.x, ge[e], r[e].e[2 * ip]);
      z[e].e[2 * ip] = fma(bm[e], r[e].e[2 * ip], z[e].e[2 * ip]);
      if (2 * ip + 1 < HE) {
        const int i1 = 2 * ip + 1 < HE ? 2 * ip + 1 : 0;
        r[e].e[i1] = fma(c2.y, ge[e], r[e].e[i1]);
        z[e].e[i1] = fma(bm[e], r[e].e[i1], z[e].e[i1]);
      }
    }
  }
#pragma unroll
  for (int ip = 0; ip < (HO + 1) / 2; ++ip) {
    const double2 c2 = so.LA2[ip];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      r[e].o[2 * ip] = fma(c2.x, go[e], r[e].o[2 * ip]);
      z[e].o[2 * ip] = fma(bm[e], r[e].o[2 * ip], z[e].o[2 * ip]);
      if (2 * ip + 1 < HO) {
        const int i1 = 2 * ip + 1 < HO ? 2 * ip + 1 : 0;
        r[e].o[i1] = fma(c2.y, go[e], r[e].o[i1]);
        z[e].o[i1] = fma(bm[e], r[e].o[i1], z[e].o[i1]);
      }
    }
  }
}

// One full RK step (all stages) with the neighbour-trace exchange.
// coef = this thread's smem column of the level: {m, q0, q1} x EPT, each a row of BD.
// The operator blocks are read from a per-stage copy (c.st[LV][s]) so that the constant
// loads depend on the stage counter and stay inside the loop as uniform loads instead of
// being hoisted out of it into (and spilled from) the register budget.
template <int NPX, int LV, int EPT>
__device__ __forceinline__ void fwd_step(const KArgs& ka, Ctx& cx, double* __restrict__ tr,
                                         const double* __restrict__ coef, EOVec<NPX> (&z)[EPT],
                                         EOVec<NPX> (&r)[EPT], long long b, double time, int n) {
  const ConstOps& c = ka.c;
  const int nst = ka.p.nstages;
#pragma unroll 1   // (fully unrolling the stages was measured 7 % slower: 5x the code, more spills)
  for (int s = 0; s < nst; ++s) {
    const StageOps& so = c.st[LV][s];
    double* tA = tr + cx.par * cx.BD;  // left-edge values  u[0]    of the thread's first element
    double* tB = tA + 2 * cx.BD;       // right-edge values u[Np-1] of the thread's last element
    tA[cx.tid] = 0.5 * (z[0].e[0] + z[0].o[0]);
    tB[cx.tid] = 0.5 * (z[EPT - 1].e[0] - z[EPT - 1].o[0]);
    trace_arrive(cx, ka.p.warp_local);
    fwd_stage_volume<NPX, EPT>(so, z, r);   // needs no neighbour data
    trace_wait(cx, ka.p.warp_local);
    double uL = tB[cx.nbL];
    double uR = tA[cx.nbR];
    cx.par ^= 1;
    double uF[EPT], uB[EPT];  // u[0], u[Np-1] of each element
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      uF[e] = 0.5 * (z[e].e[0] + z[e].o[0]);
      uB[e] = 0.5 * (z[e].e[0] - z[e].o[0]);
    }
    if (!(cx.flags & CX_PERIODIC)) {
      if (cx.flags & CX_FIRST) uL = inflow_value(ka.p, b, time, n, s, c.rkc[s]);
      if (cx.flags & CX_LAST) uR = uB[EPT - 1];
    }
    double g0[EPT], g1[EPT], bm[EPT];
    const double rkb = c.bsig[s];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double left = (e == 0) ? uL : uB[e - 1];
      const double right = (e == EPT - 1) ? uR : uF[e + 1];
      g0[e] = (uF[e] - left) * coef[(size_t)(1 * EPT + e) * cx.BD];
      g1[e] = (uB[e] - right) * coef[(size_t)(2 * EPT + e) * cx.BD];
      bm[e] = rkb * coef[(size_t)e * cx.BD];
    }
    fwd_stage_surface<NPX, EPT>(so, z, r, g0, g1, bm);
  }
}

// Reverse of one RK step: stages s = last..0  
//   lk += rkb*lu ; lu += dt*L^T lk ; lk *= rka,   dt*L^T lk = m*Dr^T lk + scatter(g),
//   g_f = dt*Fscale_f*c_f * (LIFT[:,f] . lk).
// Carried in the scaled even/odd form (mu = T^-T lu, w = m * T^-T lk):
//   w += (rkb*m) mu ; G = {LS.we + LA.wo, LS.we - LA.wo} ; gam_f = q_f G_f ;
//   mu_e += DO^T wo ; mu_o += DE^T we ; a0 = gam0 - gam1[left], aN = gam1 - gam0[right] ;
//   mu_e[0] += (a0+aN)/2 ; mu_o[0] += (a0-aN)/2 ; w *= rka  (the last as a stage scaling, see ConstOps).
template <int NPX, int LV, int EPT>
__device__ __forceinline__ void adj_step(const KArgs& ka, Ctx& cx, double* __restrict__ tr,
                                         const double* __restrict__ coef, EOVec<NPX> (&mu)[EPT],
                                         EOVec<NPX> (&w)[EPT]) {
  constexpr int HE = EO<NPX>::HE, HO = EO<NPX>::HO;
  const ConstOps& c = ka.c;
#pragma unroll 1
  for (int s = ka.p.nstages - 1; s >= 0; --s) {
    const StageOps& so = c.sta[s];
    const double rkb = c.bsga[s];
    double gam0[EPT], gam1[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double bm = rkb * coef[(size_t)e * cx.BD];
#pragma unroll
      for (int i = 0; i < HE; ++i) w[e].e[i] = fma(bm, mu[e].e[i], w[e].e[i]);
#pragma unroll
      for (int i = 0; i < HO; ++i) w[e].o[i] = fma(bm, mu[e].o[i], w[e].o[i]);
    }
    {
      double Ge[EPT], Go[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) Ge[e] = Go[e] = 0.0;
#pragma unroll
      for (int ip = 0; ip < (HE + 1) / 2; ++ip) {
        const double2 c2 = so.LS2[ip];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          Ge[e] = fma(c2.x, w[e].e[2 * ip], Ge[e]);
          if (2 * ip + 1 < HE) Ge[e] = fma(c2.y, w[e].e[2 * ip + 1 < HE ? 2 * ip + 1 : 0], Ge[e]);
        }
      }
#pragma unroll
      for (int ip = 0; ip < (HO + 1) / 2; ++ip) {
        const double2 c2 = so.LA2[ip];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          Go[e] = fma(c2.x, w[e].o[2 * ip], Go[e]);
          if (2 * ip + 1 < HO) Go[e] = fma(c2.y, w[e].o[2 * ip + 1 < HO ? 2 * ip + 1 : 0], Go[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        gam0[e] = (Ge[e] + Go[e]) * coef[(size_t)(1 * EPT + e) * cx.BD];
        gam1[e] = (Ge[e] - Go[e]) * coef[(size_t)(2 * EPT + e) * cx.BD];
      }
    }
    double* tA = tr + cx.par * cx.BD;
    double* tB = tA + 2 * cx.BD;
    tA[cx.tid] = gam0[0];
    tB[cx.tid] = gam1[EPT - 1];
    trace_arrive(cx, ka.p.warp_local);
    // volume part (needs no neighbour data): mu_e += DO^T wo, mu_o += DE^T we (row i of the
    // block times w[i], accumulated straight into mu); the rka scaling of w lives in the blocks
#pragma unroll
    for (int i = 0; i < HO; ++i) {
#pragma unroll
      for (int jp = 0; jp < (HE + 1) / 2; ++jp) {
        const double2 c2 = so.DO2[i * HP + jp];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          mu[e].e[2 * jp] = fma(c2.x, w[e].o[i], mu[e].e[2 * jp]);
          if (2 * jp + 1 < HE) {
            const int j1 = 2 * jp + 1 < HE ? 2 * jp + 1 : 0;
            mu[e].e[j1] = fma(c2.y, w[e].o[i], mu[e].e[j1]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < HE; ++i) {
#pragma unroll
      for (int jp = 0; jp < (HO + 1) / 2; ++jp) {
        const double2 c2 = so.DE2[i * HP + jp];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          mu[e].o[2 * jp] = fma(c2.x, w[e].e[i], mu[e].o[2 * jp]);
          if (2 * jp + 1 < HO) {
            const int j1 = 2 * jp + 1 < HO ? 2 * jp + 1 : 0;
            mu[e].o[j1] = fma(c2.y, w[e].e[i], mu[e].o[j1]);
          }
        }
      }
    }
    trace_wait(cx, ka.p.warp_local);
    double gam1L = tB[cx.nbL];  // right-face term of the left neighbour
    double gam0R = tA[cx.nbR];  // left-face term of the right neighbour
    cx.par ^= 1;
    if (!(cx.flags & CX_PERIODIC)) {
      if (cx.flags & CX_FIRST) gam1L = 0.0;
      if (cx.flags & CX_LAST) gam0R = 0.0;
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double a0 = gam0[e] - ((e == 0) ? gam1L : gam1[e - 1]);
      const double aN = gam1[e] - ((e == EPT - 1) ? gam0R : gam0[e + 1]);
      mu[e].e[0] = fma(0.5, a0 + aN, mu[e].e[0]);
      mu[e].o[0] = fma(0.5, a0 - aN, mu[e].o[0]);
    }
  }
}

// Deterministic per-trajectory sum of one value per thread: the thread that owns element 0
// adds the KT partials in index order (bit-stable across grid / batch sizes).
static __device__ __noinline__ double traj_sum(double* red, int tid, int KT, bool first, double v) {
  __syncthreads();
  red[tid] = v;
  __syncthreads();
  double s = 0.0;
  if (first) {
    for (int j = 0; j < KT; ++j) s += red[tid + j];
  }
  __syncthreads();
  return s;
}

// zf = P~ z : even and odd blocks of the prolongation
template <int NP, int EPT>
__device__ __forceinline__ void prolong_eo(const ProlongOps& po, const EOVec<NP> (&z)[EPT],
                                           EOVec<NP + 1> (&f)[EPT]) {
#pragma unroll
  for (int i = 0; i < EO<NP + 1>::HE; ++i) {
    double acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) acc[e] = po.PE[i * HM] * z[e].e[0];
#pragma unroll
    for (int j = 1; j < EO<NP>::HE; ++j) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) acc[e] = fma(po.PE[i * HM + j], z[e].e[j], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) f[e].e[i] = acc[e];
  }
#pragma unroll
  for (int i = 0; i < EO<NP + 1>::HO; ++i) {
    double acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) acc[e] = po.PO[i * HM] * z[e].o[0];
#pragma unroll
    for (int j = 1; j < EO<NP>::HO; ++j) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) acc[e] = fma(po.PO[i * HM + j], z[e].o[j], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) f[e].o[i] = acc[e];
  }
}

// ---------------------------------------------------------------------------------------
// The march kernel.  DO_FWD: forward phase (optionally writing fine-residual checkpoints,
// RESID); DO_ADJ: adjoint phase + indicator.  Fused = both.  EPT = elements per thread.
// A thread owns elements k0 .. k0+EPT-1 of trajectory t_local of the CTA's group;
// KT = K/EPT threads make one trajectory.
// Shared memory (doubles, after a 16-byte mbarrier header), BD = blockDim.x:
//   tr[4][BD]          trace exchange, double buffered {left[2], right[2]} (also reduction pad)
//   coef[2][3][EPT][BD] per-element {m, q0, q1} for the primal and the enriched level
//   big[NPF][EPT][BD]  forward: parked coarse state / sigma;  adjoint: TMA landing tile
// Park / checkpoint row order of an even/odd state: e[0..HE), then o[0..HO).
// ---------------------------------------------------------------------------------------
// BDT > 0: blockDim.x is the compile-time constant BDT (all shared-memory strides fold into
// immediate offsets); BDT = 0: any block size.
template <int NP, int EPT, int BDT, bool DO_FWD, bool RESID, bool DO_ADJ>
__global__ void __launch_bounds__(MAXBD / EPT, 1) march_kernel(const __grid_constant__ KArgs ka) {
  constexpr int NPF = NP + 1;
  const MarchParams& p = ka.p;
  const ConstOps& c = ka.c;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int BD = BDT > 0 ? BDT : (int)blockDim.x;
  const int tid = threadIdx.x;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sm_tr = reinterpret_cast<double*>(smem_raw + 16);
  double* sm_coef = sm_tr + 4 * BD + tid;            // + ((lv*3 + c)*EPT + e)*BD
  double* sm_big = sm_tr + (4 + 6 * EPT) * BD;       // + (row*EPT + e)*BD + tid
  const size_t cstride = (size_t)EPT * BD;           // row stride of a park / tile column

  Ctx cx;
  cx.tid = tid;
  cx.BD = BD;
  cx.par = 0;
  const int K = p.K;
  const int KT = K / EPT;
  const int t_local = tid / KT;
  const int kt = tid - t_local * KT;
  const int k0 = kt * EPT;
  const bool in_tile = (t_local < p.tpc);
  cx.flags = (p.bc == BC_PERIODIC ? CX_PERIODIC : 0);
  cx.nbL = tid;
  cx.nbR = tid;
  if (in_tile) {  // padding threads keep self-neighbours and no boundary role
    if (kt == 0) cx.flags |= CX_FIRST;
    if (kt == KT - 1) cx.flags |= CX_LAST;
    cx.nbL = (kt == 0) ? tid + KT - 1 : tid - 1;
    cx.nbR = (kt == KT - 1) ? tid - (KT - 1) : tid + 1;
  }

  const size_t tile = (size_t)NPF * EPT * BD;  // doubles per checkpoint tile
  const uint32_t tile_bytes = (uint32_t)(tile * sizeof(double));
  uint32_t land_phase = 0u;  // completed phases of the landing buffer's mbarrier

  cx.trbar = &mbar[1];
  cx.trphase = 0u;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);                  // TMA landing tile
    mbar_init(&mbar[1], (uint32_t)(BD / 32));  // trace exchange: one arrival per warp
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

#pragma unroll 1
  for (int g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
    const long long b = (long long)g * p.tpc + t_local;
    const bool active = in_tile && (b < p.B);
    const long long bs = active ? b : 0;  // in-bounds index for per-trajectory parameter reads
    {
      // per-element stage coefficients -> shared memory (face coefficients
      // c_f = (a*nx - (1-alpha)|a*nx|)/2, nx = -1,+1; AdvecRHS1D.m:11; du(mapO) = 0, :16)
      const double a = (active && p.a_arr) ? p.a_arr[b] : p.a;
      const double dt = (active && p.dt_arr) ? p.dt_arr[b] : p.dt;
      const double sg = (a > 0.0) ? 1.0 : ((a < 0.0) ? -1.0 : 0.0);
      const double e0 = 0.5 * (-1.0 - (1.0 - p.alpha) * sg);  // c_0 / a
      const double e1 = 0.5 * (1.0 - (1.0 - p.alpha) * sg);   // c_1 / a
#pragma unroll
      for (int lv = 0; lv < 2; ++lv) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          double m = 0.0, q0 = 0.0, q1 = 0.0;
          if (in_tile && (lv == 0 || RESID || DO_ADJ)) {
            const int k = k0 + e;
            const bool outflow_face = !(cx.flags & CX_PERIODIC) && (k == K - 1);
            const double rx = p.rxk[lv][k];
            m = -a * rx * dt;
            q0 = -p.fs0[lv][k] * e0 / rx;
            q1 = outflow_face ? 0.0 : -p.fs1[lv][k] * e1 / rx;
          }
          sm_coef[(size_t)((lv * 3 + 0) * EPT + e) * BD] = m;
          sm_coef[(size_t)((lv * 3 + 1) * EPT + e) * BD] = q0;
          sm_coef[(size_t)((lv * 3 + 2) * EPT + e) * BD] = q1;
        }
      }
    }
    const size_t slot = p.ckpt_by_block ? (size_t)blockIdx.x : (size_t)g;
    double* ck = p.ckpt ? p.ckpt + slot * (size_t)p.S * tile : nullptr;
    const size_t gofs = (size_t)bs * NP * K + k0;  // this thread's column in [B][NP][K]

    double u[EPT][NP];  // nodal terminal state (adjoint terminal condition / J)
    // ------------------------------------------------------------------ forward phase
    if (DO_FWD) {
      EOVec<NP> z[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
#pragma unroll
        for (int i = 0; i < NP; ++i) u[e][i] = active ? p.u0[gofs + (size_t)i * K + e] : 0.0;
        z[e].from_nodal(u[e]);
      }
      double* hist = (p.hist && active) ? p.hist + (size_t)b * (p.S + 1) * NP * K + k0 : nullptr;
      if (hist) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NP; ++i) hist[(size_t)i * K + e] = u[e][i];
        }
      }
      double time = p.t0;
      double* park = sm_big + tid;  // element e, row i at park[(i*EPT + e)*BD]
      if (RESID) {  // P u^0 waits in the park for the first fine step
        EOVec<NPF> f[EPT];
        prolong_eo<NP, EPT>(c.pr[1], z, f);
#pragma unroll
        for (int e = 0; e < EPT; ++e) f[e].store(park + (size_t)e * BD, cstride);
      }
#pragma unroll 1
      for (int n = 0; n < p.S; ++n) {
        if (RESID) {
          // fine one-step image of the injected coarse state: sigma = Phi_f(P u^n).
          // park holds P u^n on entry; the coarse state takes its place during the fine step.
          EOVec<NPF> f[EPT], rf[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            f[e].load(park + (size_t)e * BD, cstride);
            rf[e].zero();
            z[e].store(park + (size_t)e * BD, cstride);
          }
          fwd_step<NPF, 1, EPT>(ka, cx, sm_tr, sm_coef + (size_t)3 * EPT * BD, f, rf, bs, time, n);
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            z[e].load(park + (size_t)e * BD, cstride);
            f[e].store(park + (size_t)e * BD, cstride);
          }
        }
        {
          // scaled RK residual; rka[0] = 0 (checked by the host) so it need not survive the
          // step boundary -- exactly `rk4a(1)*resu` of the mlx
          EOVec<NP> r[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) r[e].zero();
          fwd_step<NP, 0, EPT>(ka, cx, sm_tr, sm_coef, z, r, bs, time, n);
        }
        time += p.dt_arr ? p.dt_arr[bs] : p.dt;  // `time = time+dt` accumulation of the mlx
        if (RESID) {
          // rho^n = P u^{n+1} - sigma  -> checkpoint tile [n][row][e][tid]  (coalesced);
          // P u^{n+1} stays in the park as the start of the next fine step.
          double* dst = ck + (size_t)n * tile + tid;
          EOVec<NPF> f[EPT];
          prolong_eo<NP, EPT>(c.pr[n & 1], z, f);
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
#pragma unroll
            for (int i = 0; i < NPF; ++i) {  // row i of the e/o order
              const size_t o = (size_t)(i * EPT + e) * BD;
              const double v = (i < EO<NPF>::HE) ? f[e].e[i < EO<NPF>::HE ? i : 0]
                                                 : f[e].o[i >= EO<NPF>::HE ? i - EO<NPF>::HE : 0];
              dst[o] = v - park[o];
              park[o] = v;
            }
          }
        }
        if (hist) {
          double* hn = hist + (size_t)(n + 1) * NP * K;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            z[e].to_nodal(u[e], 0.5);
#pragma unroll
            for (int i = 0; i < NP; ++i) hn[(size_t)i * K + e] = u[e][i];
          }
        }
      }
#pragma unroll
      for (int e = 0; e < EPT; ++e) z[e].to_nodal(u[e], 0.5);
      if (p.uT && active) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NP; ++i) p.uT[gofs + (size_t)i * K + e] = u[e][i];
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
#pragma unroll
        for (int i = 0; i < NP; ++i) u[e][i] = active ? p.uT_in[gofs + (size_t)i * K + e] : 0.0;
      }
    }

    // ------------------------------------------------------------------ adjoint phase
    if (DO_ADJ) {
      // checkpoints were written through the generic proxy; TMA reads them through the
      // async proxy -> fence, then make every thread's writes visible to the issuing thread.
      if (DO_FWD) {
        __threadfence();
        fence_proxy_async();
      }
      __syncthreads();  // also: everyone is done with `park` (aliases the landing tile)
      double* land = sm_big;
      if (tid == 0 && p.S >= 1) {
        mbar_expect_tx(&mbar[0], tile_bytes);
        tma_bulk_g2s(land, ck + (size_t)(p.S - 1) * tile, tile_bytes, &mbar[0]);
      }
      // terminal condition lam^S = dJ_f/du at P u^S, and J of the coarse solution
      EOVec<NPF> mu[EPT];
      double jpart = 0.0;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int k = k0 + e;
        double lu[NPF];
        if (p.func == FUNC_LINEAR) {
#pragma unroll
          for (int i = 0; i < NPF; ++i) lu[i] = in_tile ? p.jw_f[(size_t)i * K + k] : 0.0;
#pragma unroll
          for (int i = 0; i < NP; ++i) jpart = fma(in_tile ? p.jw_c[(size_t)i * K + k] : 0.0, u[e][i], jpart);
        } else {
          const double jacC = in_tile ? 1.0 / p.rxk[0][k] : 0.0, jacF = in_tile ? 1.0 / p.rxk[1][k] : 0.0;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            double acc = c.Mref[0][i * NP] * u[e][0];
#pragma unroll
            for (int j = 1; j < NP; ++j) acc = fma(c.Mref[0][i * NP + j], u[e][j], acc);
            jpart = fma(u[e][i], jacC * acc, jpart);
          }
          double uf[NPF];
#pragma unroll
          for (int i = 0; i < NPF; ++i) {
            double acc = c.P[i * NP] * u[e][0];
#pragma unroll
            for (int j = 1; j < NP; ++j) acc = fma(c.P[i * NP + j], u[e][j], acc);
            uf[i] = acc;
          }
#pragma unroll
          for (int i = 0; i < NPF; ++i) {
            double acc = c.Mref[1][i * NPF] * uf[0];
#pragma unroll
            for (int j = 1; j < NPF; ++j) acc = fma(c.Mref[1][i * NPF + j], uf[j], acc);
            lu[i] = 2.0 * jacF * acc;
          }
        }
        // mu = T^-T lam : halves of the mirrored sums / differences
        mu[e].from_nodal(lu);
#pragma unroll
        for (int i = 0; i < NPF / 2; ++i) {
          mu[e].e[i] *= 0.5;
          mu[e].o[i] *= 0.5;
        }
      }
      const double Jtot = traj_sum(sm_tr, tid, KT, (cx.flags & CX_FIRST) != 0, jpart);
      if (p.J && active && (cx.flags & CX_FIRST)) p.J[b] = Jtot;

      EOVec<NPF> w[EPT];
      double eta[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        w[e].zero();
        eta[e] = 0.0;
      }
#pragma unroll 1
      for (int n = p.S - 1; n >= 0; --n) {
        // eta[k] += lam^{n+1} . rho^n ; the tile is consumed at once, so one landing buffer
        // suffices: the refill for step n-1 flies during the five stages of this step.
        mbar_wait(&mbar[0], land_phase & 1u);
        ++land_phase;
        const double* rho = land + tid;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < EO<NPF>::HE; ++i)
            eta[e] = fma(mu[e].e[i], rho[(size_t)(i * EPT + e) * BD], eta[e]);
#pragma unroll
          for (int i = 0; i < EO<NPF>::HO; ++i)
            eta[e] = fma(mu[e].o[i], rho[(size_t)((EO<NPF>::HE + i) * EPT + e) * BD], eta[e]);
        }
        __syncthreads();  // every thread has consumed the landing tile
        if (tid == 0 && n >= 1) {
          mbar_expect_tx(&mbar[0], tile_bytes);
          tma_bulk_g2s(land, ck + (size_t)(n - 1) * tile, tile_bytes, &mbar[0]);
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) w[e].zero();  // w~ restarts every step (rka[0] = 0)
        adj_step<NPF, 1, EPT>(ka, cx, sm_tr, sm_coef + (size_t)3 * EPT * BD, mu, w);
      }
      if (active) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          if (p.eta) p.eta[(size_t)b * K + k0 + e] = eta[e];
          if (p.lam0) {
            double lu[NPF];
            mu[e].to_nodal(lu, 1.0);  // lam = T^T mu
            double* l0 = p.lam0 + (size_t)b * NPF * K + k0 + e;
#pragma unroll
            for (int i = 0; i < NPF; ++i) l0[(size_t)i * K] = lu[i];
          }
        }
      }
    }
    __syncthreads();  // smem (coef / park / landing tile / traces) free before the next group
  }
}
#endif  // __CUDACC__ && DGADJ_DEVICE_CODE

// host-callable launcher exported by each per-order object (dgadj_march_np.cu, -DDGADJ_NP=n)
typedef cudaError_t (*march_launch_fn)(int variant, int ept, int grid, int block, cudaStream_t stream,
                                       const KArgs* ka);

}  // namespace dgadj
