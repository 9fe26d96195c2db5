// dgadj_kernels.cuh -- persistent sm_100a kernels for the batched 1-D nodal-DG march.
//
// Mapping: one thread owns EPT adjacent elements of one trajectory for the whole march; their
// Np modal (Legendre) coefficients and the RK residual live in registers.  A CTA owns `tpc`
// trajectories (tpc*K <= blockDim) and loops over trajectory groups (persistent grid).
// The modal operators and RK coefficients travel as a __grid_constant__ kernel parameter, i.e.
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

// Checkpoint tiles leave the forward phase either as coalesced STG (0, the default) or through bulk-TMA
// from a double-buffered park (1; see STORE_HOOK at fwd_step).  Compile-time, because even the untaken
// hook costs issue slots in the stage loop; the host enables p.tma_store only in a build that has it.
// MEASURED (B200, config 2, profiles/r2_march_fused_fullsize{,_tma_store_build}_ncu.json): the TMA build
// removes the lg_throttle stalls as intended (2.9 % -> 0.07 % of the samples) but is 5 % SLOWER overall
// (1038 ms vs 990 ms per launch; fp64 pipe 67.0 % vs 70.0 %): the residual now takes an STS per value (the
// MIO queue the stage loop already leans on) instead of an STG, the hooks add 1.6 % executed instructions
// to the stage loops, and the doubled park costs 4 registers of the 254.  The STG form stays the product.
#ifndef DGADJ_TMA_STORE_PATH
#define DGADJ_TMA_STORE_PATH 0
#endif

namespace dgadj {

constexpr int MAXNP = 10;      // enriched space of N = 8
constexpr int MAXSTAGES = 5;
constexpr int MAXBD = 1024;
// CTAs of MAXBD/EPT threads per SM the march kernels are compiled for: at N = 8 the state fills the
// register file (1); N <= 2 fits 128 registers without spilling and gains 5-8 % from twice the warps
// (measured, K = 64: N = 1 3.02 -> 3.27e11, N = 2 2.62 -> 2.77e11 updates/s; N = 4, 5 spill and lose 15-40 %)
__host__ __device__ constexpr int march_min_ctas(int NP, int EPT) { return (NP <= 3 && EPT == 4) ? 2 : 1; }

constexpr int HM = 5;  // max half dimension of the even/odd blocks ((MAXNP+1)/2)

// Even/odd blocks of a nodal operator set (used by the Burgers kernels, whose flux is nodal):
//   even-out = DE * odd-in  (+ LS * (g0+g1)),   odd-out = DO * even-in  (+ LA * (g0-g1))
// stored as double2 pairs: entry (i, j) is DE2[i*HP + j/2].{x,y}[j & 1].
constexpr int HP = 3;  // pairs per row ((HM+1)/2)
struct alignas(16) StageOps {
  double2 DE2[HM * HP];  // i < HE, j < HO
  double2 DO2[HM * HP];  // i < HO, j < HE
  double2 LS2[HP];
  double2 LA2[HP];
};

// The advection march runs in the orthonormal Legendre (modal) basis  u^ = V^-1 u :
//   * D^ = V^-1 Dr V is strictly upper triangular and couples only modes of opposite parity:
//     floor(Np^2/4) non-zeros instead of Np^2 (Np = 10: 25 instead of 100);
//   * V^-1 LIFT = V^T E: column f holds the basis values at the faces, p_i = P~_i(+1) and
//     P~_i(-1) = (-1)^i p_i -- the same Np numbers give the traces u(-1), u(+1) and the lift;
//   * the prolongation to order N+1 is the injection (u^, 0): no arithmetic;
//   * V^T M V = I: J = int u^2 is J_k |u^|^2.
// Non-zeros of D^ row by row: (i, j) with j = i+1, i+3, ... < Np.
constexpr int MAXNZ = 26;  // 25 for Np = 10, padded to a multiple of 2
__host__ __device__ constexpr int nz_row_offset(int NPX, int i) {
  int o = 0;
  for (int q = 0; q < i; ++q) o += (NPX - q) / 2;
  return o;
}
__host__ __device__ constexpr int nz_index(int NPX, int i, int j) { return nz_row_offset(NPX, i) + (j - i - 1) / 2; }
__host__ __device__ constexpr int nz_count(int NPX) { return nz_row_offset(NPX, NPX); }

struct alignas(16) ModalStage {
  double D[MAXNZ];
};
// Stage scaling.  The low-storage recurrences  r_s = rka_s r_{s-1} + R_s  (forward) and
// w_s = rka_{s+1} w_{s+1} + bm_s mu_s  (adjoint) are carried as r = sig_s r~, w = sga_s w~ with
// sig_0 = 1, sig_s = rka_s sig_{s-1} and sga_last = 1, sga_s = rka_{s+1} sga_{s+1}: the rka
// multiply disappears (r~_s = r~_{s-1} + R_s / sig_s); the factors 1/sig_s, sga_s are folded
// into per-stage copies of D^ -- copies that exist anyway, because indexing the constants by
// the stage counter is what keeps ptxas from hoisting (and then spilling) the loop-invariant
// constant loads out of the stage loop.
struct alignas(16) ConstOps {
  ModalStage st[2][MAXSTAGES];    // [level][stage]: D^ / sig_s            (forward sweeps)
  ModalStage sta[MAXSTAGES];      // enriched level: D^ * sga_s            (adjoint sweep)
  double p[2][MAXNP];             // [level] p_i = P~_i(+1)
  double isig[MAXSTAGES];         // 1 / sig_s
  double sga[MAXSTAGES];
  double bsig[MAXSTAGES];         // rkb_s * sig_s   (state update  z += bsig_s m r~)
  double bsga[MAXSTAGES];         // rkb_s / sga_s   (adjoint update w~ += bsga_s m mu)
  double V[MAXNP * MAXNP];        // primal Vandermonde, row stride Np   (u = V u^ : outputs)
  double iV[MAXNP * MAXNP];       // its inverse                          (u^ = V^-1 u : inputs)
  double iVf[MAXNP * MAXNP];      // inverse of the enriched Vandermonde  (lam0 = V_f^-T mu)
  double rka[MAXSTAGES], rkb[MAXSTAGES], rkc[MAXSTAGES];
};

struct MarchParams {
  long long B;
  int K, S, tpc, ngroups, nstages, bc, inflow, func;
  int warp_local;         // K/EPT divides 32: trajectories never straddle warps
  double alpha, a, dt, t0;
  const double* a_arr;
  const double* dt_arr;
  const double* rxk[2];   // [level][K]   per-element rx (mean of rx(:,k))
  const double* fs0[2];   // [level][K]   Fscale(1,k)
  const double* fs1[2];   // [level][K]   Fscale(2,k)
  const double* jw_c;     // [NP][K]   modal weights V^T jw of the linear functional
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
  // windowed marches (dgadj_fwd_adj_windowed): a window's adjoint sweep starts from the modal adjoint
  // state the later window left (mu_in, [B][NPF][K]) instead of the functional's terminal condition,
  // leaves its own (mu_out), continues the indicator sums already in `eta` (eta_acc), and counts
  // its steps from n0 (time-dependent inflow data)
  const double* mu_in;
  double* mu_out;
  int eta_acc;
  int n0;
  int in_modal, out_modal;   // the input state / uT are modal coefficients (window hand-over: no V^-1 V round trip)
  const int* npk;            // hp (HP kernels only): [K] modes of each element's primal space, <= NP; the enriched
                             // space has one more.  Modes beyond an element's count are held at zero (padded hp:
                             // in the orthonormal modal basis a lower order IS the truncated space)
  int tma_store;             // forward checkpoints leave through bulk-TMA from a double-buffered park (host: fits, not warp_local)
  int vec_io;                // every [B][Np][K] pointer is 16-byte aligned and EPT is even: 128-bit state I/O
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
// nbuf = 2: the park / residual tile is double buffered (bulk-TMA checkpoint stores, see march_kernel)
__host__ __device__ constexpr size_t march_smem_bytes(int NP, int EPT, int BD, int variant, int nbuf = 1) {
  const int nbig = (variant == VAR_FWD) ? 0 : (NP + 1) * nbuf;
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
// shared -> global bulk copy (SASS UBLKCP.S.G... the store direction), tracked by bulk async-groups
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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
  double* ck;         // this CTA's checkpoint slot (tiles [S][NPF][EPT][BD])
  double* big;        // park / residual tile area in shared memory
};

#ifndef DGADJ_SPLIT_BARRIER
#define DGADJ_SPLIT_BARRIER 1
#endif
// Trace exchange synchronisation.  Split form: a warp *arrives* as soon as its traces are in
// shared memory and *waits* only when it needs its neighbours' -- the volume terms (most of a
// stage) sit in between, so warps rarely block.  DGADJ_SPLIT_BARRIER=0 keeps a plain
// __syncthreads() at the arrive point (for A/B measurements).
// Measured alternatives (B200, N=8, K=1024): with the exchange synchronisation removed
// altogether (wrong results) the kernel is 12 % faster -- the ceiling for any scheme; point-to-point
// variants (per-warp flags with a software spin; per-warp mbarrier pairs awaited only by the
// edge lanes) were 30-50 % SLOWER: divergent waits and the extra live registers (spills at 254
// registers/thread) cost more than the looser coupling gains.  Folding 1/sig_s and sga_s into
// per-stage copies of p (three DMUL fewer per update) was 1.2 % slower: ten more constant loads
// per stage and level outweigh the multiplies.  Splitting the trace sums into two partial sums per
// parity (shorter dependent DFMA chains, two more DADD per element-stage) was 1.2 % slower too: the
// chains are not what the pipe waits for.  Rotating the stage loop so that the outer elements' surface
// update and traces come first and the arrive follows them at once (inner elements and the volume
// terms after it: half the stretch between a warp's wait and its next arrive, same operations) was
// 4 % slower: two elements at a time halve the independent chains in exactly those phases; moving one
// diagonal of the volume terms in front of the trace sums and the last ones behind the exchange (so
// that the scalar chains share a basic block with independent DFMA) was 1 % slower.  Not built: a uniform-mesh variant with {m, q0, q1} as
// kernel constants instead of shared-memory columns (-3 LDS, -1 DMUL per element-stage) -- the
// reference's rx / Fscale of a "uniform" mesh differ between elements by 1e-12..1e-11 (cancellation
// noise of J = Dr*x), so one constant for all elements would leave the 1e-12 parity envelope.
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
    case INFLOW_TABLE: return p.uin_table[(n + p.n0) * p.nstages + s];
    default: return 0.0;
  }
}

// modal coefficients of one element; rows <-> a strided column (park / landing / checkpoint tile)
template <int NPX>
struct MVec {
  double v[NPX];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NPX; ++i) v[i] = 0.0;
  }
  __device__ __forceinline__ void load(const double* col, size_t stride) {
#pragma unroll
    for (int i = 0; i < NPX; ++i) v[i] = col[(size_t)i * stride];
  }
  __device__ __forceinline__ void store(double* col, size_t stride) const {
#pragma unroll
    for (int i = 0; i < NPX; ++i) col[(size_t)i * stride] = v[i];
  }
};

// traces of a modal state: u(-1) = Se - So, u(+1) = Se + So, S{e,o} = sum over even / odd modes
template <int NPX>
__device__ __forceinline__ void traces(const double (&pv)[MAXNP], const MVec<NPX>& z, double& uF, double& uB) {
  double Se = pv[0] * z.v[0], So = (NPX > 1) ? pv[1] * z.v[1] : 0.0;
#pragma unroll
  for (int i = 2; i < NPX; i += 2) Se = fma(pv[i], z.v[i], Se);
#pragma unroll
  for (int i = 3; i < NPX; i += 2) So = fma(pv[i], z.v[i], So);
  uF = Se - So;
  uB = Se + So;
}

// One RK step (all stages) in the scaled modal form.  Reference update (utils/AdvecRHS1D.m:11,19
// + the LSERK4 loop of utils/One_code.mlx), with m = -a*rx*dt per element and r = resu/m:
//     r^ = rka r^ + D^ u^ + p o {se | so} ;  u^ += (rkb m) r^
//     g0 = (u(-1) - uL) q0,  g1 = (u(+1) - uR) q1,  q_f = dt Fscale_f c_f / m,
//     se = g1 + g0 (even modes), so = g1 - g0 (odd modes)
// (the rka multiply is absorbed by the stage scaling described at ConstOps).
// coef = this thread's smem column of the level: {m, q0, q1} x EPT, each a row of BD.
// STORE_HOOK (bulk-TMA checkpoint stores, p.tma_store): the residual tile of step n-1 was completed in
// the park buffer (n-1)&1 by every thread before it arrived at this step's first trace barrier, so
//   LV = 1 (the fine step, first of a step): after the first wait, thread 0 hands that buffer to the TMA unit;
//   LV = 0 (the coarse step, second):        before its first arrive, thread 0 waits until the TMA unit has
//                                            read the buffer -- every thread that passes this barrier may
//                                            overwrite it (it does at the top of step n+1).
// No barrier is added: the store rides on the exchanges the stages need anyway.
// HP: nm[e] = number of modes of element e in this space; the surface term (the only one that feeds a mode from
// below) skips the modes beyond it, so they stay exactly zero.
template <int NPX, int LV, int EPT, bool STORE_HOOK = false, int TILE_ELEMS = 0, bool HP = false>
__device__ __forceinline__ void fwd_step(const KArgs& ka, Ctx& cx, double* __restrict__ tr,
                                         const double* __restrict__ coef, MVec<NPX> (&z)[EPT],
                                         MVec<NPX> (&r)[EPT], long long b, double time, int n,
                                         const int (&nm)[EPT]) {
  const ConstOps& c = ka.c;
  const int nst = ka.p.nstages;
#pragma unroll 1   // (fully unrolling the stages was measured 7 % slower: 5x the code, more spills)
  for (int s = 0; s < nst; ++s) {
    const ModalStage& so = c.st[LV][s];
    double uF[EPT], uB[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) traces<NPX>(c.p[LV], z[e], uF[e], uB[e]);
    double* tA = tr + cx.par * cx.BD;  // left-edge values  u(-1) of the thread's first element
    double* tB = tA + 2 * cx.BD;       // right-edge values u(+1) of the thread's last element
    tA[cx.tid] = uF[0];
    tB[cx.tid] = uB[EPT - 1];
    if (DGADJ_TMA_STORE_PATH && STORE_HOOK && LV == 0 && s == 0 && ka.p.tma_store && cx.tid == 0) tma_store_wait_read();
    trace_arrive(cx, ka.p.warp_local);
    // volume terms (no neighbour data): r^_i += sum_{j = i+1, i+3, ...} D^_ij u^_j, walked by
    // diagonals (j - i = 1, 3, ...) so that consecutive DFMA hit different accumulators: no
    // dependent chain is ever issued back to back (DFMA latency ~20 cycles, 2 warps/scheduler)
#pragma unroll
    for (int d = 1; d < NPX; d += 2) {
#pragma unroll
      for (int i = 0; i + d < NPX; ++i) {
        const double cij = so.D[nz_index(NPX, i, i + d)];
#pragma unroll
        for (int e = 0; e < EPT; ++e) r[e].v[i] = fma(cij, z[e].v[i + d], r[e].v[i]);
      }
    }
    trace_wait(cx, ka.p.warp_local);
    if (DGADJ_TMA_STORE_PATH && STORE_HOOK && LV == 1 && s == 0 && n > 0 && ka.p.tma_store && cx.tid == 0) {
      const size_t tile = (size_t)TILE_ELEMS * cx.BD;
      tma_bulk_s2g(cx.ck + (size_t)(n - 1) * tile, cx.big + (size_t)((n - 1) & 1) * tile, (uint32_t)(tile * sizeof(double)));
    }
    double uL = tB[cx.nbL];
    double uR = tA[cx.nbR];
    cx.par ^= 1;
    if (!(cx.flags & CX_PERIODIC)) {
      if (cx.flags & CX_FIRST) uL = inflow_value(ka.p, b, time, n, s, c.rkc[s]);
      if (cx.flags & CX_LAST) uR = uB[EPT - 1];
    }
    const double isig = c.isig[s], rkb = c.bsig[s];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double left = (e == 0) ? uL : uB[e - 1];
      const double right = (e == EPT - 1) ? uR : uF[e + 1];
      const double g0 = (uF[e] - left) * coef[(size_t)(1 * EPT + e) * cx.BD];
      const double g1 = (uB[e] - right) * coef[(size_t)(2 * EPT + e) * cx.BD];
      const double se = (g1 + g0) * isig, sd = (g1 - g0) * isig;
      const double bm = rkb * coef[(size_t)e * cx.BD];
#pragma unroll
      for (int i = 0; i < NPX; ++i) {
        if (HP && i >= nm[e]) continue;
        r[e].v[i] = fma(c.p[LV][i], (i & 1) ? sd : se, r[e].v[i]);
        z[e].v[i] = fma(bm, r[e].v[i], z[e].v[i]);
      }
    }
  }
}

// Reverse of one RK step: stages s = last..0  
//   lk += rkb*lu ; lu += dt*L^T lk ; lk *= rka,   dt*L^T lk = m*Dr^T lk + scatter(g),
//   g_f = dt*Fscale_f*c_f * (LIFT[:,f] . lk).
// Carried in the scaled modal form (mu = V^T lu, w = m V^T lk, w = sga_s w~):
//   w~ += (bsga_s m) mu ; Gse = sum_even p_i w_i, Gso = sum_odd p_i w_i ;
//   gam0 = q0 (Gse - Gso), gam1 = q1 (Gse + Gso) ; a0 = gam0 - gam1[left], aN = gam1 - gam0[right] ;
//   mu_j += sum_{i < j, i+j odd} D^_ij w_i + p_j {aN + a0 | aN - a0}.
template <int NPX, int LV, int EPT, bool HP = false>
__device__ __forceinline__ void adj_step(const KArgs& ka, Ctx& cx, double* __restrict__ tr,
                                         const double* __restrict__ coef, MVec<NPX> (&mu)[EPT],
                                         MVec<NPX> (&w)[EPT], const int (&nm)[EPT]) {
  const ConstOps& c = ka.c;
#pragma unroll 1
  for (int s = ka.p.nstages - 1; s >= 0; --s) {
    const ModalStage& so = c.sta[s];
    const double rkb = c.bsga[s], sga = c.sga[s];
    double gam0[EPT], gam1[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double bm = rkb * coef[(size_t)e * cx.BD];
#pragma unroll
      for (int i = 0; i < NPX; ++i) w[e].v[i] = fma(bm, mu[e].v[i], w[e].v[i]);
      double Gd, Gs;  // (Gse - Gso), (Gse + Gso)
      traces<NPX>(c.p[LV], w[e], Gd, Gs);
      gam0[e] = Gd * (sga * coef[(size_t)(1 * EPT + e) * cx.BD]);
      gam1[e] = Gs * (sga * coef[(size_t)(2 * EPT + e) * cx.BD]);
    }
    double* tA = tr + cx.par * cx.BD;
    double* tB = tA + 2 * cx.BD;
    tA[cx.tid] = gam0[0];
    tB[cx.tid] = gam1[EPT - 1];
    trace_arrive(cx, ka.p.warp_local);
    // volume part (no neighbour data): mu_j += D^_ij w_i, walked by diagonals (see fwd_step)
#pragma unroll
    for (int d = 1; d < NPX; d += 2) {
#pragma unroll
      for (int i = 0; i + d < NPX; ++i) {
        const double cij = so.D[nz_index(NPX, i, i + d)];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          if (HP && i + d >= nm[e]) continue;   // (the transposed derivative feeds HIGHER modes: held at zero beyond the element's space)
          mu[e].v[i + d] = fma(cij, w[e].v[i], mu[e].v[i + d]);
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
      const double ae = aN + a0, ao = aN - a0;
#pragma unroll
      for (int i = 0; i < NPX; ++i) {
        if (HP && i >= nm[e]) continue;
        mu[e].v[i] = fma(c.p[LV][i], (i & 1) ? ao : ae, mu[e].v[i]);
      }
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

// The thread's EPT adjacent elements of the rows of a [.][NPX][K] field: row i at base + i*K, elements
// contiguous.  vec (uniform; host-checked 16-byte alignment, EPT even): 128-bit accesses -- a warp then
// reads whole 128-byte lines instead of touching 32 sectors for 8 useful bytes each.
template <int NPX, int EPT>
__device__ __forceinline__ void load_rows(const double* base, size_t K, bool active, int vec, double (&u)[EPT][NPX]) {
  if (EPT % 2 == 0 && vec) {
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
#pragma unroll
      for (int e = 0; e < EPT; e += 2) {
        const double2 t = active ? *reinterpret_cast<const double2*>(base + (size_t)i * K + e) : make_double2(0.0, 0.0);
        u[e][i] = t.x;
        u[e + 1 < EPT ? e + 1 : e][i] = t.y;
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
#pragma unroll
      for (int i = 0; i < NPX; ++i) u[e][i] = active ? base[(size_t)i * K + e] : 0.0;
    }
  }
}
template <int NPX, int EPT>
__device__ __forceinline__ void store_rows(double* base, size_t K, int vec, const double (&u)[EPT][NPX]) {
  if (EPT % 2 == 0 && vec) {
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
#pragma unroll
      for (int e = 0; e < EPT; e += 2)
        *reinterpret_cast<double2*>(base + (size_t)i * K + e) = make_double2(u[e][i], u[e + 1 < EPT ? e + 1 : e][i]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
#pragma unroll
      for (int i = 0; i < NPX; ++i) base[(size_t)i * K + e] = u[e][i];
    }
  }
}

// dense Np x Np change of basis at the ends of a march (once per trajectory, not a hot loop)
template <int N1, bool TRANSPOSED>
__device__ __forceinline__ void apply_matrix(const double* M, const double (&x)[N1], double (&y)[N1]) {
#pragma unroll
  for (int i = 0; i < N1; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < N1; ++j) acc = fma(TRANSPOSED ? M[j * N1 + i] : M[i * N1 + j], x[j], acc);
    y[i] = acc;
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
// BDT > 0: blockDim.x is the compile-time constant BDT (all shared-memory strides fold into
// immediate offsets); BDT = 0: any block size.
// ---------------------------------------------------------------------------------------
template <int NP, int EPT, int BDT, bool DO_FWD, bool RESID, bool DO_ADJ, bool HP = false>
__global__ void __launch_bounds__(MAXBD / EPT, march_min_ctas(NP, EPT)) march_kernel(const __grid_constant__ KArgs ka) {
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
    if (DGADJ_TMA_STORE_PATH) {
      cx.ck = ck;
      cx.big = sm_big;
    }
    const size_t gofs = (size_t)bs * NP * K + k0;  // this thread's column in [B][NP][K]

    int nmc[EPT], nmf[EPT];   // hp: modes of the thread's elements in the primal / the enriched space
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      nmc[e] = (HP && in_tile) ? p.npk[k0 + e] : NP;
      nmf[e] = nmc[e] + 1;
    }
    MVec<NP> z[EPT];  // primal modal state (at the end of the forward phase: u^(T))
    // ------------------------------------------------------------------ forward phase
    {
      // nodal input (u0, or the terminal primal of an adjoint-only call) -> modal
      const double* src = DO_FWD ? p.u0 : p.uT_in;
      double un[EPT][NP];
      load_rows<NP, EPT>(src + gofs, (size_t)K, active, p.vec_io, un);
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        if (p.in_modal) {
#pragma unroll
          for (int i = 0; i < NP; ++i) z[e].v[i] = un[e][i];
        } else {
          apply_matrix<NP, false>(c.iV, un[e], z[e].v);
        }
        if (HP) {   // the element's own space: the L2 projection of the input onto it
#pragma unroll
          for (int i = 0; i < NP; ++i) z[e].v[i] = (i < nmc[e]) ? z[e].v[i] : 0.0;
        }
      }
    }
    if (DO_FWD) {
      double* hist = (p.hist && active) ? p.hist + (size_t)b * (p.S + 1) * NP * K + k0 : nullptr;
      if (hist) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NP; ++i) hist[(size_t)i * K + e] = p.u0[gofs + (size_t)i * K + e];
        }
      }
      double time = p.t0 + (p.n0 ? p.n0 * (p.dt_arr ? p.dt_arr[bs] : p.dt) : 0.0);
#pragma unroll 1
      for (int n = 0; n < p.S; ++n) {
        // element e, row i at park[(i*EPT + e)*BD]; with bulk-TMA checkpoint stores the park alternates
        // between two buffers: step n's residual tile leaves from buffer n&1 while step n+1 fills the other
        double* park = sm_big + tid + ((DGADJ_TMA_STORE_PATH && RESID && p.tma_store && (n & 1)) ? tile : (size_t)0);
        if (RESID) {
          // fine one-step image of the injected coarse state: sigma = Phi_f(P u^n); in the modal
          // basis P u^n is (u^, 0).  The coarse state waits in the park meanwhile.
          MVec<NPF> f[EPT], rf[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
#pragma unroll
            for (int i = 0; i < NP; ++i) f[e].v[i] = z[e].v[i];
            f[e].v[NP] = 0.0;
            rf[e].zero();
            z[e].store(park + (size_t)e * BD, cstride);
          }
          fwd_step<NPF, 1, EPT, true, NPF * EPT, HP>(ka, cx, sm_tr, sm_coef + (size_t)3 * EPT * BD, f, rf, bs, time, n, nmf);
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            z[e].load(park + (size_t)e * BD, cstride);
            f[e].store(park + (size_t)e * BD, cstride);
          }
        }
        {
          // scaled RK residual; rka[0] = 0 (checked by the host) so it need not survive the
          // step boundary -- exactly `rk4a(1)*resu` of the mlx
          MVec<NP> r[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) r[e].zero();
          fwd_step<NP, 0, EPT, RESID, NPF * EPT, HP>(ka, cx, sm_tr, sm_coef, z, r, bs, time, n, nmc);
        }
        time += p.dt_arr ? p.dt_arr[bs] : p.dt;  // `time = time+dt` accumulation of the mlx
        if (RESID) {
          // rho^n = P u^{n+1} - sigma  -> checkpoint tile [n][row][e][tid]
          if (DGADJ_TMA_STORE_PATH && p.tma_store) {
            // completed in place in the park (already the tile's layout); the TMA unit takes it to the ring
            // after the next trace barrier (STORE_HOOK in fwd_step) -- no store instruction per value
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
#pragma unroll
              for (int i = 0; i < NPF; ++i) {
                const size_t o = (size_t)(i * EPT + e) * BD;
                park[o] = ((i < NP) ? z[e].v[i < NP ? i : 0] : 0.0) - park[o];
              }
            }
            fence_proxy_async();   // generic-proxy writes -> visible to the async proxy (the TMA read)
          } else {
            double* dst = ck + (size_t)n * tile + tid;   // coalesced STG
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
#pragma unroll
              for (int i = 0; i < NPF; ++i) {
                const size_t o = (size_t)(i * EPT + e) * BD;
                dst[o] = ((i < NP) ? z[e].v[i < NP ? i : 0] : 0.0) - park[o];
              }
            }
          }
        }
        if (hist) {
          double* hn = hist + (size_t)(n + 1) * NP * K;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            double u[NP];
            apply_matrix<NP, false>(c.V, z[e].v, u);
#pragma unroll
            for (int i = 0; i < NP; ++i) hn[(size_t)i * K + e] = u[i];
          }
        }
      }
      if (DGADJ_TMA_STORE_PATH && RESID && p.tma_store && p.S >= 1) {
        // the last residual tile: no later trace barrier to ride on
        __syncthreads();
        if (tid == 0) {
          tma_bulk_s2g(ck + (size_t)(p.S - 1) * tile, sm_big + (size_t)((p.S - 1) & 1) * tile, tile_bytes);
          tma_store_wait_all();   // all tiles of this trajectory are in global memory (the adjoint phase /
        }                         // the next group's park writes follow a __syncthreads)
      }
      if (p.uT && active) {
        double un[EPT][NP];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          if (p.out_modal) {
#pragma unroll
            for (int i = 0; i < NP; ++i) un[e][i] = z[e].v[i];
          } else {
            apply_matrix<NP, false>(c.V, z[e].v, un[e]);
          }
        }
        store_rows<NP, EPT>(p.uT + gofs, (size_t)K, p.vec_io, un);
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
      // terminal condition mu^S = V_f^T dJ_f/du at P u^S, and J of the coarse solution
      MVec<NPF> mu[EPT];
      double jpart = 0.0;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int k = k0 + e;
        if (p.func == FUNC_LINEAR) {
#pragma unroll
          for (int i = 0; i < NPF; ++i) mu[e].v[i] = in_tile ? p.jw_f[(size_t)i * K + k] : 0.0;
#pragma unroll
          for (int i = 0; i < NP; ++i) jpart = fma(in_tile ? p.jw_c[(size_t)i * K + k] : 0.0, z[e].v[i], jpart);
        } else {
          // J = int u^2 dx = sum_k J_k |u^_k|^2 in the orthonormal basis; dJ_f/du^ = 2 J_k (u^, 0)
          const double jacC = in_tile ? 1.0 / p.rxk[0][k] : 0.0, jacF = in_tile ? 1.0 / p.rxk[1][k] : 0.0;
          double n2 = 0.0;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            n2 = fma(z[e].v[i], z[e].v[i], n2);
            mu[e].v[i] = 2.0 * jacF * z[e].v[i];
          }
          mu[e].v[NP] = 0.0;
          jpart = fma(jacC, n2, jpart);
        }
      }
      if (HP) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NPF; ++i) mu[e].v[i] = (i < nmf[e]) ? mu[e].v[i] : 0.0;
        }
      }
      const double Jtot = traj_sum(sm_tr, tid, KT, (cx.flags & CX_FIRST) != 0, jpart);
      if (p.J && active && (cx.flags & CX_FIRST)) p.J[b] = Jtot;
      if (p.mu_in && active) {   // a window of a longer march: continue the later window's adjoint
#pragma unroll
        for (int e = 0; e < EPT; ++e) mu[e].load(p.mu_in + (size_t)b * NPF * K + k0 + e, (size_t)K);
      }

      MVec<NPF> w[EPT];
      double eta[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) eta[e] = (p.eta_acc && active) ? p.eta[(size_t)b * K + k0 + e] : 0.0;
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
          for (int i = 0; i < NPF; ++i) eta[e] = fma(mu[e].v[i], rho[(size_t)(i * EPT + e) * BD], eta[e]);
        }
        __syncthreads();  // every thread has consumed the landing tile
        if (tid == 0 && n >= 1) {
          mbar_expect_tx(&mbar[0], tile_bytes);
          tma_bulk_g2s(land, ck + (size_t)(n - 1) * tile, tile_bytes, &mbar[0]);
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) w[e].zero();  // w~ restarts every step (rka[0] = 0)
        adj_step<NPF, 1, EPT, HP>(ka, cx, sm_tr, sm_coef + (size_t)3 * EPT * BD, mu, w, nmf);
      }
      if (active) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          if (p.eta) p.eta[(size_t)b * K + k0 + e] = eta[e];
          if (p.mu_out) mu[e].store(p.mu_out + (size_t)b * NPF * K + k0 + e, (size_t)K);
        }
        if (p.lam0) {
          double lu[EPT][NPF];
#pragma unroll
          for (int e = 0; e < EPT; ++e) apply_matrix<NPF, true>(c.iVf, mu[e].v, lu[e]);  // lam = V_f^-T mu
          store_rows<NPF, EPT>(p.lam0 + (size_t)b * NPF * K + k0, (size_t)K, p.vec_io, lu);
        }
      }
    }
    __syncthreads();  // smem (coef / park / landing tile / traces) free before the next group
  }
}
#endif  // __CUDACC__ && DGADJ_DEVICE_CODE

// host-callable launcher exported by each per-order object (dgadj_march_np.cu, -DDGADJ_NP=n)
typedef cudaError_t (*march_launch_fn)(int variant, int ept, int grid, int block, size_t smem, cudaStream_t stream,
                                       const KArgs* ka);

}  // namespace dgadj
