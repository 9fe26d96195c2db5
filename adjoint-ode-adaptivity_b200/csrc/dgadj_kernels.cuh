// dgadj_kernels.cuh -- persistent sm_100a kernels for the batched 1-D nodal-DG march.
//
// Mapping: one thread owns one element of one trajectory for the whole march; its Np nodal
// values, the RK residual and its metric terms live in registers.  A CTA owns `tpc`
// trajectories (tpc*K <= blockDim) and loops over trajectory groups (persistent grid).
// Dr / LIFT / P / Mref / RK coefficients sit in __constant__ memory and are read as
// DFMA constant-bank operands (loops fully unrolled on the template order).  Neighbour
// traces cross threads through a double-buffered shared-memory pair (one barrier per stage).
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

struct ConstOps {
  double Dr[2][MAXNP * MAXNP];    // [level][i*NPX + j]; level 0 = primal (NP), 1 = enriched (NP+1)
  double LIFT[2][MAXNP * 2];      // [level][i*2 + f]
  double Mref[2][MAXNP * MAXNP];  // reference mass matrices inv(V V')
  double P[MAXNP * MAXNP];        // prolongation [NPF][NP]
  double rka[MAXSTAGES], rkb[MAXSTAGES], rkc[MAXSTAGES];
};
__constant__ ConstOps c;

struct MarchParams {
  long long B;
  int K, S, tpc, ngroups, nstages, bc, inflow, func;
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
  double* uT;             // forward: written; adjoint-only: read via uT_in
  const double* uT_in;
  double* hist;           // [B][S+1][NP][K] or null
  double* ckpt;           // tiles [slot][S][NPF][BD]
  int ckpt_by_block;      // 1: slot = blockIdx.x (fused ring); 0: slot = group
  double* J;
  double* lam0;
  double* eta;
};

enum { BC_INFLOW = 0, BC_PERIODIC = 1 };
enum { INFLOW_ZERO = 0, INFLOW_SIN_AT = 1, INFLOW_SIN_AAT = 2, INFLOW_TABLE = 3 };
enum { FUNC_LINEAR = 0, FUNC_INT_U2 = 1 };

// ---------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + bulk TMA (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
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

// ---------------------------------------------------------------------------------------
// per-thread context
// ---------------------------------------------------------------------------------------
struct Ctx {
  int tid, BD, k, K, nbL, nbR;
  int nstages, inflow;
  bool is_first, is_last, periodic;  // element position inside its trajectory
  double* trA;                       // smem [2][BD]  left-edge values  (u[0]   / g0)
  double* trB;                       // smem [2][BD]  right-edge values (u[Np-1]/ g1)
  int par;                           // trace double-buffer parity
  const double* uin_table;
};

__device__ __forceinline__ double inflow_value(const Ctx& cx, double a, double t, int n, int s) {
  switch (cx.inflow) {
    case INFLOW_SIN_AT: return -sin(a * t);
    case INFLOW_SIN_AAT: return -sin(a * a * t);
    case INFLOW_TABLE: return cx.uin_table[n * cx.nstages + s];
    default: return 0.0;
  }
}

// One RK stage of  resu = rka*resu + dt*rhs(u);  u += rkb*resu  for one element, with
//   dt*rhs[i] = m * (Dr u)[i] + LIFT[i][0]*g0 + LIFT[i][1]*g1,   m = -a*rx*dt,
//   g0 = (u[0]-uL)*f0, g1 = (u[Np-1]-uR)*f1,  f = dt*Fscale*c   (utils/AdvecRHS1D.m:11,19).
template <int NPX, int LV>
__device__ __forceinline__ void fwd_stage(double (&u)[NPX], double (&res)[NPX], double uL, double uR,
                                          double m, double f0, double f1, double rka, double rkb) {
  const double g0 = (u[0] - uL) * f0;
  const double g1 = (u[NPX - 1] - uR) * f1;
#pragma unroll
  for (int i = 0; i < NPX; ++i) {
    double acc = c.Dr[LV][i * NPX] * u[0];
#pragma unroll
    for (int j = 1; j < NPX; ++j) acc = fma(c.Dr[LV][i * NPX + j], u[j], acc);
    double sf = c.LIFT[LV][i * 2] * g0;
    sf = fma(c.LIFT[LV][i * 2 + 1], g1, sf);
    res[i] = fma(rka, res[i], fma(m, acc, sf));
  }
#pragma unroll
  for (int i = 0; i < NPX; ++i) u[i] = fma(rkb, res[i], u[i]);
}

// One full RK step (all stages) with the neighbour-trace exchange.  One barrier per stage.
template <int NPX, int LV>
__device__ __forceinline__ void fwd_step(Ctx& cx, double (&u)[NPX], double (&res)[NPX], double m,
                                         double f0, double f1, double a, double time, double dt, int n) {
  for (int s = 0; s < cx.nstages; ++s) {
    double* tA = cx.trA + cx.par * cx.BD;
    double* tB = cx.trB + cx.par * cx.BD;
    tA[cx.tid] = u[0];
    tB[cx.tid] = u[NPX - 1];
    __syncthreads();
    double uL = tB[cx.nbL];
    double uR = tA[cx.nbR];
    cx.par ^= 1;
    if (!cx.periodic) {
      if (cx.is_first) uL = inflow_value(cx, a, time + c.rkc[s] * dt, n, s);
      if (cx.is_last) uR = u[NPX - 1];
    }
    fwd_stage<NPX, LV>(u, res, uL, uR, m, f0, f1, c.rka[s], c.rkb[s]);
  }
}

// Reverse of one RK step: stages s = last..0
//   lk += rkb*lu ; lu += dt*L^T lk ; lk *= rka       
//   dt*L^T lk = m*Dr^T lk + scatter(g),  g_f = f_f * (LIFT[:,f] . lk)
template <int NPX, int LV>
__device__ __forceinline__ void adj_step(Ctx& cx, double (&lu)[NPX], double (&lk)[NPX], double m,
                                         double f0, double f1) {
  for (int s = cx.nstages - 1; s >= 0; --s) {
    const double rka = c.rka[s], rkb = c.rkb[s];
#pragma unroll
    for (int i = 0; i < NPX; ++i) lk[i] = fma(rkb, lu[i], lk[i]);
    double g0 = c.LIFT[LV][0] * lk[0], g1 = c.LIFT[LV][1] * lk[0];
#pragma unroll
    for (int i = 1; i < NPX; ++i) {
      g0 = fma(c.LIFT[LV][i * 2], lk[i], g0);
      g1 = fma(c.LIFT[LV][i * 2 + 1], lk[i], g1);
    }
    g0 *= f0;
    g1 *= f1;
    double* tA = cx.trA + cx.par * cx.BD;
    double* tB = cx.trB + cx.par * cx.BD;
    tA[cx.tid] = g0;
    tB[cx.tid] = g1;
    __syncthreads();
    double g1L = tB[cx.nbL];  // right-face term of the left neighbour
    double g0R = tA[cx.nbR];  // left-face term of the right neighbour
    cx.par ^= 1;
    if (!cx.periodic) {
      if (cx.is_first) g1L = 0.0;
      if (cx.is_last) g0R = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
      double acc = c.Dr[LV][i] * lk[0];
#pragma unroll
      for (int j = 1; j < NPX; ++j) acc = fma(c.Dr[LV][j * NPX + i], lk[j], acc);
      lu[i] = fma(m, acc, lu[i]);
    }
    lu[0] += g0 - g1L;
    lu[NPX - 1] += g1 - g0R;
#pragma unroll
    for (int i = 0; i < NPX; ++i) lk[i] *= rka;
  }
}

// Deterministic per-trajectory sum of one value per element: element 0's thread adds the K
// partials in index order (bit-stable across grid / batch sizes).
__device__ __forceinline__ double traj_sum(Ctx& cx, double* red, double v) {
  __syncthreads();
  red[cx.tid] = v;
  __syncthreads();
  double s = 0.0;
  if (cx.is_first) {
    for (int j = 0; j < cx.K; ++j) s += red[cx.tid + j];
  }
  return s;
}

struct Smem {
  double* trA;
  double* trB;
  double* red;
  double* big;     // park [NPF][BD]  (forward)  /  land [2][NPF][BD] (adjoint)
  uint64_t* mbar;  // [2]
};

template <int NP>
__device__ __forceinline__ Smem carve_smem(unsigned char* base, int BD) {
  Smem s;
  s.mbar = reinterpret_cast<uint64_t*>(base);
  double* d = reinterpret_cast<double*>(base + 16);
  s.trA = d;
  s.trB = d + 2 * BD;
  s.red = d + 4 * BD;
  s.big = d + 5 * BD;
  return s;
}
template <int NP>
__host__ __device__ constexpr size_t smem_bytes(int BD, bool fwd_resid, bool adj) {
  size_t big = 0;
  if (fwd_resid) big = (size_t)(NP + 1) * BD;
  if (adj) big = (size_t)2 * (NP + 1) * BD;
  return 16 + sizeof(double) * (5 * (size_t)BD + big);
}

// ---------------------------------------------------------------------------------------
// The march kernel.  DO_FWD: forward phase (optionally writing fine-residual checkpoints,
// RESID); DO_ADJ: adjoint phase + indicator.  Fused = both.
// ---------------------------------------------------------------------------------------
template <int NP, bool DO_FWD, bool RESID, bool DO_ADJ, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) march_kernel(const MarchParams p) {
  constexpr int NPF = NP + 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int BD = blockDim.x;
  const int tid = threadIdx.x;
  Smem sm = carve_smem<NP>(smem_raw, BD);

  Ctx cx;
  cx.tid = tid;
  cx.BD = BD;
  cx.K = p.K;
  const int t_local = tid / p.K;
  cx.k = tid - t_local * p.K;
  cx.is_first = (cx.k == 0);
  cx.is_last = (cx.k == p.K - 1);
  cx.periodic = (p.bc == BC_PERIODIC);
  cx.nbL = cx.is_first ? tid + p.K - 1 : tid - 1;
  cx.nbR = cx.is_last ? tid - (p.K - 1) : tid + 1;
  const bool in_tile = (t_local < p.tpc);
  if (!in_tile) {  // padding threads: keep smem indices valid
    cx.nbL = tid;
    cx.nbR = tid;
    cx.is_first = false;
    cx.is_last = false;
  }
  cx.nstages = p.nstages;
  cx.inflow = p.inflow;
  cx.trA = sm.trA;
  cx.trB = sm.trB;
  cx.par = 0;
  cx.uin_table = p.uin_table;

  const size_t tile = (size_t)NPF * BD;  // doubles per checkpoint tile
  uint32_t land_uses[2] = {0u, 0u};      // completed phases of each landing buffer's mbarrier

  if (DO_ADJ) {
    if (tid == 0) {
      mbar_init(&sm.mbar[0], 1);
      mbar_init(&sm.mbar[1], 1);
    }
    fence_proxy_async();
    __syncthreads();
  }

  // metric terms of this thread's element (shared mesh across the batch)
  double rxk[2] = {0, 0}, fs0[2] = {0, 0}, fs1[2] = {0, 0};
  if (in_tile) {
#pragma unroll
    for (int lv = 0; lv < 2; ++lv) {
      if (lv == 1 && !(RESID || DO_ADJ)) break;
      rxk[lv] = p.rxk[lv][cx.k];
      fs0[lv] = p.fs0[lv][cx.k];
      fs1[lv] = p.fs1[lv][cx.k];
    }
  }

  for (int g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
    const long long b = (long long)g * p.tpc + t_local;
    const bool active = in_tile && (b < p.B);
    const double a = (active && p.a_arr) ? p.a_arr[b] : p.a;
    const double dt = (active && p.dt_arr) ? p.dt_arr[b] : p.dt;
    // face coefficients c = (a*nx - (1-alpha)|a*nx|)/2, nx = -1,+1   (AdvecRHS1D.m:11)
    const double c0 = (-a - (1.0 - p.alpha) * fabs(a)) * 0.5;
    const double c1 = (a - (1.0 - p.alpha) * fabs(a)) * 0.5;
    const bool outflow_face = (!cx.periodic) && cx.is_last;  // du(mapO) = 0, AdvecRHS1D.m:16
    const double mC = -a * rxk[0] * dt, f0C = dt * fs0[0] * c0, f1C = outflow_face ? 0.0 : dt * fs1[0] * c1;
    const double mF = -a * rxk[1] * dt, f0F = dt * fs0[1] * c0, f1F = outflow_face ? 0.0 : dt * fs1[1] * c1;
    const size_t slot = p.ckpt_by_block ? (size_t)blockIdx.x : (size_t)g;
    double* ck = p.ckpt ? p.ckpt + slot * (size_t)p.S * tile : nullptr;

    double u[NP];
    // ------------------------------------------------------------------ forward phase
    if (DO_FWD) {
      const double* u0 = p.u0 + (size_t)b * NP * p.K + cx.k;
#pragma unroll
      for (int i = 0; i < NP; ++i) u[i] = active ? u0[(size_t)i * p.K] : 0.0;
      double res[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) res[i] = 0.0;
      double* hist = (p.hist && active) ? p.hist + (size_t)b * (p.S + 1) * NP * p.K + cx.k : nullptr;
      if (hist) {
#pragma unroll
        for (int i = 0; i < NP; ++i) hist[(size_t)i * p.K] = u[i];
      }
      double time = p.t0;
      double* park = sm.big + tid;  // thread-private column park[i*BD]
      for (int n = 0; n < p.S; ++n) {
        if (RESID) {
          // fine one-step image of the injected coarse state: sigma = Phi_f(P u^n)
          double uf[NPF], rf[NPF];
#pragma unroll
          for (int i = 0; i < NPF; ++i) {
            double acc = c.P[i * NP] * u[0];
#pragma unroll
            for (int j = 1; j < NP; ++j) acc = fma(c.P[i * NP + j], u[j], acc);
            uf[i] = acc;
            rf[i] = 0.0;
          }
#pragma unroll
          for (int i = 0; i < NP; ++i) park[(size_t)i * BD] = u[i];
          fwd_step<NPF, 1>(cx, uf, rf, mF, f0F, f1F, a, time, dt, n);
#pragma unroll
          for (int i = 0; i < NP; ++i) u[i] = park[(size_t)i * BD];
#pragma unroll
          for (int i = 0; i < NPF; ++i) park[(size_t)i * BD] = uf[i];
        }
        fwd_step<NP, 0>(cx, u, res, mC, f0C, f1C, a, time, dt, n);
        time += dt;  // `time = time+dt` accumulation of the mlx
        if (RESID) {
          // rho^n = P u^{n+1} - sigma  -> checkpoint tile [n][i][tid]  (coalesced)
          double* dst = ck + (size_t)n * tile + tid;
#pragma unroll
          for (int i = 0; i < NPF; ++i) {
            double acc = -park[(size_t)i * BD];
#pragma unroll
            for (int j = 0; j < NP; ++j) acc = fma(c.P[i * NP + j], u[j], acc);
            dst[(size_t)i * BD] = acc;
          }
        }
        if (hist) {
          double* hn = hist + (size_t)(n + 1) * NP * p.K;
#pragma unroll
          for (int i = 0; i < NP; ++i) hn[(size_t)i * p.K] = u[i];
        }
      }
      if (p.uT && active) {
        double* uT = p.uT + (size_t)b * NP * p.K + cx.k;
#pragma unroll
        for (int i = 0; i < NP; ++i) uT[(size_t)i * p.K] = u[i];
      }
    } else {
      const double* uT = p.uT_in + (size_t)b * NP * p.K + cx.k;
#pragma unroll
      for (int i = 0; i < NP; ++i) u[i] = active ? uT[(size_t)i * p.K] : 0.0;
    }

    // ------------------------------------------------------------------ adjoint phase
    if (DO_ADJ) {
      // checkpoints were written through the generic proxy; TMA reads them through the
      // async proxy -> fence, then make every thread's writes visible to the issuing thread.
      if (DO_FWD) {
        __threadfence();
        fence_proxy_async();
      }
      __syncthreads();  // also: everyone is done with `park` (aliases the landing buffers)
      const uint32_t tile_bytes = (uint32_t)(tile * sizeof(double));
      double* land0 = sm.big;
      double* land1 = sm.big + tile;
      if (tid == 0) {
        if (p.S >= 1) {
          mbar_expect_tx(&sm.mbar[0], tile_bytes);
          tma_bulk_g2s(land0, ck + (size_t)(p.S - 1) * tile, tile_bytes, &sm.mbar[0]);
        }
        if (p.S >= 2) {
          mbar_expect_tx(&sm.mbar[1], tile_bytes);
          tma_bulk_g2s(land1, ck + (size_t)(p.S - 2) * tile, tile_bytes, &sm.mbar[1]);
        }
      }
      // terminal condition lam^S = dJ_f/du at P u^S, and J of the coarse solution
      double lu[NPF], lk[NPF];
      double jpart = 0.0;
      if (p.func == FUNC_LINEAR) {
#pragma unroll
        for (int i = 0; i < NPF; ++i) lu[i] = in_tile ? p.jw_f[(size_t)i * p.K + cx.k] : 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) jpart = fma(in_tile ? p.jw_c[(size_t)i * p.K + cx.k] : 0.0, u[i], jpart);
      } else {
        const double jacC = 1.0 / rxk[0], jacF = 1.0 / rxk[1];
        double uf[NPF];
#pragma unroll
        for (int i = 0; i < NPF; ++i) {
          double acc = c.P[i * NP] * u[0];
#pragma unroll
          for (int j = 1; j < NP; ++j) acc = fma(c.P[i * NP + j], u[j], acc);
          uf[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NPF; ++i) {
          double acc = c.Mref[1][i * NPF] * uf[0];
#pragma unroll
          for (int j = 1; j < NPF; ++j) acc = fma(c.Mref[1][i * NPF + j], uf[j], acc);
          lu[i] = in_tile ? 2.0 * jacF * acc : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          double acc = c.Mref[0][i * NP] * u[0];
#pragma unroll
          for (int j = 1; j < NP; ++j) acc = fma(c.Mref[0][i * NP + j], u[j], acc);
          jpart = fma(u[i], in_tile ? jacC * acc : 0.0, jpart);
        }
      }
#pragma unroll
      for (int i = 0; i < NPF; ++i) lk[i] = 0.0;
      const double Jtot = traj_sum(cx, sm.red, jpart);
      if (p.J && active && cx.is_first) p.J[b] = Jtot;

      double eta = 0.0;
      for (int n = p.S - 1; n >= 0; --n) {
        const int it = p.S - 1 - n;
        const int buf = it & 1;
        const uint32_t parity = (land_uses[buf] + (uint32_t)(it >> 1)) & 1u;
        mbar_wait(&sm.mbar[buf], parity);
        const double* rho = (buf ? land1 : land0) + tid;
#pragma unroll
        for (int i = 0; i < NPF; ++i) eta = fma(lu[i], rho[(size_t)i * BD], eta);
        __syncthreads();  // every thread has consumed this landing buffer
        if (tid == 0 && n >= 2) {
          mbar_expect_tx(&sm.mbar[buf], tile_bytes);
          tma_bulk_g2s(buf ? land1 : land0, ck + (size_t)(n - 2) * tile, tile_bytes, &sm.mbar[buf]);
        }
        adj_step<NPF, 1>(cx, lu, lk, mF, f0F, f1F);
      }
      land_uses[0] += (uint32_t)((p.S + 1) >> 1);
      land_uses[1] += (uint32_t)(p.S >> 1);
      if (active) {
        if (p.eta) p.eta[(size_t)b * p.K + cx.k] = eta;
        if (p.lam0) {
          double* l0 = p.lam0 + (size_t)b * NPF * p.K + cx.k;
#pragma unroll
          for (int i = 0; i < NPF; ++i) l0[(size_t)i * p.K] = lu[i];
        }
      }
      __syncthreads();  // landing buffers / red free before the next group reuses them
    }
  }
}

// ---------------------------------------------------------------------------------------
// rank / refine flag:  one CTA per trajectory, stable descending rank of |eta| by counting
// (rank_k = #{j : |eta_j| > |eta_k|  or (== and j < k)}), exact and deterministic.
// ---------------------------------------------------------------------------------------
__global__ void rank_kernel(long long B, int K, const double* __restrict__ eta, int topk,
                            int32_t* __restrict__ order, uint8_t* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* ae = reinterpret_cast<double*>(smem_raw);
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) ae[k] = fabs(eta[(size_t)b * K + k]);
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const double v = ae[k];
      int r = 0;
      for (int j = 0; j < K; ++j) {
        const double w = ae[j];
        r += (w > v) || (w == v && j < k);
      }
      if (order) order[(size_t)b * K + r] = k;
      if (flags) flags[(size_t)b * K + k] = (r < topk) ? 1 : 0;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// batch reduction of the indicators in a fixed order: CTA j owns element column k = j;
// partial sums over fixed-size batch chunks are combined in chunk order.
// sums[k] = sum_b |eta[b][k]|;  sums[K..K+3] = { sum|eta|, sum eta^2, max|eta|, sum J }.
// ---------------------------------------------------------------------------------------
__global__ void reduce_cols_kernel(long long B, int K, const double* __restrict__ eta,
                                   double* __restrict__ colsum, double* __restrict__ colsq,
                                   double* __restrict__ colmax) {
  // grid = ceil(K/32) x 1, block = 32 x 8: thread (x,y) strides over b = y, y+8, ... in order
  __shared__ double s1[8][33], s2[8][33], s3[8][33];
  const int k = blockIdx.x * 32 + threadIdx.x;
  double a1 = 0, a2 = 0, a3 = 0;
  if (k < K) {
    for (long long b = threadIdx.y; b < B; b += 8) {
      const double v = fabs(eta[(size_t)b * K + k]);
      a1 += v;
      a2 = fma(v, v, a2);
      a3 = fmax(a3, v);
    }
  }
  s1[threadIdx.y][threadIdx.x] = a1;
  s2[threadIdx.y][threadIdx.x] = a2;
  s3[threadIdx.y][threadIdx.x] = a3;
  __syncthreads();
  if (threadIdx.y == 0 && k < K) {
    for (int y = 1; y < 8; ++y) {
      a1 += s1[y][threadIdx.x];
      a2 += s2[y][threadIdx.x];
      a3 = fmax(a3, s3[y][threadIdx.x]);
    }
    colsum[k] = a1;
    colsq[k] = a2;
    colmax[k] = a3;
  }
}
__global__ void reduce_final_kernel(long long B, int K, const double* __restrict__ colsq,
                                    const double* __restrict__ colmax, const double* __restrict__ J,
                                    double* __restrict__ sums) {
  // single thread block, thread 0 does the K-length ordered sums; J summed by 256 ordered lanes
  __shared__ double sj[256];
  double aj = 0;
  if (J) {
    for (long long b = threadIdx.x; b < B; b += 256) aj += J[b];
  }
  sj[threadIdx.x] = aj;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0, t3 = 0, tj = 0;
    for (int k = 0; k < K; ++k) {
      t1 += sums[k];
      t2 += colsq[k];
      t3 = fmax(t3, colmax[k]);
    }
    for (int i = 0; i < 256; ++i) tj += sj[i];
    sums[K + 0] = t1;
    sums[K + 1] = t2;
    sums[K + 2] = t3;
    sums[K + 3] = tj;
  }
}

// ---------------------------------------------------------------------------------------
// register-only DFMA peak microbenchmark (roofline denominator)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) dfma_peak_kernel(double* out, int iters, double x, double y) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      a0 = fma(a0, x, y);
      a1 = fma(a1, x, y);
      a2 = fma(a2, x, y);
      a3 = fma(a3, x, y);
      a4 = fma(a4, x, y);
      a5 = fma(a5, x, y);
      a6 = fma(a6, x, y);
      a7 = fma(a7, x, y);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[0] = s;  // keep the chain alive
}

}  // namespace dgadj
