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

namespace dgadj {

struct BurgersArgs {
  long long B;
  int K, S, periodic, limit;
  double dt;
  const double* dt_arr;
  const double* rxk;   // [K]
  const double* fs0;   // [K]
  const double* fs1;   // [K]
  const double* xc;    // [NP][K]  x - x0 (SlopeLimitLin.m:11-12)
  const double* hk;    // [K]      x(Np,k) - x(1,k)
  const double* u0;    // [B][NP][K]
  double* uT;          // [B][NP][K]
  double* hist;        // [B][S+1][NP][K] or null
  unsigned char* flags;  // [B][S][K] bit s = limited after stage s, or null
  double* maxvel;      // [B][S][5] or null
  StageOps so;
  double aw[MAXNP];    // cell average weights  V(1,1)*invV(1,:)
  double sl[MAXNP];    // slope weights         Dr(1,:)*V(:,1:2)*invV(1:2,:)
  double rka[5], rkb[5];
};

__device__ __forceinline__ double minmod3(double a, double b, double c) {
  // utils/minmod.m:7-11: s = sum(sign(v))/3; |s| == 1 -> s * min|v|, else 0
  const double sa = (a > 0.0) - (a < 0.0), sb = (b > 0.0) - (b < 0.0), sc = (c > 0.0) - (c < 0.0);
  const double s = (sa + sb + sc) / 3.0;
  if (fabs(s) == 1.0) return s * fmin(fabs(a), fmin(fabs(b), fabs(c)));
  return 0.0;
}

template <int NP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) burgers_kernel(const __grid_constant__ BurgersArgs p) {
  constexpr int HE = (NP + 1) / 2, HO = NP / 2;
  __shared__ double trL[2][MAXT], trR[2][MAXT], avg[MAXT];
  __shared__ double wmax[2][32];
  const int tid = threadIdx.x, K = p.K;
  const int lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
  const bool in = tid < K;
  const int k = in ? tid : 0;
  const int nbL = in ? (tid == 0 ? K - 1 : tid - 1) : tid;
  const int nbR = in ? (tid == K - 1 ? 0 : tid + 1) : tid;
  const bool first = in && tid == 0, last = in && tid == K - 1;
  const double rx = in ? p.rxk[k] : 0.0, fs0 = in ? p.fs0[k] : 0.0, fs1 = in ? p.fs1[k] : 0.0;
  const double h = in ? p.hk[k] : 1.0;

  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    const double dt = p.dt_arr ? p.dt_arr[b] : p.dt;
    double u[NP], res[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      u[i] = in ? p.u0[((size_t)b * NP + i) * K + k] : 0.0;
      res[i] = 0.0;
    }
    int par = 0;
    // limiter applied to the element's current state; returns 1 if the cell was limited
    auto limiter = [&]() -> int {
      double v = 0.0;
#pragma unroll
      for (int i = 0; i < NP; ++i) v = fma(p.aw[i], u[i], v);
      __syncthreads();
      avg[tid] = v;
      __syncthreads();
      if (!in) return 0;
      double vm = avg[nbL], vp = avg[nbR];
      if (!p.periodic) {       // quirk C-16: ghost averages copy the end cells
        if (first) vm = v;
        if (last) vp = v;
      }
      const double ue1 = u[0], ue2 = u[NP - 1];
      const double ve1 = v - minmod3(v - ue1, v - vm, vp - v);
      const double ve2 = v + minmod3(ue2 - v, v - vm, vp - v);
      if (!(fabs(ve1 - ue1) > 1.0e-8 || fabs(ve2 - ue2) > 1.0e-8)) return 0;
      double d = 0.0;
#pragma unroll
      for (int i = 0; i < NP; ++i) d = fma(p.sl[i], u[i], d);
      const double ux = (2.0 / h) * d;
      const double slope = minmod3(ux, (vp - v) / h, (v - vm) / h);
#pragma unroll
      for (int i = 0; i < NP; ++i) u[i] = v + p.xc[(size_t)i * K + k] * slope;
      return 1;
    };
    if (p.limit) limiter();
    double* hist = (p.hist && in) ? p.hist + ((size_t)b * (p.S + 1) * NP) * K + k : nullptr;
    if (hist) {
#pragma unroll
      for (int i = 0; i < NP; ++i) hist[(size_t)i * K] = u[i];
    }
    for (int n = 0; n < p.S; ++n) {
      unsigned fl = 0u;
#pragma unroll 1
      for (int s = 0; s < 5; ++s) {
        // ---- exchange 1: traces and the mesh-wide max|u|
        double m = 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) m = fmax(m, fabs(u[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        trL[par][tid] = u[0];
        trR[par][tid] = u[NP - 1];
        if (lane == 0) wmax[par][wid] = m;
        __syncthreads();
        double maxvel = 0.0;
        for (int w = 0; w < nw; ++w) maxvel = fmax(maxvel, wmax[par][w]);
        double uL = trR[par][nbL], uR = trL[par][nbR];
        par ^= 1;
        if (!p.periodic) {
          if (first) uL = u[0];
          if (last) uR = u[NP - 1];
        }
        if (p.maxvel && tid == 0) p.maxvel[((size_t)b * p.S + n) * 5 + s] = maxvel;
        // flux = nx (u-^2 - u+^2)/4 - C/2 (u- - u+),  g_f = Fscale_f * flux_f
        const double g0 = fs0 * (-((u[0] * u[0] - uL * uL) / 2.0) / 2.0 - maxvel / 2.0 * (u[0] - uL));
        const double g1 = fs1 * (((u[NP - 1] * u[NP - 1] - uR * uR) / 2.0) / 2.0 - maxvel / 2.0 * (u[NP - 1] - uR));
        const double ge = g0 + g1, go = g0 - g1;
        // f = u^2/2 in the even/odd basis
        double fe[HE], fo[HO > 0 ? HO : 1];
#pragma unroll
        for (int i = 0; i < NP / 2; ++i) {
          const double a = u[i] * u[i] / 2.0, c = u[NP - 1 - i] * u[NP - 1 - i] / 2.0;
          fe[i] = a + c;
          fo[i] = a - c;
        }
        if (NP & 1) fe[NP / 2] = u[NP / 2] * u[NP / 2] / 2.0;
        const double rka = p.rka[s], rkb = p.rkb[s];
        // rows: even part E_i = -rx (DE fo)_i + LS_i ge, odd part O_i = -rx (DO fe)_i + LA_i go
        double E[HE], O[HO > 0 ? HO : 1];
#pragma unroll
        for (int i = 0; i < HE; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < HO; ++j) {
            const double2 c2 = p.so.DE2[i * HP + j / 2];
            acc = fma((j & 1) ? c2.y : c2.x, fo[j], acc);
          }
          const double2 l2 = p.so.LS2[i / 2];
          E[i] = fma(-rx, acc, ((i & 1) ? l2.y : l2.x) * ge);
        }
#pragma unroll
        for (int i = 0; i < HO; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < HE; ++j) {
            const double2 c2 = p.so.DO2[i * HP + j / 2];
            acc = fma((j & 1) ? c2.y : c2.x, fe[j], acc);
          }
          const double2 l2 = p.so.LA2[i / 2];
          O[i] = fma(-rx, acc, ((i & 1) ? l2.y : l2.x) * go);
        }
        // back to nodal: rhs_i = (E_i + O_i)/2, rhs_{N-i} = (E_i - O_i)/2, mid = E_mid
#pragma unroll
        for (int i = 0; i < NP / 2; ++i) {
          const double r0 = 0.5 * (E[i] + O[i]), r1 = 0.5 * (E[i] - O[i]);
          res[i] = fma(rka, res[i], dt * r0);
          res[NP - 1 - i] = fma(rka, res[NP - 1 - i], dt * r1);
        }
        if (NP & 1) res[NP / 2] = fma(rka, res[NP / 2], dt * E[NP / 2]);
#pragma unroll
        for (int i = 0; i < NP; ++i) u[i] = fma(rkb, res[i], u[i]);
        // ---- exchange 2: cell averages, limiter
        if (p.limit) fl |= (unsigned)limiter() << s;
      }
      if (p.flags && in) p.flags[((size_t)b * p.S + n) * K + k] = (unsigned char)fl;
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
  }
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

extern "C" int dgadj_burgers_forward(dgadj_handle* h, int64_t B, int32_t S, double dt, const double* dt_dev,
                                     int32_t limit, const double* invV_host, const double* V_host,
                                     const double* x_host, const double* u0_dev, double* uT_dev,
                                     double* hist_dev, uint8_t* flags_dev, double* maxvel_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || S < 0 || !u0_dev) return fail(h, DGADJ_ERR_INVALID, "bad burgers arguments");
  if (!h->ops_set) return fail(h, DGADJ_ERR_STATE, "dgadj_set_operators has not been called");
  if (limit && (!invV_host || !V_host || !x_host)) return fail(h, DGADJ_ERR_INVALID, "the limiter needs V, invV and x");
  if (h->nstages != 5) return fail(h, DGADJ_ERR_UNSUPPORTED, "the Burgers march is LSERK4 only");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Np = h->Np, K = h->K;
  BurgersArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.K = K;
  a.S = S;
  a.periodic = (h->cfg.bc == DGADJ_BC_PERIODIC);
  a.limit = limit ? 1 : 0;
  a.dt = dt;
  a.dt_arr = dt_dev;
  a.rxk = h->d_mesh[0][0];
  a.fs0 = h->d_mesh[0][1];
  a.fs1 = h->d_mesh[0][2];
  a.u0 = u0_dev;
  a.uT = uT_dev;
  a.hist = hist_dev;
  a.flags = flags_dev;
  a.maxvel = maxvel_dev;
  a.so = h->base_ops[0];
  for (int s = 0; s < 5; ++s) {
    a.rka[s] = h->cops.rka[s];
    a.rkb[s] = h->cops.rkb[s];
  }
  if (limit) {
    // aw = V(1,1)*invV(1,:) (SlopeLimitN.m:9);  sl = Dr(1,:)*V(:,1:2)*invV(1:2,:) (SlopeLimitN.m:27,
    // SlopeLimitLin.m:16);  xc = x - x0, h = x(Np,:) - x(1,:) (SlopeLimitLin.m:10-12)
    std::vector<double> xc((size_t)Np * K), hk(K);
    for (int i = 0; i < Np; ++i) a.aw[i] = V_host[0] * invV_host[i];
    const double* Dr = h->Dr_nodal;
    for (int i = 0; i < Np; ++i) {
      double s = 0.0;
      for (int j = 0; j < Np; ++j)
        s += Dr[j] * (V_host[(size_t)j * Np + 0] * invV_host[0 * Np + i] + (Np > 1 ? V_host[(size_t)j * Np + 1] * invV_host[1 * Np + i] : 0.0));
      a.sl[i] = s;
    }
    for (int k = 0; k < K; ++k) {
      const double hh = x_host[(size_t)(Np - 1) * K + k] - x_host[k];
      const double x0 = x_host[k] + hh / 2;
      hk[k] = hh;
      for (int i = 0; i < Np; ++i) xc[(size_t)i * K + k] = x_host[(size_t)i * K + k] - x0;
    }
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
  } else {
    a.xc = h->d_mesh[0][0];  // never read
    a.hk = h->d_mesh[0][0];
  }
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
