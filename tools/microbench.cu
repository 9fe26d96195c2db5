// tools/microbench.cu -- fp64 pipe ceilings for the instruction mixes of the march kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
struct CB { double c[5][64]; };   // five copies, indexed by a loop counter so loads stay in the loop

// mode A: register operands only, NCH independent chains
template <int NCH>
__global__ void __launch_bounds__(1024, 1) k_reg(double* out, int iters, double x, double y) {
  double a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}
// mode B: one uniform constant load per USE DFMAs (constants indexed by a runtime-varying copy id)
template <int NCH, int USE>
__global__ void __launch_bounds__(1024, 1) k_cst(const __grid_constant__ CB cb, double* out, int iters, double y) {
  double a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const double* c = cb.c[it % 5];
#pragma unroll
    for (int r = 0; r < 64; ++r) {
#pragma unroll
      for (int u = 0; u < USE; ++u) a[(r * USE + u) % NCH] = fma(a[(r * USE + u) % NCH], c[r], y);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}
// mode C: fp64 tensor-core MMA (mma.sync m8n8k4), NCH independent accumulator chains per warp,
// optionally interleaved with DFMA chains (are they separate pipes on B200?)
template <int NCH, int NFMA>
__global__ void __launch_bounds__(1024, 1) k_dmma(double* out, int iters, double x, double y) {
  double c0[NCH], c1[NCH], f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < NFMA; ++i) f[i] = threadIdx.x + i;
  const double a = x, b = 1.0 + 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
#pragma unroll
      for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], x, y);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NFMA; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}
// mode D: the march kernel's instruction mix without its dependencies: per 64 constants (1 LDCU : 4
// DFMA = 256 DFMA) NLDS shared-memory loads whose values feed DFMAs, NSTS stores and NIMAD integer
// multiply-adds -- is the fp64 pipe held back by the *mix* at 8 warps per SM, or by what the real
// kernel waits for (barriers, exchange latencies)?
template <int NCH, int NLDS, int NSTS, int NIMAD, int FRESH = 1, int KIND = 0>
__global__ void __launch_bounds__(256, 1) k_mix(const __grid_constant__ CB cb, double* out, int iters, double y) {
  __shared__ double sm[4096];
  double a[NCH];
  int k[4] = {(int)threadIdx.x, 1, 2, 3};
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-9 * i;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const double* c = cb.c[it % 5];
    double add[NLDS > 0 ? NLDS : 1];
#pragma unroll
    for (int l = 0; l < NLDS; ++l) add[l] = sm[(threadIdx.x + 64 * l + it) & 4095];
#pragma unroll
    for (int r = 0; r < 64; ++r) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = (r * 4 + u) % NCH;
        a[q] = fma(a[q], c[r], (NLDS > 0 && (r * 4 + u) < NLDS) ? add[r * 4 + u] : y);
      }
      if (r < NIMAD) {   // four independent integer chains; KIND picks the instruction class
        int& kk = k[r & 3];
        if (KIND == 0) kk = kk * 3 + it;                                   // IMAD
        else if (KIND == 1) kk = kk ^ (it + r);                            // LOP3 (ALU)
        else if (KIND == 2) kk = __funnelshift_l(kk, it, 3);               // SHF (ALU)
        else if (KIND == 3) asm volatile("mov.b32 %0, %1;" : "=r"(kk) : "r"(k[(r + 1) & 3]));   // register move
        else if (KIND == 4) kk = kk * it + r;                              // IMAD, register multiplier
      }
      // FRESH: store the accumulator the DFMA just above wrote; otherwise one written 16 DFMAs ago
      if (r < NSTS) sm[(threadIdx.x + 2048 + 32 * r) & 4095] = a[(r * 4 + (FRESH ? 3 : 3 + NCH - 16)) % NCH];
    }
  }
  double s = k[0] + k[1] + k[2] + k[3];
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}
template <typename F>
double run(F launch, double flops_per_iter_per_thread, int threads, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return flops_per_iter_per_thread * iters * (double)threads * 148 / (best * 1e-3) / 1e12;
}
int main() {
  double* out; cudaMalloc(&out, 8);
  CB cb; for (int i = 0; i < 5; ++i) for (int j = 0; j < 64; ++j) cb.c[i][j] = 0.999 + 1e-6 * j;
  const int IT = 4000;
  printf("reg  1024thr 8ch : %.2f TF\n", run([&](int it){ k_reg<8><<<148,1024>>>(out,it,0.999,1e-9); }, 2.0*16*8, 1024, IT));
  printf("reg   512thr 8ch : %.2f TF\n", run([&](int it){ k_reg<8><<<148,512>>>(out,it,0.999,1e-9); }, 2.0*16*8, 512, IT));
  printf("reg   512thr 4ch : %.2f TF\n", run([&](int it){ k_reg<4><<<148,512>>>(out,it,0.999,1e-9); }, 2.0*16*4, 512, IT));
  printf("reg   512thr 2ch : %.2f TF\n", run([&](int it){ k_reg<2><<<148,512>>>(out,it,0.999,1e-9); }, 2.0*16*2, 512, IT));
  printf("reg   256thr 8ch : %.2f TF\n", run([&](int it){ k_reg<8><<<148,256>>>(out,it,0.999,1e-9); }, 2.0*16*8, 256, IT));
  printf("cst 1:1 512thr 8ch: %.2f TF\n", run([&](int it){ k_cst<8,1><<<148,512>>>(cb,out,it,1e-9); }, 2.0*64*1, 512, IT));
  printf("cst 1:2 512thr 8ch: %.2f TF\n", run([&](int it){ k_cst<8,2><<<148,512>>>(cb,out,it,1e-9); }, 2.0*64*2, 512, IT));
  printf("cst 1:4 512thr 8ch: %.2f TF\n", run([&](int it){ k_cst<8,4><<<148,512>>>(cb,out,it,1e-9); }, 2.0*64*4, 512, IT));
  printf("cst 1:2 1024thr 8ch: %.2f TF\n", run([&](int it){ k_cst<8,2><<<148,1024>>>(cb,out,it,1e-9); }, 2.0*64*2, 1024, IT));
  printf("cst 1:1 1024thr 8ch: %.2f TF\n", run([&](int it){ k_cst<8,1><<<148,1024>>>(cb,out,it,1e-9); }, 2.0*64*1, 1024, IT));
  printf("mix 256thr 32ch, DFMA + LDCU only (1:4)              : %.2f TF\n", run([&](int it){ k_mix<32,0,0,0><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +9 STS +18 IMAD per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,9,18,1,4><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +18 IMAD (dependent chains, imm) per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,0,18><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +18 LOP3         per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,0,18,1,1><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +18 SHF          per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,0,18,1,2><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +18 MOV          per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,0,18,1,3><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +18 IMAD (reg)   per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,0,18,1,4><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +9 STS (stale)   per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,9,0,0><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS +9 STS (fresh)   per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,9,0,1><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +22 LDS               per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,22,0,0><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 32ch +44 LDS +18 STS +36 IMAD per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<32,44,18,36,1,4><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  printf("mix 256thr 8ch  +22 LDS +9 STS +18 IMAD per 256 DFMA: %.2f TF\n", run([&](int it){ k_mix<8,22,9,18,1,4><<<148,256>>>(cb,out,it,1e-9); }, 2.0*64*4, 256, IT));
  // m8n8k4: 8*8*4*2 = 512 flop per warp instruction = 16 flop per thread
  printf("dmma 512thr 4ch       : %.2f TF (mma only)\n", run([&](int it){ k_dmma<4,0><<<148,512>>>(out,it,0.999,1e-9); }, 16.0*8*4, 512, IT));
  printf("dmma 1024thr 4ch      : %.2f TF (mma only)\n", run([&](int it){ k_dmma<4,0><<<148,1024>>>(out,it,0.999,1e-9); }, 16.0*8*4, 1024, IT));
  printf("dmma+dfma 512thr 4+8  : %.2f TF (mma 16*4 + fma 2*8 per thread-iter)\n", run([&](int it){ k_dmma<4,8><<<148,512>>>(out,it,0.999,1e-9); }, (16.0*4+2.0*8)*8, 512, IT));
  printf("dmma+dfma 1024thr 4+8 : %.2f TF\n", run([&](int it){ k_dmma<4,8><<<148,1024>>>(out,it,0.999,1e-9); }, (16.0*4+2.0*8)*8, 1024, IT));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
