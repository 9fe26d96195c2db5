// dgadj_march_np.cu -- one translation unit per primal order: compiled 8 times with
// -DDGADJ_NP=2..9 (N = 1..8) so the per-order kernels build in parallel.  Exports
// dgadj::march_launch_np<NP>(variant, grid, block, stream, args).
#define DGADJ_DEVICE_CODE 1
#include "dgadj_kernels.cuh"

#ifndef DGADJ_NP
#error "compile with -DDGADJ_NP=<Np>"
#endif

namespace dgadj {

template <int NP, int EPT, int BDT, bool F, bool R, bool A, bool HP = false>
static cudaError_t launch_bd(int variant, int grid, int block, size_t smem, cudaStream_t stream, const KArgs* ka) {
  // the opt-in shared-memory size is a per-device function attribute: remembered per (instantiation,
  // device), so that a process with handles on several GPUs sets it on each of them
  static bool attr_set[64] = {};
  if (block > MAXBD / EPT) return cudaErrorInvalidConfiguration;
  if (smem < march_smem_bytes(NP, EPT, block, variant)) return cudaErrorInvalidConfiguration;
  auto kern = march_kernel<NP, EPT, BDT, F, R, A, HP>;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  kern<<<grid, block, smem, stream>>>(*ka);
  return cudaGetLastError();
}

#define DGADJ_CAT2(a, b) a##b
#define DGADJ_CAT(a, b) DGADJ_CAT2(a, b)

// the hot variants (forward, fused) also exist with the two common block sizes baked in
template <int NP, int EPT, bool F, bool R, bool A>
static cudaError_t launch_one(int variant, int grid, int block, size_t smem, cudaStream_t stream, const KArgs* ka) {
  if (F && R == A) {
    if (block == MAXBD / EPT) return launch_bd<NP, EPT, MAXBD / EPT, F, R, A>(variant, grid, block, smem, stream, ka);
    if (block == MAXBD / EPT / 2) return launch_bd<NP, EPT, MAXBD / EPT / 2, F, R, A>(variant, grid, block, smem, stream, ka);
  }
  return launch_bd<NP, EPT, 0, F, R, A>(variant, grid, block, smem, stream, ka);
}

// hp (per-element orders, p.npk set): the forward march and the fused march, any block size (BDT = 0)
template <int EPT>
static cudaError_t launch_ept_hp(int variant, int grid, int block, size_t smem, cudaStream_t stream, const KArgs* ka) {
  switch (variant) {
    case VAR_FWD: return launch_bd<DGADJ_NP, EPT, 0, true, false, false, true>(variant, grid, block, smem, stream, ka);
#if DGADJ_NP + 1 <= 10
    case VAR_FUSED: return launch_bd<DGADJ_NP, EPT, 0, true, true, true, true>(variant, grid, block, smem, stream, ka);
#endif
    default: return cudaErrorNotSupported;
  }
}

template <int EPT>
static cudaError_t launch_ept(int variant, int grid, int block, size_t smem, cudaStream_t stream, const KArgs* ka) {
  switch (variant) {
    case VAR_FWD: return launch_one<DGADJ_NP, EPT, true, false, false>(variant, grid, block, smem, stream, ka);
#if DGADJ_NP + 1 <= 10  // the enriched space must fit MAXNP
    case VAR_FWD_RESID: return launch_one<DGADJ_NP, EPT, true, true, false>(variant, grid, block, smem, stream, ka);
    case VAR_ADJ: return launch_one<DGADJ_NP, EPT, false, false, true>(variant, grid, block, smem, stream, ka);
    case VAR_FUSED: return launch_one<DGADJ_NP, EPT, true, true, true>(variant, grid, block, smem, stream, ka);
#endif
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t DGADJ_CAT(march_launch_np, DGADJ_NP)(int variant, int ept, int grid, int block, size_t smem,
                                                 cudaStream_t stream, const KArgs* ka) {
  if (ka->p.npk) {
    if (ept == 1) return launch_ept_hp<1>(variant, grid, block, smem, stream, ka);
    if (ept == 2) return launch_ept_hp<2>(variant, grid, block, smem, stream, ka);
    if (ept == 4) return launch_ept_hp<4>(variant, grid, block, smem, stream, ka);
    return cudaErrorInvalidValue;
  }
  if (ept == 1) return launch_ept<1>(variant, grid, block, smem, stream, ka);
  if (ept == 2) return launch_ept<2>(variant, grid, block, smem, stream, ka);
  if (ept == 4) return launch_ept<4>(variant, grid, block, smem, stream, ka);
  return cudaErrorInvalidValue;
}

}  // namespace dgadj
