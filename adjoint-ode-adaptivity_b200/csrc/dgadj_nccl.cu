// dgadj_nccl.cu -- the one cross-GPU exchange of the path: the K+4 indicator
// partials of dgadj_reduce_indicators combined over the ranks of the caller's NCCL communicator,
// feeding the batch-mean refinement rule of python/Main_variable_params.py:340-341 when the batch
// is sharded over GPUs.  All-gather over NVLink + a fixed rank-order sum on every rank: the
// result has the same bits on every rank and for every reduction topology (a ring/tree all-reduce
// would not guarantee that).  With one partial per rank the batch sum still depends on the NUMBER of
// ranks (each partial is summed over its own shard); the *_blocks entry points reduce fixed-size
// blocks of trajectories and combine them in global block order, so that 1/2/4/8-GPU runs of one
// global batch give the same bits and refine the same element.
//
// NCCL is resolved at run time (dlsym on the process, then dlopen of libnccl.so.2): the library
// has no link-time dependency on it, and inside a torch process the already-loaded NCCL -- the one
// the caller's communicator belongs to -- is the one that gets used.
#include <dlfcn.h>
#include <nccl.h>

#include "dgadj_internal.h"

namespace dgadj {

struct NcclApi {
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommCount)(const ncclComm_t, int*);
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static const NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a = {};
    void* lib = nullptr;
    if (!dlsym(RTLD_DEFAULT, "ncclAllGather")) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    auto sym = [&](const char* n) -> void* {
      void* p = dlsym(RTLD_DEFAULT, n);
      if (!p && lib) p = dlsym(lib, n);
      return p;
    };
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.CommCount = (decltype(a.CommCount))sym("ncclCommCount");
    a.CommUserRank = (decltype(a.CommUserRank))sym("ncclCommUserRank");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.AllGather && a.CommCount && a.CommUserRank && a.GetErrorString;
    return a;
  }();
  return api;
}

// out[j] = parts[0][j] (+ | max) parts[1][j] ... in rank order; entry K+2 is a maximum
__global__ void combine_partials_kernel(const double* __restrict__ parts, int G, int K, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K + 4) return;
  double acc = parts[j];
  for (int r = 1; r < G; ++r) {
    const double v = parts[(size_t)r * (K + 4) + j];
    acc = (j == K + 2) ? fmax(acc, v) : acc + v;
  }
  out[j] = acc;
}

}  // namespace dgadj

extern "C" int dgadj_allreduce_indicators(dgadj_handle* h, void* nccl_comm, int32_t K, double* sums_dev, void* stream) {
  using namespace dgadj;
  if (!h) return DGADJ_ERR_INVALID;
  if (!nccl_comm || K <= 0 || !sums_dev) return fail(h, DGADJ_ERR_INVALID, "bad allreduce_indicators arguments");
  const NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(h, DGADJ_ERR_STATE, "NCCL (libnccl.so.2) could not be resolved in this process");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  int G = 0;
  ncclResult_t r = nc.CommCount(comm, &G);
  if (r != ncclSuccess) return fail(h, DGADJ_ERR_CUDA, "ncclCommCount: %s", nc.GetErrorString(r));
  if (G == 1) return DGADJ_OK;
  const size_t need = (size_t)G * (K + 4) * sizeof(double);
  if (need > h->nccl_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->nccl_scratch);
    h->nccl_scratch = nullptr;
    h->nccl_bytes = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->nccl_scratch, need));
    h->nccl_bytes = need;
  }
  r = nc.AllGather(sums_dev, h->nccl_scratch, (size_t)(K + 4), ncclDouble, comm, st);
  if (r != ncclSuccess) return fail(h, DGADJ_ERR_CUDA, "ncclAllGather: %s", nc.GetErrorString(r));
  const int block = 128;
  combine_partials_kernel<<<(K + 4 + block - 1) / block, block, 0, st>>>(h->nccl_scratch, G, K, sums_dev);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return DGADJ_OK;
}

// Count-independent form: every rank holds nblk_local rows of K+4 block partials (dgadj_reduce_indicator_blocks
// over its shard, whose length is a whole number of blocks); the rows of all ranks are all-gathered in rank
// order -- which is global block order for the contiguous shards of sharding.shard_range -- and summed in that
// order on every rank.  nccl_comm == NULL (or a communicator of one rank): the local rows only.
extern "C" int dgadj_allreduce_indicator_blocks(dgadj_handle* h, void* nccl_comm, int32_t K, int32_t nblk_local,
                                                const double* parts_dev, double* sums_dev, void* stream) {
  using namespace dgadj;
  if (!h) return DGADJ_ERR_INVALID;
  if (K <= 0 || nblk_local <= 0 || !parts_dev || !sums_dev) return fail(h, DGADJ_ERR_INVALID, "bad allreduce_indicator_blocks arguments");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  int G = 1;
  const NcclApi* nc = nullptr;
  if (nccl_comm) {
    nc = &nccl_api();
    if (!nc->ok) return fail(h, DGADJ_ERR_STATE, "NCCL (libnccl.so.2) could not be resolved in this process");
    ncclResult_t r = nc->CommCount((ncclComm_t)nccl_comm, &G);
    if (r != ncclSuccess) return fail(h, DGADJ_ERR_CUDA, "ncclCommCount: %s", nc->GetErrorString(r));
  }
  const double* rows = parts_dev;
  if (G > 1) {
    const size_t need = (size_t)G * nblk_local * (K + 4) * sizeof(double);
    if (need > h->nccl_bytes) {
      CUDA_TRY(h, cudaDeviceSynchronize());
      cudaFree(h->nccl_scratch);
      h->nccl_scratch = nullptr;
      h->nccl_bytes = 0;
      CUDA_TRY(h, cudaMalloc((void**)&h->nccl_scratch, need));
      h->nccl_bytes = need;
    }
    ncclResult_t r = nc->AllGather(parts_dev, h->nccl_scratch, (size_t)nblk_local * (K + 4), ncclDouble, (ncclComm_t)nccl_comm, st);
    if (r != ncclSuccess) return fail(h, DGADJ_ERR_CUDA, "ncclAllGather: %s", nc->GetErrorString(r));
    rows = h->nccl_scratch;
  }
  const int block = 128;
  combine_partials_kernel<<<(K + 4 + block - 1) / block, block, 0, st>>>(rows, G * nblk_local, K, sums_dev);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return DGADJ_OK;
}
