// dgadj_internal.h -- handle layout and error helpers shared by the translation units of
// libdgadj.so.  Not part of the public ABI (include/dgadj.h).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dgadj.h"
#include "dgadj_kernels.cuh"

using namespace dgadj;

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
struct LaunchPlan {
  int ept, KT, tpc, block, ngroups, grid;
  size_t smem;
  size_t tile;  // doubles per checkpoint tile
  int tma_store;  // residual tiles leave through bulk-TMA from a double-buffered park (smem includes it)
};

struct dgadj_handle {
  dgadj_config cfg;
  int Np, NpF, K, nstages;
  bool ops_set, enr_set, jw_set;
  ConstOps cops;
  StageOps base_ops[2];  // even/odd blocks of the nodal operators (Burgers kernels)
  double Dnz[2][MAXNZ];   // unscaled non-zeros of the modal derivative, both levels
  double Vhost[2][MAXNP * MAXNP];  // host copies of the Vandermonde matrices
  double* d_mesh[2][3];  // [level]{rx, fs0, fs1} each [K]
  double* d_nodal[2][2];  // [level]{Dr[Np*Np], LIFT[Np*2]} nodal copies (dgadj_rhs)
  double* d_jwc;          // nodal primal weights (Burgers adjoint)
  double* d_P;            // nodal prolongation P[NpF][Np] (initial-data term of the indicator)
  double* d_jwm_c;        // modal weights V^T jw of the linear functional, primal / enriched
  double* d_jwm_f;
  double* d_uin;
  int uin_n;
  int* d_npk;             // hp: [K] modes per element of the primal space, or null (uniform order)
  double* d_Php;          // hp: projected prolongations P_n, n = 0..Np (dgadj_ic_indicator)
  double* ring;
  size_t ring_bytes;
  double* red_scratch;
  size_t red_bytes;
  void* fd_scratch;     // dgadj_fd_awr: coarse states of the thread form; the FD loops' tables / states
  size_t fd_bytes;
  void* fd_tbl;         // dgadj_fd_awr: interpolation tables of the last mesh (kept while the same steps come back)
  size_t fd_tbl_bytes, fd_tbl_dt_cap;
  double* fd_tbl_dt;    //   host copy of the steps the tables were built from
  int fd_tbl_n, fd_tbl_rf;
  double* tdg_scratch;  // dgadj_tdg_*adapt_loop*: templates, element blocks, mesh history
  size_t tdg_bytes;
  // dgadj_tdg_march / _adjoint / _adjoint_rec / _err_contribution: the per-element constant blocks of the last few
  // meshes stay on the device (host copy kept for the comparison), so a repeated call is a kernel launch and nothing
  // else -- no upload, no stream synchronisation
  struct TdgConstSlot {
    double* dev;
    double* host;
    size_t cap, n;
    unsigned long long stamp;
  } tdg_cache[4];
  unsigned long long tdg_clock;
  double* bg_scratch;   // dgadj_burgers_forward: limiter geometry
  size_t bg_bytes;
  double* bgs_scratch;  // dgadj_burgers_adjoint: per-CTA stage states
  size_t bgs_bytes;
  double* nccl_scratch; // dgadj_allreduce_indicators: gathered partials [ranks][K+4]
  size_t nccl_bytes;
  double Dr_nodal[MAXNP * MAXNP];  // host copies of the primal nodal Dr / LIFT
  double LIFT_nodal[MAXNP * 2];
  double Dr_nodal_f[MAXNP * MAXNP];  // ... and of the enriched space, and the nodal prolongation
  double LIFT_nodal_f[MAXNP * 2];
  double P_host[MAXNP * MAXNP];
  double* bgf_consts;   // dgadj_burgers_fwd_adj: element widths + functional weights
  size_t bgf_consts_bytes;
  int sm_count, cc_major, cc_minor;
  size_t total_mem;
  int tune_ept, tune_block, tune_grid;
  // host pipeline
  cudaStream_t s_in, s_k, s_out;
  cudaEvent_t ev_in[2], ev_k[2], ev_out[2];
  double* dbuf[2];
  size_t dbuf_bytes;
  double* pin[2];
  size_t pin_bytes;
  bool pipe_init;
  char err[512];
  int64_t launches;
};

static inline int fail(dgadj_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}
#define CUDA_TRY(h, call)                                                                      \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail((h), e_ == cudaErrorMemoryAllocation ? DGADJ_ERR_NOMEM : DGADJ_ERR_CUDA,      \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

