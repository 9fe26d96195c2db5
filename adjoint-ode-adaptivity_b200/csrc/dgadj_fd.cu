// dgadj_fd.cu -- the finite-difference path of python/Main_finite_difference.py, batched over
// initial conditions on a shared time mesh (the batch pattern of Main_variable_params.py:330-344):
//   forwardSolve (:34-51)  explicit Euler on the coarse mesh
//   adjSolve     (:54-76)  discrete adjoint on the ref_factor-refined mesh.  The reference solves
//                          (JF^T - I) v = -k densely; JF is sub-diagonal, so that system is the
//                          backward recurrence v_N = 0, v_i = k_i + jf_i v_{i+1}
//   errEst       (:79-94)  res[n] = u_f[n] - fwdUpdate(u_f, dt_f, n);  err = res * v
//   driver       (:270-277, :337)  |err|, drop two entries, windows of ref_factor-1 at stride
//                          ref_factor, argmax (lowest index on ties)
// One thread marches one trajectory; the mesh-dependent interpolation tables (np.interp of
// interpU, :24-31) are built once on the host and shared by the batch.  All arithmetic is
// written unfused (__dmul_rn / __dadd_rn) to mirror NumPy's separate multiply and add.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dgadj_internal.h"

namespace dgadj {

enum { FD_ODE_SIN = 0, FD_ODE_LINEAR = 1 };
enum { FD_FUNC_INT_U = 0, FD_FUNC_U_N = 1, FD_FUNC_INT_U2 = 2 };
constexpr int FD_MAX_REF = 16;

struct FdTables {
  const int* jc;       // [nf+1] coarse interval of fine node i (np.interp's binary search)
  const double* dx;    // [nf+1] t_fine[i] - t_coarse[jc[i]]
  const double* den;   // [n]    t_coarse[j+1] - t_coarse[j]
  const double* dtf;   // [nf]   fine steps
  const double* dtn;   // [n]    coarse steps
  const unsigned char* exact;  // [nf+1] 1: fine node coincides with a coarse node (returns fp[j])
};

__device__ __forceinline__ double fd_interp(const double* __restrict__ uc, long long stride, const FdTables& t,
                                            int i, int n) {
  const int j = t.jc[i];
  const double uj = uc[(size_t)j * stride];
  if (t.exact[i] || j >= n) return uj;
  // np.interp: slope = (fp[j+1]-fp[j])/(xp[j+1]-xp[j]);  res = slope*(x-xp[j]) + fp[j]
  const double slope = __ddiv_rn(__dadd_rn(uc[(size_t)(j + 1) * stride], -uj), t.den[j]);
  return __dadd_rn(__dmul_rn(slope, t.dx[i]), uj);
}

__global__ void fd_awr_kernel(long long B, int n, int rf, int ode, int func, FdTables t,
                              const double* __restrict__ u0, double* __restrict__ uc /*[n+1][B] scratch*/,
                              double* __restrict__ u_out, double* __restrict__ v_out,
                              double* __restrict__ err_out, double* __restrict__ steps_out,
                              int* __restrict__ idx_out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nf = n * rf;
  // ---- forwardSolve: u[m] = fwdUpdate(u, dt, m)
  double* ub = uc + b;  // column b of the [n+1][B] scratch (coalesced across the warp)
  double u = u0[b];
  ub[0] = u;
  if (u_out) u_out[(size_t)b * (n + 1)] = u;
  for (int m = 1; m <= n; ++m) {
    const double dt = t.dtn[m - 1];
    u = (ode == FD_ODE_LINEAR) ? __dmul_rn(__dadd_rn(1.0, dt), u) : __dadd_rn(u, __dmul_rn(sin(u), dt));
    ub[(size_t)m * B] = u;
    if (u_out) u_out[(size_t)b * (n + 1) + m] = u;
  }
  // ---- one backward sweep over the fine mesh: adjoint recurrence, residual, window sums
  double v_next = 0.0;                                  // v[nf] = v0 = 0
  double uf_next = fd_interp(ub, B, t, nf, n);          // u_fine[nf]
  if (v_out) v_out[(size_t)b * (nf + 1) + nf] = 0.0;
  double win[FD_MAX_REF];
  double best = -1.0;
  int best_idx = 0;
  for (int i = nf - 1; i >= 0; --i) {
    const double uf = fd_interp(ub, B, t, i, n);
    const double dt = t.dtf[i];
    double k, jf, upd;
    if (ode == FD_ODE_LINEAR) {
      jf = __dadd_rn(1.0, dt);                          // getJF :118-119
      upd = __dmul_rn(__dadd_rn(1.0, dt), uf);          // fwdUpdate :112-113
    } else {
      jf = __dadd_rn(1.0, __dmul_rn(cos(uf), dt));      // getJF :138-139
      upd = __dadd_rn(uf, __dmul_rn(sin(uf), dt));      // fwdUpdate :131-132
    }
    if (func == FD_FUNC_INT_U2) k = __dmul_rn(__dmul_rn(2.0, uf), dt);   // getK :225-227
    else if (func == FD_FUNC_INT_U) k = dt;                              // :153-155
    else k = (i == nf - 1) ? 1.0 : 0.0;                                  // :162-165
    const double err = __dmul_rn(__dadd_rn(uf_next, -upd), v_next);      // err[i+1] = res[i+1]*v[i+1]
    if (err_out) err_out[(size_t)b * (nf + 1) + i + 1] = err;
    // window r covers fine entries 2 + r*rf + q, q = 0..rf-2  (:270-277)
    const int pos = i + 1 - 2;
    if (pos >= 0) {
      const int r = pos / rf, q = pos - r * rf;
      if (q < rf - 1) {
        win[q] = fabs(err);
        if (q == 0) {                                   // window complete: sum in ascending order
          double s = win[0];
          for (int qq = 1; qq < rf - 1; ++qq) s = __dadd_rn(s, win[qq]);
          if (steps_out) steps_out[(size_t)b * n + r] = s;
          if (s >= best) {                              // descending r: >= keeps the lowest index
            best = s;
            best_idx = r;
          }
        }
      }
    }
    const double v = __dadd_rn(k, __dmul_rn(jf, v_next));
    if (v_out) v_out[(size_t)b * (nf + 1) + i] = v;
    v_next = v;
    uf_next = uf;
  }
  if (err_out) err_out[(size_t)b * (nf + 1)] = __dmul_rn(0.0, v_next);   // res_u[0] = 0
  if (idx_out) idx_out[b] = best_idx;
}


// ------------------------------------------------------------------------------------------------------
// Small batches: ONE WARP per trajectory.  A thread-per-trajectory launch of a few thousand trajectories
// leaves one warp per SM to run n + 2 nf dependent sin / cos evaluations (118 us at B = 4096, n = 31).
// Only two recurrences are sequential -- the forward Euler steps and v_i = k_i + jf_i v_{i+1} -- and of
// those only the first costs transcendentals; interpolation, jf / fwdUpdate (the 2 nf sin / cos), the
// residuals, the window sums and the argmax are independent per fine node and go lane-strided through
// the warp's shared arrays.  Every value is produced by the arithmetic of fd_awr_kernel in the same
// order (the recurrences run redundantly in all lanes), so the results are bit-identical.
//   shared per warp: uc[n+1] | uf[nf+1] | jf[nf] | rs[nf+1] (residual, then |err|) | v[nf+1]
// ------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t fd_warp_doubles(int n, int rf) { return (size_t)(n + 1) + 4 * ((size_t)n * rf + 1); }

__global__ void fd_awr_warp_kernel(long long B, int n, int rf, int ode, int func, FdTables t,
                                   const double* __restrict__ u0, double* __restrict__ u_out, double* __restrict__ v_out,
                                   double* __restrict__ err_out, double* __restrict__ steps_out, int* __restrict__ idx_out) {
  extern __shared__ double fdsm[];
  const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const long long b = (long long)blockIdx.x * wpc + (threadIdx.x >> 5);
  if (b >= B) return;   // (whole warps; no CTA-wide barrier below)
  const int nf = n * rf;
  double* uc = fdsm + (size_t)(threadIdx.x >> 5) * fd_warp_doubles(n, rf);
  double* uf = uc + (n + 1);
  double* jfa = uf + (nf + 1);
  double* rs = jfa + (nf + 1);
  double* va = rs + (nf + 1);
  // ---- forwardSolve (sequential; all lanes carry the same value)
  double u = u0[b];
  if (lane == 0) uc[0] = u;
  for (int m = 1; m <= n; ++m) {
    const double dt = t.dtn[m - 1];
    u = (ode == FD_ODE_LINEAR) ? __dmul_rn(__dadd_rn(1.0, dt), u) : __dadd_rn(u, __dmul_rn(sin(u), dt));
    if (lane == 0) uc[m] = u;
  }
  __syncwarp();
  if (u_out)
    for (int m = lane; m <= n; m += 32) u_out[(size_t)b * (n + 1) + m] = uc[m];
  // ---- fine nodes: interpolated state, Jacobian factor, residual  res[i+1] = u_f[i+1] - fwdUpdate(u_f[i])
  for (int i = lane; i <= nf; i += 32) uf[i] = fd_interp(uc, 1, t, i, n);
  __syncwarp();
  for (int i = lane; i < nf; i += 32) {
    const double ufi = uf[i], dt = t.dtf[i];
    double jf, upd;
    if (ode == FD_ODE_LINEAR) {
      jf = __dadd_rn(1.0, dt);
      upd = __dmul_rn(__dadd_rn(1.0, dt), ufi);
    } else {
      jf = __dadd_rn(1.0, __dmul_rn(cos(ufi), dt));
      upd = __dadd_rn(ufi, __dmul_rn(sin(ufi), dt));
    }
    jfa[i] = jf;
    rs[i + 1] = __dadd_rn(uf[i + 1], -upd);
  }
  // ---- adjoint recurrence (sequential, two flops per step; all lanes)
  double v = 0.0;
  if (lane == 0) va[nf] = 0.0;
  __syncwarp();
  for (int i = nf - 1; i >= 0; --i) {
    const double dt = t.dtf[i];
    double k;
    if (func == FD_FUNC_INT_U2) k = __dmul_rn(__dmul_rn(2.0, uf[i]), dt);
    else if (func == FD_FUNC_INT_U) k = dt;
    else k = (i == nf - 1) ? 1.0 : 0.0;
    v = __dadd_rn(k, __dmul_rn(jfa[i], v));
    if (lane == 0) va[i] = v;
  }
  __syncwarp();
  // ---- err[i] = res[i] * v[i]; |err| stays in rs for the windows
  for (int i = lane; i <= nf; i += 32) {
    const double e = (i == 0) ? __dmul_rn(0.0, va[0]) : __dmul_rn(rs[i], va[i]);
    if (err_out) err_out[(size_t)b * (nf + 1) + i] = e;
    if (v_out) v_out[(size_t)b * (nf + 1) + i] = va[i];
    rs[i] = fabs(e);
  }
  __syncwarp();
  // ---- windows (entries 2 + r rf + q, q = 0..rf-2, summed in ascending order), argmax with the lowest index
  double best = -1.0;
  int best_idx = 0x7fffffff;
  for (int r = lane; r < n; r += 32) {
    double sum = rs[2 + r * rf];
    for (int q = 1; q < rf - 1; ++q) sum = __dadd_rn(sum, rs[2 + r * rf + q]);
    if (steps_out) steps_out[(size_t)b * n + r] = sum;
    if (sum > best) {   // ascending r: strict keeps the lowest index
      best = sum;
      best_idx = r;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    if (ob > best || (ob == best && oi < best_idx)) {
      best = ob;
      best_idx = oi;
    }
  }
  if (idx_out && lane == 0) idx_out[b] = (best_idx == 0x7fffffff) ? 0 : best_idx;
}

// launches the warp form when the batch is small and a warp's arrays fit; else one thread per trajectory
static bool fd_uses_warp(const dgadj_handle* h, long long B, int n, int rf) {
  const bool warp = h->tune_block == 32 || (h->tune_block != 1 && B <= 16384);
  return warp && fd_warp_doubles(n, rf) * sizeof(double) <= 48 * 1024;
}
static cudaError_t fd_launch_awr(dgadj_handle* h, long long B, int n, int rf, int ode, int func, const FdTables& t,
                                 const double* u0, double* uc, double* u_out, double* v_out, double* err_out,
                                 double* steps_out, int* idx_out, cudaStream_t st) {
  const size_t per_warp = fd_warp_doubles(n, rf) * sizeof(double);
  if (fd_uses_warp(h, B, n, rf)) {
    int wpc = (int)((48 * 1024) / per_warp);
    wpc = wpc > 4 ? 4 : wpc;
    fd_awr_warp_kernel<<<(unsigned)((B + wpc - 1) / wpc), wpc * 32, wpc * per_warp, st>>>(B, n, rf, ode, func, t, u0, u_out, v_out,
                                                                                        err_out, steps_out, idx_out);
  } else {
    const int block = 128;
    fd_awr_kernel<<<(unsigned)((B + block - 1) / block), block, 0, st>>>(B, n, rf, ode, func, t, u0, uc, u_out, v_out, err_out,
                                                                        steps_out, idx_out);
  }
  return cudaGetLastError();
}

}  // namespace dgadj

extern "C" int dgadj_fd_awr(dgadj_handle* h, int64_t B, int32_t n, int32_t ref_factor, int32_t ode,
                            int32_t functional, const double* dt_host, const double* u0_dev, double* u_dev,
                            double* v_dev, double* err_fine_dev, double* err_steps_dev, int32_t* ref_idx_dev,
                            void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || n <= 0 || !dt_host || !u0_dev) return fail(h, DGADJ_ERR_INVALID, "bad fd_awr arguments");
  if (ref_factor < 3 || ref_factor > FD_MAX_REF)
    return fail(h, DGADJ_ERR_UNSUPPORTED, "ref_factor must be in [3, %d] (the reference requires > 2)", FD_MAX_REF);
  if (ode != FD_ODE_SIN && ode != FD_ODE_LINEAR) return fail(h, DGADJ_ERR_INVALID, "unknown ode %d", ode);
  if (functional < FD_FUNC_INT_U || functional > FD_FUNC_INT_U2) return fail(h, DGADJ_ERR_INVALID, "unknown functional %d", functional);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nf = n * ref_factor;
  // ---- the tables of this mesh: kept on the device (their own buffer) for as long as the same steps come back -- a
  // repeated call is a kernel launch, no table build, no upload, no synchronisation
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_dx = 0, o_den = o_dx + al(sizeof(double) * (nf + 1)), o_dtf = o_den + al(sizeof(double) * n),
               o_dtn = o_dtf + al(sizeof(double) * nf), o_jc = o_dtn + al(sizeof(double) * n),
               o_ex = o_jc + al(sizeof(int) * (nf + 1)), tbl_bytes = o_ex + al(nf + 1);
  const bool hit = h->fd_tbl && h->fd_tbl_n == n && h->fd_tbl_rf == ref_factor &&
                   memcmp(h->fd_tbl_dt, dt_host, sizeof(double) * n) == 0;
  if (!hit) {
    // host tables: refineAll (:16-21), cumulative times, np.interp's interval search (:24-31)
    std::vector<double> dtf(nf), tc(n + 1), tf(nf + 1), dx(nf + 1), den(n);
    std::vector<int> jc(nf + 1);
    std::vector<unsigned char> exact(nf + 1);
    for (int j = 0; j < n; ++j)
      for (int f = 0; f < ref_factor; ++f) dtf[j * ref_factor + f] = dt_host[j] / ref_factor;
    tc[0] = 0.0;
    for (int j = 0; j < n; ++j) tc[j + 1] = tc[j] + dt_host[j];   // np.cumsum: sequential
    tf[0] = 0.0;
    for (int i = 0; i < nf; ++i) tf[i + 1] = tf[i] + dtf[i];
    for (int j = 0; j < n; ++j) den[j] = tc[j + 1] - tc[j];
    for (int i = 0; i <= nf; ++i) {
      const double x = tf[i];
      int j;
      if (x >= tc[n]) {
        j = n;  // x beyond / at the last node: np.interp returns fp[-1] (right fill = fp[-1])
      } else {
        j = 0;  // largest j with tc[j] <= x
        int lo = 0, hi = n;
        while (lo < hi) {
          const int mid = (lo + hi + 1) / 2;
          if (tc[mid] <= x) lo = mid; else hi = mid - 1;
        }
        j = lo;
      }
      jc[i] = j;
      dx[i] = x - tc[j];
      exact[i] = (j >= n || tc[j] == x) ? 1 : 0;
    }
    h->fd_tbl_n = 0;   // (no valid entry until everything below is queued)
    if (tbl_bytes > h->fd_tbl_bytes || (size_t)n > h->fd_tbl_dt_cap) {
      CUDA_TRY(h, cudaDeviceSynchronize());
      cudaFree(h->fd_tbl);
      free(h->fd_tbl_dt);
      h->fd_tbl = nullptr;
      h->fd_tbl_dt = nullptr;
      h->fd_tbl_bytes = h->fd_tbl_dt_cap = 0;
      h->fd_tbl_dt = (double*)malloc(sizeof(double) * n);
      if (!h->fd_tbl_dt) return fail(h, DGADJ_ERR_NOMEM, "host copy of the FD steps");
      h->fd_tbl_dt_cap = n;
      CUDA_TRY(h, cudaMalloc(&h->fd_tbl, tbl_bytes));
      h->fd_tbl_bytes = tbl_bytes;
    }
    unsigned char* tb = (unsigned char*)h->fd_tbl;
    // (stream-ordered after the kernels that may still read the old tables; the pageable sources are staged by the
    //  driver before cudaMemcpyAsync returns)
    CUDA_TRY(h, cudaMemcpyAsync(tb + o_dx, dx.data(), sizeof(double) * (nf + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(tb + o_den, den.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(tb + o_dtf, dtf.data(), sizeof(double) * nf, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(tb + o_dtn, dt_host, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(tb + o_jc, jc.data(), sizeof(int) * (nf + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(tb + o_ex, exact.data(), nf + 1, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));   // the host vectors go out of scope
    memcpy(h->fd_tbl_dt, dt_host, sizeof(double) * n);
    h->fd_tbl_n = n;
    h->fd_tbl_rf = ref_factor;
  }
  const unsigned char* tb = (const unsigned char*)h->fd_tbl;
  FdTables t;
  t.dx = (const double*)(tb + o_dx);
  t.den = (const double*)(tb + o_den);
  t.dtf = (const double*)(tb + o_dtf);
  t.dtn = (const double*)(tb + o_dtn);
  t.jc = (const int*)(tb + o_jc);
  t.exact = tb + o_ex;
  // ---- coarse states [n+1][B] of the thread-per-trajectory form
  double* uc = nullptr;
  if (!fd_uses_warp(h, B, n, ref_factor)) {
    const size_t need = sizeof(double) * (size_t)(n + 1) * (size_t)B;
    if (need > h->fd_bytes) {
      CUDA_TRY(h, cudaDeviceSynchronize());
      cudaFree(h->fd_scratch);
      h->fd_scratch = nullptr;
      h->fd_bytes = 0;
      CUDA_TRY(h, cudaMalloc(&h->fd_scratch, need));
      h->fd_bytes = need;
    }
    uc = (double*)h->fd_scratch;
  }
  CUDA_TRY(h, fd_launch_awr(h, B, n, ref_factor, ode, functional, t, u0_dev, uc, u_dev, v_dev, err_fine_dev, err_steps_dev,
                            ref_idx_dev, st));
  h->launches++;
  return DGADJ_OK;
}

// =======================================================================================================
// Device-resident adaptive loop of python/Main_finite_difference.py:263-343, batched
// with the batch-mean rule of python/Main_variable_params.py:340-341: per iteration the interpolation
// tables of the current mesh (refineAll :16-21, np.interp's interval search :24-31 -- the host routine of
// dgadj_fd_awr above, restated for one device thread: the meshes have at most a few hundred nodes),
// forwardSolve / adjSolve / errEst / window sums (:266-277), batch mean, argmax step and midpoint
// insertion (:336-341).  One call enqueues every iteration; nothing is read back in between.
// =======================================================================================================
namespace dgadj {

struct FdTablesRW {
  int* jc;
  double* dx;
  double* den;
  double* dtf;
  double* dtn;
  unsigned char* exact;
};

__global__ void fd_tables_kernel(int n, int rf, const double* __restrict__ times, FdTablesRW t, double* __restrict__ tc,
                                 double* __restrict__ tf) {
  // thread 0: the sequential sums (np.diff, np.cumsum); then all threads: the interval search per fine node
  const int nf = n * rf;
  if (threadIdx.x == 0) {
    for (int j = 0; j < n; ++j) t.dtn[j] = times[j + 1] - times[j];                 // dt_n = np.diff(times)
    for (int j = 0; j < n; ++j)
      for (int f = 0; f < rf; ++f) t.dtf[j * rf + f] = t.dtn[j] / rf;              // refineAll
    tc[0] = 0.0;
    for (int j = 0; j < n; ++j) tc[j + 1] = tc[j] + t.dtn[j];
    tf[0] = 0.0;
    for (int i = 0; i < nf; ++i) tf[i + 1] = tf[i] + t.dtf[i];
    for (int j = 0; j < n; ++j) t.den[j] = tc[j + 1] - tc[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= nf; i += blockDim.x) {
    const double x = tf[i];
    int j;
    if (x >= tc[n]) {
      j = n;
    } else {
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (tc[mid] <= x) lo = mid; else hi = mid - 1;
      }
      j = lo;
    }
    t.jc[i] = j;
    t.dx[i] = x - tc[j];
    t.exact[i] = (j >= n || tc[j] == x) ? 1 : 0;
  }
}

// batch mean of err_steps[b][r] per coarse step (fixed-order tree)
__global__ void fd_loop_mean_kernel(long long B, int n, const double* __restrict__ steps, double* __restrict__ mean) {
  __shared__ double sm[256];
  const int r = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  for (long long b = tid; b < B; b += 256) acc += steps[(size_t)b * n + r];
  sm[tid] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) sm[tid] += sm[tid + o];
    __syncthreads();
  }
  if (tid == 0) mean[r] = sm[0] / (double)B;
}

__global__ void fd_refine_kernel(int n, const double* __restrict__ ind, const double* __restrict__ times,
                                 double* __restrict__ times_next, int* __restrict__ ref_idx, double* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int best = 0;
  double bv = ind[0], tot = 0.0;
  for (int k = 0; k < n; ++k) {       // np.argmax: first maximum
    const double v = ind[k];
    tot += v;
    if (v > bv) {
      bv = v;
      best = k;
    }
  }
  *ref_idx = best;
  if (total) *total = tot;
  if (times_next) {
    for (int j = 0; j <= best; ++j) times_next[j] = times[j];
    times_next[best + 1] = (times[best] + times[best + 1]) / 2.0;
    for (int j = best + 1; j <= n; ++j) times_next[j + 1] = times[j];
  }
}

}  // namespace dgadj

extern "C" int dgadj_fd_adapt_loop(dgadj_handle* h, int64_t B, int32_t iters, int32_t n0, int32_t ref_factor, int32_t ode,
                                   int32_t functional, const double* times0_host, const double* u0_dev,
                                   double* times_hist_dev, double* err_hist_dev, int32_t* ref_idx_dev,
                                   double* total_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || iters < 0 || n0 < 1 || !times0_host || !u0_dev || !times_hist_dev || !err_hist_dev || !ref_idx_dev)
    return fail(h, DGADJ_ERR_INVALID, "bad fd_adapt_loop arguments");
  if (ref_factor < 3 || ref_factor > FD_MAX_REF)
    return fail(h, DGADJ_ERR_UNSUPPORTED, "ref_factor must be in [3, %d] (the reference requires > 2)", FD_MAX_REF);
  if (ode != FD_ODE_SIN && ode != FD_ODE_LINEAR) return fail(h, DGADJ_ERR_INVALID, "unknown ode %d", ode);
  if (functional < FD_FUNC_INT_U || functional > FD_FUNC_INT_U2) return fail(h, DGADJ_ERR_INVALID, "unknown functional %d", functional);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nmax = n0 + iters, nfmax = nmax * ref_factor, W = nmax + 2;
  // scratch: tables (sized for the last mesh) | tc | tf | mean | coarse states [nmax+1][B] | err_steps [B][nmax]
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_dx = 0, o_den = o_dx + al(sizeof(double) * (nfmax + 1)), o_dtf = o_den + al(sizeof(double) * nmax),
               o_dtn = o_dtf + al(sizeof(double) * nfmax), o_tc = o_dtn + al(sizeof(double) * nmax),
               o_tf = o_tc + al(sizeof(double) * (nmax + 1)), o_jc = o_tf + al(sizeof(double) * (nfmax + 1)),
               o_ex = o_jc + al(sizeof(int) * (nfmax + 1)), o_uc = o_ex + al(nfmax + 1),
               o_steps = o_uc + al(sizeof(double) * (size_t)(nmax + 1) * (size_t)B),
               need = o_steps + al(sizeof(double) * (size_t)nmax * (size_t)B);
  if (need > h->fd_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->fd_scratch);
    h->fd_scratch = nullptr;
    h->fd_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->fd_scratch, need));
    h->fd_bytes = need;
  }
  unsigned char* base = (unsigned char*)h->fd_scratch;
  FdTablesRW tw = {(int*)(base + o_jc), (double*)(base + o_dx), (double*)(base + o_den), (double*)(base + o_dtf),
                   (double*)(base + o_dtn), base + o_ex};
  FdTables t = {tw.jc, tw.dx, tw.den, tw.dtf, tw.dtn, tw.exact};
  double* tc = (double*)(base + o_tc);
  double* tf = (double*)(base + o_tf);
  double* uc = (double*)(base + o_uc);
  double* steps = (double*)(base + o_steps);
  CUDA_TRY(h, cudaMemcpyAsync(times_hist_dev, times0_host, (size_t)(n0 + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  for (int it = 0; it <= iters; ++it) {
    const int n = n0 + it;
    const double* times = times_hist_dev + (size_t)it * W;
    fd_tables_kernel<<<1, 128, 0, st>>>(n, ref_factor, times, tw, tc, tf);
    CUDA_TRY(h, fd_launch_awr(h, B, n, ref_factor, ode, functional, t, u0_dev, uc, nullptr, nullptr, nullptr, steps, nullptr, st));
    double* mrow = err_hist_dev + (size_t)it * nmax;
    fd_loop_mean_kernel<<<n, 256, 0, st>>>(B, n, steps, mrow);
    fd_refine_kernel<<<1, 32, 0, st>>>(n, mrow, times, it == iters ? nullptr : times_hist_dev + (size_t)(it + 1) * W,
                                       ref_idx_dev + it, total_dev ? total_dev + it : nullptr);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 4;
  }
  return DGADJ_OK;
}

// =======================================================================================================
// Per-trajectory meshes: the reference's own single-trajectory loop (python/Main_finite_difference.py:263-343)
// for every trajectory of a batch at once -- each thread runs the WHOLE loop on its own mesh: forwardSolve,
// the interpolation of interpU (np.interp's interval search on its own cumulative times), adjSolve's
// recurrence, errEst, the window sums, np.argmax and the midpoint insertion.  Same unfused arithmetic as
// fd_awr_kernel; scratch arrays are [index][B] (coalesced across the warp).
// =======================================================================================================
namespace dgadj {

__global__ void fd_adapt_pt_kernel(long long B, int iters, int n0, int rf, int ode, int func,
                                   const double* __restrict__ times0, const double* __restrict__ u0,
                                   double* __restrict__ times /*[nmax+1][B]*/, double* __restrict__ tc /*[nmax+1][B]*/,
                                   double* __restrict__ tf /*[nfmax+1][B]*/, double* __restrict__ uc /*[nmax+1][B]*/,
                                   int* __restrict__ ref_hist /*[B][iters+1]*/, double* __restrict__ tot_hist /*[B][iters+1]*/,
                                   double* __restrict__ times_out /*[B][nmax+1]*/) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const size_t sB = (size_t)B;
  double* t = times + b;
  double* c = tc + b;
  double* f = tf + b;
  double* ub = uc + b;
  for (int j = 0; j <= n0; ++j) t[(size_t)j * sB] = times0[j];
  for (int it = 0; it <= iters; ++it) {
    const int n = n0 + it, nf = n * rf;
    // cumulative times of the coarse and the refined mesh (np.diff, refineAll, np.cumsum: sequential sums)
    c[0] = 0.0;
    f[0] = 0.0;
    for (int j = 0; j < n; ++j) {
      const double dt = __dadd_rn(t[(size_t)(j + 1) * sB], -t[(size_t)j * sB]);
      c[(size_t)(j + 1) * sB] = __dadd_rn(c[(size_t)j * sB], dt);
      const double dtf = __ddiv_rn(dt, (double)rf);
      for (int q = 0; q < rf; ++q) f[(size_t)(j * rf + q + 1) * sB] = __dadd_rn(f[(size_t)(j * rf + q) * sB], dtf);
    }
    // forwardSolve
    double u = u0[b];
    ub[0] = u;
    for (int m = 1; m <= n; ++m) {
      const double dt = __dadd_rn(t[(size_t)m * sB], -t[(size_t)(m - 1) * sB]);
      u = (ode == FD_ODE_LINEAR) ? __dmul_rn(__dadd_rn(1.0, dt), u) : __dadd_rn(u, __dmul_rn(sin(u), dt));
      ub[(size_t)m * sB] = u;
    }
    const double tcn = c[(size_t)n * sB];
    auto interp = [&](int i) -> double {   // np.interp(t_fine[i], t_coarse, u)
      const double x = f[(size_t)i * sB];
      int j;
      if (x >= tcn) {
        j = n;
      } else {
        int lo = 0, hi = n;
        while (lo < hi) {
          const int mid = (lo + hi + 1) / 2;
          if (c[(size_t)mid * sB] <= x) lo = mid; else hi = mid - 1;
        }
        j = lo;
      }
      const double tj = c[(size_t)j * sB];
      const double uj = ub[(size_t)j * sB];
      if (j >= n || tj == x) return uj;
      const double den = __dadd_rn(c[(size_t)(j + 1) * sB], -tj);
      const double slope = __ddiv_rn(__dadd_rn(ub[(size_t)(j + 1) * sB], -uj), den);
      return __dadd_rn(__dmul_rn(slope, __dadd_rn(x, -tj)), uj);
    };
    // one backward sweep: adjoint recurrence, residual, window sums, argmax (as fd_awr_kernel)
    double v_next = 0.0, uf_next = interp(nf), win[FD_MAX_REF], best = -1.0, tot = 0.0;
    int best_idx = 0;
    for (int i = nf - 1; i >= 0; --i) {
      const double uf = interp(i);
      const int jc = i / rf;
      const double dtc = __dadd_rn(t[(size_t)(jc + 1) * sB], -t[(size_t)jc * sB]);
      const double dt = __ddiv_rn(dtc, (double)rf);
      double k, jf, upd;
      if (ode == FD_ODE_LINEAR) {
        jf = __dadd_rn(1.0, dt);
        upd = __dmul_rn(__dadd_rn(1.0, dt), uf);
      } else {
        jf = __dadd_rn(1.0, __dmul_rn(cos(uf), dt));
        upd = __dadd_rn(uf, __dmul_rn(sin(uf), dt));
      }
      if (func == FD_FUNC_INT_U2) k = __dmul_rn(__dmul_rn(2.0, uf), dt);
      else if (func == FD_FUNC_INT_U) k = dt;
      else k = (i == nf - 1) ? 1.0 : 0.0;
      const double err = __dmul_rn(__dadd_rn(uf_next, -upd), v_next);
      const int pos = i + 1 - 2;
      if (pos >= 0) {
        const int r = pos / rf, q = pos - r * rf;
        if (q < rf - 1) {
          win[q] = fabs(err);
          if (q == 0) {
            double s = win[0];
            for (int qq = 1; qq < rf - 1; ++qq) s = __dadd_rn(s, win[qq]);
            tot += s;
            if (s >= best) {
              best = s;
              best_idx = r;
            }
          }
        }
      }
      v_next = __dadd_rn(k, __dmul_rn(jf, v_next));
      uf_next = uf;
    }
    ref_hist[(size_t)b * (iters + 1) + it] = best_idx;
    if (tot_hist) tot_hist[(size_t)b * (iters + 1) + it] = tot;
    if (it < iters) {   // midpoint insertion (:338-341)
      const double mid = __ddiv_rn(__dadd_rn(t[(size_t)best_idx * sB], t[(size_t)(best_idx + 1) * sB]), 2.0);
      for (int j = n; j > best_idx; --j) t[(size_t)(j + 1) * sB] = t[(size_t)j * sB];
      t[(size_t)(best_idx + 1) * sB] = mid;
    }
  }
  const int nlast = n0 + iters;
  if (times_out)
    for (int j = 0; j <= nlast; ++j) times_out[(size_t)b * (nlast + 1) + j] = t[(size_t)j * sB];
}

}  // namespace dgadj

extern "C" int dgadj_fd_adapt_loop_pt(dgadj_handle* h, int64_t B, int32_t iters, int32_t n0, int32_t ref_factor, int32_t ode,
                                      int32_t functional, const double* times0_host, const double* u0_dev,
                                      double* times_out_dev, int32_t* ref_hist_dev, double* tot_hist_dev, void* stream) {
  if (!h) return DGADJ_ERR_INVALID;
  if (B <= 0 || iters < 0 || n0 < 1 || !times0_host || !u0_dev || !ref_hist_dev)
    return fail(h, DGADJ_ERR_INVALID, "bad fd_adapt_loop_pt arguments");
  if (ref_factor < 3 || ref_factor > FD_MAX_REF)
    return fail(h, DGADJ_ERR_UNSUPPORTED, "ref_factor must be in [3, %d] (the reference requires > 2)", FD_MAX_REF);
  if (ode != FD_ODE_SIN && ode != FD_ODE_LINEAR) return fail(h, DGADJ_ERR_INVALID, "unknown ode %d", ode);
  if (functional < FD_FUNC_INT_U || functional > FD_FUNC_INT_U2) return fail(h, DGADJ_ERR_INVALID, "unknown functional %d", functional);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nmax = (size_t)n0 + iters, nfmax = nmax * ref_factor;
  const size_t n_d = (size_t)(n0 + 1) + (size_t)B * (3 * (nmax + 1) + (nfmax + 1));
  const size_t need = n_d * sizeof(double) + 64;
  if (need > h->fd_bytes) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    cudaFree(h->fd_scratch);
    h->fd_scratch = nullptr;
    h->fd_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->fd_scratch, need));
    h->fd_bytes = need;
  }
  double* d = (double*)h->fd_scratch;
  double* t0 = d;
  double* times = t0 + (n0 + 1);
  double* tc = times + (size_t)B * (nmax + 1);
  double* uc = tc + (size_t)B * (nmax + 1);
  double* tf = uc + (size_t)B * (nmax + 1);
  CUDA_TRY(h, cudaMemcpyAsync(t0, times0_host, (size_t)(n0 + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  const int block = (B <= 16384) ? 32 : 128;
  fd_adapt_pt_kernel<<<(unsigned)((B + block - 1) / block), block, 0, st>>>(B, iters, n0, ref_factor, ode, functional, t0, u0_dev,
                                                                            times, tc, tf, uc, ref_hist_dev, tot_hist_dev, times_out_dev);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return DGADJ_OK;
}
