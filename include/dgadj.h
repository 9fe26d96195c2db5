/* dgadj.h -- C-ABI of libdgadj.so: B200-native (sm_100a) batched 1-D nodal-DG forward march,
 * reverse-time discrete adjoint march and per-element adjoint-weighted error indicator.
 *
 * The reference (wglao/Adjoint-ODE-Adaptivity) has NO native / FFI layer: its hot path is
 * MATLAB + NumPy source.  Each entry point below therefore cites the reference *source
 * routine* it replaces; INTEGRATION.md shows the bindings a maintainer adds on the
 * reference side: ctypes (python/galerkin.py, python/Main_finite_difference.py) and MATLAB
 * loadlibrary (utils/, matlab/).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary.
 *   - every function returns DGADJ_OK (0) or a negative dgadj_status; nothing throws.
 *     dgadj_last_error(h) gives a human-readable message for the last failure on h.
 *   - `*_dev` pointers are CUDA device pointers on the handle's device, caller-owned and
 *     borrowed for the call; work is enqueued on `stream` (a cudaStream_t cast to void*,
 *     NULL = the legacy default stream) and is asynchronous -- the caller synchronises.
 *     Exceptions, which block the host until their constants are on the device (they take per-call
 *     HOST arrays of mesh constants): dgadj_burgers_forward / dgadj_burgers_adjoint, and
 *     dgadj_fwd_adj_windowed (scratch is released on return).  dgadj_tdg_march / _adjoint / _adjoint_rec /
 *     _err_contribution keep the constant blocks of the last four meshes on the device, dgadj_fd_awr the
 *     tables of the last one: a call on a mesh that is still there is a kernel launch only; a new mesh
 *     costs the table build and one pageable upload (the host array may be released on return).  The fused / loop entry points (dgadj_fwd_adj, dgadj_burgers_fwd_adj,
 *     dgadj_*_adapt_loop*) stage their constants with stream-ordered copies and do not block.
 *   - scratch buffers belong to the handle: use a handle from ONE stream at a time (a second call on
 *     another stream may overwrite constants a running kernel still reads).
 *   - `*_host` entry points take host pointers (pinned or pageable), copy H2D in chunks,
 *     run the same kernels overlapped with the copies, copy D2H and synchronise before
 *     returning.
 *   - all matrices are fp64, row-major; fields are [B][Np][K] (node-major inside a
 *     trajectory: element index fastest), the batched form of galerkin.py's (Np, K) arrays
 *     (python/galerkin.py:216).
 *   - a handle is bound to one device and is not thread-safe (one handle per GPU / rank).
 *   - there is NO CPU fallback: dgadj_create fails with DGADJ_ERR_NO_DEVICE unless the
 *     device has compute capability 10.x.
 */
#ifndef DGADJ_H
#define DGADJ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGADJ_VERSION 100 /* 0.1.0 */

typedef struct dgadj_handle dgadj_handle;

typedef enum {
  DGADJ_OK = 0,
  DGADJ_ERR_INVALID = -1,     /* bad argument / shape */
  DGADJ_ERR_NO_DEVICE = -2,   /* no sm_100 device (no CPU fallback exists) */
  DGADJ_ERR_CUDA = -3,        /* a CUDA runtime call failed; see dgadj_last_error */
  DGADJ_ERR_UNSUPPORTED = -4, /* valid request outside the compiled envelope (N, K, ...) */
  DGADJ_ERR_STATE = -5,       /* operators / weights not set before a march */
  DGADJ_ERR_NOMEM = -6
} dgadj_status;

enum { DGADJ_BC_INFLOW = 0,    /* utils/AdvecRHS1D.m:14-16 (inflow Dirichlet + free outflow) */
       DGADJ_BC_PERIODIC = 1 };/* BASELINE config 1/2 text; not in the reference */
enum { DGADJ_INFLOW_ZERO = 0,
       DGADJ_INFLOW_SIN_AT = 1,   /* uin = -sin(a t)    utils/AdvecRHS1D.m:14 */
       DGADJ_INFLOW_SIN_AAT = 2,  /* uin = -sin(a a t)  utils/One_code.mlx (quirk C-2) */
       DGADJ_INFLOW_TABLE = 3 };  /* uin[n*nstages+s] supplied by the caller */
enum { DGADJ_FUNC_LINEAR = 0,  /* J = sum jw o u(T)  (J = int psi u dx; getK 'J=int(u)') */
       DGADJ_FUNC_INT_U2 = 1 };/* J = int u(T)^2 dx  (getK 'J=int(u^2)', Main_finite_difference.py:225) */
enum { DGADJ_SCHEME_LSERK4 = 0,  /* utils/Globals1D.m:20-34 */
       DGADJ_SCHEME_EULER = 1 }; /* matlab/fwd_euler_march.m / forwardSolve semantics */

typedef struct {
  int32_t device;     /* CUDA device ordinal */
  int32_t N;          /* primal polynomial order, Np = N+1, 1 <= N <= 8 (N = 9: forward march only) */
  int32_t K;          /* elements per mesh, 1 <= K <= 1024 */
  int32_t bc;         /* DGADJ_BC_* */
  int32_t inflow;     /* DGADJ_INFLOW_* (bc = inflow only) */
  int32_t functional; /* DGADJ_FUNC_* */
  int32_t scheme;     /* DGADJ_SCHEME_* */
  int32_t reserved;
  double alpha;       /* flux parameter of AdvecRHS1D.m:9-11: 1 = central (reference), 0 = upwind */
} dgadj_config;

int dgadj_version(void);

/* Replaces: the `Globals1D` state shared by every reference routine (utils/Globals1D.m:3-17). */
int dgadj_create(const dgadj_config* cfg, dgadj_handle** out);
void dgadj_destroy(dgadj_handle* h);
const char* dgadj_last_error(const dgadj_handle* h);

/* Operators of the primal space, as produced by StartUp1D (utils/StartUp1D.m:9-33) /
 * BaseGalerkin1D.startUp1D (python/galerkin.py:199-237).  Host pointers, copied.
 *   Dr[Np*Np], LIFT[Np*2], V[Np*Np] (Vandermonde1D.m: V(i,j) = P~_{j-1}(r_i), orthonormal
 *   Legendre), rx[Np*K], Fscale[2*K].
 * The march runs in the modal basis of V: V^-1 Dr V must be the (parity-sparse, strictly upper
 * triangular) Legendre derivative and V^-1 LIFT = V^T E -- true for every operator set
 * StartUp1D produces; anything else is rejected with DGADJ_ERR_UNSUPPORTED.               */
int dgadj_set_operators(dgadj_handle* h, int Np, int K, const double* Dr, const double* LIFT,
                        const double* V, const double* rx, const double* Fscale);

/* (A refinement loop keeps ONE handle: K may be any mesh size up to cfg.K, the capacity the handle was created
 * with; setting the operators of a mesh of another size voids the enriched operators and functional weights of
 * the old one, which are then set again.)                                                                  */

/* Operators of the enriched space (order N+1; matlab/MAIN.m:34 solves the adjoint at Ns+1)
 * plus the nodal prolongation P[NpF*Np] = V_{N+1}(:,1:Np) inv(V_N).                      */
int dgadj_set_enriched(dgadj_handle* h, int NpF, const double* DrF, const double* LIFTF,
                       const double* VF, const double* rxF, const double* FscaleF,
                       const double* P);

/* Weights of a linear terminal functional J = sum_{i,k} jw[i,k] u[i,k] in both spaces
 * (jw_c[Np*K], jw_f[NpF*K]); used when cfg.functional == DGADJ_FUNC_LINEAR.              */
int dgadj_set_functional_weights(dgadj_handle* h, const double* jw_c, const double* jw_f);

/* hp (per-element orders Ns(k), matlab/MAIN.m:21,141, applied to the DG-in-space march):
 * nodes_per_element_host[K], each in [2, Np]; NULL returns to the uniform order.  Fields keep the handle's
 * [B][Np][K] layout on its LGL nodes (an element of lower order holds its polynomial's values there; on input it
 * is L2-projected onto its own space); the kernels hold the modes beyond an element's space at zero -- in the
 * orthonormal modal basis a lower order IS the truncated space, so this is the hp scheme, not an approximation
 * of it.  The enriched space of the adjoint / indicator has one order more per element.  Built for
 * dgadj_forward (without checkpoints / history), dgadj_fwd_adj(_host), dgadj_fwd_adj_windowed and
 * dgadj_ic_indicator (the coarse data enter through the L2 projection onto each element's space);
 * DGADJ_ERR_UNSUPPORTED elsewhere (two-call adjoint, dgadj_rhs).                                       */
int dgadj_set_element_orders(dgadj_handle* h, const int32_t* nodes_per_element_host);

/* Caller-supplied inflow values uin[S*nstages] (cfg.inflow == DGADJ_INFLOW_TABLE).       */
int dgadj_set_inflow_table(dgadj_handle* h, int n, const double* uin);

/* Launch-shape overrides for tuning sweeps (0 = automatic): elements per thread (1, 2 or 4),
 * target threads per CTA, CTAs in the persistent grid.  For the time-DG march and the FD path
 * block_threads = 1 / 32 forces the thread-per-trajectory / warp-per-trajectory kernel (automatic:
 * a warp per trajectory up to 16 384 trajectories).                                        */
int dgadj_set_tuning(dgadj_handle* h, int32_t elems_per_thread, int32_t block_threads,
                     int32_t grid_ctas);

/* Per-trajectory advection speed / time step: if a_dev (dt_dev) is NULL the scalar is used. */
typedef struct {
  int64_t B;            /* trajectories in this call */
  int32_t S;            /* time steps */
  int32_t reserved;
  double t0;            /* start time (reference: time = 0) */
  double a;             /* advection speed (AdvecRHS1D.m:1 argument `a`) */
  double dt;            /* time step (One_code.mlx CFL rule, computed by the host) */
  const double* a_dev;  /* [B] or NULL */
  const double* dt_dev; /* [B] or NULL */
} dgadj_march_args;

/* Forward march.  Replaces: the `for tstep ... for INTRK=1:5` LSERK4 loop of
 * utils/One_code.mlx with utils/AdvecRHS1D.m:8-19 inlined (and, with DGADJ_SCHEME_EULER,
 * the explicit-Euler march matlab/fwd_euler_march.m / Main_finite_difference.py:34-51).
 *   u0_dev[B][Np][K] -> uT_dev[B][Np][K]; hist_dev[B][S+1][Np][K] (or NULL) receives every
 *   step's state (the reference hands the primal to the adjoint in memory, adj_march.m:4);
 *   ckpt_dev (or NULL) receives the opaque forward checkpoints the adjoint consumes
 *   (dgadj_ckpt_bytes(h, B, S) bytes).                                                   */
int dgadj_forward(dgadj_handle* h, const dgadj_march_args* args, const double* u0_dev,
                  double* uT_dev, double* hist_dev, void* ckpt_dev, void* stream);
int64_t dgadj_ckpt_bytes(dgadj_handle* h, int64_t B, int32_t S);

/* Reverse-time adjoint march + per-element indicator.  Replaces: matlab/adj_march.m:67-118
 * (adjoint one order higher, err(k) = v_k' * residual) and errEst
 * (python/Main_finite_difference.py:79-94), for the DG-in-space march.
 *   uT_dev[B][Np][K] (terminal primal, for J and the terminal condition), ckpt_dev from
 *   dgadj_forward -> J_dev[B], lam0_dev[B][NpF][K] (dJ/du0 in the enriched space, or NULL),
 *   eta_dev[B][K] (signed; consumers take abs, matlab/MAIN.m:51).                        */
int dgadj_adjoint(dgadj_handle* h, const dgadj_march_args* args, const double* uT_dev,
                  const void* ckpt_dev, double* J_dev, double* lam0_dev, double* eta_dev,
                  void* stream);

/* Fused forward + adjoint + indicator: one persistent CTA marches a group of trajectories
 * forward, then immediately backward, through a per-CTA checkpoint ring owned by the
 * handle (so live checkpoints are #CTAs x S x state, never B x S).  Any output may be NULL. */
int dgadj_fwd_adj(dgadj_handle* h, const dgadj_march_args* args, const double* u0_dev,
                  double* uT_dev, double* J_dev, double* lam0_dev, double* eta_dev,
                  void* stream);

/* The same march for step counts whose residual ring does not fit the device (dgadj_fwd_adj needs
 * #CTAs x S tiles; DGADJ_ERR_NOMEM beyond): two-level checkpointing.  Per chunk of `batch_chunk`
 * trajectories (0 = as many as fit) the coarse march runs once keeping the state at the start of
 * every `window` steps; then, windows in reverse, the window is marched again with its residual
 * checkpoints and swept by the adjoint, which hands its state and the indicator sums to the
 * earlier window.  About 1.3x the work of dgadj_fwd_adj; the results are the same (bit for bit
 * with periodic or time-independent inflow data).  Synchronous (scratch is released on return). */
int dgadj_fwd_adj_windowed(dgadj_handle* h, const dgadj_march_args* args, int32_t window,
                           int64_t batch_chunk, const double* u0_dev, double* uT_dev, double* J_dev,
                           double* lam0_dev, double* eta_dev, void* stream);

/* Same as dgadj_fwd_adj with HOST buffers (chunked H2D, kernels, D2H, sync).
 * a_host / dt_host: [B] or NULL.                                                         */
int dgadj_fwd_adj_host(dgadj_handle* h, const dgadj_march_args* args, const double* a_host,
                       const double* dt_host, const double* u0_host, double* uT_host,
                       double* J_host, double* lam0_host, double* eta_host);
int dgadj_forward_host(dgadj_handle* h, const dgadj_march_args* args, const double* a_host,
                       const double* dt_host, const double* u0_host, double* uT_host,
                       double* hist_host);

/* One evaluation of the semi-discrete right-hand side, rhsu = AdvecRHS1D(u, time, a)
 * (utils/AdvecRHS1D.m:1-20), in the primal (level 0) or enriched (level 1) space:
 *   u_dev[B][Np_level][K] -> rhs_dev[B][Np_level][K]; a_dev[B] or NULL (then `a`).       */
int dgadj_rhs(dgadj_handle* h, int64_t B, int32_t level, const double* u_dev, double t, double a,
              const double* a_dev, double* rhs_dev, void* stream);

/* Initial-data term of the indicator.  The march's eta weighs the residual of the *evolution*:
 * the coarse and the enriched solution both start from P u0.  Given the initial data at the
 * enriched nodes too, u0f_dev[B][NpF][K], and the adjoint at t = 0 (lam0 of dgadj_fwd_adj /
 * dgadj_adjoint), this adds   eta[b][k] += sum_i lam0[b][i][k] ((P u0)[i][k] - u0f[b][i][k]),
 * after which sum_k eta[b][k] = J_f(P u_c(T)) - J_f(enriched march of the true initial data): the
 * whole coarse-vs-enriched difference of the functional (SURVEY App. E.5 conventions).        */
int dgadj_ic_indicator(dgadj_handle* h, int64_t B, const double* u0_dev, const double* u0f_dev,
                       const double* lam0_dev, double* eta_dev, void* stream);

/* Refine flag / ranking.  Replaces: ref_i = find(abs(err)==max(abs(err))) (matlab/MAIN.m:137),
 * np.argmax(err_steps) (Main_finite_difference.py:337) and sort(...,'descend') (MAIN.m:99).
 *   eta_dev[B][K] -> order_dev[B][K] int32 (elements by descending |eta|, ties by lowest
 *   index; or NULL), flags_dev[B][K] uint8 (1 on the topk elements; or NULL).            */
int dgadj_rank(dgadj_handle* h, int64_t B, int32_t K, const double* eta_dev, int32_t topk,
               int32_t* order_dev, uint8_t* flags_dev, void* stream);

/* Shared-mesh refinement on the device (matlab/MAIN.m:137-141; python/Main_finite_difference.py:336-341; the
 * batch rule of python/Main_variable_params.py:340-341 feeds it the batch-reduced indicator): the `topk`
 * elements with the largest |ind_dev[k]| (lowest index on ties) are split at their midpoints.
 *   v_x_dev[K+1] -> [K+topk+1] in place (the array must have room); refined_dev[topk] (or NULL) = the
 *   elements that were split, ascending.                                                                  */
int dgadj_refine_shared(dgadj_handle* h, int32_t K, const double* ind_dev, int32_t topk, double* v_x_dev,
                        int32_t* refined_dev, void* stream);

/* Batch reduction feeding the cross-GPU all-reduce (python/Main_variable_params.py:340,
 * jnp.mean(err_refine, axis=0)): sums_dev[K+4] = { sum_b |eta[b][k]| (k<K), sum|eta|,
 * sum eta^2, max|eta|, sum_b J[b] } in a fixed, grid-independent summation order.        */
int dgadj_reduce_indicators(dgadj_handle* h, int64_t B, int32_t K, const double* eta_dev,
                            const double* J_dev, double* sums_dev, void* stream);

/* The cross-GPU step of that rule when the batch is sharded over ranks (one handle, one GPU, one
 * NCCL rank each): sums_dev[K+4] (this rank's partials, in place) -> the combination over all
 * ranks of nccl_comm (an ncclComm_t of the caller's NCCL): all-gather + sum in rank order, entry
 * K+2 a maximum -- the same bits on every rank and for every reduction topology.  (Each rank's
 * partial is summed over its own shard, so a different NUMBER of ranks changes the summation order
 * of the batch sum at rounding level; dgadj_reduce_indicator_blocks / dgadj_allreduce_indicator_blocks
 * below are the count-independent form.)  Asynchronous on `stream`.  NCCL is resolved at run time. */
int dgadj_allreduce_indicators(dgadj_handle* h, void* nccl_comm, int32_t K, double* sums_dev, void* stream);

/* Count-independent form of the two calls above (hard part 6: "fixed-order reduction
 * for bit-stable ranking").  The batch is reduced in blocks of rows_per_block trajectories:
 *   parts_dev[nblk][K+4], nblk = ceil(B / rows_per_block), row j = the sums of dgadj_reduce_indicators
 *   over trajectories [j rows_per_block, (j+1) rows_per_block).
 * dgadj_allreduce_indicator_blocks all-gathers the rows of every rank (nblk_local each; rank order =
 * global block order for contiguous shards) and adds them in that order on every rank -> sums_dev[K+4].
 * When every shard is a whole number of blocks, the result has the same bits for 1, 2, 4 or 8 GPUs (and
 * for one GPU taking the batch in several calls).  nccl_comm may be NULL: the local rows only.      */
int dgadj_reduce_indicator_blocks(dgadj_handle* h, int64_t B, int32_t K, int64_t rows_per_block,
                                  const double* eta_dev, const double* J_dev, double* parts_dev, void* stream);
int dgadj_allreduce_indicator_blocks(dgadj_handle* h, void* nccl_comm, int32_t K, int32_t nblk_local,
                                     const double* parts_dev, double* sums_dev, void* stream);

/* Finite-difference path of python/Main_finite_difference.py, batched over initial conditions
 * on a shared time mesh: forwardSolve (:34-51) -> adjSolve (:54-76, on the ref_factor-refined
 * mesh) -> errEst (:79-94) -> per-step window sums and argmax (:270-277, :337).
 *   ode: 0 = du/dt = sin(u) (:128-140), 1 = du/dt = u (:110-121);
 *   functional: 0 = J = int u, 1 = J = u_N, 2 = J = int u^2 (getK, :153-227);
 *   dt_host[n] coarse steps (host); u0_dev[B] -> u_dev[B][n+1], v_dev[B][n*rf+1],
 *   err_fine_dev[B][n*rf+1] (signed), err_steps_dev[B][n], ref_idx_dev[B] int32 (0-based
 *   np.argmax of err_steps; the reference refines element ref_idx, :337-341).  Any output
 *   may be NULL.                                                                           */
enum { DGADJ_FD_ODE_SIN = 0, DGADJ_FD_ODE_LINEAR = 1 };
enum { DGADJ_FD_FUNC_INT_U = 0, DGADJ_FD_FUNC_U_N = 1, DGADJ_FD_FUNC_INT_U2 = 2 };
int dgadj_fd_awr(dgadj_handle* h, int64_t B, int32_t n, int32_t ref_factor, int32_t ode,
                 int32_t functional, const double* dt_host, const double* u0_dev, double* u_dev,
                 double* v_dev, double* err_fine_dev, double* err_steps_dev, int32_t* ref_idx_dev,
                 void* stream);

/* DG-in-time ODE march (matlab/dg_march.m:1-80; u' = sin(u) with Newton, or the linear branch
 * u' = u) and its reverse-time DG adjoint with the per-element indicator err(k)
 * (matlab/adj_march.m:1-122), batched over the initial value on a shared mesh of Ks elements.
 * Orders may differ between elements (Ns(k), matlab/MAIN.m:21,141): Np / Np_primal / nq are
 * the mesh maxima and each element's block carries its own node count.  The per-element
 * constant blocks (everything fem_setup.m:1-41 and the polyfit/polyval interpolation produce;
 * layout and padding rules in csrc/dgadj_tdg.cu) are built by the host.
 *   march:   y0_dev[B] -> y_dev[B][Ks][Np], its_dev[B][Ks] (Newton iterations; or NULL)
 *   adjoint: y_dev (primal, Np_primal) -> v_dev[B][Ks][Np_primal+1] (or NULL), err_dev[B][Ks]
 *            (signed; matlab/MAIN.m:51 takes abs); y0_hard = the `y0 = 1` of adj_march.m:9, the
 *            initial value the first element's residual is measured against; y0_dev[B] (or NULL)
 *            overrides it per trajectory -- a batch of initial values needs its own.
 *   adjoint_rec: matlab/adj_rec.m:18-71 (linear branch) -- the adjoint solved at the primal
 *            order and reconstructed to order N+1 through the Radau points of
 *            utils/Globals1D.m:37-42 (N <= 4); v_dev[B][Ks][Np_primal+1] = values at
 *            [Radau points; t_{k+1}], err_dev as above.  */
int dgadj_tdg_march(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np, int32_t nq, int32_t linear,
                    double tol, int32_t maxit, const double* elem_consts_host, const double* y0_dev,
                    double* y_dev, int32_t* its_dev, void* stream);
int dgadj_tdg_adjoint(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np_primal, int32_t nq,
                      int32_t linear, double y0_hard, const double* y0_dev,
                      const double* elem_consts_host, const double* y_dev, double* v_dev,
                      double* err_dev, void* stream);
int dgadj_tdg_adjoint_rec(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np_primal, double y0_hard,
                          const double* y0_dev, const double* elem_consts_host, const double* y_dev,
                          double* v_dev, double* err_dev, void* stream);
/* matlab/err_contribution.m:1-50 (exact-adjoint error contributions of the linear model problem;
 * unused by the reference, MAIN.m:50): err_dev[B][Ks] = cvec_k . y_dev[b][k][:] (+ u(1) - 1 on the
 * first element, :42-43); cvec_host[Ks][Np] = the quadrature of a(t)(phi_j - phi_j')(t) over
 * element k, built by the host.                                                             */
int dgadj_tdg_err_contribution(dgadj_handle* h, int64_t B, int32_t Ks, int32_t Np,
                               const double* cvec_host, const double* y_dev, double* err_dev,
                               void* stream);

/* The adaptive refinement loop of matlab/MAIN.m:29-166 with the mesh on the device, batched
 * over initial values on a shared mesh with the batch-mean indicator (python/Main_variable_params.py:340-341):
 * per iteration dg_march (MAIN.m:32), adj_march at order n+1 with the indicator (MAIN.m:34), mean_b |err|
 * (MAIN.m:51), argmax element (lowest index on ties) and midpoint insertion (MAIN.m:137-141).  ONE call
 * enqueues all iters+1 solves on `stream`; nothing is read back in between (iteration `it` has Ks0 + it
 * elements whatever gets refined).  Uniform order n = Np - 1.  Element blocks (layout of dgadj_tdg_march /
 * dgadj_tdg_adjoint) are affine in the element width h: block_k = T0 + h_k T1, the two templates per kind
 * built by the host.
 *   times_hist_dev[(iters+1)][Ks0+iters+2]: row `it` = the mesh of iteration `it` (Ks0+it+1 entries);
 *   err_hist_dev[(iters+1)][Ks0+iters]:     row `it` = mean_b |err[b][k]|;
 *   ref_idx_dev[iters+1]:                   element refined after iteration `it` (0-based);
 *   stats_dev[(iters+1)][2] (or NULL):      {mean_b y(T), sum_k mean|err|};
 *   istats_dev[(iters+1)][3] (or NULL):     {max Newton iterations, element solves that hit maxit without
 *                                            converging (dg_march.m:69-73 prints those), non-finite indicators}
 *   y_last_dev[B][Ks0+iters][Np] (or NULL): the primal of the last solve.                                    */
typedef struct {
  int64_t B;
  int32_t iters, Ks0, Np, nq_march, nq_adj, linear, maxit;
  int32_t y0_per_trajectory;   /* 1: every trajectory's first-element residual is measured against its own y0;
                                  0: against y0_hard (the `y0 = 1` of adj_march.m:9) */
  double tol, y0_hard;
  const double* times0_host;                               /* [Ks0+1] */
  const double* march_T0_host; const double* march_T1_host;
  const double* adj_T0_host; const double* adj_T1_host;
} dgadj_tdg_loop_args;
int dgadj_tdg_adapt_loop(dgadj_handle* h, const dgadj_tdg_loop_args* args, const double* y0_dev,
                         double* times_hist_dev, double* err_hist_dev, int32_t* ref_idx_dev, double* stats_dev,
                         int32_t* istats_dev, double* y_last_dev, void* stream);

/* The same for the finite-difference loop of python/Main_finite_difference.py:263-343: per iteration the
 * interpolation tables of the current mesh, forwardSolve / adjSolve / errEst / window sums (dgadj_fd_awr),
 * batch mean of err_steps, argmax step (np.argmax, :337) and midpoint insertion (:338-341) -- all on the device.
 *   times_hist_dev[(iters+1)][n0+iters+2], err_hist_dev[(iters+1)][n0+iters], ref_idx_dev[iters+1],
 *   total_dev[iters+1] (or NULL) = sum of the mean step indicators (the loop's `err`, :263).              */
int dgadj_fd_adapt_loop(dgadj_handle* h, int64_t B, int32_t iters, int32_t n0, int32_t ref_factor, int32_t ode,
                        int32_t functional, const double* times0_host, const double* u0_dev, double* times_hist_dev,
                        double* err_hist_dev, int32_t* ref_idx_dev, double* total_dev, void* stream);

/* Per-trajectory meshes (build plan step 8): the reference's single-trajectory loops run for
 * every trajectory of a batch at once, each on ITS OWN mesh (no batch rule) -- matlab/MAIN.m:29-166 and
 * python/Main_finite_difference.py:263-343.  All iterations in one call, meshes on the device.
 *   tdg: times_dev[B][Ks0+iters+2] in (every row the initial mesh) / out (every row that trajectory's final mesh);
 *        ref_hist_dev[B][iters+1] the element each trajectory refined per iteration; tot_hist_dev[B][iters+1] (or
 *        NULL) = sum_k |err_k|; y_last_dev[B][Ks0+iters][Np], its_last_dev[B][Ks0+iters] (or NULL): the last solve.
 *        Uses args->{B, iters, Ks0, Np, nq_*, linear, maxit, tol, templates}; y0 is per trajectory.
 *   fd:  times_out_dev[B][n0+iters+1] (or NULL) the final meshes; ref_hist_dev / tot_hist_dev as above.       */
int dgadj_tdg_adapt_loop_pt(dgadj_handle* h, const dgadj_tdg_loop_args* args, const double* y0_dev, double* times_dev,
                            int32_t* ref_hist_dev, double* tot_hist_dev, double* y_last_dev, int32_t* its_last_dev,
                            void* stream);
int dgadj_fd_adapt_loop_pt(dgadj_handle* h, int64_t B, int32_t iters, int32_t n0, int32_t ref_factor, int32_t ode,
                           int32_t functional, const double* times0_host, const double* u0_dev, double* times_out_dev,
                           int32_t* ref_hist_dev, double* tot_hist_dev, void* stream);

/* Inviscid Burgers forward march (LSERK4) with SlopeLimitN (utils/SlopeLimitN.m:9-32,
 * SlopeLimitLin.m:10-18, minmod.m:6-12) applied to the initial state and after every stage,
 * on the handle's primal operators (dgadj_set_operators) and boundary type (periodic, or
 * zero-jump ends); writes the forward checkpoints the adjoint needs.  The reference has no
 * Burgers right-hand side: it is build-specified (local Lax-Friedrichs).
 *   V_host / invV_host [Np*Np], x_host [Np*K]: StartUp1D arrays the limiter uses;
 *   u0_dev[B][Np][K] -> uT_dev; checkpoints (each may be NULL): hist_dev[B][S+1][Np][K] states,
 *   lim_dev[B][S][K] uint16 (bits 0-4: cell limited after stage s; bits 5+2s, 6+2s: winning
 *   minmod argument), lim0_dev[B][K] uint8 (same code for the pass on the initial state),
 *   amax_dev[B][S][5] int32 (+-(flat index + 1) of max|u| per stage, sign of u there),
 *   maxvel_dev[B][S][5] (max|u| per stage).
 *   limit: 0 none, 1 SlopeLimitN (detect troubled cells, limit those: utils/SlopeLimitN.m:9-32),
 *   2 SlopeLimit1 (limit every cell: utils/SlopeLimit1.m:10-22); tvb_M: the M of the TVB minmod
 *   (utils/minmodB.m:6-11) used for the slope in SlopeLimitLin; 0 = plain minmod, the reference's call. */
int dgadj_burgers_forward(dgadj_handle* h, int64_t B, int32_t S, double dt, const double* dt_dev,
                          int32_t limit, double tvb_M, const double* invV_host, const double* V_host,
                          const double* x_host, const double* u0_dev, double* uT_dev,
                          double* hist_dev, uint16_t* lim_dev, uint8_t* lim0_dev, int32_t* amax_dev,
                          double* maxvel_dev, void* stream);

/* Discrete adjoint of that march for J = sum jw o u(T) (jw_host[Np*K]): the limiter and
 * max|u| are transposed on the branches the forward run recorded (hard
 * part 5).  Consumes all five checkpoint arrays -> lam0_dev[B][Np][K] = dJ/du0, J_dev[B].   */
int dgadj_burgers_adjoint(dgadj_handle* h, int64_t B, int32_t S, double dt, const double* dt_dev,
                          const double* invV_host, const double* V_host, const double* x_host,
                          const double* jw_host, const double* hist_dev, const uint16_t* lim_dev,
                          const uint8_t* lim0_dev, const int32_t* amax_dev, const double* maxvel_dev,
                          double* lam0_dev, double* J_dev, void* stream);

/* BASELINE config 3 as ONE call: the limited Burgers march above, its reverse-time discrete adjoint and the
 * per-element error indicator, fused in one persistent kernel (one CTA takes a trajectory forward and then
 * immediately backward).  Forward states are checkpointed into a per-CTA ring owned by the handle
 * (#CTAs x S x state -- never B x S, so T past shock formation fits at the full batch) and streamed back
 * with bulk-TMA; nothing else is recorded: the adjoint phase takes each step again from its checkpoint
 * with the same stage routine (the same bits, hence the same limiter flags, minmod branches, wave speeds
 * and argmax) and transposes it.
 *   indicator = 0: the adjoint of the march itself.  lam0_dev[B][Np][K] = dJ/du0 (what
 *     dgadj_burgers_adjoint gives), J = sum jw o u(T); eta_dev must be NULL.
 *   indicator = 1 (needs dgadj_set_enriched and the enriched-space arrays below): conventions of
 *     matlab/MAIN.m:34 (adjoint one order higher), matlab/adj_march.m:103-117 (err(k) = v_k' * residual)
 *     and errEst (python/Main_finite_difference.py:79-94), in the nonlinear form of SURVEY App. E.5:
 *       rho^n = P u^{n+1} - Phi_f(P u^n)   (Phi_f: the same limited LSERK4 step at order N+1)
 *       lam_f^n = Phi_f'(P u^n)^T lam_f^{n+1},  lam_f^S = jwF   (frozen branches of those steps)
 *       eta_dev[b][k] = sum_n lam_f^{n+1}_k . rho^n_k   (signed);  lam0_dev[B][Np+1][K] = lam_f^0.
 *   nlim_dev[B][2] (or NULL): limiter activations (cell, stage) of the forward march, and of the steps the
 *     adjoint phase takes again (indicator = 0: the same steps, the same number; 1: the enriched steps);
 *   status_dev[B] (or NULL): per-trajectory status word, bit 0 = a non-finite value in u(T).
 * Any output may be NULL.  Shapes whose stage states exceed shared memory (DGADJ_ERR_UNSUPPORTED) go
 * through dgadj_burgers_forward / dgadj_burgers_adjoint.                                              */
typedef struct {
  int64_t B;              /* trajectories */
  int32_t S;              /* LSERK4 steps */
  int32_t limit;          /* 0 none, 1 SlopeLimitN, 2 SlopeLimit1 (as dgadj_burgers_forward) */
  int32_t indicator;      /* 0 / 1, see above */
  int32_t reserved;
  double dt;              /* time step; dt_dev[B] overrides it per trajectory when not NULL */
  const double* dt_dev;
  double tvb_M;           /* M of utils/minmodB.m:6-11; 0 = plain minmod */
  const double* invV_host; const double* V_host; const double* x_host;      /* primal StartUp1D arrays */
  const double* jw_host;                                                     /* [Np*K] weights of J */
  const double* invVF_host; const double* VF_host; const double* xF_host;   /* enriched space (indicator) */
  const double* jwF_host;                                                    /* [(Np+1)*K] */
} dgadj_burgers_args;
int dgadj_burgers_fwd_adj(dgadj_handle* h, const dgadj_burgers_args* args, const double* u0_dev, double* uT_dev,
                          double* J_dev, double* lam0_dev, double* eta_dev, int32_t* nlim_dev,
                          uint32_t* status_dev, void* stream);
/* Launch shape of that kernel for a batch of B: elements per thread, threads per CTA, CTAs, dynamic shared
 * memory, bytes of the state ring per time step (ring = that x S).                                     */
int dgadj_burgers_plan(dgadj_handle* h, int64_t B, int32_t indicator, int32_t* ept, int32_t* block,
                       int32_t* grid, int64_t* smem_bytes, int64_t* ring_bytes_per_step);

/* Per-trajectory status word of a march (the reference's only failure report is the
 * "not converged" print of matlab/dg_march.m:69-73): status_dev[b] = DGADJ_STATUS_NOT_CONVERGED if any of the
 * trajectory's its_per_trajectory Newton counts exceeds maxit (dg_march stops at maxit + 1 iterations),
 * | DGADJ_STATUS_NON_FINITE if any of its values_per_trajectory doubles (a state, an adjoint, indicators) is
 * NaN or Inf.  Either array may be NULL.  (dgadj_burgers_fwd_adj and the adaptive loops report the same
 * conditions themselves.)                                                                                */
#define DGADJ_STATUS_NOT_CONVERGED 1u
#define DGADJ_STATUS_NON_FINITE 2u
int dgadj_march_status(dgadj_handle* h, int64_t B, int64_t values_per_trajectory, const double* values_dev,
                       int32_t its_per_trajectory, const int32_t* its_dev, int32_t maxit, uint32_t* status_dev,
                       void* stream);

/* Register-only DFMA microbenchmark: sustained fp64 FMA-pipe peak of the handle's device
 * in TFLOP/s (the roofline denominator SURVEY section 8(d) asks to be measured).        */
int dgadj_measure_dfma_peak(dgadj_handle* h, double seconds, double* tflops_out,
                            double* sm_clock_mhz_out);

/* Device properties used by the host to size batches. */
int dgadj_device_info(dgadj_handle* h, int32_t* sm_count, int64_t* total_mem, int32_t* cc_major,
                      int32_t* cc_minor);

/* Number of kernel launches issued through this handle since creation (bench accounting). */
int64_t dgadj_launch_count(const dgadj_handle* h);

/* Launch shape the handle would use for a batch of B trajectories (diagnostics / bench):
 * elements per thread, threads per CTA, trajectories per CTA, CTAs, dynamic smem bytes.   */
int dgadj_plan(dgadj_handle* h, int64_t B, int32_t fused, int32_t* ept, int32_t* block,
               int32_t* tpc, int32_t* grid, int64_t* smem_bytes);

/* Host-only utilities (no device needed; used by the CPU test-suite): the even/odd
 * (symmetric / antisymmetric) blocks of a nodal operator set Dr[Np*Np] / LIFT[Np*2], which the
 * Burgers kernels use for Dr f.  Outputs are [5*5] / [5] arrays (row stride 5); *violation =
 * largest entry that must vanish by symmetry, relative to the largest operator entry.     */
int dgadj_host_eo_operators(int Np, const double* Dr, const double* LIFT, double* DE, double* DO,
                            double* LS, double* LA, double* violation);
/* The modal operators the advection march runs on: Dnz[26] = non-zeros of V^-1 Dr V row by row
 * ((i, j), j = i+1, i+3, ...), p[Np] = P~_i(+1), iV[Np*Np] = V^-1; *violation as above.   */
int dgadj_host_modal_operators(int Np, const double* Dr, const double* LIFT, const double* V,
                               double* Dnz, double* p, double* iV, double* violation);

#ifdef __cplusplus
}
#endif
#endif /* DGADJ_H */
