/*
 * qoc_b200.h -- C ABI of the B200-native piecewise-constant GRAPE evaluator.
 *
 * Drop-in boundary for the exp-based GRAPE path of olof3/QuantumOptimalControl.jl.  The reference has no
 * FFI of its own; the boundary is the set of Julia call signatures below, and every entry point names the
 * reference function (file:line) it replaces.  INTEGRATION.md shows the Julia `ccall` shim a maintainer adds.
 *
 * Conventions (identical to what the reference passes around):
 *   - all matrices are COLUMN-MAJOR; complex numbers are interleaved (re, im) doubles, i.e. Julia ComplexF64,
 *     C `double _Complex`, `cuDoubleComplex`.  A "c128 d x d" argument is therefore `const double*` of
 *     length 2*d*d.
 *   - A0 and A[j] are ALREADY multiplied by dt ("propagation with unity time step",
 *     src/gradient_computations.jl:1; callers pass A0dt.., examples/ipopt_callbacks_exp.jl:16).
 *   - u is nc x Nt (x batch) Float64, column-major: u[j + nc*(k + Nt*b)].
 *   - the caller owns every host buffer; the handle owns all device memory.
 *   - a handle is not thread-safe (like the reference's `cache`); calls are synchronous on return unless the
 *     name ends in _device (those enqueue on the given stream and return).
 *   - no exceptions cross the ABI: every call returns a qoc_status; qoc_last_error() gives the detail text.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device of compute capability 10.x.
 */
#ifndef QOC_B200_H
#define QOC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qoc_handle qoc_handle;

typedef enum {
  QOC_OK = 0,
  QOC_ERR_INVALID = 1,      /* bad argument (NULL pointer, non-positive size, unknown enum)                      */
  QOC_ERR_DIMENSION = 2,    /* "A0 and x0 have incompatiable dimensions", src/gradient_computations.jl:84-87    */
  QOC_ERR_STALE_CACHE = 3,  /* "Cache data from other control signal u", src/gradient_computations.jl:37-39     */
  QOC_ERR_UNSUPPORTED = 4,  /* size / mode not covered by the kernels that exist (never a silent fallback)      */
  QOC_ERR_CUDA = 5,         /* CUDA runtime failure; text in qoc_last_error                                      */
  QOC_ERR_NO_DEVICE = 6,    /* no sm_100 device: the product path fails loudly                                   */
  QOC_ERR_NOT_FINITE = 7,   /* J came out NaN/Inf                                                                */
  QOC_ERR_SINGULAR = 8      /* Pade denominator numerically singular (pivot == 0)                               */
} qoc_status;

/* dUkdp_order of grape_sensitivity (src/gradient_computations.jl:35): 1..4 = the reference's truncated
 * Taylor series of expm_jacobian! (:177-213); QOC_ORDER_FRECHET = exact Frechet derivative through the
 * block-triangular augmented matrix [[X,A_j],[0,X]] (north-star mode; not available in the reference). */
#define QOC_ORDER_FRECHET 0

/* built-in terminal costs (so that J and the terminal costate are formed on the device) */
typedef enum {
  QOC_COST_INFIDELITY = 0, /* J = 1 - |tr(T'x)|^2/n^2, dJ_dx = (-2 Omega/n^2) T   src/penalty_fcns.jl:15-24         */
  QOC_COST_ABS_TRACE = 1,  /* J = 1 - |tr(T'x)|,       dJ_dx = -(Omega/|Omega|) T  test/test_gradient_computation.jl:24-25 */
  QOC_COST_NONE = 2,       /* no built-in cost: caller evaluates Jfinal / dJfinal_dx (arbitrary closures) on the
                              host from x_final and passes lambda_final to qoc_gradient                         */
  QOC_COST_ZCAL = 3        /* setup_infidelity_zcalibrated (src/penalty_fcns.jl:27-42): m = diag(T'x), four columns,
                              J = 1 - F^2/16 with F = abs_sum_phase_calibrated(m) maximised over the virtual-Z phase by
                              the reference's golden-section search (src/fidelities.jl:81-137, tolerance 1e-9),
                              dJ_dx = (-2F/16) T Diagonal(dF_dm) from its rrule (:48-56).  Evaluated per pulse on the
                              device: x_final never crosses the bus.  m != 4 -> QOC_ERR_DIMENSION (:28-30).   */
} qoc_cost;

typedef struct {
  int32_t d;        /* Hilbert-space dimension: A0, A[j] are d x d                                              */
  int32_t m;        /* number of propagated state columns: x0 is d x m.  m <= 64; more than 8 columns (the reference's
                     * full-propagator use, m = d, src/penalty_fcns.jl:14) are swept in chunks that share one K1 pass     */
  int32_t nc;       /* number of controls: length(A)                                                            */
  int32_t nt;       /* number of time slices: size(u, 2)                                                        */
  int32_t batch;    /* number of independent pulses evaluated per call (1 = the reference's case)               */
  int32_t order;    /* 1..4 or QOC_ORDER_FRECHET                                                                */
  int32_t cost;     /* qoc_cost                                                                                 */
  int32_t n;        /* normalisation n of setup_infidelity (src/penalty_fcns.jl:15); <= 0 means m               */
  int32_t device;   /* CUDA device ordinal                                                                      */
  /* running state penalty L(x) = mu * sum |x[pen_rows, pen_cols]|^2 summed over all Nt+1 states
   * (src/penalty_fcns.jl:1-11, examples/ipopt_callbacks_exp.jl:18); n_pen_rows == 0 disables it.
   * Index lists are 0-BASED here (the Julia shim subtracts 1).                                                  */
  int32_t n_pen_rows;
  int32_t n_pen_cols;
  const int32_t* pen_rows;
  const int32_t* pen_cols;
  double mu;
  int32_t store_costates; /* != 0: keep lambda_k for qoc_get_costates (debug / parity tests)                    */
  int32_t reserved;
} qoc_problem;

/* ---- lifetime ------------------------------------------------------------------------------------------- */

/* replaces setup_grape_cache(A0, x0, u_size)  src/gradient_computations.jl:79-96.
 * A0: c128 d x d; A: nc consecutive c128 d x d; x0: c128 d x m; T: c128 d x m target (may be NULL when
 * cost == QOC_COST_NONE).  Uploads the constants and allocates x, lambda, U_k, dU_k/du_j, dJdu on the device. */
int qoc_create(const qoc_problem* prob, const double* A0, const double* A, const double* x0, const double* T,
               qoc_handle** out);
int qoc_destroy(qoc_handle* h);

/* grape_sensitivity takes dUkdp_order per call (src/gradient_computations.jl:35) and the cost closures are
 * built after the cache (examples/ipopt_callbacks_exp.jl:9 vs examples/zz_coupling_ipopt_exp.jl:16): both can
 * be changed on an existing handle.  Changing the order invalidates cached Jacobians (recomputed on demand
 * from the cached u); changing the cost does not touch the cached propagation.                               */
int qoc_set_order(qoc_handle* h, int order);
int qoc_set_cost(qoc_handle* h, int cost, const double* T, int n);
/* The reference's propagate computes the matrix exponentials only (src/gradient_computations.jl:17-25) and
 * grape_sensitivity adds the Jacobians (:65-74).  Default (on = 0): the same split -- qoc_propagate runs K1 without
 * Jacobians (an f-only call, e.g. a line search, costs about a third of an evaluation) and qoc_gradient re-runs K1 with them
 * on the cached u.  on != 0: qoc_propagate produces them together with U_k (they share the Pade powers), which is
 * cheaper when f_grad always follows f.  qoc_eval always does both in one pass.                                  */
int qoc_set_eager_jacobians(qoc_handle* h, int on);
/* Optional promise |u[j,k]| <= umax[j] for every later call (an optimiser with box bounds knows them:
 * examples/zz_coupling_ipopt_exp.jl:54-56).  On the general path (d > 28) the Pade degree and the number of squarings are
 * chosen per launch from a bound on ||X_k||_1; without the promise that bound is max_k |u_jk|, read back from the device
 * (one small synchronisation per call); with it qoc_eval_device never synchronises (and can be captured in a CUDA graph).
 * The promise is checked on the device: a violation makes the next synchronising call fail with QOC_ERR_INVALID.
 * umax: nc doubles; NULL withdraws the promise.  Smaller dimensions choose (degree, s) per slice on the device: no effect.  */
int qoc_set_control_bounds(qoc_handle* h, const double* umax);

/* ---- the hot path, host buffers in / host buffers out ------------------------------------------------------ */

/* replaces propagate(A0, A, u, x0, cache) + Jfinal(x[end]) + sum(L, x)
 * (src/gradient_computations.jl:2-32; examples/ipopt_callbacks_exp.jl:16-18).
 * J_out: batch doubles (may be NULL, and is left untouched when cost == QOC_COST_NONE and no penalty);
 * x_final_out: c128 d x m x batch (may be NULL).  Records u as "the u used for the computations" (:12).       */
int qoc_propagate(qoc_handle* h, const double* u, double* J_out, double* x_final_out);

/* replaces grape_sensitivity(A0, A, dJfinal_dx, u, x0, cache; dUkdp_order, dL_dx)  (:35-77).
 * u: if non-NULL it must equal the last propagated u, else QOC_ERR_STALE_CACHE (:37-39).
 * lambda_final: c128 d x m x batch = dJfinal_dx(x[end]) evaluated by the caller; NULL selects the built-in
 * cost of the handle.  dJdu_out: nc x Nt x batch doubles.                                                      */
int qoc_gradient(qoc_handle* h, const double* u, const double* lambda_final, double* dJdu_out);

/* propagate + gradient in one call (what f followed by f_grad computes, examples/ipopt_callbacks_exp.jl:11-31) */
int qoc_eval(qoc_handle* h, const double* u, double* J_out, double* dJdu_out);

/* ---- the same, device buffers (inputs already resident in HBM), asynchronous on `stream` ------------------ */
/* d_u: nc x Nt x batch doubles in device memory; d_J: batch doubles; d_dJdu: nc x Nt x batch doubles.
 * stream: a cudaStream_t cast to void* (NULL = default stream).  Skips the stale-u bookkeeping.              */
int qoc_eval_device(qoc_handle* h, const double* d_u, double* d_J, double* d_dJdu, void* stream);

/* ---- pulse parameterisation around the path -------------------------------------------------------------- */
/* The immediate caller of the path maps spline coefficients to the pulse and the gradient back
 * (examples/ipopt_callbacks_exp.jl:13-14 `u = transpose(B*c)` and :28 `dJdc = B'*transpose(dJdu)`).
 * qoc_set_basis uploads B (Nt x ns, column-major: basis functions at the slice midpoints,
 * examples/zz_coupling_ipopt_exp.jl:29-38); qoc_eval_coeffs evaluates J and dJ/dc for coefficients
 * c (ns x nc x batch, column-major) with both skinny products on the device, so only ns*nc numbers cross the bus.  */
int qoc_set_basis(qoc_handle* h, const double* B, int ns);
int qoc_eval_coeffs(qoc_handle* h, const double* c, double* J_out, double* dJdc_out);

/* ---- time-segment sharding of ONE long pulse across ranks (one process per GPU) --------------------------- */
/* A handle created with nt = the LOCAL number of slices evaluates the slices [k_first, k_first+nt) of a longer
 * pulse.  Phase 1 computes the local U_k, dU_k/du_j and the rank propagator S_p = U_last ... U_first
 * (d x d, c128 column-major, device memory).  The caller all-gathers the S_p (NCCL), forms the boundary state
 * x_start and boundary costate lambda_end for its segment, and phase 2 finishes the local sweeps.
 * (src/gradient_computations.jl:27-29 and :52-58 are the serial loops this replaces.)                         */
int qoc_shard_phase1_device(qoc_handle* h, const double* d_u, double* d_S_out, void* stream);
/* d_x_start: c128 d x m (state entering the local segment); writes d_x_end: c128 d x m (state leaving it).   */
int qoc_shard_forward_device(qoc_handle* h, const double* d_x_start, double* d_x_end, void* stream);
/* d_lambda_end: c128 d x m costate at the end of the local segment; writes the local gradient columns
 * d_dJdu (nc x nt_local) and d_lambda_start (c128 d x m, costate entering the segment from the right).       */
int qoc_shard_backward_device(qoc_handle* h, const double* d_lambda_end, double* d_dJdu, double* d_lambda_start,
                              void* stream);

/* Running state penalty (n_pen_rows > 0) under time sharding.  The costate recurrence is then affine (src/gradient_computations.jl
 * :47-49, :55-57): over the local segment  lambda_start = S_p' lambda_end + c_p.  After qoc_shard_forward_device this call writes
 * c_p (c128 d x m) and the local sum_k L(x_k) over the nt_local + 1 local states (1 double, may be NULL).  The caller exchanges
 * the c_p next to the S_p and walks the boundary costates down from the last rank; a boundary state belongs to two ranks, so
 * it subtracts L and dL_dx of x_start(p), p >= 1, once (quantumoptimalcontrol.jl_b200/sharding.py does exactly this).
 * lambda_end handed to qoc_shard_backward_device is, as for qoc_gradient, the costate BEFORE dL_dx of the last local state is
 * added.  Returns QOC_ERR_INVALID without a penalty, QOC_ERR_STALE_CACHE before the forward call.                           */
int qoc_shard_affine_device(qoc_handle* h, double* d_c_out, double* d_Jpen_out, void* stream);

/* Phase 2 in one call for the built-in costs (no running penalty: QOC_ERR_UNSUPPORTED, use the three calls above): d_S_all = the nranks all-gathered rank propagators (c128 d x d each,
 * rank order); computes x_start / J / lambda_end on the device (redundantly on every rank, src/penalty_fcns.jl:15-24)
 * and runs the local boundary scan and sweeps.  d_J: 1 double, d_dJdu: nc x nt_local doubles (device memory).        */
int qoc_shard_phase2_device(qoc_handle* h, const double* d_S_all, int nranks, int rank, double* d_J, double* d_dJdu,
                            void* stream);

/* ---- one process, several GPUs (n_gpus / shard kind) ----------------------------------------------- */
/* The reference has no multi-device notion; its two loops over pulses (a multistart driver's) and over slices
 * (src/gradient_computations.jl:27-29, :52-58, :65-74) are what shards.  A Julia / C host cannot run one process per GPU
 * under torchrun, so the library drives the devices itself: one host thread enqueues, every rank has its own stream.
 *   QOC_SHARD_BATCH: prob->batch pulses block-partitioned over the ranks (rank p: pulses [p B / P, (p+1) B / P)), no exchange.
 *   QOC_SHARD_TIME:  ONE pulse (batch == 1, built-in cost), rank p owns the slices [p Nt / P, (p+1) Nt / P).  The rank
 *                    propagators S_p are stored by a kernel of rank p straight into every rank's buffer through NVLink peer
 *                    mappings (cudaMemcpyPeerAsync when a pair has none); phase 2 starts when the P events have fired; every
 *                    rank copies its gradient segment to its slice of the caller's host buffer.  No NCCL, no MPI.
 * devices: n_ranks CUDA ordinals (NULL = 0 .. n_ranks-1); an ordinal may repeat (several ranks on one GPU).
 * qoc_sharded_eval: u and dJdu_out are nc x Nt x batch host arrays exactly as for qoc_eval; J_out: batch doubles.       */
typedef struct qoc_sharded qoc_sharded;
typedef enum { QOC_SHARD_BATCH = 0, QOC_SHARD_TIME = 1 } qoc_shard_kind;
int qoc_create_sharded(const qoc_problem* prob, const double* A0, const double* A, const double* x0, const double* T,
                       int n_ranks, const int* devices, int shard_kind, qoc_sharded** out);
int qoc_sharded_destroy(qoc_sharded* s);
int qoc_sharded_set_order(qoc_sharded* s, int order);
int qoc_sharded_eval(qoc_sharded* s, const double* u, double* J_out, double* dJdu_out);
int qoc_sharded_ranks(const qoc_sharded* s);
/* device time (ms) of the slowest rank in the last qoc_sharded_eval: H2D of its inputs .. D2H of its results, CUDA events */
double qoc_sharded_last_ms(const qoc_sharded* s);
const char* qoc_sharded_last_error(const qoc_sharded* s); /* s may be NULL: error of the last failed qoc_create_sharded */

/* ---- cache getters (the reference returns x as its result, :31; parity tests read the rest) --------------- */
int qoc_get_states(qoc_handle* h, double* x_out);        /* c128 d x m x (Nt+1) x batch  = cache.x              */
int qoc_get_costates(qoc_handle* h, double* lam_out);    /* c128 d x m x (Nt+1) x batch  = cache.lambda          */
int qoc_get_propagators(qoc_handle* h, double* U_out);   /* c128 d x d x Nt x batch      = cache.Uk_vec          */
int qoc_get_jacobians(qoc_handle* h, double* dU_out);    /* c128 d x d x nc x Nt x batch = dUkdu of :61,:67      */

/* ---- diagnostics ---------------------------------------------------------------------------------------- */
const char* qoc_status_string(int status);
const char* qoc_last_error(const qoc_handle* h); /* h may be NULL: error of the last failed qoc_create        */
/* number of kernel launches issued by the last propagate/gradient/eval call on this handle                   */
int qoc_last_launch_count(const qoc_handle* h);
/* device time in ms of the named stage of the last *_device / eval call, measured with CUDA events on the
 * launching stream; stage: 0 = K1 (expm + Jacobians + segment products), 1 = K2 (boundary scan + cost),
 * 2 = K3 (sweeps + gradient contraction).  Requires qoc_set_profiling(h, 1).                                 */
int qoc_set_profiling(qoc_handle* h, int on);
double qoc_stage_ms(const qoc_handle* h, int stage);
/* algorithmic flops (SURVEY.md 8d F_alg) of the last call, with the Pade degree / squarings actually chosen
 * per slice, summed over slices and pulses.                                                                  */
double qoc_last_alg_flops(const qoc_handle* h);
/* flops EXECUTED by K1's DMMA tile loops in the last call: zero-padded (8 NT)^2 x (4 KS) real tile products, three per
 * complex product (3M scheme), one per product on the real-Hamiltonian path.  0 on the general (d > 28) path.
 * A diagnostic next to qoc_last_alg_flops: the roofline fraction is quoted on the algorithmic figure, this one says how
 * busy the FP64 tensor pipe actually was.                                                                        */
double qoc_last_exec_flops(const qoc_handle* h);
int qoc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QOC_B200_H */
