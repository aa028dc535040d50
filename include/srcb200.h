/*
 * srcb200.h -- C ABI of the B200-native soft-robot-control hot path (libsrcb200.so).
 *
 * This is the drop-in boundary for the ONE data-parallel hot path of StanfordASL/soft-robot-control:
 * batched evaluation + linearization of the reduced-order models (SSM, TPWL), the batched iLQR solve and the
 * POD Gram contraction.  The reference has no FFI (it is pure Python); each entry point below names the
 * reference Python call(s) it replaces (paths relative to the reference root).  The Python classes in
 * soft-robot-control_b200/ (importable as `sofacontrol_b200`) bind these symbols with ctypes.
 *
 * Conventions
 *   - every function returns int: 0 = OK, <0 = argument/shape error (SRCB200_E_*), >0 = cudaError_t value;
 *     srcb200_last_error_string() describes the last failure on the calling thread.
 *   - all `const double*` / `double*` / `int32_t*` data arguments are DEVICE pointers unless named `h_*`;
 *     matrices are dense row-major FP64.  The library never allocates or frees device memory, never takes
 *     ownership and never synchronises the host: work is enqueued on `stream` (a cudaStream_t, may be NULL).
 *   - model structs hold device pointers owned by the caller; they are passed by pointer from host memory and
 *     copied by value into the launch.
 *   - state layout: reduced state x = [v; q] (sofacontrol/utils.py:129-142).
 */
#ifndef SRCB200_H
#define SRCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCB200_ABI_VERSION 3

/* error codes (<0) */
#define SRCB200_E_NULL        (-1)  /* required pointer is NULL */
#define SRCB200_E_DIM         (-2)  /* unsupported / inconsistent dimensions */
#define SRCB200_E_METHOD      (-3)  /* unknown discretisation / tpwl method (reference raises RuntimeError) */
#define SRCB200_E_WORKSPACE   (-4)  /* workspace too small */
#define SRCB200_E_NOGPU       (-5)  /* no sm_100 device / kernel image not loadable */

/* discretisation methods: sofacontrol/tpwl/tpwl.py:272-297, sofacontrol/SSM/ssm.py:279-301 */
#define SRCB200_DISCR_FE   0
#define SRCB200_DISCR_BE   1
#define SRCB200_DISCR_BIL  2
#define SRCB200_DISCR_ZOH  3   /* TPWL only: srcb200_zoh_batch on the bank (pre_discretize) or per evaluation */
#define SRCB200_DISCR_NONE 4   /* model is already discrete: SSM discrete=True, TPWL pre-discretised bank */

int         srcb200_abi_version(void);
const char* srcb200_last_error_string(void);
/* 0 if a compute-capability 10.x device is current and the sm_100a image loads */
int         srcb200_device_check(void);

/* ------------------------------------------------------------------------------------------------------------
 * SSM polynomial reduced-order model  (sofacontrol/SSM/ssm.py:18-344)
 * ---------------------------------------------------------------------------------------------------------- */
#define SRCB200_SSM_MAX_N      8
#define SRCB200_SSM_MAX_M      16
#define SRCB200_SSM_MAX_ORDER  4
#define SRCB200_SSM_MAX_FEAT   128

typedef struct srcb200_ssm_model {
    int32_t n;             /* state_dim  (ssm.py:33) */
    int32_t m;             /* input_dim  (ssm.py:34) */
    int32_t nz;            /* output_dim (ssm.py:35); the reference applies the output basis to x, so nz == n */
    int32_t order;         /* ROM_order == SSM_order (ssm.py:36-37) */
    int32_t nfeat;         /* number of monomials of degree 1..order (83 for n=6, order=3) */
    int32_t discr_method;  /* SRCB200_DISCR_FE|BE|BIL, or SRCB200_DISCR_NONE when r_coeff/B_r hold rd_coeff/Bd */
    const double*  r_coeff;  /* n  x nfeat : reduced dynamics (ssm.py:167-168 / 177-178) */
    const double*  w_coeff;  /* nz x nfeat : reduced -> observed, C_map (ssm.py:170-171) */
    const double*  v_coeff;  /* n  x nfeat : observed -> reduced, W_map (ssm.py:173-174) */
    const double*  B_r;      /* n  x m */
    const double*  z_ref;    /* nz */
    const uint8_t* mono;     /* nfeat x SRCB200_SSM_MAX_ORDER variable indices (sorted), 0xFF padded:
                                the basis order of SSM.get_poly_basis (ssm.py:158-164) */
} srcb200_ssm_model;

/* Replaces SSMDynamics.get_jacobians / get_continuous_jacobians / get_discrete_jacobians (ssm.py:198-225),
 * get_observer_jacobians (ssm.py:228-235) and x_to_zfyf (ssm.py:105-111) for `count` states at once.
 *   x (count x n), u (count x m)  ->  A (count x n x n), B (count x n x m), d (count x n)   [discretised with dt]
 *                                     H (count x nz x n), c (count x nz) = C(x) - H x, z (count x nz) = C(x)+z_ref
 * Any output pointer may be NULL.  If dt < 0 the continuous (A_c, B_c, d_c) are returned. */
int srcb200_ssm_eval_linearize_batch(const srcb200_ssm_model* mdl, int64_t count, const double* x, const double* u,
                                     double dt, double* A, double* B, double* d, double* H, double* c, double* z,
                                     void* stream);

/* Polynomial maps on `count` points: which = 0: out = w_coeff phi(in) (+ z_ref if add_ref)   [C_map / x_to_zfyf]
 *                                    which = 1: out = v_coeff phi(in - z_ref if add_ref)      [W_map / compute_RO_state]
 *                                    which = 2: out = r_coeff phi(in) + B_r u  (u (count x m) may be NULL)
 *                                               [reduced_dynamics / reduced_dynamics_discrete, ssm.py:167-178] */
int srcb200_ssm_map_batch(const srcb200_ssm_model* mdl, int32_t which, int32_t add_ref, int64_t count,
                          const double* in, const double* u, double* out, void* stream);

/* Replaces SSM.rollout (ssm.py:134-156) for `batch` independent trajectories:
 *   x0 (batch x n), u (batch x N x m)  ->  x (batch x (N+1) x n), z (batch x (N+1) x nz) = C_map(x) + z_ref.
 * z may be NULL. */
int srcb200_ssm_rollout_batch(const srcb200_ssm_model* mdl, int64_t batch, int32_t N, const double* x0,
                              const double* u, double dt, double* x, double* z, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * TPWL model  (sofacontrol/tpwl/tpwl.py:14-342)
 * ---------------------------------------------------------------------------------------------------------- */
#define SRCB200_TPWL_NN        0
#define SRCB200_TPWL_WEIGHTING 1

typedef struct srcb200_tpwl_model {
    int32_t n;             /* state_dim = 2 r (tpwl.py:38) */
    int32_t m;             /* input_dim */
    int32_t nz;            /* output_dim (rows of H), 0 if no output model */
    int32_t P;             /* num_points (tpwl.py:27) */
    int32_t method;        /* SRCB200_TPWL_NN | SRCB200_TPWL_WEIGHTING (tpwl.py:244,251) */
    int32_t discr_method;  /* applied per evaluation to the selected/blended (A,B,d); SRCB200_DISCR_NONE when the
                              bank below is already discrete (pre_discretize, tpwl.py:299-322) */
    double  wq, wv;        /* dist_weights['q'], ['v'] (tpwl.py:166-167) */
    double  beta;          /* beta_weighting (tpwl.py:189) */
    const double* qT;      /* r x P : stored positions, TRANSPOSED so that consecutive points are contiguous */
    const double* vT;      /* r x P : stored velocities, transposed */
    const double* A;       /* P x n x n bank (continuous A_c, or pre-discretised A_d) */
    const double* B;       /* P x n x m */
    const double* d;       /* P x n */
    const double* H;       /* nz x n  (Hf @ V, tpwl.py:86-89) or NULL */
    const double* z_ref;   /* nz or NULL */
} srcb200_tpwl_model;

/* Replaces TPWL.calc_nearest_point (tpwl.py:160-168) for `count` states: idx (count) int32, BIT-EXACT with
 * np.argmin of the numpy distances (numpy's pairwise summation order is reproduced); dist (count) optional. */
int srcb200_tpwl_nearest_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, int32_t* idx,
                               double* dist, void* stream);

/* Replaces TPWL.calc_weighting_factors (tpwl.py:170-191): w (count x P). */
int srcb200_tpwl_weights_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, double* w,
                               void* stream);

/* Replaces TPWLATV.get_jacobians (tpwl.py:236-270) + discretize_dynamics (tpwl.py:272-297, fe/be/bil) for `count`
 * states: A (count x n x n), B (count x n x m), d (count x n), idx (count) [nn only, may be NULL].
 * dt < 0 returns the continuous / stored matrices untouched.  workspace: srcb200_tpwl_linearize_workspace(). */
size_t srcb200_tpwl_linearize_workspace(const srcb200_tpwl_model* mdl, int64_t count);
int srcb200_tpwl_linearize_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, double dt,
                                 double* A, double* B, double* d, int32_t* idx, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* Replaces TPWL.rollout / TPWLATV.update_state (tpwl.py:193-216, 226-234, 336-339) for `batch` trajectories:
 *   x0 (batch x n), u (batch x N x m) -> x (batch x (N+1) x n), z (batch x (N+1) x nz) (NULL allowed),
 *   idx (batch x N) nearest point per step (nn only, NULL allowed). */
size_t srcb200_tpwl_rollout_workspace(const srcb200_tpwl_model* mdl, int64_t batch);
int srcb200_tpwl_rollout_batch(const srcb200_tpwl_model* mdl, int64_t batch, int32_t N, const double* x0,
                               const double* u, double dt, double* x, double* z, int32_t* idx, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Replaces TPWL.x_to_zfyf(x, zf=True) (tpwl.py:115-126): z (count x nz) = H x + z_ref. */
int srcb200_tpwl_output_batch(const srcb200_tpwl_model* mdl, int64_t count, const double* x, double* z, void* stream);

/* Replaces TPWLATV.discretize_dynamics for a whole bank (pre_discretize, tpwl.py:299-322), methods fe/be/bil:
 * (A_c, B_c, d_c) (P x ..) -> (A_d, B_d, d_d).  In-place allowed. */
int srcb200_discretize_batch(int32_t n, int32_t m, int32_t discr_method, int64_t count, double dt,
                             const double* A_c, const double* B_c, const double* d_c, double* A_d, double* B_d,
                             double* d_d, void* stream);

/* Zero-order-hold discretisation of a batch (sofacontrol/utils.py:302-335 zoh_affine: expm of the (n+m+1)^2 augmented
 * matrix; TPWLATV.discretize_dynamics('zoh') / pre_discretize, tpwl.py:291-322).  Pade-13 scaling and squaring. */
size_t srcb200_zoh_workspace(int32_t n, int32_t m, int64_t count);
int srcb200_zoh_batch(int32_t n, int32_t m, int64_t count, double dt, const double* A_c, const double* B_c,
                      const double* d_c, double* A_d, double* B_d, double* d_d, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * iLQR  (sofacontrol/lqr/ilqr.py:6-300, sofacontrol/lqr/config.py:1-31)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct srcb200_ilqr_config {   /* field for field iLQRConfig (lqr/config.py:1-31) */
    int32_t max_iter;
    int32_t include_input_var_constraint;
    int32_t do_linesearch;
    int32_t regularize;
    int32_t state_regularization;
    int32_t counter_limit;
    double  epsilon;
    double  alpha0, alpha_scaling, improv_lb, improv_ub, alpha_min;
    double  rho0, drho0, rho_scaling, rho_increase_fp, rho_max, rho_min;
} srcb200_ilqr_config;
/* Non-PD Q_uu~ (ilqr.py:276-299), reproduced literally: with `regularize` rho is raised, the backward sweep stops at
 * that step (K_t = k_t = 0 for every t at or below it), rho is lowered once as after a complete sweep and the line
 * search runs with those gains -- the reference never restarts the sweep (its comment at ilqr.py:287 says it does;
 * the `break` only leaves the for loop).  Without `regularize` the non-PD matrix is inverted as it is. */

#define SRCB200_ILQR_MODEL_SSM  0
#define SRCB200_ILQR_MODEL_TPWL 1

/* per-problem status bits (output) */
#define SRCB200_ILQR_ST_CONVERGED   1   /* 0 <= J_prev - J < epsilon (ilqr.py:109-115) */
#define SRCB200_ILQR_ST_MAXITER     2   /* left the loop on nbr_iter > max_iter (ilqr.py:54) */
#define SRCB200_ILQR_ST_ABANDONED   4   /* counter_limit consecutive line-search failures (ilqr.py:98-103) */
#define SRCB200_ILQR_ST_NONPD       8   /* some backward sweep met a non-PD Q_uu~ (ilqr.py:282-287) */
#define SRCB200_ILQR_ST_NONFINITE  16   /* cost became NaN/Inf */

typedef struct srcb200_ilqr_problem {
    int64_t batch;           /* independent problems */
    int32_t N;               /* planning_horizon */
    int32_t gauss_newton;    /* 0: constant model.H in the cost derivatives, exactly ilqr.py:177-196;
                                1: H_t = dC/dx at x_t (SSM models; the H-property adapter of SURVEY.md App. C.2) */
    double  dt;
    const double* x0;        /* batch x n */
    const double* u_init;    /* batch x N x m warm start, or NULL for zeros (ilqr.py:46-47) */
    const double* z_target;  /* batch x (N+1) x nz (set_target, ilqr.py:21-22); stride 0 allowed via shared_target */
    const double* u_last;    /* batch x m (set_u_last, ilqr.py:24-25) or NULL for zeros */
    const double* Q;         /* nz x nz   (cost_params.Q)  shared by the batch */
    const double* R;         /* m x m     (cost_params.R)  */
    const double* Qf;        /* nz x nz   (cost_params.Qf) */
    const double* H_const;   /* nz x n constant output matrix when gauss_newton == 0 (model.H); NULL = zeros */
    int32_t shared_target;   /* 1: z_target is a single (N+1) x nz trajectory used by every problem */
    int32_t _pad;
} srcb200_ilqr_problem;

typedef struct srcb200_ilqr_result {
    double*  x;           /* batch x (N+1) x n */
    double*  u;           /* batch x N x m */
    double*  K;           /* batch x N x m x n : gains of the LAST backward pass (ilqr.py:107) */
    double*  cost;        /* batch : cost of the returned trajectory */
    double*  cost0;       /* batch : cost of the initial rollout (NULL allowed) */
    double*  rho;         /* batch : final rho (NULL allowed) */
    int32_t* iterations;  /* batch : nbr_iter on exit */
    int32_t* status;      /* batch : SRCB200_ILQR_ST_* bits */
    int32_t* trials;      /* batch : total number of line-search forward passes (NULL allowed) */
    double*  trace;       /* optional batch x (max_iter+1) x 4 : per iteration {cost after, alpha accepted (0 if
                             failed), rho after the backward pass, horizon index of the failed PD test of that
                             backward pass or -1}; NULL allowed */
} srcb200_ilqr_result;

/* Device scratch of one solve_batch call: per problem two trajectory records (accepted / trial), the feed-forward
 * gains, the line-search scalars and 8 doubles of saved solver state, plus the task queues of the persistent kernels
 * (a problem is suspended after every iteration and resumed by the next free warp / CTA).  Contents need not be
 * initialised or preserved between calls. */
size_t srcb200_ilqr_workspace_bytes(int32_t model_kind, const void* model, const srcb200_ilqr_problem* prob);

/* Replaces iLQR.ilqr_computation (ilqr.py:27-107) for prob->batch independent problems.  Persistent kernel: one
 * warp (Trunk / Diamond SSM shape) or one CTA (every other model) runs one iteration of one problem at a time.
 * model_kind selects the struct behind `model` (srcb200_ssm_model / srcb200_tpwl_model). */
int srcb200_ilqr_solve_batch(int32_t model_kind, const void* model, const srcb200_ilqr_config* cfg,
                             const srcb200_ilqr_problem* prob, const srcb200_ilqr_result* res, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Replaces iLQR.forward_pass (ilqr.py:117-162): x_prev/u_prev nominal, alpha, K (may be NULL), k (may be NULL)
 *   -> x, u, cost (batch), A (batch x N x n x n), B (batch x N x n x m), d (batch x N x n)  (A/B/d NULL allowed) */
int srcb200_ilqr_forward_pass(int32_t model_kind, const void* model, const srcb200_ilqr_config* cfg,
                              const srcb200_ilqr_problem* prob, const double* x_prev, const double* u_prev,
                              double alpha, const double* K, const double* k, double* x, double* u, double* cost,
                              double* A, double* B, double* d, void* workspace, size_t workspace_bytes,
                              void* stream);

/* Replaces iLQR.dlqr_recursion (ilqr.py:219-300): nominal x,u and its linearisation A,B -> K, k, Q_u, Q_uu and
 * the updated (rho, drho) (batch each, in/out); pd_fail_step (batch, NULL allowed): horizon index at which the PD
 * test failed and the sweep stopped (K, k zero from there down, Q_u / Q_uu zero below it), -1 if it never failed. */
int srcb200_ilqr_backward_pass(int32_t model_kind, const void* model, const srcb200_ilqr_config* cfg,
                               const srcb200_ilqr_problem* prob, const double* x, const double* u,
                               const double* A, const double* B, double* K, double* k, double* Q_u, double* Q_uu,
                               double* rho, double* drho, int32_t* pd_fail_step, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * POD  (sofacontrol/mor/pod.py:181-200, 9-78)
 * ---------------------------------------------------------------------------------------------------------- */
/* G (ns x ns, row-major, ldg) (+)= X^T X for a row-block X (nf x ns, row-major, ldx) -- the Gram matrix whose
 * eigen-decomposition replaces np.linalg.svd in compute_POD (pod.py:191): S^2 = eig(G), U = X V S^-1.
 * FP64 tensor-core (DMMA) SYRK-style contraction; both triangles are written.  accumulate != 0 adds to G. */
int srcb200_pod_gram(int64_t nf, int64_t ns, const double* X, int64_t ldx, double* G, int64_t ldg,
                     int32_t accumulate, void* stream);

/* C (M x N, ldc) = alpha * A (M x K, lda) * B (K x N, ldb), FP64 DMMA.  Used for U = X (V S^-1) (pod.py:198),
 * POD.compute_RO_state / compute_FO_state / compute_RO_matrix (pod.py:22-72) on batches, and the TPWL weighted
 * bank blend  W (batch x P) * bank (P x (n*n+n*m+n))  (tpwl.py:246-248).  transA != 0 uses A^T (A is K x M). */
int srcb200_dgemm(int32_t transA, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, double* C, int64_t ldc, void* stream);

/* Eigen-decomposition of a small symmetric positive semi-definite matrix A (n x n, n <= 160, row-major, lda) by
 * one-sided Jacobi in one CTA: evals (n, descending), V (n x n, row-major, ldv, column j = eigenvector j).  The
 * Rayleigh-Ritz / orthonormalisation step of the leading-eigenpair solver on the snapshot Gram matrix that replaces
 * the full np.linalg.svd of compute_POD (pod.py:191-199: only the modes the energy rule keeps are needed, and
 * sum(S^2) = trace(G)).  info (device int, may be NULL): sweeps used, -1 if the sweep limit was hit. */
int srcb200_sym_eig_psd(int32_t n, const double* A, int64_t lda, double* evals, double* V, int64_t ldv,
                        int32_t* info, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Callers either side of the hot path, batched.  All matrices row-major, device pointers.
 * ---------------------------------------------------------------------------------------------------------- */
/* DiscreteEKFObserver.predict_state (sofacontrol/tpwl/observer.py:94-104) for `batch` independent filters:
 *   x <- A_d x + B_d u + d_d ;  Sigma <- (A_d Sigma) A_d^T + W.
 * A_d (batch, n, n), B_d (batch, n, m), d_d (batch, n): the linearisation at x (srcb200_tpwl_linearize_batch);
 * u (batch, m); W (n, n) shared; x (batch, n) and Sigma (batch, n, n) are updated in place. */
int srcb200_ekf_predict_batch(int32_t n, int32_t m, int64_t batch, const double* A_d, const double* B_d,
                              const double* d_d, const double* u, const double* W, double* x, double* Sigma,
                              void* stream);
/* DiscreteEKFObserver.update_state (observer.py:106-126):  y_r = y - y_ref ; S = (C Sigma) C^T + V ;
 *   K = (Sigma C^T) inv(S) ; x <- x + K (y_r - C x) ; Sigma <- (I - K C) Sigma.
 * C (p, n), V (p, p), y_ref (p) (NULL = zeros) shared; y (batch, p); x, Sigma in place. */
int srcb200_ekf_update_batch(int32_t n, int32_t p, int64_t batch, const double* C, const double* V,
                             const double* y_ref, const double* y, double* x, double* Sigma, void* stream);
/* Infinite-horizon discrete LQR (sofacontrol/lqr/lqr.py:6-31) for `batch` systems A (batch, n, n), B (batch, n, m);
 * Q (n, n), R (m, m) shared (shared_cost != 0) or per system.  K (batch, m, n) with u = +K x, P (batch, n, n).
 *   mode 0: `solve_riccati` literally -- value iteration from P = 0 until ||L - L_old||_F <= tol (reference: 1e-4);
 *           iterations (batch, NULL allowed) = number of passes.
 *   mode 1: `dare` -- the stabilising solution to working precision (structure-preserving doubling; tol is the
 *           relative change of P at which it stops, e.g. 1e-15), then K = -inv(B^T P B + R) (B^T P A). */
int srcb200_dlqr_riccati_batch(int32_t n, int32_t m, int64_t batch, const double* A, const double* B,
                               const double* Q, const double* R, int32_t shared_cost, double tol, int32_t max_iter,
                               int32_t mode, double* K, double* P, int32_t* iterations, void* stream);
/* TrajTrackingLQR.perform_dlqr_recursion (sofacontrol/lqr/traj_tracking_lqr.py:31-44) on `batch` trajectories of
 * `steps` linearisations A (batch, steps, n, n), B (batch, steps, n, m) in time order:  P_T = Q, backwards
 *   K_i = -solve(R + B^T P B, B^T P A) ; P <- Q + K^T R K + (A + B K)^T P (A + B K).
 * K (batch, steps, m, n), P (batch, steps + 1, n, n), both in time order. */
int srcb200_tvlqr_batch(int32_t n, int32_t m, int32_t steps, int64_t batch, const double* A, const double* B,
                        const double* Q, const double* R, double* K, double* P, void* stream);
/* One stored TPWL point per entry from reduced second-order matrices: extract_AB (sofacontrol/utils.py:251-286, dense
 * branch) and the affine term of add_continuous_TPWL (sofacontrol/tpwl/tpwl_utils.py:263-276):
 *   A_c = [[-inv(M) D, -inv(M) K], [I, 0]] ; B_c = [[inv(M) H], [0]] ; d_c = [solve(M, f + K q) ; 0].
 * K, D, M (count, r, r), H (count, r, m), f, q (count, r; both NULL with d_c NULL for extract_AB alone)
 * -> A_c (count, 2r, 2r), B_c (count, 2r, m), d_c (count, 2r). */
int srcb200_tpwl_bank_point_batch(int32_t r, int32_t m, int64_t count, const double* K, const double* D,
                                  const double* M, const double* H, const double* f, const double* q, double* A_c,
                                  double* B_c, double* d_c, void* stream);
/* GuSTO.compute_accuracy (sofacontrol/scp/gusto.py:203-223) for `batch` trajectories of N linearisation points:
 * (f_k, A_k, B_k) continuous dynamics at the previous iterate (x_k, u_k), f at the candidate (x, u);
 *   error = sum_i dt ||s o (f_i - fa_i)||_2, approx = sum_i dt ||s o fa_i||_2, fa_i = f_k,i + A_k,i dx_i + B_k,i du_i,
 *   rho = error / (J + approx).   f_k, f (batch, N, n); A_k (batch, N, n, n); B_k (batch, N, n, m);
 * x, x_k (batch, N + 1, n); u, u_k (batch, N, m); f_scale (n) or NULL; J (batch) or NULL; outputs (batch). */
int srcb200_gusto_accuracy_batch(int32_t n, int32_t m, int32_t N, int64_t batch, double dt, const double* f_k,
                                 const double* A_k, const double* B_k, const double* f, const double* x,
                                 const double* x_k, const double* u, const double* u_k, const double* f_scale,
                                 const double* J, double* rho, double* error, double* approx, void* stream);
/* Receding-horizon bookkeeping between two solves (the warm-start hooks of sofacontrol/lqr/ilqr.py:24-25, 46-47 and
 * the target window of tpwl/controllers.py:185-198), control step k of T:  u_applied = u_plan[:, 0] (also logged to
 * u_log (batch, T, m) if not NULL), u_warm = [u_plan[:, 1:], u_plan[:, -1]], z_window = z_ref[:, k+1 : k+2+N].
 * u_plan, u_warm (batch, N, m); z_ref (batch, T + N + 1, nz); z_window (batch, N + 1, nz). */
int srcb200_mpc_shift_batch(int64_t batch, int32_t N, int32_t m, int32_t nz, int32_t T, int32_t k,
                            const double* u_plan, const double* z_ref, double* u_warm, double* u_applied,
                            double* z_window, double* u_log, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRCB200_H */
