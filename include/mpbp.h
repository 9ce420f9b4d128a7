/*
 * mpbp.h -- C ABI of the B200-native block-preconditioner hot path (libmpbp.so).
 *
 * Drop-in boundary for the hot path of abarret/mp-block-preconditioners: applying the
 * approximate-commutator (BFBt) block preconditioner inside the Krylov solve of the two-phase
 * variable-viscosity MAC-grid Stokes system.  Each entry point cites the reference interface it
 * replaces (paths are relative to the reference checkout).
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = argument error (MPBP_E_*), >0 = CUDA/NCCL failure;
 *     mpbp_last_error_string() describes the last failure on the calling thread.
 *   - all `double*` vector arguments are DEVICE pointers borrowed for the duration of the call
 *     (the *_host entry points take HOST pointers and copy inside the call).
 *   - vectors are ordered [u_n | v_n | u_s | v_s | p], each field row-major over the rank's slab
 *     of `rows_local x n` cells (preconditioner.py:100-106, utils.py:178-208); with one rank the
 *     slab is the whole n x n grid and the layout is exactly the reference's.
 *   - `stream` is a cudaStream_t passed as void*; NULL = the legacy default stream.
 *   - a plan is not thread-safe; distinct plans may be used concurrently.
 */
#ifndef MPBP_H
#define MPBP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPBP_E_ARG (-1)      /* bad argument */
#define MPBP_E_STATE (-2)    /* plan not in a state that allows the call */
#define MPBP_E_NOMEM (-3)    /* workspace too small */
#define MPBP_E_UNSUPPORTED (-4)

/* sub-solver kinds: what fills the ilupp.ILUTPreconditioner slots of solve.py:251 and :254 */
#define MPBP_SUB_JACOBI 0 /* `sweeps` damped Jacobi sweeps from x=0 (solve.py:149-159, :262, :268) */
#define MPBP_SUB_MG 1     /* `cycles` V(nu1,nu2) cycles, the "multigrid PC with Jacobi smoother" of solve.py:266,:274 */

#define MPBP_SIDE_LEFT 0  /* scipy.sparse.linalg.gmres semantics (solve.py:12, :221) */
#define MPBP_SIDE_RIGHT 1 /* flexible right preconditioning, pyamg.krylov.fgmres semantics (solve.py:285) */

typedef struct mpbp_plan mpbp_plan;

typedef struct mpbp_config {
  /* MultiphaseBlockPreconditioner(n, xi, eta_n, eta_s), preconditioner.py:18-24 */
  int n;
  double xi, eta_n, eta_s;
  /* get_big_A_matrix(c, d_u, d_p=1.0, d_div=-1.0), preconditioner.py:299 */
  double c, d_u, d_p, d_div;
  /* cell-centred theta_n (n*n doubles, HOST, global grid); NULL = analytic thn of preconditioner.py:9-11 */
  const double* theta_host;
  /* slab decomposition: rank owns grid rows [rank*n/nranks, (rank+1)*n/nranks) */
  int rank, nranks;
  const void* nccl_unique_id; /* 128-byte ncclUniqueId (mpbp_nccl_unique_id) when nranks > 1 */
  /* definition of the approximate solves F~^-1 (solve.py:251,:258,:274) and (GtG)~^-1 (solve.py:254,:265,:271) */
  int F_kind, P_kind;
  int F_sweeps, P_sweeps; /* MPBP_SUB_JACOBI */
  int F_cycles, P_cycles; /* MPBP_SUB_MG */
  double omega;           /* Jacobi damping */
  int nu1, nu2;           /* pre / post smoothing sweeps */
  int n_coarse;           /* coarsest grid size (dense inverse there) */
  int cheb;               /* 1: Chebyshev-accelerate the V-cycles over [lmin, lmax] */
  double lmin, lmax;
  int project;            /* 1: remove the mean from every (GtG)~^-1 result (constant null space, solve.py:260-264) */
  int operators_only;     /* 1: no multigrid hierarchy (operator applies only; sub-solves return MPBP_E_STATE) */
  int dist_min_n;         /* nranks>1: multigrid levels with n < dist_min_n are replicated on every rank (all-gather
                             of the restricted residual) instead of slab-distributed; 0 = default (512) */
  /* optional caller-owned DEVICE workspace (e.g. a torch tensor); NULL = the plan cudaMallocs */
  void* workspace;
  size_t workspace_bytes;
} mpbp_config;

/* fills defaults: the reference's default run (solve.py:291-297) with the multigrid sub-solver */
int mpbp_config_default(mpbp_config* cfg);
/* device bytes a plan built from cfg needs (excluding Krylov bases) */
int mpbp_plan_workspace_bytes(const mpbp_config* cfg, size_t* bytes);
/* replaces MultiphaseBlockPreconditioner.__init__ + get_big_A_matrix + the setup of
 * solve_with_approx_schur_pc (preconditioner.py:18-24, :299-349; solve.py:243-254) */
int mpbp_plan_create(mpbp_plan** plan, const mpbp_config* cfg);
int mpbp_plan_destroy(mpbp_plan* plan);
const char* mpbp_last_error_string(void);
/* hash of the sources the loaded binary was compiled from (the host side refuses a binary that does not match the tree) */
const char* mpbp_build_id(void);
int mpbp_nccl_unique_id(void* out128);
/* queries */
int mpbp_plan_rows_local(const mpbp_plan* plan); /* rows of the slab owned by this rank */
int mpbp_plan_num_levels(const mpbp_plan* plan);
long long mpbp_plan_launches(const mpbp_plan* plan); /* kernels launched by the plan so far */

/* ---- operator applies (each replaces one dense np.matmul of the reference) ---- */
/* y = A x, 5N -> 5N.  `A @ xk` solve.py:166; np.matmul(A,u_vec) apply.py:72; inside fgmres solve.py:285 */
int mpbp_apply_A(mpbp_plan*, const double* x, double* y, void* stream);
/* y = F x, 4N -> 4N.  F of preconditioner.py:337 (implicit in F_inv, solve.py:251) */
int mpbp_apply_F(mpbp_plan*, const double* x, double* y, void* stream);
/* y = G p, N -> 4N.  np.matmul(G, x_p) solve.py:273; G of preconditioner.py:313 */
int mpbp_apply_G(mpbp_plan*, const double* p, double* y, void* stream);
/* r = D w (+ add if non-NULL), 4N -> N.  np.matmul(D, Finv_v) + v[4N:] solve.py:259; D un-negated, preconditioner.py:311,:349 */
int mpbp_apply_D(mpbp_plan*, const double* w, const double* add, double* r, void* stream);
/* y = (-D G) p, N -> N.  Gt_G of solve.py:247 */
int mpbp_apply_GtG(mpbp_plan*, const double* p, double* y, void* stream);
/* y = (-D F G) p, N -> N.  np.matmul(Gt_F_G, x_a) solve.py:267; Gt_F_G of solve.py:248-249 */
int mpbp_apply_GtFG(mpbp_plan*, const double* p, double* y, void* stream);

/* ---- relaxation and the approximate solves ---- */
/* solve.Jacobi (solve.py:149-159) with damping omega on F (4N) / GtG (N): `sweeps` sweeps starting from x (in place) */
int mpbp_jacobi_F(mpbp_plan*, const double* b, double* x, int sweeps, double omega, void* stream);
int mpbp_jacobi_P(mpbp_plan*, const double* b, double* x, int sweeps, double omega, void* stream);
/* one V(nu1,nu2) cycle from a zero initial guess: x = V b */
int mpbp_vcycle_F(mpbp_plan*, const double* b, double* x, void* stream);
int mpbp_vcycle_P(mpbp_plan*, const double* b, double* x, void* stream);
/* x = F~^-1 b and x = (GtG)~^-1 b as configured: `F_inv @ .` solve.py:258,:274; `Gt_G_factorization @ .` solve.py:265,:271 */
int mpbp_solve_F(mpbp_plan*, const double* b, double* x, void* stream);
int mpbp_solve_P(mpbp_plan*, const double* b, double* x, void* stream);

/* ---- the preconditioner apply: approx_schur_op(v), solve.py:257-277 (v is not modified) ---- */
int mpbp_precond_apply(mpbp_plan*, const double* v, double* z, void* stream);
/* same with HOST buffers (h2d + apply + d2h inside the call): what LinearOperator.matvec sees, solve.py:280-281 */
int mpbp_precond_apply_host(mpbp_plan*, const double* v_host, double* z_host, void* stream);
/* communication probe for the multi-GPU throughput sweep (BASELINE.json configs[4]: "halo + allreduce scaling"): average
 * microseconds of one level-0 halo exchange of a 5-field vector and of one scalar all-reduce (0 with one rank) */
int mpbp_comm_probe(mpbp_plan*, int reps, double* halo_us, double* allreduce_us, void* stream);
/* algorithmic bytes of one precond apply / one A apply under the current configuration (SURVEY 8d accounting) */
int mpbp_precond_bytes(const mpbp_plan*, double* bytes);

/* ---- Krylov vector kernels (np.dot / np.linalg.norm / axpy inside gmres; utils.py:7-17) ---- */
/* results are written to HOST doubles (the call synchronises the stream); len = local length, summed over ranks */
int mpbp_dot(mpbp_plan*, const double* x, const double* y, size_t len, double* result, void* stream);
int mpbp_nrm2(mpbp_plan*, const double* x, size_t len, double* result, void* stream);
int mpbp_axpy(mpbp_plan*, double alpha, const double* x, double* y, size_t len, void* stream);
/* out[k] = <V_k, w>, k < nvec; V_k = V + k*ld */
int mpbp_multi_dot(mpbp_plan*, const double* V, size_t ld, int nvec, const double* w, size_t len, double* out, void* stream);
/* w += sum_k alpha[k] V_k (alpha on HOST) */
int mpbp_multi_axpy(mpbp_plan*, const double* V, size_t ld, int nvec, const double* alpha, double* w, size_t len, void* stream);
/* weighted_L1, weighted_L2 (weight w scalar) and max_norm of a-b, utils.py:7-17: out[0..2] on HOST */
int mpbp_wnorms(mpbp_plan*, const double* a, const double* b, size_t len, double w, double* out3, void* stream);

/* ---- manufactured solution / RHS of solve.main (solve.py:52-78, utils.py:159-210) on the device ---- */
int mpbp_fill_manufactured(mpbp_plan*, double* u_vec, double* b_vec, double b_p_sign, void* stream);

/* ---- Krylov solve: fgmres(A, b, M=approx_schur, tol, maxiter) solve.py:285 / scipy gmres solve.py:12 ---- */
typedef struct mpbp_gmres_opts {
  double rtol;       /* tol=1e-8, solve.py:285 */
  int restart;       /* inner iterations per cycle */
  int maxiter;       /* LEFT: max outer cycles (scipy); RIGHT: max total inner iterations (pyamg) */
  int side;          /* MPBP_SIDE_LEFT / MPBP_SIDE_RIGHT */
  int use_precond;   /* 0: M=None (solve.py:207) */
  int x0_nonzero;    /* 1: x holds the initial guess on entry */
  int force_iters;   /* >0: ignore convergence and run exactly this many inner iterations (benchmarking) */
  void* workspace;   /* optional caller-owned device memory for the Krylov bases */
  size_t workspace_bytes;
  /* per-iteration callback of pyamg's fgmres as the reference uses it, callback=print_true_res_norm(A, b) at
   * solve.py:285 / :163-169: after every inner iteration the current iterate x_k = x0 + Z y_k is formed in xk_buf
   * (DEVICE, 5N doubles, caller-owned; one multi-axpy, no extra preconditioner apply), the stream is synchronised and
   * iter_cb(cb_user, k, relative recurrence residual) is called; a non-zero return stops the solve.  RIGHT side only. */
  void* xk_buf;
  int (*iter_cb)(void* cb_user, int iteration, double rel_residual);
  void* cb_user;
} mpbp_gmres_opts;
int mpbp_gmres_opts_default(mpbp_gmres_opts*);
int mpbp_gmres_workspace_bytes(const mpbp_plan*, const mpbp_gmres_opts*, size_t* bytes);
/* hist[k] = relative (preconditioned, LEFT / recurrence, RIGHT) residual after inner iteration k+1;
 * *info: 0 converged, >0 = maxiter reached (scipy semantics) */
int mpbp_gmres(mpbp_plan*, const double* b, double* x, const mpbp_gmres_opts*, double* hist_host, int hist_cap,
               int* n_iters, int* info, void* stream);
/* Spectral diagnostics as a by-product of the solve (replaces the dense compute_preconditioned_A + PETSc/SLEPc
 * get_eigenvals analysis of solve.py:103-200, :304-309): the upper Hessenberg matrix H_k = V_{k+1}^T (A M^-1) V_k
 * (RIGHT; M A for LEFT) of the LAST Arnoldi cycle of the last mpbp_gmres call on this plan, before the Givens
 * rotations.  Written row-major into H ((k+1) x k, leading dimension ldh >= k); *k receives the cycle length.
 * Its eigenvalues (Ritz values) approximate the outer spectrum of the preconditioned operator. */
int mpbp_gmres_last_hessenberg(const mpbp_plan*, double* H_host, int ldh, int* k);
/* HOST b in, HOST x out: what `fgmres(A, b_vec, M=...)` is to the reference's caller */
int mpbp_gmres_host(mpbp_plan*, const double* b_host, double* x_host, const mpbp_gmres_opts*, double* hist_host,
                    int hist_cap, int* n_iters, int* info, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPBP_H */
