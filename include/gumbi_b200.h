/* gumbi_b200 -- C ABI of the B200-native exact-GP inference core.
 *
 * The reference (JohnGoertz/Gumbi @ 27bdbee) has no FFI: its hot path leaves the repo through
 * Python calls into PyMC (gumbi/regression/pymc/GP.py).  Each entry point below names the reference
 * call it replaces; the Python-side binding a maintainer would add is in INTEGRATION.md
 * (gumbi_b200/_lib.py is that binding, via ctypes).
 *
 * Conventions
 *   - every function returns int: 0 = ok; > 0 = LAPACK-style "leading minor of order k is not
 *     positive definite" (first failing pivot, 1-based); < 0 = argument / CUDA / NCCL error, text
 *     via gb2_last_error().
 *   - plain pointers + sizes only; no exceptions, no Python or torch types cross this boundary.
 *   - host-pointer entry points copy in/out synchronously; *_dev entry points take device pointers
 *     on the handle's device and are ordered on the handle's stream (they return after the stream
 *     has drained unless documented otherwise).
 *   - the caller owns every buffer it passes; the handle owns all device memory it allocates.
 *   - one handle = one GPU = one pair of CUDA streams; a handle is not re-entrant, different handles may be driven from different
 *     threads.  Process-wide state is limited to the error text of a failed gb2_create / gb2_nccl_unique_id (handle == NULL in
 *     gb2_last_error) and the NCCL function table bound by the first multi-GPU call.
 *   - matrices are C-contiguous (row-major) float64, indices int32, exactly as numpy hands them
 *     over from Regressor.get_shaped_data (gumbi/regression/base.py:435-471).
 */
#ifndef GUMBI_B200_H
#define GUMBI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB2_ABI_VERSION 1

/* continuous_kernel= of PymcGP.fit (gumbi/regression/pymc/GP.py:266, :664-684) */
enum gb2_kind {
    GB2_EXPQUAD = 0,     /* pm.gp.cov.ExpQuad      exp(-r^2/2)                               */
    GB2_MATERN52 = 1,    /* pm.gp.cov.Matern52     (1+sqrt5 r+5/3 r^2) exp(-sqrt5 r)          */
    GB2_MATERN32 = 2,    /* pm.gp.cov.Matern32     (1+sqrt3 r) exp(-sqrt3 r)                  */
    GB2_MATERN12 = 3,    /* pm.gp.cov.Matern12     exp(-r)                                    */
    GB2_EXPONENTIAL = 4  /* pm.gp.cov.Exponential  exp(-r/2)                                  */
};

enum gb2_precision {
    GB2_FP64 = 0,        /* everything IEEE fp64 (DMMA tensor path for the contractions)      */
    GB2_TF32 = 1         /* tf32 tensor-core Cholesky trailing update (tcgen05), fp64 elsewhere */
};

#define GB2_MAX_TERMS 4
#define GB2_MAX_D 16
#define GB2_MAX_LIN 8
#define GB2_MAX_COREG 3
#define GB2_MAX_P 16

/* One additive term  (eta^2 k_cont(ls) [+ tau Linear(c)]) * prod_f Coregion_f
 * = what PymcGP._construct_kernels builds per GP (GP.py:711-727, :739-750).            */
typedef struct gb2_term {
    int32_t kind;                         /* enum gb2_kind                                      */
    int32_t d;                            /* # active continuous columns (<= GB2_MAX_D)         */
    int32_t cont_idx[GB2_MAX_D];          /* active_dims of the stationary kernel (GP.py:410)   */
    double ls[GB2_MAX_D];                 /* lengthscale per active column (ARD=False: repeat)  */
    double eta;                           /* cov = eta**2 * k  (GP.py:409-410)                  */
    int32_t n_lin;                        /* # linear columns (0 = no Linear kernel)            */
    int32_t lin_idx[GB2_MAX_LIN];         /* active_dims of pm.gp.cov.Linear (GP.py:453)        */
    double c[GB2_MAX_LIN];                /* Linear offset c                                    */
    double tau;                           /* tau * Linear                                       */
    int32_t n_coreg;                      /* # Coregion factors multiplying this term           */
    int32_t coreg_col[GB2_MAX_COREG];     /* column of X holding the level index (GP.py:462)    */
    int32_t coreg_P[GB2_MAX_COREG];       /* # levels                                           */
    const double* coreg_B[GB2_MAX_COREG]; /* host ptr, P*P row-major, B = W W^T + diag(kappa)   */
} gb2_term;

typedef struct gb2_kernel {
    int32_t n_terms;                      /* 1, or 1 + #categorical dims when additive=True     */
    gb2_term terms[GB2_MAX_TERMS];
    double sigma;                         /* pm.gp.cov.WhiteNoise(sigma)  (GP.py:560-561)       */
    int32_t noise_col;                    /* -1, or column for the Output_noise Coregion (:569) */
    int32_t noise_P;
    const double* noise_B;                /* host ptr, P*P (only the diagonal is used)          */
    double jitter;                        /* pm.gp.util.stabilize JITTER_DEFAULT = 1e-6         */
} gb2_kernel;

typedef struct gb2_handle gb2_handle;

int gb2_abi_version(void);

/* Lifetime.  device = CUDA ordinal.  Replaces nothing in the reference (it keeps no state: F8). */
int gb2_create(gb2_handle** out, int device, int precision);
int gb2_destroy(gb2_handle* h);
const char* gb2_last_error(const gb2_handle* h); /* h may be NULL: last create() error */

/* Training data X:(N,D_in), y:(N,) as produced by Regressor.get_shaped_data (base.py:435-471),
 * which PymcGP.build_model hands to gp.marginal_likelihood("ml", X=X, y=y, ...) (GP.py:521,580). */
int gb2_set_train(gb2_handle* h, const double* X, int64_t N, int32_t D_in, const double* y);
int gb2_set_train_dev(gb2_handle* h, const double* dX, int64_t N, int32_t D_in, const double* dy);

/* Hyper-parameters = the point=self.MAP argument of Marginal.predict (GP.py:845-847) /
 * one L-BFGS-B iterate of pm.find_MAP (GP.py:811).                                              */
int gb2_set_kernel(gb2_handle* h, const gb2_kernel* k);

/* K(X,X)+Knoise+jitter build -> Cholesky -> v = L^-1 y.  Replaces the first half of
 * Marginal._build_conditional and MvNormal.logp's factorisation (GP.py:580, :845).             */
int gb2_factorize(gb2_handle* h);

/* log p(y | X, theta) = -N/2 log 2pi - sum log L_ii - 1/2 |v|^2 ; needs gb2_factorize.
 * Replaces the "ml" term evaluated by pm.find_MAP (GP.py:580,811).                             */
int gb2_mll(gb2_handle* h, double* out);

/* Value and gradient of log p(y | X, theta) w.r.t. every entry of gb2_kernel -- what pm.find_MAP's L-BFGS-B
 * (GP.py:809-811) obtains from PyTensor's reverse mode through Cholesky; here the closed form
 * 1/2 sum_ij (alpha alpha^T - K^-1)_ij dK_ij/dtheta evaluated on device.  Needs gb2_factorize.
 * grad_out has GB2_GRAD_LEN doubles:
 *   term t at offset t*GB2_GRAD_TERM:  [0, MAX_D) d/d ls[k] | [MAX_D] d/d eta | [MAX_D+1, +MAX_LIN) d/d c[l] |
 *                                      [MAX_D+1+MAX_LIN] d/d tau | then MAX_COREG tables of MAX_P*MAX_P: d/d coreg_B[f][p*P+q]
 *                                      (row-major with the factor's own P; B entries treated as independent)
 *   [GB2_GRAD_SIGMA] d/d sigma | [GB2_GRAD_NOISE_B, +MAX_P*MAX_P) d/d noise_B[p*P+q] (diagonal only)
 * The chain rule onto W, kappa (B = W W^T + diag kappa) and onto log-transformed variables is the caller's (tiny).   */
#define GB2_GRAD_TERM (GB2_MAX_D + 2 + GB2_MAX_LIN + GB2_MAX_COREG * GB2_MAX_P * GB2_MAX_P)
#define GB2_GRAD_SIGMA (GB2_MAX_TERMS * GB2_GRAD_TERM)
#define GB2_GRAD_NOISE_B (GB2_GRAD_SIGMA + 1)
#define GB2_GRAD_LEN (GB2_GRAD_NOISE_B + GB2_MAX_P * GB2_MAX_P)
int gb2_mll_grad(gb2_handle* h, double* mll_out, double* grad_out);
/* alpha = K^-1 y (length N) as computed by the gb2_mll_grad call on the current factorisation; the Kronecker-aware
 * multi-output solve (gumbi_b200/kron.py) needs it for the gradient w.r.t. the output Coregion and the output noise.   */
int gb2_get_alpha(gb2_handle* h, double* alpha_out);

/* Posterior mean/variance at Xs:(M,D_in) -- replaces PymcGP.predict (GP.py:837-849):
 * Marginal.predict(Xs, point=MAP, diag=True, pred_noise=with_noise).  Needs gb2_factorize.     */
int gb2_predict(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var);
int gb2_predict_dev(gb2_handle* h, const double* dXs, int64_t M, int32_t pred_noise, double* dmean, double* dvar);

/* gb2_factorize + gb2_predict in ONE pass -- what a single PymcGP.predict call costs in the reference, which rebuilds K,
 * re-factorises and solves every time (GP.py:845-847).  The prediction points are appended as extra rows of the
 * factor: their solve K(X*,X) L^-T is carried by the factorisation's own panel solves and trailing updates instead of a
 * separate triangular solve afterwards.  Same results as the two calls up to summation order; leaves the handle factorised.
 * fp64, replicated storage, M * round_up(N+1,128) * 8 bytes <= 8 GiB.  Multi-GPU: collective, each rank passes its own points. */
int gb2_factorize_predict(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var);
int gb2_factorize_predict_dev(gb2_handle* h, const double* dXs, int64_t M, int32_t pred_noise, double* dmean, double* dvar);

/* Posterior mean and FULL covariance at Xs:(M,D_in): cov:(M,M) row-major = K(X*,X*) - A^T A (+ noise diag if pred_noise).
 * Replaces the distribution gp.conditional(var_name, points_array) builds for draw_point_samples / draw_grid_samples
 * (GP.py:861-979; pm.gp.Marginal.conditional, diag=False).  M <= 32768.                                                */
int gb2_predict_full(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* cov);

/* Sparse FITC approximation.  Replaces pm.gp.MarginalSparse(approx="FITC") as gumbi builds it for sparse=True:
 * gp.marginal_likelihood("ml", X=X, Xu=Xu, y=y, sigma=sigma) (gumbi/regression/pymc/GP.py:571-578, :589-591) and the
 * MarginalApprox conditional behind predict (GP.py:845-847).  Xu:(m,D_in) host, row-major = the inducing points
 * (pm.gp.util.kmeans_inducing_points, GP.py:572; computed by the caller).  Uses the training set of gb2_set_train and the
 * kernel of gb2_set_kernel with its scalar sigma (a noise Coregion is refused: the reference reverts to the scalar sigma for
 * sparse models, GP.py:573-577).  fp64, single GPU, round_up(N,128) * round_up(m+1,128) * 8 bytes <= 8 GiB.
 * gb2_fitc_factorize returns 0, > 0 (Kuu + jitter I not positive definite: first failing pivot) or < 0 (error).             */
int gb2_fitc_factorize(gb2_handle* h, const double* Xu, int64_t m);
int gb2_fitc_mll(gb2_handle* h, double* out);
int gb2_fitc_predict(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var);

/* Test hooks: copy the dense objects back (row-major, n x n with n = N).  The strict upper
 * triangle is returned as zero for L and mirrored for K.                                        */
int gb2_get_K(gb2_handle* h, double* K_out);   /* rebuilds K+Knoise+jitter into scratch; O(N^2)  */
int gb2_get_L(gb2_handle* h, double* L_out);
int gb2_get_v(gb2_handle* h, double* v_out);   /* v = L^-1 y, length N                           */
/* Measurement aid: with set_option("trace", 1), every block step k of the factorisation writes six %globaltimer stamps (ns) at
 * out[6k + i]: panel stream -- 0 diagonal kernel eligible, 1 diagonal kernel done, 2 panel solve done, 3 next-column update done;
 * main stream -- 4 bulk trailing update eligible, 5 bulk trailing update (+ fused predict rows) done.  n = 6 * number of steps. */
int gb2_get_trace(gb2_handle* h, uint64_t* out, int64_t n);

/* Device-side milliseconds (CUDA events on the handle's stream) of the phases of the most recent
 * gb2_factorize / gb2_predict*: out[0]=feature prep, [1]=K build, [2]=Cholesky(+v),
 * [3]=K* build, [4]=triangular solve, [5]=mean/var reduction, [6]=#kernel launches of the last
 * factorize, [7]=#kernel launches of the last predict.                                          */
#define GB2_N_TIMINGS 8
int gb2_get_timings(gb2_handle* h, double* out);

/* Stream-ordered timing marks for benchmarks: gb2_mark(h, slot) records a CUDA event (slot 0..3) on the handle's
 * main stream; gb2_elapsed_ms synchronises on mark b and returns the device time between marks a and b.          */
int gb2_mark(gb2_handle* h, int slot);
int gb2_elapsed_ms(gb2_handle* h, int a, int b, double* ms);

/* Multi-GPU factorisation (north_star: "shard K by row-blocks across the 8xB200 box ... NCCL-over-NVLink ... panel broadcast
 * of the block Cholesky").  Nothing in the reference corresponds to this (it is single-process).  One process per GPU; each
 * process creates its own handle, rank 0 obtains an id with gb2_nccl_unique_id and ships the 128 bytes to the other ranks by
 * any means (the Python host side uses torch.distributed), then every rank calls gb2_dist_init.  Afterwards gb2_factorize is
 * collective: 128-row blocks of K are owned block-cyclically, each rank builds and updates its own rows, the diagonal block
 * is broadcast and the panel all-gathered every block step, and on return every rank holds the complete factor, so
 * gb2_predict* stays a local call (ranks predict disjoint slices of the grid).  NCCL is bound with dlopen at the first call. */
int gb2_nccl_unique_id(char* out128);
int gb2_dist_init(gb2_handle* h, int rank, int world, const char* unique_id128);
int gb2_dist_finalize(gb2_handle* h);
/* Gather `count` doubles from every rank (rank-major) on the handle's stream: the per-rank slices of the posterior.  */
int gb2_dist_allgather_dev(gb2_handle* h, const double* dsend, double* drecv, int64_t count);

/* Options; returns <0 if the name is unknown.
 *   "shard_storage" 0|1  multi-GPU: every rank stores only the row blocks of the factor it owns (N beyond one GPU's HBM);
 *                        gb2_predict* then becomes a collective over ALL points (same Xs on every rank, full result on every
 *                        rank).  Must be set identically on all ranks before gb2_factorize.  fp64 only.
 *   "p2p"           1|0  multi-GPU panel exchange through NVLink peer mappings (default) or NCCL broadcast + all-gather
 *   "tf32_nb"       0..16  GB2_TF32 factor-panel width in 128-column blocks (0 = auto); "tf32_leaf" 1..16 fp64 leaf width of the solve
 *   "lookahead"     1|0  panel look-ahead on a second stream;  "chain_on_panel", "kbuild_v1", "kbuild_persist", "kbuild_occ": ablations
 *   "solve_streams" 1..4 fp64 predict solve: row slabs of the prediction points on this many concurrent streams (default 4)
 *   "fused_group"   1|2|4|8  gb2_factorize_predict: column blocks per bulk update of the prediction rows (default 4)
 *   "trace"         0|1  record the per-step timeline read by gb2_get_trace (measurement aid, off by default)
 *   "fp64_panel"    -1..16  fp64 two-level blocking -- panels of this many 128-column blocks are factored with updates restricted to the
 *                        panel, then applied to the trailing matrix by one update of depth 128*value (-1 = auto: 16 from padded N >= 16384,
 *                        plain algorithm below; 0 = always plain)
 *   "bulk_persistent" 0|1|n  bulk trailing updates of the factorisation: one CTA per tile (0, default), persistent grid (1), or a persistent
 *                        grid of n CTAs
 *   process-wide ablations of the fp64 GEMM: "dgemm_tma" usage mask (7 = TMA-staged kernel everywhere, default; 0 = cp.async kernel),
 *   "dgemm_persistent" 0|1|2, "dgemm_deep" depth threshold of the 64x128 tile variant of the cp.async kernel, "dgemm_fence" 1|0
 *   (0 removes the generic->async proxy fence at the release of a shared-memory stage: reproduces the round-2 race, diagnostic only)
 * Environment: GB2_OPTS="name=value,..." applies options to every handle of the process at gb2_create (measurement scripts).  */
int gb2_set_option(gb2_handle* h, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* GUMBI_B200_H */
