/*
 * skb.h -- C ABI of the B200-native GP-posterior + acquisition engine ("skb" = scikit-optimize on Blackwell).
 *
 * The reference (scikit-optimize 0.9.0) is pure Python and has no FFI layer; its boundary for this path is duck
 * typing.  Every entry point below therefore names the reference interface it stands in for
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add at those call sites.
 *
 * Conventions
 *   - plain C, plain pointers and sizes, no torch / C++ types; all matrices row-major (C-contiguous) float64.
 *   - every function returns SKB_OK (0) or a negative skb_status; skb_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing throws across the ABI.
 *   - a skb_state owns device memory on ONE GPU and is not thread-safe; different states are independent.
 *   - "dev" entry points take DEVICE pointers on the state's GPU and enqueue on `stream` (a cudaStream_t passed
 *     as void*; NULL = the CUDA default stream, skb_state_stream(st) = the state's own stream) without
 *     synchronising the host.  Scratch buffers belong to the state: do not overlap two calls on one state.
 *     "host" entry points take HOST pointers, stage through pinned buffers, and return when results are in place.
 *   - there is no CPU fallback: without a usable sm_100 device skb_state_create fails with SKB_ERR_NO_DEVICE.
 */
#ifndef SKB_H_
#define SKB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_ABI_VERSION 1

typedef enum {
    SKB_OK = 0,
    SKB_ERR_INVALID = -1,       /* bad argument (NULL, negative size, unknown enum) */
    SKB_ERR_UNSUPPORTED = -2,   /* outside the supported grammar / size limits */
    SKB_ERR_NO_DEVICE = -3,     /* no CUDA device, or not sm_100 */
    SKB_ERR_CUDA = -4,          /* a CUDA runtime call failed; see skb_last_error() */
    SKB_ERR_NOMEM = -5,
    SKB_ERR_NOT_POSDEF = -6     /* skb_fit_finalize_host: K(theta) has no Cholesky factor */
} skb_status;

/* Stationary base kernels of the hot path: sklearn RBF.__call__ (sklearn kernels.py:1558-1570) and
 * Matern.__call__ (:1713-1729) as subclassed by skopt/learning/gaussian_process/kernels.py:68-200. */
typedef enum {
    SKB_KERNEL_RBF = 0,
    SKB_KERNEL_MATERN12 = 1,    /* nu = 0.5 */
    SKB_KERNEL_MATERN32 = 2,    /* nu = 1.5 */
    SKB_KERNEL_MATERN52 = 3,    /* nu = 2.5 */
    /* HammingKernel.__call__ (skopt/learning/gaussian_process/kernels.py:317-407), the kernel cook_estimator picks for
     * all-categorical spaces (skopt/utils.py:377-378): k(x, y) = exp(-sum_k length_scale[k] * [x_k != y_k]) on the
     * UNSCALED coordinates (length_scale[] of the descriptor carries the weights).  Value path only: the reference
     * has no gradient_x for it (utils.py:320-330 routes such models to acq_optimizer="sampling"), so
     * skb_score_grad_* and skb_fit_create return SKB_ERR_UNSUPPORTED for this kind. */
    SKB_KERNEL_HAMMING = 4
} skb_kernel_kind;

/* Acquisition functions of skopt/acquisition.py: gaussian_ei (:232-321), gaussian_pi (:149-229),
 * gaussian_lcb (:90-146).  Values follow _gaussian_acquisition (:20-87): -EI, -PI, LCB, all to be MINIMISED. */
/* Arithmetic of the O(n^2)-per-candidate variance contraction of skb_score_*.  Both meet the 1e-9 parity gate.
 *   FP64       : FP64 tensor-core MMA (DMMA), bit-reproducible reference path of this library (default).
 *   SPLIT_INT8_P7 : operands split into 7 signed 8-bit digit planes (55-bit fixed point per row / global scale),
 *                products accumulated exactly in int32 on tcgen05 (kind::i8), recombined in FP64 (csrc/ozaki.cuh);
 *                2e-14 of the prior variance from the exact contraction.
 *   SPLIT_INT8 : the same on the leading 6 planes (47-bit operands, 21 instead of 28 digit-plane products; 4e-12 of the
 *                prior variance; the kernel values feeding it are computed to that accuracy only, csrc/kcross.cuh) WHEN a
 *                self-check run once per state - 256 probe candidates scored with FP64 and with 6 planes - agrees to 1e-10
 *                on the variance (relative to the prior variance) and on the mean (relative to max(|mean|, y_std)) AND an
 *                a-priori bound computed from the state (skb_state_split_bounds) is below 2.5e-10; otherwise 7 planes.  The check runs on the first call with at least 32768 candidates (smaller calls on
 *                an undecided state contract 7 planes); skb_state_split_info reports the choice and the measured
 *                difference.  The gradient entry points always contract 7 planes. */
typedef enum { SKB_PRECISION_FP64 = 0, SKB_PRECISION_SPLIT_INT8 = 1, SKB_PRECISION_SPLIT_INT8_P7 = 2 } skb_precision;

typedef enum { SKB_ACQ_EI = 0, SKB_ACQ_PI = 1, SKB_ACQ_LCB = 2, SKB_NUM_ACQ = 3 } skb_acq_kind;

/* Everything GaussianProcessRegressor.predict reads from a fitted model (skopt gpr.py:315-356), flattened.
 *   kernel_(x, X_i) = amplitude * base(|| (x - X_i) / length_scale ||) + bias        (White is 0 off-diagonal)
 *   kernel_.diag(x) = amplitude + bias + white
 * L_inv is the inverse of the lower Cholesky factor est.L_ (row-major, entries above the diagonal ignored), so
 * that K_inv_ (gpr.py:223-224) = L_inv^T L_inv and  k^T K_inv_ k = || L_inv k ||^2.  The caller keeps ownership
 * of all host arrays; they are copied during skb_state_create. */
typedef struct {
    int32_t n_train;
    int32_t n_dims;
    int32_t kernel;              /* skb_kernel_kind */
    int32_t reserved;
    double amplitude;
    double bias;
    double white;
    double y_mean;               /* est.y_train_mean_ */
    double y_std;                /* est.y_train_std_ */
    const double* length_scale;  /* [n_dims] */
    const double* X_train;       /* [n_train * n_dims] == est.X_train_ */
    const double* alpha;         /* [n_train]          == est.alpha_ */
    const double* L_inv;         /* [n_train * n_train] lower triangular */
    const double* K_inv;         /* [n_train * n_train] == est.K_inv_ (symmetric); optional: NULL disables the
                                    gradient entry points (they contract with K_inv like gpr.py:353-354) */
} skb_state_desc;

/* acq_func_kwargs of _gaussian_acquisition (acquisition.py:32-35) and y_opt = min(yi) (optimizer.py:558). */
typedef struct {
    double y_opt;
    double xi;                   /* default 0.01 */
    double kappa;                /* default 1.96; ignored when kappa_inf != 0 */
    int32_t kappa_inf;           /* kappa == "inf": LCB = -std (acquisition.py:138,144) */
    int32_t acq_mask;            /* bit i set -> evaluate skb_acq_kind i */
    int32_t top_k;               /* number of (value, index) pairs to return per requested acquisition, >= 0 */
    int32_t precision;           /* skb_precision: how the variance contraction k^T K^-1 k is evaluated (value path) */
} skb_acq_params;

/* Outputs of one scoring call over m candidates.  Any pointer may be NULL (that output is skipped).
 *   mean/std      == predict(X, return_std=True)                       (gpr.py:316-343)
 *   acq[a]        == _gaussian_acquisition(X, est, y_opt, a)           (acquisition.py:20-87)
 *   topk_val/idx  == values[order[:top_k]], order[:top_k] + index_offset, order = lexsort by (value, index):
 *                    element 0 is np.argmin (first minimum, optimizer.py:564); the set is argsort[:k] (:570).
 *                    NaN sorts last; first_nan[a] receives the lowest global index holding NaN or -1. */
typedef struct {
    double* mean;                /* [m] */
    double* std;                 /* [m] */
    double* acq[SKB_NUM_ACQ];    /* [m] each */
    double* topk_val[SKB_NUM_ACQ];   /* [top_k] each */
    int64_t* topk_idx[SKB_NUM_ACQ];  /* [top_k] each */
    int64_t* first_nan[SKB_NUM_ACQ]; /* [1] each */
    int64_t* n_var_clamped;      /* [1] number of candidates whose variance was < 0 and set to 0 (gpr.py:336-340) */
} skb_score_out;

/* Gradient outputs (batched extension of gpr.py:345-362 + acquisition.py:218-227,305-319: the reference
 * evaluates one point per call; here every row of X gets its own gradient). */
typedef struct {
    double* mean;                /* [m] */
    double* std;                 /* [m] */
    double* mean_grad;           /* [m * n_dims] */
    double* std_grad;            /* [m * n_dims] */
    double* acq[SKB_NUM_ACQ];        /* [m] each */
    double* acq_grad[SKB_NUM_ACQ];   /* [m * n_dims] each */
} skb_grad_out;

typedef struct skb_state skb_state;

int skb_abi_version(void);
const char* skb_last_error(void);
/* Number of visible CUDA devices whose compute capability is 10.x; <0 on error. */
int skb_device_count(void);

/* Replaces: the attribute reads est.X_train_/alpha_/K_inv_/kernel_/y_train_mean_/y_train_std_ at gpr.py:315-356. */
int skb_state_create(const skb_state_desc* desc, int device, skb_state** out);
void skb_state_destroy(skb_state* st);
/* Add est.K_inv_ ([n_train * n_train], host) to a state created without it: enables the gradient entry points.
 * (A BO loop in 'sampling' mode never needs it; the Python mirror attaches it on the first gradient call.) */
int skb_state_attach_kinv(skb_state* st, const double* K_inv);
/* Digit planes SKB_PRECISION_SPLIT_INT8 contracts on this state (6 or 7; 0 = not decided yet, the self-check runs on the
 * first split-integer call) and the relative variance difference the self-check measured (-1 before it ran). */
int skb_state_split_info(skb_state* st, int32_t* planes, double* probe_rel_diff);
/* The a-priori half of that decision (-1 before it ran): bounds, from the row scales of L_inv, the kernel scale, |alpha|
 * and the size of the scaled training rows, on |var(6 planes) - var(exact)| / prior variance and on
 * |mean(6 planes) - mean(exact)| / y_std (8 standard deviations of the rounding model in csrc/skb.cu, split_error_bounds).
 * Six planes are contracted only when both bounds are <= 2.5e-10 AND the measured probe difference is <= 1e-10. */
int skb_state_split_bounds(skb_state* st, double* bound_var, double* bound_mean);
/* Run the self-check of SKB_PRECISION_SPLIT_INT8 now (on the state's own stream, waits for it) instead of on the first
 * large call; a no-op once the state has decided. */
int skb_state_decide_split(skb_state* st);
/* cudaStream_t owned by the state (so callers can record events on it). */
void* skb_state_stream(skb_state* st);
int skb_state_sync(skb_state* st);
/* Upper bound (bytes) of scratch memory the scoring calls may allocate; 0 keeps the default (2 GiB). */
int skb_state_set_scratch_limit(skb_state* st, uint64_t bytes);

/* Replaces: _gaussian_acquisition(X, model, y_opt, acq_func, acq_func_kwargs) for each requested acquisition
 * (optimizer.py:557-560, acquisition.py:20-87) composed with model.predict(X, return_std=True) (gpr.py:239-343),
 * plus np.argmin / np.argsort[:k] of the values (optimizer.py:564,570).  X: [m * n_dims] candidates. */
int skb_score_dev(skb_state* st, const double* X, int64_t m, int64_t index_offset,
                  const skb_acq_params* params, const skb_score_out* out, void* stream);
int skb_score_host(skb_state* st, const double* X, int64_t m, int64_t index_offset,
                   const skb_acq_params* params, const skb_score_out* out);

/* Replaces: model.predict(X) mean-only (optimizer.py:539 gp_hedge gains; gpr.py:315-318). */
int skb_predict_mean_dev(skb_state* st, const double* X, int64_t m, double* mean, void* stream);
int skb_predict_mean_host(skb_state* st, const double* X, int64_t m, double* mean);

/* Replaces: gaussian_acquisition_1D(x, model, y_opt, acq_func, acq_func_kwargs) (acquisition.py:7-17) and
 * predict(x, return_std=True, return_mean_grad=True, return_std_grad=True) (gpr.py:345-362), batched over rows. */
int skb_score_grad_dev(skb_state* st, const double* X, int64_t m, const skb_acq_params* params,
                       const skb_grad_out* out, void* stream);
int skb_score_grad_host(skb_state* st, const double* X, int64_t m, const skb_acq_params* params,
                        const skb_grad_out* out);

/* Replaces: the arithmetic of gaussian_ei / gaussian_pi / gaussian_lcb (acquisition.py:129-146, 211-229, 295-321)
 * for a posterior produced by ANY regressor (the duck-typed `model.predict(X, return_std=True[, grads])` boundary).
 * Host pointers; mu_grad / std_grad ([m * n_dims]) may be NULL (then acq_grad is not written). */
int skb_acq_from_posterior_host(int device, int64_t m, int32_t n_dims, const double* mu, const double* std,
                                const double* mu_grad, const double* std_grad, const skb_acq_params* params,
                                double* const acq[3], double* const acq_grad[3]);

/* Replaces: the per-second composition of _gaussian_acquisition for "EIps" / "PIps" (acquisition.py:61-80):
 * out = acq * exp(-mu_t + std_t^2/2) and its gradient, given the objective model's acquisition and the log-time
 * model's posterior.  Host pointers; pass out_grad == NULL to skip gradients. */
int skb_ps_combine_host(int device, int64_t m, int32_t n_dims, const double* acq, const double* acq_grad,
                        const double* mu_t, const double* std_t, const double* mu_t_grad, const double* std_t_grad,
                        double* out, double* out_grad);

/* Replaces: np.argmin(values) / np.argsort(values)[:k] (optimizer.py:564,570) for a value vector that lives on the
 * host (used after skb_ps_combine_host).  Same ordering rules as skb_score_out.topk_*. */
int skb_topk_host(int device, const double* values, int64_t m, int64_t index_offset, int32_t top_k, double* topk_val,
                  int64_t* topk_idx, int64_t* first_nan);

/* predict(X, return_cov=True) (gpr.py:315-325): mean [m] (may be NULL) and quad [m * m] = K_trans . K_inv . K_trans^T (row-major,
 * kernel units; the caller forms (kernel(X) - quad) * y_std^2).  Needs K_inv (skb_state_attach_kinv); all m points form one
 * batch, limited by the state's scratch (SKB_ERR_UNSUPPORTED beyond it: a joint covariance is asked for a handful of points). */
int skb_predict_cov_host(skb_state* st, const double* X, int64_t m, double* mean, double* quad);

/* The same ranking for a value vector that already lives on the state's GPU (e.g. the acq[] output of
 * skb_score_grad_dev): enqueued on `stream`, no host synchronisation; uses the state's scratch like the scoring calls. */
int skb_topk_dev(skb_state* st, const double* values, int64_t m, int64_t index_offset, int32_t top_k, double* topk_val,
                 int64_t* topk_idx, int64_t* first_nan, void* stream);

/* Cross-rank half of the same ranking when the candidate set is sharded over GPUs (the reference has no
 * multi-device path - this is np.argmin / np.argsort[:k] of optimizer.py:564,570 over the CONCATENATED shards).
 * A top-k record is n_acq * (2 top_k + 1) 8-byte words:
 *     [n_acq][top_k] values (f64) | [n_acq][top_k] global indices (i64, -1 = padding) | [n_acq] first NaN index (i64, -1)
 * i.e. the topk_val / topk_idx / first_nan pointers of skb_score_out aimed into ONE buffer, so that a rank's result for
 * all requested acquisitions travels in one all-gather.  `gathered` = n_records such records back to back (device);
 * `merged` (device) receives one record: per acquisition the top_k smallest (value, index) pairs of all records (ties to
 * the lowest index) and the lowest NaN index.  One launch, no host synchronisation; skb_score_dev with m == 0 pads its
 * outputs (NaN, -1, -1) so that empty shards merge as nothing. */
int skb_topk_merge_records_dev(int device, const void* gathered, int32_t n_records, int32_t n_acq, int32_t top_k,
                               void* merged, void* stream);

/* The same exchange + merge as ONE kernel over NVLink peer memory instead of an all-gather followed by
 * skb_topk_merge_records_dev: every rank stores its record into every peer's buffer, raises a flag there, waits for the
 * world's flags in its own buffer and merges (csrc/topk.cuh, topk_exchange_peer_kernel).  The caller provides the peer
 * mapping: each rank allocates skb_peer_buffer_bytes() of SYMMETRIC memory, zeroed before the first exchange, and passes
 * every rank's buffer as it is mapped on this device (torch.distributed._symmetric_memory: rendezvous(...).buffer_ptrs).
 * seq = 1, 2, 3, ... counts the exchanges of the group and must be the same on all ranks.  `merged` (device) receives
 * n_acq * (2 top_k + 1) + 1 words: the merged record, then a status word (0, or 1 if a peer did not arrive within
 * timeout_ms - the merged record is then meaningless).  Requires world <= 8 and n_acq * (2 top_k + 1) <= 1024. */
typedef struct skb_peer_group {
    void* buffers[8];     /* [world] device pointers, buffers[rank] is the own one */
    int32_t rank, world;
} skb_peer_group;
int64_t skb_peer_buffer_bytes(void);
int skb_topk_exchange_peer_dev(int device, const skb_peer_group* group, uint64_t seq, const void* record, int32_t n_acq,
                               int32_t top_k, void* merged, int32_t timeout_ms, void* stream);

/* ---- candidate generation on the device (the step that feeds the path: optimizer.py:552-553) ----
 * Replaces: `space.rvs(n_points, random_state=rng)` + `space.transform(...)` (skopt/space/space.py:874-903, 942-974;
 * transformers.py:243-276) for "normalize"-warped Real (uniform prior), Integer and Categorical dimensions.
 *
 * skb_mt19937_uniform_dev continues NumPy's legacy MT19937 stream: `key` / `pos` are RandomState.get_state()[1:3]
 * (host, in/out); `count` doubles are written to DEVICE memory d_out in exactly the order and with exactly the bits
 * of `rng.uniform(0, 1, count)`, and key / pos come back advanced (feed them to RandomState.set_state).
 * skb_candidates_generate_dev draws m * n_dims uniforms from the same stream (dimension by dimension, the reference's
 * draw order) and maps them to the candidate matrix d_X [m][n_dims] with the reference's floating-point operations,
 * bit for bit; key / pos come back advanced.
 * skb_candidates_generate_jump_dev is the parallel form of the same draw: one CTA per dimension, each started from its own
 * point of the stream by an MT19937 jump (csrc/rng.cuh).  The caller supplies `window` = the next 624 RAW words of the
 * generator and, for the columns j = 1 .. n_dims - 1, `polys[j - 1]` = z^(2 m j - 1) mod phi(z) as 624 words (bit b of word
 * k = coefficient of z^(32 k + b)); `tail` receives the raw words at stream offsets [2 m n_dims - 624, 2 m n_dims + 624)
 * counted from the first word drawn, from which the caller rebuilds the state NumPy would hold (skopt_b200/mt_jump.py
 * does both).  Needs 2 m >= 624.  The candidate matrix is bit-identical to skb_candidates_generate_dev's.
 * skb_gather_rows_dev copies the k rows idx[] of d_X to the host (arg-min row / L-BFGS starts). */
typedef enum { SKB_DIM_REAL_UNIFORM = 0, SKB_DIM_INTEGER_UNIFORM = 1 } skb_dim_kind;
typedef struct {
    int32_t kind;                /* skb_dim_kind; Categorical dimensions are INTEGER_UNIFORM over [0, n_categories-1] */
    int32_t reserved;
    double low, high;            /* bounds in the original space */
} skb_dim_desc;
int skb_mt19937_uniform_dev(int device, uint32_t* key, int32_t* pos, int64_t count, double* d_out, void* stream);
int skb_candidates_generate_dev(int device, uint32_t* key, int32_t* pos, int64_t m, int32_t n_dims,
                                const skb_dim_desc* dims, double* d_X, void* stream);
int skb_candidates_generate_jump_dev(int device, const uint32_t* window, const uint32_t* polys, int64_t m, int32_t n_dims,
                                     const skb_dim_desc* dims, double* d_X, uint32_t* tail, void* stream);
int skb_gather_rows_dev(int device, const double* d_X, const int64_t* idx, int32_t k, int32_t n_dims, double* rows,
                        void* stream);

/* ---- GP fit objective on the device (opt-in, the default fit stays on the host) ----
 * Replaces: sklearn GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True) (_gpr.py:541-657), the
 * function the L-BFGS-B hyper-parameter search of `fit` evaluates (reached from skopt gpr.py:195), for the kernel
 * grammar  amplitude * {RBF | Matern}(length_scale[n_dims]) + WhiteKernel(noise):
 *   theta = [log amplitude, log length_scale[0..n_dims), log noise]   (natural logs; log noise = -inf means no noise)
 *   lml   = -1/2 y.K^-1 y - sum log diag chol(K) - n/2 log 2 pi,  K = kernel(X) + alpha I
 *   grad  = d lml / d theta, [n_dims + 2] in the order of theta (the caller drops fixed hyper-parameters and sums the
 *           length-scale entries of an isotropic kernel).
 * A K that is not positive definite gives lml = -inf and a zero gradient, like the reference.  The dK/dtheta tensor
 * of the reference (n x n x (n_dims + 2)) is never formed.  Every step is this library's own kernel: the kernel matrix,
 * a blocked right-looking Cholesky, the triangular inverse by recursive doubling, alpha = L^-T (L^-1 y),
 * K^-1 = L^-T L^-1 and the gradient contraction (csrc/fit.cuh, csrc/chol.cuh); no cuSOLVER / cuBLAS is loaded. */
typedef struct skb_fit skb_fit;
typedef struct {
    int32_t n_train, n_dims;
    int32_t kernel;            /* skb_kernel_kind */
    int32_t reserved;
    double alpha;              /* GaussianProcessRegressor.alpha: jitter added to the diagonal of K */
    const double* X_train;     /* [n_train * n_dims] host, row-major */
    const double* y_train;     /* [n_train] host, already normalised (y_train_ of the estimator) */
} skb_fit_desc;
int skb_fit_create(const skb_fit_desc* desc, int device, skb_fit** out);
void skb_fit_destroy(skb_fit* fit);
int skb_fit_lml_host(skb_fit* fit, const double* theta, int32_t want_grad, double* lml, double* grad);
/* `count` hyper-parameter vectors at once (thetas [count][n_dims + 2], lml [count], grad [count][n_dims + 2]): the
 * evaluations run concurrently on separate streams, each bit-identical to a skb_fit_lml_host call.  Used to advance the
 * L-BFGS-B restarts of one fit (sklearn _gpr.py:316-337 runs them one after the other) in lock step. */
int skb_fit_lml_batch_host(skb_fit* fit, int32_t count, const double* thetas, int32_t want_grad, double* lml,
                           double* grad);
/* End of the search: factorise K(theta*) once more on the device and hand back what `fit` stores on the estimator.
 * Replaces the tail of sklearn's fit (_gpr.py:341-367: K = kernel_(X_train_) + alpha I, L_ = cholesky(K),
 * alpha_ = cho_solve(L_, y_train_)).  lml [1], alpha_out [n_train] and L_out [n_train * n_train] (row-major, lower
 * triangular, zeros above the diagonal like SciPy's) are host arrays; L_out may be NULL.  Returns SKB_ERR_NOT_POSDEF when
 * K is not positive definite (the reference raises LinAlgError there).  The factor stays on the device for
 * skb_state_create_from_fit. */
int skb_fit_finalize_host(skb_fit* fit, const double* theta, double* lml, double* alpha_out, double* L_out);
/* Scoring state straight from the finalised fit: replaces skopt gpr.py:223-224 (L_inv, K_inv_ = L_inv^T L_inv on the
 * host) and the upload of skb_state_create - the triangular inverse and K^-1 = L^-T L^-1 (cuBLAS trsm + syrk) and their packed
 * FP64 tiles / digit planes are produced on the device from the resident factor, X_train and alpha.  `desc` supplies
 * the scalars and length_scale of the PREDICTION kernel (after skopt silenced the noise term, gpr.py:199-220);
 * its X_train / alpha / L_inv / K_inv pointers are ignored.  Requires a successful skb_fit_finalize_host. */
int skb_state_create_from_fit(skb_fit* fit, const skb_state_desc* desc, skb_state** out);

/* Per-kernel device timing for roofline reports.  When enabled, every launch group is bracketed by CUDA events on
 * the launching stream.  skb_state_get_timing waits for them, sums the elapsed milliseconds per kind
 * (0 = K1 cross-kernel + mean, 1 = K2 triangular contraction + acquisition, 2 = top-k, 3 = gradient kernels),
 * counts the spans, and resets the accumulation. */
int skb_state_set_timing(skb_state* st, int enabled);
int skb_state_get_timing(skb_state* st, double* ms_by_kind, int64_t* launches_by_kind, int n_kinds);

/* Roofline denominators of this path measured on `device` now (MEASURED_PEAKS.json has HBM and bf16 only): the issue
 * rate of tcgen05.mma kind::i8 at M=128, N=256 (MACs per clock per SM from the SM cycle counter, and TOP/s = 2 x MACs / s
 * over all SMs from CUDA events, operands resident in shared memory) and the FP64 tensor (DMMA m8n8k4) rate in TFLOP/s.
 * Two ~1 ms kernels, run four times each; synchronises the device. */
int skb_measure_peaks(int device, double* i8_mac_per_clk_sm, double* i8_tops, double* fp64_dmma_tflops);

/* Diagnostics of the 6-plane contraction kernel (tools/k2o_diag.py, profiles/r02_k2o_diagnosis.md): `reps` launches of the kernel
 * alone over `tiles64` 64-candidate tiles of whatever the scratch holds, with parts switched off (dbg bit 0: no epilogue work,
 * bit 1: no operand copies, bit 2: no accumulator hand-over, bit 4: no stage barriers; 8: complete kernel, only timed).
 * Returns milliseconds per launch and the mean SM cycles a CTA spent.  Not used by any product path. */
int skb_debug_k2o(skb_state* st, int dbg, int64_t tiles64, int reps, double* ms_per_launch, double* cycles_per_cta);

/* Number of kernel launches issued through this library by the calling process so far (bench.py reports it). */
int64_t skb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SKB_H_ */
