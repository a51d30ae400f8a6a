/*
 * liesel_b200.h — C ABI of the B200-native hot path behind Liesel's model interface.
 *
 * What this boundary replaces in the reference (liesel 0.5.1, paths relative to the
 * reference checkout):
 *
 *   - the value the kernels read through `ModelInterface.log_prob`
 *     (src/liesel/goose/types.py:72-102) after `update_state`, i.e. what
 *     `LieselInterface.update_state/log_prob` (src/liesel/goose/interface.py:284-313)
 *     obtain from `Model.update_state` (src/liesel/model/model.py:2786-2865), and its
 *     gradient as taken by `jax.value_and_grad` inside NUTS/HMC
 *     (src/liesel/goose/nuts.py:193-194,254-260; src/liesel/goose/hmc.py:168-169,231-238);
 *   - the IWLS score / information / Cholesky of `IWLSKernel._score/_chol_info`
 *     (src/liesel/goose/iwls.py:206-243), for which the reference offers the
 *     `chol_info_fn` hook (src/liesel/goose/iwls.py:130-145);
 *   - the quadratic form of `tau2_gibbs_kernel` (src/liesel/model/distreg.py:277-297);
 *   - the B-spline basis of `basis_matrix` (src/liesel/contrib/splines.py:109-198).
 *
 * Conventions. Every pointer named in a *_desc struct or passed to a compute call is a
 * DEVICE pointer unless stated otherwise. Compute calls are asynchronous on `stream`,
 * allocate nothing, never synchronise and never abort: they return a status code
 * (0 = LSLB_OK). A plan is immutable after creation and may be used from several host
 * threads at once, each with its own workspace. Data pointers in the model description
 * are BORROWED: the caller owns them and must keep them alive while the plan is in use.
 *
 * Chains are batched: `C` chains are evaluated per call against one copy of the data.
 * `params` is [C x P] row-major (one flat position per chain, coefficient blocks in the
 * order of `lslb_model_desc.blocks`), `aux` is [C x A] row-major (one tau2 per penalised
 * block that has `tau2_slot >= 0`).
 */
#ifndef LIESEL_B200_H
#define LIESEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSLB_VERSION 100
#define LSLB_MAX_BLOCKS 16
#define LSLB_MAX_IWLS_P 64
#define LSLB_MAX_PEERS 16
#define LSLB_PEER_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

typedef struct CUstream_st* lslb_stream_t; /* == cudaStream_t */
typedef struct lslb_plan lslb_plan;
typedef struct lslb_peer lslb_peer; /* peer-memory exchange group of the ranks of one box */

enum lslb_status {
  LSLB_OK = 0,
  LSLB_ERR_INVALID = 1,     /* bad argument / unsupported model structure */
  LSLB_ERR_WORKSPACE = 2,   /* workspace too small */
  LSLB_ERR_CUDA = 3,        /* a CUDA runtime call failed (see lslb_last_cuda_error) */
  LSLB_ERR_UNSUPPORTED = 4, /* valid request this build does not implement */
  LSLB_ERR_COMM = 5         /* NCCL missing or a collective failed (see lslb_last_cuda_error) */
};

enum lslb_dtype { LSLB_F32 = 0, LSLB_F64 = 1 };

/* Response families (reference: tfd.<Dist>(...).log_prob(y) in Dist.update,
 * src/liesel/model/nodes.py:1139-1152). */
enum lslb_family {
  LSLB_NORMAL = 0,    /* predictor 0 = loc (identity link), predictor 1 = log(scale) */
  LSLB_BERNOULLI = 1, /* predictor 0 = logits */
  LSLB_POISSON = 2    /* predictor 0 = log_rate */
};

/* Additive terms of a predictor (reference: smooth Calc `jnp.dot(X, beta)`,
 * src/liesel/model/distreg.py:105,175, summed at :238-240). */
enum lslb_kind {
  LSLB_DENSE = 0,    /* X [n x ld] row-major, ncoef columns used */
  LSLB_SPLINE = 1,   /* cubic B-spline in banded form: start[n], weights[n x 4] */
  LSLB_GROUP = 2,    /* random intercept b[index[i]]; rows sorted by index */
  LSLB_INTERCEPT = 3 /* all-ones column; nothing is read */
};

enum lslb_prior {
  LSLB_PRIOR_NONE = 0,
  LSLB_PRIOR_NORMAL = 1, /* elementwise Normal(m, s): distreg.py:102 */
  LSLB_PRIOR_MVND = 2    /* MultivariateNormalDegenerate.from_penalty(0, tau2, K, rank):
                            distreg.py:166-172, distributions/mvn_degen.py:209-287,436-444 */
};

typedef struct lslb_block_desc {
  int32_t kind;       /* enum lslb_kind */
  int32_t predictor;  /* 0 or 1 */
  int32_t ncoef;      /* p (dense), k (spline), G (group), 1 (intercept) */
  int32_t prior;      /* enum lslb_prior */
  const void* data;   /* dense: X; spline: weights [n x 4]; else NULL */
  const int32_t* index; /* spline: start [n]; group: idx [n] (ascending); else NULL */
  int64_t ld;         /* dense: row stride of X in elements */
  double prior_m;     /* Normal prior mean */
  double prior_s;     /* Normal prior standard deviation */
  const void* K;      /* MVND: penalty matrix [ncoef x ncoef] row-major, model dtype */
  double rank;        /* MVND: rank of K (distreg.py:161) */
  double log_pdet;    /* MVND: log pseudo-determinant of K (mvn_degen.py:32-56) */
  int32_t tau2_slot;  /* MVND: column of `aux` holding tau2, or -1 for tau2_fixed */
  int32_t has_ig;     /* MVND: add InverseGamma(ig_a, ig_b).log_prob(tau2) (distreg.py:162) */
  double tau2_fixed;
  double ig_a;
  double ig_b;
  const void* center; /* spline: column means cbar[ncoef] (model dtype) or NULL. When set the term
                         is the column-centred basis (B - 1 cbar^T) beta, i.e. the dense centred
                         design matrix a Liesel user passes to add_np_smooth for identifiability,
                         kept in banded form plus a rank-one correction. */
} lslb_block_desc;

typedef struct lslb_model_desc {
  int32_t dtype;        /* enum lslb_dtype: type of y, data, K, params, aux and outputs */
  int32_t family;       /* enum lslb_family */
  int64_t n;            /* rows */
  const void* y;        /* response [n] */
  double fixed_scale;   /* Normal without a predictor-1 block: constant scale */
  double const_term;    /* cached log-probs of Dist nodes that never change */
  int32_t nblocks;
  int32_t reserved;
  const lslb_block_desc* blocks; /* HOST array of nblocks entries */
} lslb_model_desc;

/* flags for compute calls */
#define LSLB_SKIP_PRIOR 1u /* likelihood part only: for ranks > 0 under row sharding, where
                              partial results are summed across ranks afterwards */

int lslb_version(void);
const char* lslb_status_string(int status);
const char* lslb_last_cuda_error(void);

/* Builds the immutable evaluation plan. Synchronous (setup time). */
int lslb_plan_create(const lslb_model_desc* desc, lslb_plan** out);
void lslb_plan_destroy(lslb_plan* plan);
int lslb_plan_num_params(const lslb_plan* plan);
int lslb_plan_num_aux(const lslb_plan* plan);
/* LSLB_F32 / LSLB_F64 of the plan's data and parameters (-1 for a NULL plan); number of
 * coefficients of block `block` (-1 when out of range). Used by bindings to check operand shapes. */
int lslb_plan_dtype(const lslb_plan* plan);
int lslb_plan_block_ncoef(const lslb_plan* plan, int block);
/* Name of the kernel variant the plan dispatches to (static string). */
const char* lslb_plan_variant(const lslb_plan* plan);
/* Symbol name (as ncu prints it, without template arguments) of the row-streaming kernel that
 * lslb_logp_grad_* launches for C chains on this plan (static string; "" for a NULL plan). */
const char* lslb_plan_kernel_name(const lslb_plan* plan, int C);

/* Bytes of device scratch a compute call on C chains needs (max over all entry points). */
size_t lslb_workspace_bytes(const lslb_plan* plan, int C);

/*
 * Log-posterior and gradient for C chains. Replaces one `update_state` + `log_prob`
 * (+ reverse-mode gradient) of the reference per chain.
 *   logp [C]; grad [C x P] or NULL for value only (mh_step, src/liesel/goose/mh.py:48-50).
 */
int lslb_logp_grad_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                       float* logp, float* grad, uint32_t flags, void* ws, size_t ws_bytes,
                       lslb_stream_t stream);
int lslb_logp_grad_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                       double* logp, double* grad, uint32_t flags, void* ws, size_t ws_bytes,
                       lslb_stream_t stream);

/*
 * IWLS quantities of ONE coefficient block (p = ncoef <= LSLB_MAX_IWLS_P) for C chains, in
 * one pass over the rows:
 *   score [C x p]     = d logp / d beta_block              (iwls.py:206-216,315)
 *   info  [C x p x p] = -d2 logp / d beta_block^2, full symmetric, WITHOUT the ridge
 *                       (iwls.py:232; observed information)
 *   chol  [C x p x p] or NULL = lower Cholesky factor of info + 1e-6*mean(diag(info))*I
 *                       (iwls.py:233-238); every entry NaN when not positive definite,
 *                       like jnp.linalg.cholesky, so IWLSKernel._safe_chol still fires.
 */
int lslb_iwls_f32(const lslb_plan* plan, int C, int block, const float* params,
                  const float* aux, float* score, float* info, float* chol, uint32_t flags,
                  void* ws, size_t ws_bytes, lslb_stream_t stream);
int lslb_iwls_f64(const lslb_plan* plan, int C, int block, const double* params,
                  const double* aux, double* score, double* info, double* chol,
                  uint32_t flags, void* ws, size_t ws_bytes, lslb_stream_t stream);

/*
 * IWLS quantities of the RANDOM-INTERCEPT block for C chains (models with one group block next to at
 * most one dense block of <= 16 columns and intercepts: the S5 family). The reference's IWLSKernel would
 * form the dense G x G matrix -jacfwd(grad(log_prob)) (iwls.py:218-243); for b[idx[i]] it is diagonal:
 *   score     [C x G] = d logp / d b                      (iwls.py:206-216,315)
 *   info_diag [C x G] = diagonal of -d2 logp / d b^2 = per-group sum of the observed information weights
 *                       + 1 / s^2 of the Normal prior, WITHOUT the ridge (iwls.py:232)
 *   loglik_group [C x G] or NULL = the group's share of the log-posterior as a function of b_g: sum of its
 *                       rows' log-likelihood terms (without the row constants) - (b_g - m)^2 / (2 s^2).
 *                       Given the other blocks the posterior of b factorises over the groups, so these
 *                       are what a group-by-group Metropolis-Hastings step needs (iwls_sampler.py).
 * No workspace. LSLB_SKIP_PRIOR leaves the prior terms out (row-sharded ranks > 0): all three outputs are
 * sums over rows, so under row sharding the caller adds them over the ranks (there is no _peer / _nccl
 * variant of this entry point; sharding BY GROUP needs no sum at all).
 */
int lslb_iwls_group_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                        float* score, float* info_diag, float* loglik_group, uint32_t flags,
                        lslb_stream_t stream);
int lslb_iwls_group_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                        double* score, double* info_diag, double* loglik_group, uint32_t flags,
                        lslb_stream_t stream);

/* Ridge + batched Cholesky alone (for info summed across ranks): chol may alias info. */
int lslb_chol_f32(int C, int p, const float* info, float* chol, lslb_stream_t stream);
int lslb_chol_f64(int C, int p, const double* info, double* chol, lslb_stream_t stream);

/* beta_block^T K beta_block per chain, out [C] (tau2_gibbs_kernel, distreg.py:291). */
int lslb_quadform_f32(const lslb_plan* plan, int C, int block, const float* params,
                      float* out, lslb_stream_t stream);
int lslb_quadform_f64(const lslb_plan* plan, int C, int block, const double* params,
                      double* out, lslb_stream_t stream);

/*
 * Metropolis-Hastings step per chain (goose/mh.py:48-68), with the state selection fused:
 *   log_acc = prop_lp - cur_lp + log_corr   (log_corr == NULL: 0);  NaN -> -inf and code 90
 *   acc = min(exp(log_acc), 1);  moved = (u <= acc)
 * On acceptance cur_lp[c] = prop_lp[c] and (when given) row c of params [C x P] is replaced by row
 * c of params_prop. u, cur_lp, prop_lp, log_corr, acc are [C]; code int32 [C]; moved uint8 [C];
 * code / acc / moved / params may be NULL.
 */
int lslb_mh_step_f32(int C, int P, const float* u, float* cur_lp, const float* prop_lp,
                     const float* log_corr, float* params, const float* params_prop, int* code,
                     float* acc, unsigned char* moved, lslb_stream_t stream);
int lslb_mh_step_f64(int C, int P, const double* u, double* cur_lp, const double* prop_lp,
                     const double* log_corr, double* params, const double* params_prop, int* code,
                     double* acc, unsigned char* moved, lslb_stream_t stream);

/*
 * IWLS proposal tail per chain (iwls.py:323-327 and :335-336; iwls_utils.py:14-70):
 *   mu   = beta + (step^2 / 2) * solve(chol, score)
 *   prop = mu + (chol/step)^{-T} z              (only when z != NULL)
 *   logq = mvn_log_prob(x, mu, chol/step)       with x = prop (z != NULL) or x_eval
 * beta, score, z, x_eval, prop are [C x p]; chol [C x p x p]; step, logq [C].
 */
int lslb_iwls_propose_f32(int C, int p, const float* beta, const float* score,
                          const float* chol, const float* step, const float* z,
                          const float* x_eval, float* prop, float* logq, lslb_stream_t stream);
int lslb_iwls_propose_f64(int C, int p, const double* beta, const double* score,
                          const double* chol, const double* step, const double* z,
                          const double* x_eval, double* prop, double* logq,
                          lslb_stream_t stream);

/*
 * The same proposal tail for a DIAGONAL information matrix (random-intercept block; info_diag from
 * lslb_iwls_group_*): chol = diag(sqrt(info_diag + 1e-6 * mean(info_diag))) (ridge of iwls.py:233-238);
 * when an entry is not positive and finite the factor is NaN as a whole in the reference and
 * IWLSKernel._safe_chol falls back to the identity with error code 2 (iwls.py:245-290): code [C] (or NULL)
 * receives 2 for such chains, 0 otherwise. beta, score, info_diag, z, x_eval, prop are [C x p].
 * logq_elem [C x p] or NULL receives the summands of logq (one univariate normal log-density per entry).
 * local_ridge != 0: entry j is ridged by 1e-6 of ITSELF instead of 1e-6 of the mean, so that the proposal
 * of an entry depends on that entry alone (group-by-group acceptance; not a reference mode).
 */
int lslb_iwls_propose_diag_f32(int C, int p, const float* beta, const float* score,
                               const float* info_diag, const float* step, const float* z,
                               const float* x_eval, float* prop, float* logq, int* code,
                               float* logq_elem, int local_ridge, lslb_stream_t stream);
int lslb_iwls_propose_diag_f64(int C, int p, const double* beta, const double* score,
                               const double* info_diag, const double* step, const double* z,
                               const double* x_eval, double* prop, double* logq, int* code,
                               double* logq_elem, int local_ridge, lslb_stream_t stream);

/*
 * Banded cubic B-spline basis (splines.py:73-106 recursion, half-open base case):
 * start [n] = first column of the 4-wide window, weights [n x 4].
 * The span is decided by comparing x against the stored knot array.
 */
int lslb_spline_basis_f32(const float* x, int64_t n, const float* knots, int nknots,
                          int32_t* start, float* weights, lslb_stream_t stream);
int lslb_spline_basis_f64(const double* x, int64_t n, const double* knots, int nknots,
                          int32_t* start, double* weights, lslb_stream_t stream);
/*
 * Any order 0..3 of the same recursion (contrib/splines.py:109-198 is order-generic). The band
 * stays FOUR wide, the layout every evaluation kernel streams: the order + 1 weights that can be
 * non-zero lie inside [start, start + 3], start = clip(span - order, 0, k - 4) with
 * k = nknots - order - 1 >= 4 basis functions; the other entries of the window are exact zeros.
 */
int lslb_spline_basis_order_f32(const float* x, int64_t n, const float* knots, int nknots, int order,
                                int32_t* start, float* weights, lslb_stream_t stream);
int lslb_spline_basis_order_f64(const double* x, int64_t n, const double* knots, int nknots, int order,
                                int32_t* start, double* weights, lslb_stream_t stream);

/*
 * Row sharding over the GPUs of one box (no counterpart in the reference, which has no
 * multi-device path: src/liesel/goose/engine.py:442-448 is a vmap over chains on one device).
 * Every rank builds its plan on ITS rows and evaluates ALL C chains; the partial sums of
 * `_reduced_sum` (src/liesel/model/model.py:50-53), of its gradient and of the IWLS score /
 * information are summed over ranks before the call returns (same stream, asynchronous).
 * LSLB_SKIP_PRIOR is set internally on ranks > 0, so priors and constants enter once.
 * All ranks must issue the same sequence of sharded calls with the same C / block.
 *
 * (1) Peer memory over NVLink/NVSwitch. The evaluation's finalize kernel writes its partials
 * straight into an exchange buffer that every other rank has mapped through CUDA IPC; one more
 * kernel publishes an epoch flag to all peers, waits for theirs and sums the W partials through
 * P2P loads in RANK ORDER, so all ranks return bitwise identical values. No library, no host
 * synchronisation. A peer group is used by one host thread at a time and its calls cannot be
 * captured into a CUDA graph (LSLB_ERR_UNSUPPORTED while the stream is capturing).
 *   lslb_peer_create : allocates this rank's buffer for results of up to `max_elems` elements
 *                      (logp+grad: C*(P+1); IWLS: C*p*(p+1)) and writes its IPC handle
 *                      (LSLB_PEER_HANDLE_BYTES, HOST memory) to `handle`; synchronous.
 *   lslb_peer_connect: `handles` = the world's handles in rank order (HOST, world x 64 bytes),
 *                      exchanged by the caller (e.g. torch.distributed.all_gather); maps the peers.
 *   lslb_peer_timed_out: *timed_out != 0 when this rank's group has FAILED: a wait for a peer
 *                      exceeded ~10 s, a peer reported failure, or lslb_peer_abort was called. A
 *                      failure is fatal for the whole group: the failing rank writes a poison value
 *                      into its flag on every peer, so each rank's current or next sharded call
 *                      returns NaN outputs (never a silent partial sum) and every later call does
 *                      too, until lslb_peer_reset. Synchronous, for diagnostics.
 *   lslb_peer_abort  : fails the group from the host side (asynchronous on `stream`). The sharded
 *                      entry points do this themselves when their evaluation cannot be launched
 *                      after the exchange slot was claimed, and then return LSLB_ERR_COMM until reset.
 *   lslb_peer_reset  : clears flags, error word and epoch of THIS rank. Collective: every rank
 *                      must call it, with a host barrier before (no sharded call in flight
 *                      anywhere) and after (nobody publishes into a buffer that is not reset yet).
 */
int lslb_peer_create(int rank, int world, size_t max_elems, lslb_peer** out, void* handle);
int lslb_peer_connect(lslb_peer* peer, const void* handles);
void lslb_peer_destroy(lslb_peer* peer);
int lslb_peer_timed_out(lslb_peer* peer, int* timed_out);
int lslb_peer_abort(lslb_peer* peer, lslb_stream_t stream);
int lslb_peer_reset(lslb_peer* peer);
int lslb_logp_grad_peer_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                            float* logp, float* grad, uint32_t flags, void* ws, size_t ws_bytes,
                            lslb_peer* peer, lslb_stream_t stream);
int lslb_logp_grad_peer_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                            double* logp, double* grad, uint32_t flags, void* ws, size_t ws_bytes,
                            lslb_peer* peer, lslb_stream_t stream);
int lslb_iwls_peer_f32(const lslb_plan* plan, int C, int block, const float* params,
                       const float* aux, float* score, float* info, float* chol, uint32_t flags,
                       void* ws, size_t ws_bytes, lslb_peer* peer, lslb_stream_t stream);
int lslb_iwls_peer_f64(const lslb_plan* plan, int C, int block, const double* params,
                       const double* aux, double* score, double* info, double* chol,
                       uint32_t flags, void* ws, size_t ws_bytes, lslb_peer* peer,
                       lslb_stream_t stream);

/*
 * (2) NCCL. `comm` is an ncclComm_t of the caller (any NCCL 2.x; the symbols are resolved from
 * the process's libnccl.so.2 at first use, LSLB_ERR_COMM if absent). logp/grad (score/info) are
 * all-reduced in place with ncclSum on `stream`; chol is computed from the reduced information.
 * lslb_nccl_unique_id / lslb_nccl_comm_create / _destroy wrap ncclGetUniqueId (128 bytes, HOST) /
 * ncclCommInitRank / ncclCommDestroy for callers without an NCCL binding of their own.
 */
int lslb_nccl_unique_id(void* id128);
int lslb_nccl_comm_create(int rank, int world, const void* id128, void** comm);
void lslb_nccl_comm_destroy(void* comm);
int lslb_logp_grad_nccl_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                            float* logp, float* grad, uint32_t flags, void* ws, size_t ws_bytes,
                            void* comm, lslb_stream_t stream);
int lslb_logp_grad_nccl_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                            double* logp, double* grad, uint32_t flags, void* ws, size_t ws_bytes,
                            void* comm, lslb_stream_t stream);
int lslb_iwls_nccl_f32(const lslb_plan* plan, int C, int block, const float* params,
                       const float* aux, float* score, float* info, float* chol, uint32_t flags,
                       void* ws, size_t ws_bytes, void* comm, lslb_stream_t stream);
int lslb_iwls_nccl_f64(const lslb_plan* plan, int C, int block, const double* params,
                       const double* aux, double* score, double* info, double* chol,
                       uint32_t flags, void* ws, size_t ws_bytes, void* comm, lslb_stream_t stream);

/*
 * Fused per-leaf bookkeeping of the batched NUTS driver (liesel_b200/nuts.py; the trajectory code
 * the reference delegates to BlackJAX, src/liesel/goose/nuts.py:254-260). `ptrs` is a HOST array of
 * 23 device pointers: q_e, r_e, g_e, q_n, r_half, g_n, lp_n, eps, inv_mass, h0, q_s, g_s, lp_s,
 * logw_s, rsum_s, ck_r, ck_s, sum_acc, n_leaf, act, turn_s, div_s, u; `ints` = {C, P, D1, odd, lo, hi}.
 *   pre : r_half = r_e + eps/2 * g_e;  q_n = q_e + eps * inv_mass * r_half
 *   post: r_n, energy error, divergence, multinomial proposal update, momentum sum, checkpointed
 *         U-turn test (odd leaves: checkpoints lo..hi; even leaves store checkpoint hi), edge update,
 *         activity mask. Chains with act == 0 are left untouched.
 */
int lslb_nuts_leaf_pre_f32(void* const* ptrs, const int* ints, lslb_stream_t stream);
int lslb_nuts_leaf_pre_f64(void* const* ptrs, const int* ints, lslb_stream_t stream);
int lslb_nuts_leaf_post_f32(void* const* ptrs, const int* ints, double div_thr, lslb_stream_t stream);
int lslb_nuts_leaf_post_f64(void* const* ptrs, const int* ints, double div_thr, lslb_stream_t stream);

/*
 * State machine of the asynchronous batched NUTS driver (liesel_b200/nuts_async.py; f1 of SURVEY 8):
 * one launch between two batched log-posterior+gradient evaluations, one warp per chain. Every chain
 * walks its own tree (BlackJAX iterative NUTS as driven by goose/nuts.py:240-264), ends its transition
 * with dual averaging (goose/da.py:69-127), the mass-matrix window sums / epoch ends (goose/mm.py:19-35,
 * goose/warmup.py:12-88), the draw and optionally the Gibbs update of smoothing variances
 * (model/distreg.py:277-297), and starts the next one in the same launch; the evaluation at q_n
 * (ptrs[29]) is the caller's. ptrs: 58 device pointers in the order of csrc/nuts_async.cu's NutsAsync;
 * ints = {C, P, max_treedepth + 1, max_treedepth, warmup, samples, epochs, gibbs blocks, A};
 * dbls = {divergence threshold, target accept, gamma, kappa, t0}. The random numbers of a step are inputs.
 */
int lslb_nuts_async_step_f32(void* const* ptrs, const int* ints, const double* dbls, lslb_stream_t stream);
int lslb_nuts_async_step_f64(void* const* ptrs, const int* ints, const double* dbls, lslb_stream_t stream);

/*
 * Measurement hook (used by bench.py; no effect on results).
 * lslb_debug_set_events: the given cudaEvent_t pair is recorded on the call's stream immediately
 *   before and after the dominant row-streaming kernel of every following lslb_logp_grad_* /
 *   lslb_iwls_* call made from this host thread (NULL, NULL switches it off).
 * (The pipe-peak micro-benchmarks and the tcgen05 probe are NOT part of this library: see
 *  include/liesel_b200_bench.h / libliesel_b200_bench.so.)
 */
int lslb_debug_set_events(void* start_event, void* stop_event);
int lslb_debug_event_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LIESEL_B200_H */
