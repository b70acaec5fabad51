/*
 * pmf_b200.h -- C-ABI of libpmf_b200.so: the B200 (sm_100a) training hot path for the
 * probabilistic matrix-factorisation models of rogeliolopezcamara/prob-matrix-factorization.
 *
 * The reference has no FFI (it is pure Python); its "operator interface" for this path is
 * the body of each model class's fit/predict/evaluate methods.  Every entry point below
 * names the reference lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers + sizes; no torch / C++ types; all functions return int:
 *     0 = PMF_OK, negative = error; pmf_last_error() gives the message (thread local).
 *   - "d_" pointers are device memory owned by the caller (e.g. torch CUDA tensors);
 *     "h_" pointers are host memory.  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  No entry point synchronises unless it says so.
 *   - factor tables are row-major float32 with a row stride `ld` (in floats) that is a
 *     multiple of 8 (one 32-byte sector) and >= K; columns K..ld-1 are kept at zero.
 *   - ids are int32; rating values float32.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PMF_B200_H
#define PMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMF_OK 0
#define PMF_EINVAL (-1)   /* bad argument */
#define PMF_ECUDA (-2)    /* CUDA runtime error */
#define PMF_ENOMEM (-3)   /* allocation failed */
#define PMF_EUNSUPPORTED (-4)

/* ---- library ---------------------------------------------------------------------- */
int pmf_version(void);                 /* 100*major + minor */
const char* pmf_last_error(void);      /* message of the last failing call on this thread */
int pmf_device_count(int* count);      /* fails (PMF_ECUDA) when no CUDA driver/device */
int pmf_row_stride(int K);             /* smallest legal `ld` for K factors */
/* Blocking device->host copy of `bytes` bytes (diagnostics / tests; synchronises the stream). */
int pmf_copy_to_host(void* h_dst, const void* d_src, int64_t bytes, void* stream);
/* Kernel-variant selection for experiments ("gamma_group", "gamma_unroll", "gamma_chunk_reduce", "topn_growth"; 0 / -1 =
 * automatic).  Per calling thread (thread-local): it affects only launches made by the thread that set it. */
int pmf_tune(const char* key, int value);
/* Asynchronous copy between device buffers by the copy engines (cudaMemcpyAsync, unified addressing: a pointer may be a
 * peer GPU's memory mapped into this process) -- used to stage per-rank row sums on their owner without occupying SMs. */
int pmf_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
/* Return the library's cached (freed but retained) device memory of the current device to the driver.  The library
 * allocates rating lists and scratch from its own stream-ordered pool (it never changes the default pool's settings)
 * and keeps up to PMF_POOL_KEEP_MB (default 8192) MB across fits.  Synchronises the device. */
int pmf_trim(void);

/* ---- a1: observation grouping ("CSR build") ----------------------------------------
 * Replaces _build_index_lists (poisson_mf_cavi.py:73-84, hpf_cavi.py:97-107,
 * gaussian_mf_cavi.py:59-76, gaussian_mf_cavi_bias.py:69-86): observations grouped by row id,
 * each row keeping ORIGINAL order (stable).  Bit-exact: perm == argsort(key, kind="stable").
 *
 * A pmf_csr is an opaque, library-owned device structure for ONE orientation:
 *   row_ptr int32[n_rows+1], perm int32[nnz] (original observation index),
 *   col int32[nnz] (id of the other side), val float32[nnz],
 * plus the work decomposition used by the pass kernels: every row is cut into segments of
 * at most seg_len observations (an empty row is one empty segment).
 */
typedef struct pmf_csr pmf_csr;

/* d_key: ids to group by; d_other: ids of the other side; d_val: ratings; all length nnz,
 * device memory.  Synchronises the stream before returning (sizes are data dependent). */
int pmf_csr_build(const int32_t* d_key, const int32_t* d_other, const float* d_val,
                  int64_t nnz, int32_t n_rows, int32_t seg_len, void* stream, pmf_csr** out);
/* Copy of rows [row_begin,row_end) as a self-contained structure (row ids stay GLOBAL via
 * pmf_csr_row_offset); used to shard the rating list by nonzero across GPUs. */
int pmf_csr_slice(const pmf_csr* src, int32_t row_begin, int32_t row_end, void* stream, pmf_csr** out);
/* Global id of local row 0 (a list built from rebased keys `id - row_offset` of one rank's / one tile's row range). */
int pmf_csr_set_row_offset(pmf_csr* csr, int32_t row_offset);
int pmf_csr_free(pmf_csr* csr);
int64_t pmf_csr_nnz(const pmf_csr* csr);
int32_t pmf_csr_rows(const pmf_csr* csr);
int32_t pmf_csr_row_offset(const pmf_csr* csr);
int32_t pmf_csr_segments(const pmf_csr* csr);
int32_t pmf_csr_multi_rows(const pmf_csr* csr);   /* rows cut into >1 segment */
int32_t pmf_csr_seg_len(const pmf_csr* csr);
const int32_t* pmf_csr_row_ptr(const pmf_csr* csr);  /* device pointers, library owned */
const int32_t* pmf_csr_perm(const pmf_csr* csr);
const int32_t* pmf_csr_col(const pmf_csr* csr);
const float* pmf_csr_val(const pmf_csr* csr);
int64_t pmf_csr_device_bytes(const pmf_csr* csr);
/* nnz-balanced, row-aligned partition of the rows into `parts` ranges:
 * h_bounds[parts+1] (host) receives the row boundaries.  Synchronises. */
int pmf_csr_partition(const pmf_csr* csr, int32_t parts, int32_t* h_bounds);

/* Routing of a (u, i, rating) list to shards / tiles (new; no reference counterpart -- the reference is single-process).
 * pmf_count_keys: d_counts[k] = number of keys equal to k (int32[n_bins], zeroed by the call); synchronises.
 * pmf_coo_partition: STABLE partition of the triples by the bucket of u (by_item == 0) or i (by_item != 0), bucket b =
 * ids in [h_bounds[b], h_bounds[b+1]) (n_buckets <= 256; every id must lie in [h_bounds[0], h_bounds[n_buckets])).
 * Original order is kept inside a bucket, so grouping a bucket afterwards gives the reference's per-row order.
 * h_offsets[n_buckets+1] receives the start of each bucket in the outputs.  Synchronises. */
int pmf_count_keys(const int32_t* d_key, int64_t n, int32_t n_bins, int32_t* d_counts, void* stream);
int pmf_coo_partition(const int32_t* d_u, const int32_t* d_i, const float* d_x, int64_t n, int32_t by_item,
                      const int32_t* h_bounds, int32_t n_buckets, int32_t* d_u_out, int32_t* d_i_out, float* d_x_out,
                      int64_t* h_offsets, void* stream);

/* ---- a2: initial state (host) ---------------------------------------------------------------------------------------
 * h_out[k] = offset + scale * E_k, k < n, where E_k is the stream np.random.Generator(PCG64).gamma(1.0, 1.0) produces from
 * the given bit-generator state -- the reference's initial draws `a + rng.gamma(1.0, 0.1, size=(R, K))`
 * (poisson_mf_cavi.py:62-63, hpf_cavi.py:71-80), bit for bit, computed by `threads` host threads (<= 0: all cores).
 * state / inc: the two 128-bit PCG64 words as {high, low} uint64 pairs (`rng.bit_generator.state["state"]`);
 * new_state_hi_lo receives the state after the n variates (what NumPy's generator would hold).  Host memory only. */
int pmf_numpy_exponential_fill(const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo, double scale, double offset,
                               int64_t n, double* h_out, int32_t threads, uint64_t* new_state_hi_lo);

/* Multi-threaded host conversions of the reference's input dtypes (DataFrame columns are int64 / float64,
 * load_data.py:93-105; NumPy state is float64): int64 -> int32 with the range of the input (callers reject ids outside
 * [0, 2^31-2]) and float64 -> float32 (round to nearest, what NumPy's astype does).  Host memory only. */
int pmf_host_i64_to_i32(const int64_t* h_in, int64_t n, int32_t* h_out, int64_t* h_min, int64_t* h_max, int32_t threads);
int pmf_host_f64_to_f32(const double* h_in, int64_t n, float* h_out, int32_t threads);
/* h_out = h_a / h_b element-wise (h_b == NULL: / b_scalar): the initial expectations E = shape / rate (hpf_cavi.py:91-95). */
int pmf_host_divide_f64(const double* h_a, const double* h_b, double b_scalar, int64_t n, double* h_out, int32_t threads);

/* ---- a3/a4: Gamma-Poisson row pass (Poisson MF and HPF-CAVI) ------------------------
 * Replaces the per-row loops poisson_mf_cavi.py:135-164 / :173-194 (+ E=a/b :167,:197) and
 * hpf_cavi.py:126-151 / :162-185 (+ :153, :158-159, :187, :192-193).  For every row r of `csr`
 * (global row R = row_offset + r), with observations t in original order:
 *     rate_t = max(<E_self[R], E_oth[col_t]>, 1e-10)
 *     shp[R] = shape_prior + E_self[R] * sum_t (val_t / rate_t) * E_oth[col_t]
 *     rte[R] = rate_prior(R) + sum_t E_oth[col_t]         rate_prior(R) = d_rate_prior_vec ?
 *     E_self[R] <- shp[R] / rte[R]   (in place; Jacobi: only row R itself reads E_self[R])
 *                                                          d_rate_prior_vec[R] : rate_prior
 * Rows without observations get (shape_prior, rate_prior(R)).
 * HPF hyper update fused when d_hyper_rate != NULL (hpf_cavi.py:158, :192):
 *     d_hyper_rate[R] = hyper_rate_prior + sum_k E_self[R,k];  d_hyper_mean[R] = hyper_shape / that
 * d_hyper_mean may alias d_rate_prior_vec (the user pass reads old E_xi, writes new E_xi).
 * d_shp / d_rte may be NULL to skip materialising the Gamma parameters.
 * No per-observation tensor is written: allocations live in registers only.
 * d_workspace: pmf_gamma_pass_workspace_bytes() bytes of device scratch (partial sums of rows
 * cut into several segments).
 */
int64_t pmf_gamma_pass_workspace_bytes(const pmf_csr* csr, int32_t ld);
int pmf_gamma_pass(const pmf_csr* csr, int32_t K, int32_t ld,
                   const float* d_E_oth, float* d_E_self, float* d_shp, float* d_rte,
                   float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                   float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                   void* d_workspace, void* stream);

/* Tiled / sharded form of the pass (new; the reference is single-process and visits each row's ratings in one go).
 * The ratings of a pass are split by the id range of the OTHER side into "tiles" -- one pmf_csr per tile -- so that the
 * part of E_oth a launch gathers from stays resident in the 126 MB L2; on several GPUs every rank additionally holds
 * only the ratings of its own user range.  The row sums [sum_t (val_t/rate_t) E_oth[col_t] | sum_t E_oth[col_t]]
 * accumulate across the launches in d_acc[(R - acc_row_base)][2*ld]:
 *   acc_flags & PMF_ACC_IN   add the sums parked by earlier tiles before using this tile's;
 *   acc_flags & PMF_ACC_OUT  park the running sums instead of finishing the row.
 * flags 0 = pmf_gamma_pass; first tile OUT, middle tiles IN|OUT, last tile IN (finishes the rows exactly as
 * pmf_gamma_pass does) -- or IN|OUT everywhere followed by pmf_gamma_combine (several GPUs).  A row with no rating in
 * a tile leaves the running sums untouched (IN|OUT), zeroes them (OUT) or finishes from them (IN). */
#define PMF_ACC_IN 1
#define PMF_ACC_OUT 2
int pmf_gamma_pass_acc(const pmf_csr* csr, int32_t K, int32_t ld,
                       const float* d_E_oth, float* d_E_self, float* d_shp, float* d_rte,
                       float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                       float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                       void* d_workspace, float* d_acc, int32_t acc_row_base, int32_t acc_flags, void* stream);
/* The per-iteration "sufficient statistics combine" of the multi-GPU item pass, fused with the Gamma
 * update and the replication of the new rows.  For rows [row_begin,row_end) -- the ones this rank owns -- the row sums
 * are read from d_mc_acc, an NVSwitch MULTICAST alias of every rank's d_acc: one multimem.ld_reduce.add per 16 bytes, the
 * ranks' partial sums are added inside the switch; the update of pmf_gamma_pass follows; the new row of E_self is
 * stored locally and, when d_mc_E_self (multicast alias of E_self) is given, to every replica with one multimem.st.
 * d_mc_acc == NULL: d_acc already holds complete sums (after an NCCL all-reduce; every rank then updates every row).
 * Callers barrier across ranks before (all partial sums parked) and after (all replicas written). */
int pmf_gamma_combine(int32_t row_begin, int32_t row_end, int32_t K, int32_t ld, const float* d_acc,
                      const float* d_mc_acc, int32_t acc_row_base, float* d_E_self, float* d_mc_E_self, float* d_shp,
                      float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                      float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior, void* stream);

/* The same combine with the ranks' row sums STAGED in local memory: every rank copies, with the copy engines (no SMs, so
 * the transfers overlap the pass kernels of the next item chunk), the sums of the rows another rank owns into that rank's
 * staging table d_stage[n_src][src_stride floats] (slot s = source rank s; row q of a slot = the owner's q-th row,
 * 2*ld floats).  For rows [row_begin,row_end) -- slot rows stage_row0 .. -- the sums are added in rank order, the local
 * ones (d_acc) taking slot self_rank's place, so the result is deterministic and independent of the owner; update and
 * replication as in pmf_gamma_combine. */
int pmf_gamma_combine_staged(int32_t row_begin, int32_t row_end, int32_t K, int32_t ld, const float* d_acc,
                             int32_t acc_row_base, const float* d_stage, int32_t n_src, int64_t src_stride,
                             int32_t self_rank, int32_t stage_row0, float* d_E_self, float* d_mc_E_self, float* d_shp,
                             float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                             float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                             void* stream);

/* Extended Poisson MF (poisson_mf_extended_cavi.py:110-164 user side, :169-216 item side):
 * x_ui ~ Poisson(phi_u psi_i theta_u . beta_i).  One pass over the rows of `csr`:
 *   shp[r] = a0 + sum_t (x_t / (E_oth[c_t] . E_self[r])) E_oth[c_t] * E_self[r]      (raw dot product, not clamped: :142)
 *   rte[r] = b0 + sum_t scale_oth[c_t] E_oth[c_t]                                    (:147-148)
 *   E_self[r] = shp / rte                                                            (:160)
 *   scale_shp[r] = a0 + sum_t x_t                                                    (:153)
 *   scale_rte[r] = b0 + sum_t scale_oth[c_t] (E_oth[c_t] . E_self_new[r])            (:163-164, uses the NEW row mean)
 *   scale_mean[r] = scale_shp / scale_rte                                            (:167)
 * Rows without observations get the prior (a0, b0) in shp/rte/scale_shp/scale_rte and keep E_self / scale_mean
 * (:112-118).  d_shp / d_rte may be NULL.  Workspace: pmf_gamma_pass_workspace_bytes.  Single GPU. */
int pmf_gamma_pass_ext(const pmf_csr* csr, int32_t K, int32_t ld, const float* d_E_oth, const float* d_scale_oth,
                       float* d_E_self, float* d_shp, float* d_rte, float* d_scale_shp, float* d_scale_rte,
                       float* d_scale_mean, float a0, float b0, void* d_workspace, void* stream);

/* out[r][:] = scale[r] * F[r][:] for rows x ld floats (folds phi / psi into the factor tables so that the extended
 * model's predictions phi_u psi_i theta_u . beta_i go through pmf_predict / pmf_eval_stats). */
int pmf_scale_rows(const float* d_F, const float* d_scale, int64_t rows, int32_t ld, float* d_out, void* stream);

/* ---- a11: textbook-HPF extras (no reference code exists: docs/Models.tex:583-726 only; PARITY UNPINNED) ----
 * pmf_gamma_geomean: G = exp(psi(shape)) / rate (the geometric-mean table, exp(E log x)), padding columns 0.
 * pmf_gamma_pass_digamma: pmf_gamma_pass with the multinomial allocation of docs/Models.tex:652-664,
 *     phi_k ∝ G_self_k G_oth_k; shape gets sum_t x_t phi_tk, rate keeps the observed-only sum of E_oth
 *     (hpf_cavi.py:151); writes E_self AND G_self.  K <= 128.
 * pmf_hpf_elbo: evidence lower bound of observed-only HPF; d_out6 (float64, zeroed by the call) receives
 *     [0] sum_obs x log sum_k G_th G_be - lgamma(x+1) - sum_k E_th E_be   [1] E log p(theta|xi)  [2] E log p(beta|eta)
 *     [3] E log p(xi)  [4] E log p(eta)  [5] entropies;   ELBO = their sum.  Row terms cover users
 *     [user_begin,user_end) and items [item_begin,item_end) (a rank's owned rows), the likelihood the ratings
 *     of `by_user`. */
int pmf_gamma_geomean(const float* d_shp, const float* d_rte, int64_t rows, int32_t K, int32_t ld, float* d_G,
                      void* stream);
int pmf_gamma_pass_digamma(const pmf_csr* csr, int32_t K, int32_t ld, const float* d_G_oth, const float* d_E_oth,
                           float* d_G_self, float* d_E_self, float* d_shp, float* d_rte, float shape_prior,
                           float rate_prior, const float* d_rate_prior_vec, float* d_hyper_rate, float* d_hyper_mean,
                           float hyper_shape, float hyper_rate_prior, void* d_workspace, void* stream);
int pmf_hpf_elbo(const pmf_csr* by_user, int32_t K, int32_t ld, const float* d_E_theta, const float* d_E_beta,
                 const float* d_G_theta, const float* d_G_beta, const float* d_shp_theta, const float* d_rte_theta,
                 const float* d_shp_beta, const float* d_rte_beta, const float* d_rate_xi, const float* d_rate_eta,
                 int32_t user_begin, int32_t user_end, int32_t item_begin, int32_t item_end, float a, float a_prime,
                 float b_prime, float c, float c_prime, float d_prime, double* d_out6, void* stream);

/* Dense U V^T top-n scoring (evaluation; the one GEMM-shaped step, so the one use of tensor cores).
 * For each of `batch_rows` user rows (row b = d_user_rows ? d_user_rows[b] : b of d_F_user) the n best items by
 *   score = float32 chain  s = s + u[k]*v[k], k = 0..K-1 (separate multiply and add),
 * ranked by (score descending, item index ascending); d_idx / d_score are [batch_rows][n].
 * tensor_cores != 0: approximate scores by tcgen05.mma (bf16 operands, fp32 accumulation in TMEM), then an exact
 * fp32 re-score of a provably sufficient candidate set, so indices are identical to tensor_cores == 0 (exact
 * CUDA-core scoring).  d_stats (int32[2], may be NULL): rows that fell back to exact scoring, candidates re-scored.
 * PARITY UNPINNED: the reference has no top-n code; semantics = oracle/pmf_oracle.py::topn. */
int64_t pmf_topn_workspace_bytes(int64_t batch_rows, int32_t n_items, int32_t K);   /* enough for every mode and n */
/* Exact requirement of one mode.  tensor_cores: 0 exact CUDA-core scoring; 1 tcgen05, fused (n <= 256 and K <= 160: a
 * persistent warp-specialised kernel tests every score against a per-row threshold in the MMA epilogue, so the
 * batch x n_items score matrix never reaches HBM and the workspace is ~32 KB per row), unfused otherwise;
 * 2 tcgen05 unfused (score matrix through HBM, 4*n_items bytes per row; kept as the comparison point). */
int64_t pmf_topn_workspace_bytes_ex(int64_t batch_rows, int32_t n_items, int32_t K, int32_t n, int32_t tensor_cores);
int pmf_topn(const float* d_F_user, const int32_t* d_user_rows, int64_t batch_rows, const float* d_F_item,
             int32_t n_items, int32_t K, int32_t ld, int32_t n, int32_t tensor_cores, int32_t* d_idx, float* d_score,
             void* d_workspace, int64_t workspace_bytes, int32_t* d_stats, void* stream);

/* ---- a8-a10: predict and evaluation -------------------------------------------------
 * predict (poisson_mf_cavi.py:221-241, hpf_cavi.py:215-231, gaussian_mf_cavi_bias.py:291-316,
 * hpf_pytorch.py:66-69,186-195): pred = <F_user[u], F_item[i]> (+ b_user[u] + b_item[i]) for
 * u < n_users and i < n_items, else 0; then + global_mean.  d_b_user/d_b_item may be NULL.
 * softplus != 0 applies torch's softplus (threshold 20) to both rows first (HPF_PyTorch).
 * d_pred is float64[n] (the dot product is accumulated in float64).
 */
int pmf_predict(const int32_t* d_users, const int32_t* d_items, int64_t n,
                const float* d_F_user, int32_t n_users, const float* d_F_item, int32_t n_items,
                int32_t K, int32_t ld, const float* d_b_user, const float* d_b_item,
                float global_mean, int32_t softplus, double* d_pred, void* stream);
/* Fused predict + error statistics (evaluate_rmse / evaluate_macro_mae, metrics.py:6-16,37-51,
 * PoissonLogPredictiveLikelihood metrics.py:53-66).  d_label[n] holds each row's index among
 * the distinct true values (0..n_labels-1, n_labels <= 64).  drop_invalid != 0 skips rows with
 * unseen ids (Gaussian, gaussian_mf_cavi_bias.py:323-324); otherwise they predict 0 and count.
 * d_out (float64, zeroed by the call) layout:
 *   [0] count  [1] sum (y-p)^2  [2] sum |y-p|  [3] sum y*log(max(p,1e-10)) - p - lgamma(y+1)
 *   [4 .. 4+n_labels) per-label sum |y-p|   [4+n_labels .. 4+2 n_labels) per-label count
 * (global_mean shifts y_true and the prediction alike, so it cancels in every statistic but [3].)
 * GaussianLogPredictiveLikelihood (metrics.py:18-35, called with the factor means only: no biases, global_mean 0) follows
 * from [0] and [1]:  -0.5 [0] log(2 pi s^2) - [1] / (2 s^2)  with s^2 = sigma^2 (the reference squares its `sigma` argument).
 * The reduction is deterministic (bit-identical results for identical inputs: the early-stopping rule compares them).
 * d_scratch: pmf_eval_stats_scratch_bytes() bytes of device memory, 8-byte aligned, private to the call's stream. */
int64_t pmf_eval_stats_scratch_bytes(void);
int pmf_eval_stats(const int32_t* d_users, const int32_t* d_items, const float* d_y,
                   const int32_t* d_label, int32_t n_labels, int64_t n,
                   const float* d_F_user, int32_t n_users, const float* d_F_item, int32_t n_items,
                   int32_t K, int32_t ld, const float* d_b_user, const float* d_b_item,
                   float global_mean, int32_t drop_invalid, double* d_out, void* d_scratch, void* stream);

/* ---- f2: device-side training loop (sweeps + validation + early stopping as one CUDA graph) -----------------------
 * Replaces the per-iteration host decision of hpf_cavi.py:196-211 / poisson_mf_cavi.py:200-217 /
 * gaussian_mf_cavi_bias.py:268-284.  pmf_loop_begin puts `stream` (not the legacy default stream) into capture INTO the
 * body of a WHILE conditional graph node; the caller then enqueues ONE iteration on that stream (pass kernels,
 * pmf_eval_stats into d_eval_out, ...) and finally pmf_loop_decide, whose kernel -- per executed iteration -- increments
 * *d_iter, stores rmse = sqrt(d_eval_out[1]/d_eval_out[0]) in d_history[*d_iter - 1] and keeps the loop going while
 * *d_iter < max_iter and the stopping rule has not fired (has_tol != 0; improvement = previous - current rmse):
 *   rule 0: stop when improvement < tol (Poisson MF / HPF, fires on negative improvement);
 *   rule 1: stop when 0 <= improvement < tol (Gaussian MF).
 * d_eval_out == NULL: no validation, exactly max_iter iterations.  pmf_loop_end ends the capture and instantiates;
 * pmf_loop_run launches the whole loop asynchronously on a stream (*d_iter must be zeroed by the caller).
 * PMF_EUNSUPPORTED when the driver lacks conditional graph nodes (callers fall back to a host loop). */
typedef struct pmf_loop pmf_loop;
int pmf_loop_begin(void* stream, pmf_loop** out);
int pmf_loop_decide(pmf_loop* loop, const double* d_eval_out, int32_t rule, double tol, int32_t has_tol, int32_t max_iter,
                    int32_t* d_iter, double* d_history, void* stream);
int pmf_loop_end(pmf_loop* loop);
int pmf_loop_run(pmf_loop* loop, void* stream);
int pmf_loop_free(pmf_loop* loop);

/* ---- a5: Gaussian MF CAVI ---------------------------------------------------------------
 * Tables per side: means m[R, ld] (ld = pmf_row_stride(K)); covariances V and second moments
 * Q = V + m m^T as packed lower triangles [R, ldq] (ldq = pmf_gauss_packed_stride(K); element (i,j),
 * i >= j, at i(i+1)/2 + j); biases b[R] (NULL for the no-bias model gaussian_mf_cavi.py).
 *
 * pmf_gauss_factor_pass replaces gaussian_mf_cavi_bias.py:132-165 / :170-201 (gaussian_mf_cavi.py:121-178):
 *     S = sum_t Q_oth[col_t];  V[R] = inv(I/eta2 + S/sigma2)   (np.linalg.inv -> float64 Cholesky inverse)
 *     m[R] = V[R] sum_t (val_t - b_self[R] - b_oth[col_t]) m_oth[col_t] / sigma2;  Q[R] = V[R] + m m^T
 * rows without observations keep their state.  In place on the self tables (Jacobi: a row reads only
 * its own bias on the self side).  d_workspace: pmf_gauss_workspace_bytes() bytes.
 * pmf_gauss_bias_pass replaces :206-232 / :237-263:
 *     b_self[R] = sum_t (val_t - b_oth[col_t] - <m_self[R], m_oth[col_t]>) / sigma2 / (1/eta_b2 + n_R/sigma2)
 * (float64 residual sums per segment, combined per row in segment order; d_workspace: the same
 * pmf_gauss_workspace_bytes() buffer, 8-byte aligned, free to reuse between the passes).
 */
int pmf_gauss_packed_stride(int K);
int64_t pmf_gauss_workspace_bytes(const pmf_csr* csr, int32_t K);
int pmf_gauss_factor_pass(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_Q_oth,
                          const float* d_b_oth, float* d_m_self, float* d_V_self, float* d_Q_self,
                          const float* d_b_self, float sigma2, float eta2, void* d_workspace, void* stream);
int pmf_gauss_bias_pass(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_m_self,
                        const float* d_b_oth, float* d_b_self, float sigma2, float eta_b2, void* d_workspace, void* stream);

/* Sharded forms (several GPUs, ratings sharded along user ranges; SURVEY.md §8e "Gaussian shards identically"): the rows of
 * `csr` get ratings from every rank, so the pass is split around the caller's NCCL all-reduce of the row statistics.
 *   phase 1: accumulate this rank's ratings into d_row_sums[n_rows][ldq + ld] (factor pass; packed second moments |
 *            residual-weighted means) or d_row_resid[n_rows] (bias pass, float64);
 *   (caller: all-reduce SUM over the ranks)
 *   phase 2: the row update of pmf_gauss_factor_pass / pmf_gauss_bias_pass from those sums; d_counts[n_rows] = ratings of
 *            each row over ALL ranks (rows with none keep their state; the bias precision uses it). */
int pmf_gauss_factor_pass_sharded(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_Q_oth,
                                  const float* d_b_oth, float* d_m_self, float* d_V_self, float* d_Q_self,
                                  const float* d_b_self, float sigma2, float eta2, void* d_workspace, float* d_row_sums,
                                  const int32_t* d_counts, int32_t phase, void* stream);
int pmf_gauss_bias_pass_sharded(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_m_self,
                                const float* d_b_oth, float* d_b_self, float sigma2, float eta_b2, void* d_workspace,
                                double* d_row_resid, const int32_t* d_counts, int32_t phase, void* stream);

/* ---- a6/a7: gradient-based HPF (HPF_PyTorch) ----------------------------------------------
 * Parameters keep the reference's shapes (hpf_pytorch.py:39-48): theta_raw (N,K), beta_raw (M,K) row-major
 * with row stride K, xi_raw (N), eta_raw (M), all float32 and unconstrained (softplus applied inside).
 *
 * pmf_hpf_map_loss_grad replaces HPF_PyTorch.loss (hpf_pytorch.py:71-184) AND its autograd backward:
 * adds the batch loss (sum, not mean) to *d_loss (float64) and ACCUMULATES d loss / d raw-parameter into the
 * dense gradient tensors (caller zero-fills them), summing over duplicate ids in the batch.  Ids are int64
 * (id_bytes = 8, what the reference's LongTensors hold) or int32.  *d_bad is set to 1 if an id is out of
 * range (torch would raise IndexError).
 * pmf_adam_dense_step replaces torch.optim.Adam's update of one tensor (compare_models.py:288,312; defaults,
 * no weight decay / amsgrad); step_size = lr/(1-beta1^t), bias_correction2_sqrt = sqrt(1-beta2^t).
 * pmf_hpf_map_predict replaces forward/predict (hpf_pytorch.py:66-69, :186-195); float32 output. */
int pmf_hpf_map_loss_grad(const void* d_users, const void* d_items, int32_t id_bytes, const float* d_ratings,
                          int64_t B, const float* d_theta_raw, const float* d_beta_raw, const float* d_xi_raw,
                          const float* d_eta_raw, const float* d_user_scale, const float* d_item_scale, int32_t N,
                          int32_t M, int32_t K, float a, float a_prime, float b_prime, float c, float c_prime,
                          float d_prime, float* d_g_theta, float* d_g_beta, float* d_g_xi, float* d_g_eta,
                          double* d_loss, int32_t* d_bad, void* stream);
int pmf_adam_dense_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                        float beta1, float beta2, float eps, float step_size, float bias_correction2_sqrt,
                        void* stream);
/* Lazy ("touch-only") Adam, equivalent to the dense update: a row without gradient in a step is not touched; the
 * zero-gradient steps it skipped are applied (in closed form) when it is next referenced, together with its deferred
 * last real step.  All arrays are device memory owned by the caller: parameters, first/second
 * moments and gradient accumulators for theta (N,K), beta (M,K), xi (N), eta (M); last_* / claim_* int32 per row,
 * zero-initialised (last = s > 0: up to date with step s-1, the gradient of step s is accumulated but not applied;
 * last = -s <= 0: up to date with step s); step_size[s] = lr/(1-beta1^s) and bc2_sqrt[s] = sqrt(1-beta2^s) for every
 * step s (1-based). */
typedef struct pmf_lazy_adam {
    float *theta, *beta, *xi, *eta;
    float *m_theta, *m_beta, *m_xi, *m_eta;
    float *v_theta, *v_beta, *v_xi, *v_eta;
    float *g_theta, *g_beta, *g_xi, *g_eta;
    int32_t *last_user, *last_item, *claim_user, *claim_item;
    const float *step_size, *bc2_sqrt;
    float beta1, beta2, eps;
    /* Closed-form catch-up (both or neither; NULL = replay skipped steps one by one).  float64, indexed by step like
     * step_size, zero beyond the last step s_max of the call:
     *   tail1[s] = sum_{s < t <= s_max} step_size[t] bc2_sqrt[t]   (beta1/sqrt(beta2))^(t-s)
     *   tail2[s] = sum_{s < t <= s_max} step_size[t] bc2_sqrt[t]^2 (beta1/beta2)^(t-s)
     * A run of J zero-gradient steps after step s then moves p by (m/sqrt(v)) [W1 - eps/sqrt(v) W2], W = tail[s] -
     * ratio^J tail[s+J], and decays m, v by beta^J (first order in eps/sqrt(v); runs where that exceeds 1e-3 are replayed). */
    const double *tail1, *tail2;
    /* with tail1 / tail2: float64 [5][n_pow] powers for run lengths J = 0 .. n_pow-1 (longer runs are replayed):
     * beta1^J, beta2^J, (beta1/sqrt(beta2))^J, (beta1/beta2)^J, beta2^(-J/2) */
    const double* pow5;
    int32_t n_pow;
} pmf_lazy_adam;
/* One pass over n (already shuffled) ratings in mini-batches of `batch`, ONE kernel per step: the first lane group to
 * reference a row settles it (deferred Adam step + catch-up, gradient zeroed), the others wait for it, then every
 * group adds its element's loss and gradients.  step0 = steps taken so far; adds the losses to *d_loss. */
int pmf_hpf_map_lazy_epoch(const pmf_lazy_adam* st, const void* d_users, const void* d_items, int32_t id_bytes,
                           const float* d_ratings, int64_t n, int64_t batch, int64_t step0, const float* d_user_scale,
                           const float* d_item_scale, int32_t N, int32_t M, int32_t K, float a, float a_prime,
                           float b_prime, float c, float c_prime, float d_prime, double* d_loss, int32_t* d_bad,
                           void* stream);
/* Bring every row up to step_now (before parameters are read from outside the lazy loop). */
int pmf_hpf_map_lazy_flush(const pmf_lazy_adam* st, int32_t N, int32_t M, int32_t K, int64_t step_now, void* stream);
int pmf_hpf_map_predict(const void* d_users, const void* d_items, int32_t id_bytes, int64_t n,
                        const float* d_theta_raw, const float* d_beta_raw, int32_t N, int32_t M, int32_t K,
                        float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMF_B200_H */
