/* sibrar_b200 -- C ABI of the B200-native (sm_100a) SingleBranchNet hot path.
 *
 * The reference (Tigxy/SiBraR---Single-Branch-Recommender) is pure Python/PyTorch and has NO FFI for this path
 * (SURVEY.md section 8b); its boundary is the Python class API of ``SingleBranchNet``
 * (``algorithms/sgd_alg.py:2009-2144``, ``algorithms/base_classes.py:87-170``).  The entry points below are what a
 * ctypes/cffi binding of that class would call; each one names the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in ``_host``;
 *   - nothing here allocates, frees or synchronises; work is enqueued on ``stream`` (a ``cudaStream_t``);
 *   - return value 0 = ok, otherwise an SBR_ERR_* code; ``sbr_last_error()`` gives a thread-local message;
 *   - "bf16" pointers are ``void*`` (raw __nv_bfloat16); row pitches (``ld*``) are in ELEMENTS.
 */
#ifndef SIBRAR_B200_H
#define SIBRAR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBR_OK 0
#define SBR_ERR_ARG 1
#define SBR_ERR_CUDA 2

/* activation ids (reference modules/polylinear.py:5-10) */
#define SBR_ACT_NONE 0
#define SBR_ACT_RELU 1
#define SBR_ACT_TANH 2
#define SBR_ACT_SIGMOID 3
#define SBR_ACT_SELU 4

/* modality source kinds for the row gather (reference algorithms/sgd_alg.py:1327-1357, 1373-1389) */
#define SBR_SRC_TABLE 0       /* fp32 table [n_rows, C] of projected features (Linear(+hidden)+act output)     */
#define SBR_SRC_CATEGORICAL 1 /* nn.Embedding: weight[cat[row]]                                               */
#define SBR_SRC_TAG 2         /* nn.EmbeddingBag(mode='mean', padding_idx=-1): mean of weight[tags[row, :]]   */

/* rec losses (reference train/rec_losses.py:116-119) */
#define SBR_LOSS_BPR 0
#define SBR_LOSS_BCE 1
#define SBR_LOSS_SSM 2

const char* sbr_last_error(void);
int sbr_version(void);

/* ------------------------------------------------------------------------------------------------ GEMM (tcgen05)
 * D[M,N] = alpha * A * B^T with bf16 operands and fp32 TMEM accumulation, fused epilogue.
 *   A: K-major  -> memory [M, K] row-major, pitch lda;  MN-major -> memory [K, M] row-major, pitch lda.
 *   B: K-major  -> memory [N, K] row-major, pitch ldb;  MN-major -> memory [K, N] row-major, pitch ldb.
 * Replaces nn.Linear forward (modules/polylinear.py:50-76; algorithms/sgd_alg.py:1356,1380,1876) and the autograd
 * dgrad / wgrad GEMMs of ``total_loss.backward()`` (train/trainer.py:221). */
typedef struct {
  const float* bias;      /* [N] added to alpha*acc before stats/activation, or NULL                             */
  int act;                /* SBR_ACT_* applied after bias                                                         */
  void* out_bf16;         /* optional bf16 output [M, N]                                                          */
  int64_t ld_bf16;
  float* out_f32;         /* optional fp32 output [M, N] ([N, M] if transpose_out)                                */
  int64_t ld_f32;
  float* colstats;        /* optional [2*N]: += column sum / sum of squares of the FINAL epilogue value, valid rows */
  int colstats_sum_only;  /* 1: only the N column sums are accumulated (e.g. a bias gradient)                     */
  int colstats_rows;      /* > 0: deterministic mode -- colstats is [colstats_rows, 2*N]; every epilogue warp writes
                             its own row of partial sums (rows = sbr_gemm_colstats_rows(M, N)), summed in a fixed order
                             by sbr_bn_finalize; 0: fp32 atomics into colstats[2*N]                                */
  const void* actgrad_y;  /* optional bf16 [M, N]: result *= act'(y) expressed through the saved output y         */
  int64_t ld_actgrad;
  int actgrad_act;
  int transpose_out;      /* fp32 output written transposed                                                       */
  int atomic_out;         /* fp32 output accumulated with atomicAdd (needed when split_k > 1)                     */
  int split_k;            /* number of partitions of the K loop (>= 1)                                            */
  int64_t split_stride;   /* > 0: partition z writes (no atomics) to out_f32 + z * split_stride; sbr_splitk_reduce  */
  float alpha;
} sbr_gemm_epilogue_t;

int sbr_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, int64_t M,
                  int64_t N, int64_t K, const sbr_gemm_epilogue_t* ep, void* stream);
int sbr_gemm_colstats_rows(int64_t M, int64_t N); /* partial-statistics rows a [M, N] GEMM writes (no split-K) */

/* out[r, c] (+)= act(sum_s partials[s * split_stride + r * ld_part + c] + bias[c]) -- the deterministic second half of
 * a split-K GEMM (split_stride > 0 above).  accumulate = 1 adds into out_f32 (gradient buffers). */
int sbr_splitk_reduce(const float* partials, int n_splits, int64_t split_stride, int64_t ld_part, int64_t rows,
                      int64_t cols, const float* bias, int act, float* out_f32, int64_t ld_f32, int accumulate,
                      void* out_bf16, int64_t ld_bf16, void* stream);

/* Same GEMM with a BIT-PACKED 0/1 A operand (K-major): bit k of row m = A_bits[m * ld_words + k / 32] >> (k % 32).
 * The multi-hot 'interactions' rows (data/Feature.py:147-150: csr.toarray() per batch in the reference) stay packed in
 * HBM (1 bit per element) and are expanded to bf16 in shared memory inside the kernel (the bit words arrive there by TMA).  ld_words is a multiple
 * of 4 (16-byte rows, A_bits 16-byte aligned) and covers whole 64-bit K blocks; bits at k >= K are zero.  Used for the forward projection (A = the entity's interaction
 * matrix) and for its wgrad (A = the transposed matrix, output written transposed). */
int sbr_gemm_bits_bf16(const uint32_t* A_bits, int64_t ld_words, const void* B, int64_t ldb, int b_mn_major, int64_t M,
                       int64_t N, int64_t K, const sbr_gemm_epilogue_t* ep, void* stream);

/* ------------------------------------------------------------------------------------------------ casts / fills */
int sbr_cast_f32_to_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                         void* stream); /* dst columns [cols, ld_dst) are zero-filled */
int sbr_transpose_f32_to_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                              void* stream); /* dst[c, r] = src[r, c] */
int sbr_transpose_f32(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                      void* stream);
int sbr_csr_to_dense_bf16(const int64_t* indptr, const int32_t* indices, const float* vals, int64_t rows, int64_t cols,
                          void* dst, int64_t ld_dst, void* stream); /* csr rows -> dense (data/Feature.py:147-150);
                                                                       vals == NULL: all stored entries are 1 */

/* ------------------------------------------------------------------------------------------------ sparse 'interactions'
 * out[r, :] = act(sum_{p in csr[r]} vals[p] * Wt[indices[p], :] + bias)  -- Linear over a sparse row without densifying
 * it (replaces data/Feature.py:147-150 + algorithms/sgd_alg.py:1380).  Wt is the TRANSPOSED weight [d, C] fp32;
 * vals == NULL means all stored entries are 1 (duplicate history rows give counts > 1 in the reference's matrices).
 * With bias == NULL and act == NONE the same kernel is the wgrad of that Linear through the transposed CSR:
 * dW[:, j] (+)= sum_{p in csrT[j]} valsT[p] * dPre[indicesT[p], :]  (transpose_out = 1: out is [C, rows];
 * accumulate = 1: out += result, like every other wgrad of the path).  out_bf16: optional bf16 copy of the result. */
int sbr_spmm_csr(const int64_t* indptr, const int32_t* indices, const float* vals, int64_t rows, const float* dense,
                 int64_t ld_dense, int64_t C, const float* bias, int act, float* out, int64_t ld_out, int transpose_out,
                 int accumulate, void* out_bf16, int64_t ld_bf16, const int32_t* row_map, int atomic, void* stream);
/* row_map / atomic (row-major fp32 output only): CSR row s adds its result into out[row_map[s]] with vector reductions
 * -- the EmbeddingBag(mean) backward of a large tag vocabulary (algorithms/sgd_alg.py:1331-1340): CSR rows are
 * <= 128-entry segments of "feature rows carrying tag t", vals = 1 / #tags of the row, dense = the bag-table gradient.
 *
 * The same with a bf16 dense operand (fp32 accumulation): W^T as the transposed bf16 shadow [d, pad8(C)] in the forward,
 * the bf16 dz table in the wgrad -- the rounding points of the tensor-core routes at half the bytes.  row_list
 * (optional): the CSR rows to compute, n_rows_dev (optional): device-side number of valid entries of row_list (the
 * "referenced rows" of a step); outputs go to the listed rows. */
int sbr_spmm_csr_bf16(const int64_t* indptr, const int32_t* indices, const float* vals, int64_t rows,
                      const void* dense_bf16, int64_t ld_dense, int64_t C, const float* bias, int act, float* out,
                      int64_t ld_out, int transpose_out, int accumulate, void* out_bf16, int64_t ld_bf16,
                      const int32_t* row_list, const int32_t* n_rows_dev, const int32_t* seg_row,
                      const int32_t* long_rows, int64_t n_long, const int32_t* out_pos, void* stream);
/* seg_row (optional): `indptr` holds SEGMENTS of at most 256 entries of the matrix rows (`rows` = number of segments), so
 * that a popular item / a heavy user is not one warp's serial loop; seg_row[s] = output row | (row has several segments
 * << 31).  Segments of such long rows are combined with atomics; for the row-major output the library clears the
 * `n_long` rows listed in long_rows first and applies bias / activation (+ the bf16 copy) to them afterwards.
 * out_pos (optional, row-major fp32 output only): matrix row -> output row (the compact positions of
 * sbr_mark_referenced).  With it every unit ADDS its raw partial sum into the (caller-cleared) output row -- no bias,
 * activation or bf16 copy; the caller finishes the rows (sbr_splitk_reduce with one slice). */

/* ------------------------------------------------------------------------------------------------ modality sampling
 * Per (row, slot) choose k distinct modalities out of n_mods (optionally slot 0 fixed to `central`), Philox keyed
 * by (seed, step).  Replaces utilities/utils.py:60-90 + algorithms/sgd_alg.py:1904-1927 (distributional parity:
 * the reference's own stream depends on PYTHONHASHSEED). */
int sbr_sample_modalities(uint8_t* mods, int64_t n_rows, int k, int n_mods, int central, uint64_t seed,
                          const int64_t* step_dev, void* stream);

/* step counter kept on the DEVICE (so that CUDA-graph replays of a step see a fresh value): *counter += 1 */
int sbr_tick(int64_t* counter_dev, void* stream);
/* start of a fused train step in ONE launch: counter0 / counter1 (nullable) += 1 and `zero_bytes` (multiple of 16) of the
 * step's accumulator arena cleared (replaces two sbr_tick launches and a memset on the step's critical path) */
int sbr_step_begin(int64_t* counter0, int64_t* counter1, void* zero, int64_t zero_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ row gather
 * One descriptor per modality of an entity. */
typedef struct {
  int kind;               /* SBR_SRC_*                                                                            */
  const int32_t* remap;   /* [n_entities] entity index -> feature row (data/Feature.py:146), or NULL = identity  */
  const float* table;     /* TABLE: [n_rows, C] fp32; CATEGORICAL/TAG: embedding weight [n_cat(+1), C] fp32       */
  float* grad;            /* same shape as `table`: gradient accumulator (atomicAdd), or NULL in forward          */
  const int32_t* codes;   /* CATEGORICAL: [n_rows] category id; TAG: [n_rows, max_tags] tag ids (pad = n_tags)    */
  int32_t max_tags;       /* TAG only                                                                             */
  int32_t pad_id;         /* TAG only                                                                             */
  int64_t n_table_rows;   /* rows of `table` / `grad` (small tables are accumulated in shared memory by the backward)  */
  int64_t key_base;       /* first segment key of this modality (sbr_gather_plan); keys = key_base + table row |
                             category | entity row (TAG)                                                          */
} sbr_modality_src_t;

/* X[r, :] = dropout(normalize(src_{mods[r]}(idx[r / k])))  written as bf16 (algorithms/sgd_alg.py:1934-1978,
 * 1865-1876) and/or fp32.  srcs_dev: device array [n_mods].  keep_mask (optional, uint8 [N, C]) overrides the
 * counter-based hash mask keyed by (row, column group, seed, step).  err_flag is set to 1 when an entity index has no feature row (KeyError at data/Feature.py:146). */
int sbr_row_gather_fwd(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx, const uint8_t* mods,
                       int64_t n_idx, int k, int C, int normalize, float p_drop, uint64_t seed,
                       const int64_t* step_dev, const uint8_t* keep_mask, void* out_bf16, int64_t ld_out,
                       float* out_f32, int64_t ld_f32, int32_t* err_flag, uint8_t* keep_bits_out, void* stream);
/* keep_bits_out (optional, uint8 [N, ceil(C / 8)]): the dropout keep decisions of this call, bit j of byte c / 8 =
 * column c -- handed to sbr_row_gather_bwd_segmented so that the backward does not regenerate the mask. */
/* backward of the above: dX (fp32 [N, C], pitch ld_dx) -> atomicAdd into the sources' grad buffers */
int sbr_row_gather_bwd(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx, const uint8_t* mods,
                       int64_t n_idx, int k, int C, int normalize, float p_drop, uint64_t seed,
                       const int64_t* step_dev, const uint8_t* keep_mask, const float* dx, int64_t ld_dx,
                       void* stream);

/* Low-contention backward of the row gather.  plan: counting sort of the N flat rows by segment key
 * (counts/cursor int32 [n_keys], offsets int32 [n_keys + 1 + ceil(n_keys / 4096)] (the tail is the scan's per-chunk
 * state), row_keys/perm/sorted_keys int32 [N]; all caller-owned scratch; `counts` must be ZERO on entry and is zero
 * again on return -- allocate it cleared once, no memset per step).  reduce: persistent blocks; one lane group (row width / 8 lanes) per chunk of consecutive SORTED rows (the
 * library picks the chunk length: one equal-length chunk per resident group, at least 8 rows; `rows_per_warp` >= 1 is
 * accepted for ABI stability and otherwise ignored) sums runs of equal keys in registers
 * (dropout-scaled dx rows), applies the L2-normalise backward once per run and atomically adds the run into the
 * source's grad buffer: a (modality, source row) that occurs n times costs ~n / run-length atomics instead of n. */
int sbr_gather_plan(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx, const uint8_t* mods,
                    int64_t n_idx, int k, int64_t n_keys, int32_t* counts, int32_t* offsets, int32_t* cursor,
                    int32_t* row_keys, int32_t* perm, int32_t* sorted_keys, void* stream);
int sbr_row_gather_bwd_segmented(const sbr_modality_src_t* srcs_dev, int n_mods, int64_t n_keys,
                                 const int32_t* offsets, const int32_t* perm, const int32_t* sorted_keys,
                                 int64_t n_rows, int C, int normalize, float p_drop, uint64_t seed,
                                 const int64_t* step_dev, const uint8_t* keep_mask, const float* dx, int64_t ld_dx,
                                 int rows_per_warp, const uint8_t* keep_bits, void* stream);

/* ------------------------------------------------------------------------------------------------ referenced rows
 * The per-step subset of feature rows a modality projection has to compute (the reference projects exactly the rows of the
 * batch: algorithms/sgd_alg.py:1949-1974 + data/Feature.py:140-162).  One sbr_ref_table_t per modality of an entity
 * (device array [n_mods]); stamp == NULL: the modality keeps the whole-table route.
 *   stamp [n_rows] int32 (0-initialised once), pos [n_rows] int32: row -> compact position of this step,
 *   list [capacity] int32: compact position -> row, count: device int32 cleared by the caller before the call;
 *   seg_first (optional, int64 [n_rows + 1]): first segment of every row of a segmented CSR matrix -- then the row's
 *   segments are appended to seg_list / seg_count as the units of work of the sparse kernels. */
typedef struct {
  int32_t* stamp;
  int32_t* pos;
  int32_t* list;
  int32_t* count;
  const int64_t* seg_first;
  int32_t* seg_list;
  int32_t* seg_count;
} sbr_ref_table_t;

/* every (entity, modality) slot of the step marks its feature row.  epoch_dev: int64 [2] owned by the caller, zero
 * before the first call (the kernel advances the epoch itself: [0] = epoch, [1] = finished blocks) */
int sbr_mark_referenced(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx, const uint8_t* mods,
                        int64_t n_idx, int k, int64_t* epoch_dev, const sbr_ref_table_t* tabs_dev, void* stream);
/* dst[slot, :] = src[list[slot], :] for slot < *count_dev, zero rows up to `capacity` (bf16 rows, 16-byte aligned) */
int sbr_gather_rows_bf16(const void* src, int64_t ld_src, const int32_t* list, const int32_t* count_dev,
                         int64_t capacity, int64_t cols, void* dst, int64_t ld_dst, void* stream);
/* wgrad of a sparse-input Linear over the listed units (rows, or segments with seg_row as in sbr_spmm_csr_bf16):
 * gT[j, :] += vals[p] * dz[pos[row], :] for every stored entry (p, j) of the unit; gT is the TRANSPOSED gradient [d, C] */
int sbr_spmm_scatter_wgrad(const int64_t* indptr, const int32_t* indices, const float* vals, const int32_t* unit_list,
                           const int32_t* n_units_dev, int64_t max_units, const int32_t* seg_row, const int32_t* pos,
                           const void* dz_bf16, int64_t ld_dz, int64_t C, float* gT, int64_t ld_gT, void* stream);
/* dst[c, r] += src[r, c], src cleared */
int sbr_transpose_add_f32(float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                          void* stream);

/* ------------------------------------------------------------------------------------------------ fused single-branch MLP
 * One persistent kernel per direction for an entity whose single-branch network is 1 or 2 Linear layers of width <= 64
 * (conf/single/algorithms/sbnet_ml1m_conf.yml, sbnet_onion18_conf.yml, ...):
 *   fwd: z[r, :] = Linear_last(act(Linear_0(dropout(normalize(src_{mods[r]}(idx[r / k]))))))  (+ act of the last layer)
 *        = _get_modality_embeddings + _embed (algorithms/sgd_alg.py:1934-1978, 1865-1877) + PolyLinear
 *        (modules/polylinear.py:50-76) without X0 / hidden activations ever reaching HBM; colstats (optional): rows of
 *        partial column sums / sums of squares of z for the BatchNorm behind the chain (sbr_bn_finalize adds them in a
 *        fixed order; sbr_mlp2_colstats_rows(n_rows) rows of 2 D floats).
 *   bwd: from dy = dL/d(output behind the optional BatchNorm) and the saved z: dz (BatchNorm backward from the column sums
 *        `sums` = [n_replicas, 2 D] of (dy, dy * xhat), then the last activation's derivative), the gradients of both
 *        weights and biases (accumulated into grad_w / grad_b, fp32 [out, in] / [out]), d gamma / d beta, and
 *        dx[r, :] = dL/dX0 (fp32, consumed by sbr_row_gather_bwd_segmented).  X0 and the hidden activation are recomputed.
 * Weights: the bf16 shadows [out, ldw] (ldw % 8 == 0). */
typedef struct {
  const void* w_bf16;
  int64_t ldw;
  const float* bias; /* may be NULL */
  int in_f, out_f;
  int act; /* SBR_ACT_* applied to this layer's output */
} sbr_mlp2_layer_t;

typedef struct {
  const sbr_modality_src_t* srcs; /* device array [n_mods] */
  int n_mods;
  const int64_t* idx;  /* [n_idx] entity indices */
  const uint8_t* mods; /* [n_idx * k] modality of every row (NULL: modality 0) */
  int64_t n_idx;
  int k, C, normalize;
  float p_drop;
  uint64_t seed;
  const int64_t* step_dev;
  const uint8_t* keep_mask; /* optional explicit dropout mask uint8 [n_idx * k, C] */
  int32_t* err_flag;
  int n_layers; /* 1 or 2 */
  sbr_mlp2_layer_t layers[2];
} sbr_mlp2_desc_t;

typedef struct {
  const float* mean_invstd; /* [2 D] of the forward */
  const float* gamma;
  const float* sums; /* [n_replicas, 2 D]: column sums of dy and of dy * xhat */
  int n_replicas;
  float* dgamma; /* += */
  float* dbeta;  /* += */
} sbr_mlp2_bn_t;

int sbr_mlp2_colstats_rows(int64_t n_rows);
int sbr_mlp2_fwd(const sbr_mlp2_desc_t* desc, int64_t n_rows, int C, float* z, int64_t ldz, float* colstats,
                 int colstats_rows, void* stream);
/* Same forward with the BatchNorm statistics finalised inside the kernel (nn.BatchNorm1d training forward,
 * modules/polylinear.py:58-62): the CTA that finishes last adds the rows of `colstats` in a fixed order and writes
 * mean_invstd [2 D] (mean | 1 / sqrt(biased var + eps)), updates the running statistics (unbiased variance) and
 * num_batches_tracked -- no separate sbr_bn_finalize launch between the chain and the scoring kernel.
 * counter: one uint32 in device memory, zero on entry, zero again on return. */
typedef struct {
  unsigned int* counter;
  float eps, momentum;
  float* mean_invstd;
  float* running_mean;          /* may be NULL */
  float* running_var;           /* may be NULL */
  int64_t* num_batches_tracked; /* may be NULL */
} sbr_mlp2_bn_tail_t;
int sbr_mlp2_fwd_bn(const sbr_mlp2_desc_t* desc, int64_t n_rows, int C, float* z, int64_t ldz, float* colstats,
                    int colstats_rows, const sbr_mlp2_bn_tail_t* tail, void* stream);
int sbr_mlp2_bwd(const sbr_mlp2_desc_t* desc, int64_t n_rows, int C, const float* dy, int64_t lddy, const float* z,
                 int64_t ldz, const sbr_mlp2_bn_t* bn, float* const* grad_w, float* const* grad_b, float* dx,
                 int64_t lddx, void* stream);
/* profiling only: with SBR_MLP2_DEBUG bit 16 the roles of block 0 of sbr_mlp2_fwd / sbr_mlp2_bwd accumulate the SM
 * cycles of every barrier wait and of the sections of their own work (16 words per role: gather, MMA, epilogue, dz);
 * returns the NUMBER of words copied to host_out (scripts/trace_mlp2.py prints them) */
int sbr_mlp2_trace_read(unsigned long long* host_out, int max_events);
/* profiling only: with SBR_MLP2_DEBUG bit 64 every CTA of an sbr_mlp2_bwd launch over more than 100 000 rows stamps
 * %globaltimer at its start and end; copies (start, end) ns pairs of the first n_ctas CTAs (returns their number) */
int sbr_mlp2_cta_times_read(unsigned long long* host_out, int n_ctas);

/* nn.EmbeddingBag(mode="mean", padding_idx=pad_id) of EVERY feature row as a dense fp32 table [n_rows, C]
 * (reference FeatureEmbedding for tag features, algorithms/sgd_alg.py:1279-1396; codes int32 [n_rows, max_tags]):
 * the model gathers from it as a SBR_SRC_TABLE source (one row per lookup, no tag loop inside the batch-sized gathers).
 * bwd: grad_weight[tag] += bag_grad[r] / #tags(r) for every tag of every row; bag_grad is cleared as it is read
 * (it is an atomic accumulator reused every step). */
int sbr_tag_bag_fwd(const int32_t* codes, int max_tags, int32_t pad_id, const float* weight, int64_t n_rows, int C,
                    float* out, void* stream);
int sbr_tag_bag_bwd(const int32_t* codes, int max_tags, int32_t pad_id, float* bag_grad, int64_t n_rows, int C,
                    float* grad_weight, int64_t n_weight_rows, void* stream);

/* table-level backward of the projection output activation: dpre = dT * act'(T) -> bf16 (+ column sums = dbias).
 * zero_dy = 1 clears dy after reading it (the gradient table is an atomicAdd accumulator reused every step). */
int sbr_actgrad_colsum(float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y, int act,
                       int64_t rows, int64_t cols, void* out_bf16, int64_t ld_out, float* out_f32, int64_t ld_out_f32,
                       float* colsum, int zero_dy, void* stream);

/* stats: [n_partials, 2*C] column sums / sums of squares written by the GEMM epilogue (n_partials = 1: accumulated
 * with atomics).  finalize adds the partial rows in a fixed order -> mean_invstd [2*C], updates running stats
 * (momentum 0.1, unbiased var) and num_batches_tracked. */
int sbr_bn_finalize(const float* stats, int n_partials, int64_t n_rows, int C, float eps, float momentum,
                    float* mean_invstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    void* stream);
/* y = act(gamma * (z - mean) * invstd + beta); z fp32 [rows, C]; writes bf16 and/or fp32.
 * eval mode: pass running stats through sbr_bn_eval_coeffs first. */
int sbr_bn_apply(const float* z, int64_t ld_z, const float* mean_invstd, const float* gamma, const float* beta,
                 int act, int64_t rows, int C, void* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32,
                 void* stream);
int sbr_bn_eval_coeffs(const float* running_mean, const float* running_var, int C, float eps, float* mean_invstd,
                       void* stream);
/* backward pass 1: dzb = dy * act'(y); sums[0:C] += dzb, sums[C:2C] += dzb * xhat (xhat from z, mean, invstd) */
int sbr_bn_bwd_reduce(const float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y, int act,
                      const float* z, int64_t ld_z, const float* mean_invstd, int64_t rows, int C, float* sums,
                      void* stream);
/* backward pass 2: dz = gamma*invstd*(dzb - mean(dzb) - xhat*mean(dzb*xhat)) -> bf16 (+fp32); dgamma, dbeta from sums.
 * sums: [n_replicas, 2*C] partial sums that are added up first (1 replica from sbr_bn_bwd_reduce). */
int sbr_bn_bwd_apply(const float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y, int act,
                     const float* z, int64_t ld_z, const float* mean_invstd, const float* gamma, const float* sums,
                     int n_replicas, int64_t rows, int C, void* dz_bf16, int64_t ld_dz, float* dz_f32,
                     int64_t ld_dz_f32, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------ losses
 * Fused modality aggregation (mean | max over k), user-item scoring, rec loss and their gradients
 * (algorithms/sgd_alg.py:1861,2114 + train/rec_losses.py:40-113).
 *   eu: fp32 [B, ku, D], ei: fp32 [B, n, ki, D]; logits out [B, n]; loss_acc[0] += rec loss (double);
 *   deu / dei written (=) with the gradient of the rec loss (NULL = forward only).
 *   ssm_shift = ln(n_items / n_neg) when the sampler is 'uniform' (rec_losses.py:104-105), else 0. */
int sbr_score_loss(const float* eu, const float* ei, int64_t B, int n, int ku, int ki, int D, int agg_max_user,
                   int agg_max_item, int loss_kind, int aggregator_sum, float ssm_shift, float* logits,
                   double* loss_acc, float* deu, float* dei, float* u_agg, float* i_agg, void* stream);
/* Backward of the scoring alone for a loss computed OUTSIDE (the reference's own loop: logits -> rec loss ->
 * loss.backward(), train/trainer.py:209-221): given d loss / d logits [B, n], writes deu / dei like sbr_score_loss. */
int sbr_score_bwd(const float* eu, const float* ei, int64_t B, int n, int ku, int ki, int D, int agg_max_user,
                  int agg_max_item, const float* dlogits, float* deu, float* dei, void* stream);

/* The same for entities whose single-branch net ends in a BatchNorm1d (algorithms/sgd_alg.py:1834-1837): the
 * kernel reads the PRE-BatchNorm values z and applies e = gamma * (z - mean) * invstd + beta on the fly, and it
 * accumulates the BatchNorm-backward column sums of the gradients it produces (sums[r][0:D] += de,
 * sums[r][D:2D] += de * xhat, r = one of n_replicas copies) -- replaces sbr_bn_apply + sbr_bn_bwd_reduce at the
 * end of the chain.  One modality slot per entity (no regularisation), D in {16, 32, 64, 128}.  bn_u / bn_i may be
 * NULL (or have z == NULL): then eu / ei hold the embeddings themselves. */
typedef struct {
  const float* z;            /* [rows, D] pre-BatchNorm values                  */
  const float* mean_invstd;  /* [2 * D] from sbr_bn_finalize                    */
  const float* gamma;        /* [D]                                             */
  const float* beta;         /* [D]                                             */
  float* sums;               /* [n_replicas, 2 * D], zeroed by the caller       */
} sbr_bn_inline_t;
int sbr_score_loss_bn(const float* eu, const sbr_bn_inline_t* bn_u, const float* ei, const sbr_bn_inline_t* bn_i,
                      int64_t B, int n, int D, int loss_kind, int aggregator_sum, float ssm_shift, float* logits,
                      double* loss_acc, float* deu, float* dei, int n_replicas, void* stream);
/* Symmetric InfoNCE (train/regularization_losses.py:8-43) between slot 0 and slot 1 of e [G, n, 2, D]:
 * contrast along n inside each of the G groups (item side: G = B, n = 1 + n_neg; user side: G = 1, n = B).
 * loss_acc[0] += weight * loss; de (+)= weight * dloss/de (accumulate=1 adds to existing gradients). */
int sbr_infonce(const float* e, int64_t G, int64_t n, int D, float temperature, float weight, double* loss_acc,
                float* de, int accumulate, float* lse_ws, void* stream);
/* Large groups (the user side: ONE group of n = B in-batch rows): the n x n logits and both gradient products are
 * tcgen05 GEMMs (sbr_gemm_bf16) and these are the memory-bound passes between them (train/regularization_losses.py:28-43
 * evaluated as  L = E0 E1^T / T,  CE over rows + CE over columns):
 *   split:   e fp32 [n, 2, D] -> bf16 [n, 3 pad8(D)] triples  A3 = [hi0 | lo0 | hi0], B3 = [hi1 | hi1 | lo1]; one GEMM over
 *            K = 3 pad8(D) then gives the logits to ~2^-16 relative
 *   lse:     lse_r[i], lse_c[j] of L fp32 [n, n] (n % 4 == 0; partial_ws: 2 * n_chunks * n floats) and
 *            loss_acc += scale * (sum_i (lse_r[i] - L_ii) + sum_j (lse_c[j] - L_jj))
 *   weights: W_ij = exp(L_ij - lse_r[i]) + exp(L_ij - lse_c[j]) - 2 delta_ij as bf16 [n, n]:
 *            dE0 = alpha W E1,  dE1 = alpha W^T E0  with alpha = weight / (n T)  (ops.infonce_gemm) */
int sbr_infonce_split(const float* e, int64_t n, int D, void* a3, void* b3, void* stream);
int sbr_infonce_lse(const float* L, int64_t n, float scale, float* lse_r, float* lse_c, float* partial_ws, int n_chunks,
                    double* loss_acc, void* stream);
int sbr_infonce_weights(const float* L, int64_t n, const float* lse_r, const float* lse_c, void* W, void* stream);

/* Bias terms of the matrix-factorisation siblings (SGDMatrixFactorization.combine_user_item_representations,
 * algorithms/sgd_alg.py:175-195): logits[b, j] += user_bias[u[b]] + item_bias[i[b, j]] + global_bias[0] (each optional)
 * and the matching gradient accumulation from d loss / d logits. */
int sbr_logit_bias_fwd(float* logits, int64_t B, int n, const int64_t* u_idx, const int64_t* i_idx,
                       const float* user_bias, const float* item_bias, const float* global_bias, void* stream);
int sbr_logit_bias_bwd(const float* dlogits, int64_t B, int n, const int64_t* u_idx, const int64_t* i_idx,
                       float* d_user_bias, float* d_item_bias, float* d_global_bias, void* stream);

/* Lower clamp of the scores of DeepMatrixFactorization (`sim[sim < mu] = mu`, algorithms/sgd_alg.py:1238-1242), in
 * place; `clamped` (optional, uint8 [n]) records where, for the backward (no gradient through a clamped score). */
int sbr_clamp_min_fwd(float* x, int64_t n, float lo, uint8_t* clamped, void* stream);
int sbr_clamp_min_bwd(float* dx, int64_t n, const uint8_t* clamped, void* stream);

/* aggregation only (eval path): out[r, :] = mean | max over k of e[r, k, :] */
int sbr_aggregate(const float* e, int64_t rows, int k, int D, int agg_max, float* out_f32, void* out_bf16,
                  int64_t ld_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------ optimizer
 * Multi-tensor Adam / AdamW (torch.optim at train/trainer.py:62-68,222-223), one launch for all parameters;
 * grads are zeroed in the same pass (zero_grad) and an optional bf16 shadow of the weight is refreshed. */
typedef struct {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* shadow_bf16;   /* optional: bf16 copy with row pitch shadow_ld (rows x cols view of the parameter) */
  int64_t numel;
  int64_t cols;        /* innermost extent (for the pitched shadow); numel % cols == 0 */
  int64_t shadow_ld;
} sbr_adam_tensor_t;
int sbr_adam_step(const sbr_adam_tensor_t* tensors_dev, int n_tensors, int64_t total_chunks,
                  const int32_t* chunk_to_tensor_dev, const int64_t* chunk_offset_dev, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int decoupled, const int64_t* step_dev,
                  float grad_scale, void* stream); /* *step_dev = 1-based index of the step being applied;
                                                      decoupled: 0 = Adam (L2 folded into the gradient), 1 = AdamW,
                                                      2 = Adagrad (torch.optim.Adagrad defaults; exp_avg_sq holds the
                                                      running sum of squares, exp_avg is left untouched) */

/* Data-parallel optimizer step (the gradient all-reduce of the data-parallel train step; the
 * reference is single-process): the SUM of the flat gradient buffers of all ranks is formed inside the NVSwitch
 * (multimem.ld_reduce on the multicast mapping of the buffer), broadcast (multimem.st) into every rank's `sum` buffer
 * and consumed by the same Adam / AdamW / Adagrad update as sbr_adam_step -- ONE kernel per rank and step, no NCCL
 * call.  All buffers are symmetric allocations (same size on every rank, mapped by every rank, multicast-bound). */
typedef struct {
  const float* flat_grads;        /* local address of this rank's flat gradient buffer [total]                       */
  const float* mc_grads;          /* multicast address of the same buffer (reads reduce over the ranks)              */
  float* sum_local;               /* local address of the summed-gradient buffer [total]                             */
  float* sum_mc;                  /* multicast address of it (stores reach every rank)                               */
  int64_t total;                  /* floats; a multiple of 4 * world                                                 */
  int32_t* const* peer_flags_dev; /* device array [world]: address of every rank's flag words int32 [2 * world]      */
  int32_t* flags_local;           /* this rank's flag words (zero at allocation, only ever grow)                     */
  int64_t* state;                 /* local int64 [4], zero at allocation: epoch, 2 release words, grid arrivals      */
  int world, rank;
} sbr_mc_comm_t;
int sbr_adam_step_mc(const sbr_adam_tensor_t* tensors_dev, int n_tensors, int64_t total_chunks,
                     const int32_t* chunk_to_tensor_dev, const int64_t* chunk_offset_dev, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int decoupled, const int64_t* step_dev,
                     float grad_scale, int apply_adam, const sbr_mc_comm_t* comm, int grid_blocks, void* stream);
/* every adam tensor's `grad` must point into flat_grads; grid_blocks: the same value on every call that shares
 * `state` (<= 2 x SM count: the blocks of the grid wait for each other); apply_adam = 0: all-reduce only (the sums
 * are left in sum_local). */


/* ------------------------------------------------------------------------------------------------ evaluation
 * Fused  scores = U * I^T  (bf16 tcgen05 GEMM, fp32 accumulate)  ->  seen-item mask (-inf)  ->  per-user top-k.
 * The [U, I] score matrix never reaches HBM.  Replaces eval/eval.py:216-220 + the topk inside rmet.calculate
 * (eval/eval.py:99-102).  Items are split into `n_splits` contiguous ranges for SM fill; the per-split survivor
 * lists are merged by sbr_topk_merge.  Ties are broken by the LOWEST item position.  Positions are
 * indices into items_in_split (+ item_offset for item-sharded multi-GPU evaluation).
 *   users bf16 [U, ldu], items bf16 [I, ldi]; seen CSR over item positions (sorted per user), may be NULL. */
int sbr_topk_workspace_bytes(int64_t U, int64_t I, int D, int k, int n_splits, int64_t* bytes_out);
/* part_keys: uint64 [n_splits, U, k], UNSORTED survivors per split as packed keys
 * (order-preserving score bits << 32 | (0xFFFFFFFF - position)), 0 = empty slot. */
int sbr_topk_scores_masked(const void* users, int64_t ldu, const void* items, int64_t ldi, int64_t U, int64_t I, int D,
                           const int64_t* seen_indptr, const int32_t* seen_indices, int k, int n_splits,
                           int32_t item_offset, uint64_t* part_keys, void* workspace, int64_t workspace_bytes,
                           void* stream);
/* exact top-k of L key lists per user: keys [L, U, k] -> sorted (descending, ties -> lowest position) scores
 * [U, k], positions [U, k] (-inf / -1 where fewer than k candidates exist) and/or packed keys [U, k].
 * Also the merge step after the NVLink all-gather of item-sharded evaluation (L = number of ranks). */
int sbr_topk_merge(const uint64_t* keys, int L, int64_t U, int k, float* out_vals, int32_t* out_idx,
                   uint64_t* out_keys, void* stream);
/* per-user metrics from ranked positions and a target CSR (sorted): ndcg, precision, recall, f_score, hitrate, ap, rr
 * for each k in ks (rmet.calculate at eval/eval.py:99-102; definitions eval/metrics.py:4-105).
 * out: fp32 [7, n_ks, U] in that metric order; item_hits (optional int32 [n_ks, I]) marks recommended items for
 * coverage. */
int sbr_metrics_at_k(const int32_t* topk_idx, int64_t U, int k, const int64_t* tgt_indptr, const int32_t* tgt_indices,
                     const int32_t* ks_dev, int n_ks, float* out, int32_t* item_hits, int64_t n_items, void* stream);

/* ------------------------------------------------------------------------------------------------ negative sampling
 * 'uniform_recbole' (data/dataloader.py:154-198): uniform with replacement over items_in_split, re-drawn while the
 * draw is a train positive of the user (binary search in the sorted train CSR).  Also draws the positives:
 * a uniformly random train interaction per slot.  out_u [B], out_i [B, 1+n_neg]. */
int sbr_sample_batch(const int32_t* coo_user, const int32_t* coo_item, int64_t nnz, const int64_t* train_indptr,
                     const int32_t* train_indices, const int32_t* items_in_split, int64_t n_items_in_split, int64_t B,
                     int n_neg, uint64_t seed, const int64_t* step_dev, int64_t* out_u, int64_t* out_i, void* stream);

/* the same for ONE batch of a shuffled epoch (DataLoader(shuffle=True) over TrainRecDataset, data/dataset.py:380-396,
 * train/trainer.py:204): slot b is interaction order[offset + b] (order: a permutation of [0, nnz) on the device), its
 * n_neg negatives are drawn as above. */
int sbr_sample_epoch_batch(const int32_t* coo_user, const int32_t* coo_item, int64_t nnz, const int64_t* order,
                           int64_t offset, const int64_t* train_indptr, const int32_t* train_indices,
                           const int32_t* items_in_split, int64_t n_items_in_split, int64_t B, int n_neg, uint64_t seed,
                           const int64_t* step_dev, int64_t* out_u, int64_t* out_i, void* stream);

/* both strategies of the reference behind one entry (data/dataset.py:361-375, data/sampling.py): `order` NULL = a random
 * train interaction per slot (sbr_sample_batch), else slot b = interaction order[offset + b].
 *   SBR_NEG_UNIFORM_RECBOLE  with replacement over items_in_split, re-drawn while a train positive (data/sampling.py:35-67)
 *   SBR_NEG_UNIFORM          n_neg DISTINCT non-positive items of the split (negative_sample_uniform +
 *                            neg_samp_vectorized_bsearch, data/sampling.py:7-32); item_pos int32 [n_items] = position of
 *                            an item inside the sorted items_in_split (NULL when the split holds every item: identity).
 * The caller checks n_choices - n_pos >= n_neg per user like the reference (ValueError). */
enum { SBR_NEG_UNIFORM_RECBOLE = 0, SBR_NEG_UNIFORM = 1 };
int sbr_sample_negatives(const int32_t* coo_user, const int32_t* coo_item, int64_t nnz, const int64_t* order,
                         int64_t offset, const int64_t* train_indptr, const int32_t* train_indices,
                         const int32_t* items_in_split, int64_t n_items_in_split, const int32_t* item_pos, int64_t B,
                         int n_neg, int strategy, uint64_t seed, const int64_t* step_dev, int64_t* out_u,
                         int64_t* out_i, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIBRAR_B200_H */
