// kernels.h -- launch wrappers for the sm_100a kernels of the batched query path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr int kScanThreads = 128;     // 4 warps per scan CTA
constexpr int kScanWarps = 4;
constexpr int kMaxGroupsPerTile = 32; // 32-slot groups per (leaf tile) work item
constexpr int kQueriesPerQuad = 8;    // queries packed into one 64-bit LUT entry of the main scan (an "oct")

// Device view of one searcher (all pointers are device pointers).
struct DevIndex {
  int distance;
  uint32_t n, d, L, B, W, dpb;
  int disjoint;
  // Squared-L2 tree-AH (TreeXHybridSMMD semantics): no bias, inv = 1.0f / mult, candidate keys are
  // (score, datapoint id) instead of (score, global slot).  See oracle/scann_oracle.c tie_index().
  int key_by_dp;
  const float* centers;       // [L][D]
  const float* centers_t;     // [D][L] (unused by the kernels, kept for debugging)
  const float* center_sqnorm; // [L] squared-L2 tokenization only
  const float* codebook;      // [B][16][dpb]
  const int32_t* block_dims;  // [B]
  const uint32_t* block_off;  // [B+1]
  const uint32_t* leaf_size;  // [L] real slots
  const uint32_t* leaf_goff;  // [L+1] offset in 32-slot groups
  const uint32_t* leaf_ntiles;// [L] work tiles per leaf
  const uint32_t* leaf_gpt;   // [L] groups per tile
  const uint32_t* codes;      // [groups][W*32] packed nibble words (see pack_codes in index.cu)
  const uint32_t* slot_dp;    // [groups*32] datapoint id per slot, 0xFFFFFFFF padding
  const uint32_t* slot_tie;   // [groups*32] slot of the datapoint in the unsharded index (NULL = identity)
  const float* dataset;       // [rows][D] f32 rows for exact reordering (NULL if none)
  const uint16_t* dataset_bf16;  // [rows][D] bf16 rows (bfloat16 reordering) when `dataset` is NULL
  const uint32_t* dp_row;     // [N] datapoint id -> row of `dataset` (NULL = identity)
  // int8 (fixed point) reordering when `dataset` and `dataset_bf16` are NULL
  const int8_t* dataset_i8;   // [rows][D]
  const float* i8_inv_mult;   // [D] 1.0f / multiplier_by_dimension
  const float* i8_dp_norm;    // [N] squared L2 norm of the original row, by datapoint id (squared L2 only)
  // int8 (fixed point) query tokenization (query_tokenization_type FIXED_POINT_INT8, prep.cu tokenize_i8_kernel):
  // non-NULL centers_i8 switches launch_tokenize / launch_tokenize_topp to it
  const int8_t* centers_i8;   // [L][D] ScalarQuantizeFloatDataset(centres, 1.0, NaN)
  const float* cen_qscale;    // [D] query scale: 1.0f / multiplier_by_dimension, times 2 for squared L2
  const float* cen_sqnorm;    // [L] float(SquaredL2Norm(float centre)) (squared L2 only)
  const float* cen_sqnorm2;   // [L] 2 * cen_sqnorm: the bias of the chunk statistic of the tokenization GEMM (prep.cu)
  float cen_i8_max_norm;      // >= max_l ||float(int8 centre l)|| (error term of the tensor-core pre-filter)
  float cen_sqnorm_max;       // >= max_l cen_sqnorm[l]
  // tensor-core tokenization (prep.cu): centres as the bf16 operand [L][tok_kp] = [hi | hi | lo | 0]
  const void* tok_b;
  uint32_t tok_kp;
  float center_max_norm;      // >= max_l ||c_l||
  // optional [nq] bytes: set to 1 for rows whose tensor-core pre-filter fell back to exact distances (the row of
  // the distance matrix then holds exact distances instead of dot products); only the index-build stage reads it
  uint8_t* tok_fallback_flag;
  // optional workspace [nq][ceil(L / 32)] floats: enables the chunk pre-selection of the tokenizer (prep.cu)
  float* tok_cmax_ws;
  // 1: callers read the [nq][L] matrix of the tokenization GEMM afterwards (the SOAR pruning of the index-build
  // stage); 0: the chunk pre-selection may skip storing it and evaluate the exact chain for whole candidate chunks
  int tok_need_rows;
};

struct ScanWork {
  // per-batch work description, all device pointers
  const int32_t* leaves;      // [nq][P] probed leaves sorted by (distance, leaf)
  const float* bias;          // [nq][P] distance to centre
  const uint8_t* lut;         // [nq][W*8*16]
  const float* mult;          // [nq]
  const float* inv_mult;      // [nq]
  int32_t* pilot_end;         // [nq] unused (the pilot publishes nothing; kept for layout stability)
  uint64_t* buf;              // [nq][cap] candidate keys
  uint32_t* cnt;              // [nq]
  uint64_t* tau;              // [nq] push only keys < tau
  uint32_t* ovf;              // [nq] overflow flag
  uint32_t* leaf_cnt;         // [L+1]
  uint32_t* leaf_eoff;        // [L+1]
  uint32_t* leaf_cur;         // [L]
  uint32_t* item_off;         // [L+1]
  uint32_t* pair_pos;         // [nq*P] position of the (query, rank) pair inside its leaf's entry list (set when counted)
  uint32_t* entry_q;          // [nq*P]
  float* entry_bias;          // [nq*P]
  uint32_t* counters;         // [8]: 0 item counter, 1 n_items, 2 n_ovf, 3 n_entries
  unsigned long long* stats;  // [4]: 0 bytes_alg, 1 pairs, 2 lookups
  uint32_t nq, P, cap, nover, quads_per_item;
  uint32_t rank_lo, rank_hi;  // ranks [rank_lo, rank_hi) of every query's leaf list go into the next work list
  uint32_t one;               // always 1; a runtime value so the scan's IMAD accumulates stay IMADs
  uint32_t stage;             // 1: the main scan stages candidates per item in shared memory (large leaves)
  const float* q_for_lut;     // [nq][D] queries: non-NULL = the pilot builds lut / mult / inv_mult itself (fused LUT build)
  uint32_t pilot_cap;         // keys the pilot buffers between selections (0 = the default, 1024 or 2 N' + 128)
  uint32_t pilot_target;      // slots the pilot samples per query before it fixes tau (4 N')
  uint32_t pilot_partial;     // 1: the pilot may stop inside a leaf once the target is reached (small leaves)
  // leaf-sharded search: only the rank that owns a query's nearest leaf (leaf % pilot_world == pilot_rank) samples for
  // tau, the others leave tau = max and the ranks all-reduce(min) it afterwards.  pilot_world <= 1: every query.
  uint32_t pilot_world, pilot_rank;
  uint32_t max_gpt;           // groups per work item of the main scan (0 = kMaxGroupsPerTile)
  uint32_t* item_leaf;        // [item_leaf_cap] leaf of each work item (written by the work-list pass; NULL = none)
  uint32_t item_leaf_cap;
  // expected queries of the batch per (non-empty leaf, probed rank): scan_prepare_phase's choice of the scan kernel
  float qpl_per_rank;
  // set by scan_prepare_phase for the next work list: 0 = octs, 1 = wide quads (scan.cu), 2 = tensor cores (scan_tc.cu)
  uint32_t scan_mode;
  uint32_t max_gpt_simt;      // the SIMT kernels' groups per item (max_gpt is overwritten for the tensor-core items)
  uint32_t rescan;            // 1: the work list holds only the queries whose buffer overflowed (sparse whatever the batch)
  uint8_t* lut_e4m3;          // [2][nq][W*128] workspace of the tensor-core scan (NULL: that path is unavailable)
};

// ---- query preparation ----
void launch_tokenize(const DevIndex& ix, const float* q, uint32_t nq, float* dist, cudaStream_t s);
void launch_topp(const DevIndex& ix, const float* dist, uint32_t nq, uint32_t P, int32_t* leaves,
                 float* bias, cudaStream_t s);
void launch_lut(const DevIndex& ix, const float* q, uint32_t nq, uint8_t* lut, float* mult,
                float* inv_mult, cudaStream_t s);
// tokenization + top-P in one call: tcgen05 GEMM pre-filter + exact refinement when the index carries the bf16
// centre operand (tokenize_tensor_path), else the SIMT pair above.  `a_ws` holds tokenize_operand_bytes(nq, d).
uint32_t tokenize_kpitch(uint32_t d);
size_t tokenize_operand_bytes(uint32_t rows, uint32_t d);
// scale (optional, [d]): rows are multiplied by it (one fp32 rounding) before the split -- the scaled queries of int8
// tokenization
cudaError_t build_tokenize_operand(const float* src, uint32_t rows, uint32_t d, int lo_term, void* out, cudaStream_t s,
                                   const float* scale = nullptr);
bool tokenize_tensor_path(const DevIndex& ix, uint32_t P);
cudaError_t launch_tokenize_topp(const DevIndex& ix, const float* q, uint32_t nq, uint32_t P, float* dist, void* a_ws,
                                 int32_t* leaves, float* bias, uint32_t* fallbacks, cudaStream_t s, int* launches);
// ---- scan ----
size_t pilot_smem_bytes(const DevIndex& ix, const ScanWork& w);
bool pilot_can_build_lut(const DevIndex& ix, const ScanWork& w);
size_t scan_smem_bytes(const DevIndex& ix, uint32_t quads_per_item);
cudaError_t launch_pilot(const DevIndex& ix, const ScanWork& w, cudaStream_t s);
// counted = the per-leaf counts of ranks [rank_lo, rank_hi) are already in leaf_cnt (the pilot kernel does that)
void launch_worklist(const DevIndex& ix, const ScanWork& w, bool only_overflowed, bool counted, cudaStream_t s, int* launches);
cudaError_t launch_scan(const DevIndex& ix, const ScanWork& w, int grid, cudaStream_t s);
// Picks the kernel of the next work list and sets scan_mode / quads_per_item / max_gpt accordingly: call after rank_lo,
// rank_hi (and rescan) are set and before launch_worklist; launch_scan follows the choice.
void scan_prepare_phase(const DevIndex& ix, ScanWork* w);
// ---- tensor-core scan (scan_tc.cu) ----
bool scan_tc_supported(const DevIndex& ix);
cudaError_t launch_scan_tc(const DevIndex& ix, const ScanWork& w, cudaStream_t s);
// *n_launched (optional) receives the number of kernels launched (1-3: small, medium, heavy-tail class)
cudaError_t launch_compact(const DevIndex& ix, const ScanWork& w, bool dedup, cudaStream_t s, int* n_launched = nullptr);
// ---- finalize ----
struct FinalizeArgs {
  const float* q;         // [nq][D]
  uint32_t nq, npre, k, out_k;
  uint32_t* out_idx;      // [nq][out_k]
  float* out_dist;        // [nq][out_k]
  // partial (sharded) outputs, optional
  uint32_t* part_ids; uint64_t* part_tie; float* part_ah; float* part_exact; uint32_t part_cap;
  // packed partial records {tie-break key u64, id u32, exact distance f32} [nq][part_cap] (sharded.cu); used when
  // part_ids is NULL
  uint4* part_rec;
  // optional [nq] score words (f2ord): partial candidates above it are dropped (sampled global threshold, sharded.cu)
  const uint32_t* part_limit;
};
// sampled global threshold (finalize.cu): every 16th score of the sorted local lists -> [nq][S]; the threshold kernel
// reads the all-gathered [world][nq][S] samples
cudaError_t launch_sample_scores(const ScanWork& w, uint32_t S, uint32_t* out, cudaStream_t s);
cudaError_t launch_sample_threshold(const uint32_t* samples, int world, uint32_t nq, uint32_t S, uint32_t nover,
                                    uint32_t* limit, cudaStream_t s);
cudaError_t launch_finalize(const DevIndex& ix, const ScanWork& w, const FinalizeArgs& a, cudaStream_t s);
cudaError_t launch_merge_partials(const DevIndex& ix, uint32_t nq, int world, int n_cand,
                                  const uint32_t* ids, const uint64_t* tie, const float* exact,
                                  uint32_t nover, uint32_t npre, uint32_t k, uint32_t* out_idx,
                                  float* out_dist, uint32_t out_k, cudaStream_t s);
cudaError_t launch_merge_topk(int distance, uint32_t nq, int world, int k_in, const uint32_t* ids, const float* dists,
                              uint32_t k, uint32_t* out_idx, float* out_dist, uint32_t out_k, cudaStream_t s,
                              bool dedup_ids = false);
// Owner-side merge of the sharded search: rec = [world][nq_slice][n_cand] packed records (each list sorted by key,
// padded with id 0xFFFFFFFF) of the nq_valid queries this rank owns; writes rows [0, nq_valid) of out_idx / out_dist.
cudaError_t launch_merge_records(const DevIndex& ix, uint32_t nq_valid, uint32_t nq_slice, int world, int n_cand,
                                 const uint4* rec, uint32_t nover, uint32_t npre, uint32_t k, uint32_t* out_idx,
                                 float* out_dist, uint32_t out_k, cudaStream_t s);
// ---- bf16 brute force (tcgen05 GEMM + fused top-k pre-filter), bruteforce.cu ----
size_t bf_query_operand_bytes(uint32_t nq, uint32_t dpitch);
cudaError_t bf_split_queries(const float* q, uint32_t nq, uint32_t d, uint32_t dpitch, void* a_operand, cudaStream_t s);
cudaError_t bf_init_state(uint32_t nq, uint32_t* cnt, uint64_t* tau, uint32_t* ovf, cudaStream_t s);
// splits = 2: A holds the hi and lo bf16 terms of the queries as two stacked row blocks (bf16 database);
// splits = 1: both operands carry their hi/lo terms concatenated along K (float database, see prep.cu)
cudaError_t bf_gemm_round(const void* a_operand, const void* db, uint32_t nq, uint32_t n_total, uint32_t dpitch,
                          uint32_t row0, uint32_t row1, const ScanWork& w, int splits, cudaStream_t s);
uint32_t bf_query_rows_pad(uint32_t nq);
// window check of the exact re-scoring (bruteforce.cu check_window): flags queries whose k-th exact distance does not
// clear "largest kept approximate distance - eps"
struct BfSafety { float eps_rel, max_row_norm; uint32_t* unsafe; uint32_t* n_unsafe; };
// xnorm != NULL: squared L2 (the candidates' keys come from the augmented dot-product pre-filter, see index.cu)
cudaError_t bf_rescore_f32(const float* q, const float* db, uint32_t nq, uint32_t d, const ScanWork& w, uint32_t kprime,
                           uint32_t k, uint32_t out_k, uint32_t id_base, uint32_t* out_idx, float* out_dist, cudaStream_t s,
                           const BfSafety* safety = nullptr, const float* xnorm = nullptr);
// squared-L2 float brute force: ||x||^2 per row in the reference's arithmetic; rows augmented with one more column
// (database: -||x||^2 / 2, queries: `last` = 1), out [n][d + 1]
cudaError_t bf_row_sqnorms(const float* x, uint32_t n, uint32_t d, float* out, cudaStream_t s);
cudaError_t bf_augment_rows(const float* x, const float* norm, float last, uint32_t n, uint32_t d, float* out, cudaStream_t s);
cudaError_t bf_rescore(const float* q, const void* db, uint32_t nq, uint32_t d, uint32_t dpitch, const ScanWork& w,
                       uint32_t kprime, uint32_t k, uint32_t out_k, uint32_t id_base, uint32_t* out_idx, float* out_dist,
                       cudaStream_t s, const BfSafety* safety = nullptr);
cudaError_t bf_max_row_norm(const void* db, bool f32, uint32_t n, uint32_t d, uint32_t pitch, float* out, cudaStream_t s);
// exact all-rows fallback for flagged queries (see bruteforce.cu)
cudaError_t bf_exact_prepare(uint32_t nq, const uint32_t* unsafe, const ScanWork& w, uint32_t* flagged, uint32_t* n_flagged,
                             cudaStream_t s);
cudaError_t bf_exact_round(const float* q, const void* db, bool f32, uint32_t d, uint32_t dpitch, const uint32_t* flagged,
                           uint32_t n_flagged, uint32_t row0, uint32_t row1, const ScanWork& w, cudaStream_t s,
                           const float* xnorm = nullptr);
cudaError_t bf_exact_emit(const uint32_t* flagged, uint32_t n_flagged, const ScanWork& w, uint32_t k, uint32_t out_k,
                          uint32_t id_base, uint32_t* out_idx, float* out_dist, cudaStream_t s, bool l2 = false);
// out[a_row * ld + b_row] = sum_k A[a_row][k] * B[b_row][k]; bf16 operands with row pitch kpitch (multiple of 64),
// a_rows_pad a multiple of 128 (padding rows readable), fp32 accumulate on tcgen05.
// cmax (optional): cmax[a_row * ld_c + b_row / 32] = max over that row's 32-column chunk of out (cbias == NULL) or of
// 2 * out - cbias[b_row] (squared L2: minus the smallest "||c||^2 - 2 <q, c>" of the chunk).
cudaError_t gemm_bf16_nt(const void* a_operand, uint32_t a_rows, uint32_t a_rows_pad, const void* b_operand,
                         uint32_t b_rows, uint32_t kpitch, float* out, uint32_t ld, cudaStream_t s,
                         float* cmax = nullptr, uint32_t ld_c = 0, const float* cbias = nullptr);
// ---- debug ----
cudaError_t launch_leaf_scores(const DevIndex& ix, const uint8_t* lut, uint32_t leaf, int16_t* out, cudaStream_t s);

}  // namespace sb
