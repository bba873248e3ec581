/*
 * scann_b200.h -- C ABI of the B200-native ScaNN batched query path.
 *
 * Drop-in boundary.  Plain pointers and sizes only; no torch,
 * pybind or C++ types cross this header.  Every entry point names the reference
 * interface it replaces (paths relative to /root/reference/scann/).
 *
 * Error convention: functions return an absl::StatusCode-numbered int
 * (0 OK, 3 INVALID_ARGUMENT, 9 FAILED_PRECONDITION, 12 UNIMPLEMENTED, 13 INTERNAL);
 * the message is available through scann_b200_last_error() (thread local).
 * CUDA failures map to INTERNAL; the library never aborts the process and has
 * no CPU fallback: without a usable CUDA device index creation fails.
 */
#ifndef SCANN_B200_H_
#define SCANN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCANN_B200_ABI_VERSION 6

enum { SCANN_B200_DOT_PRODUCT = 0, SCANN_B200_SQUARED_L2 = 1 };

enum {
  SCANN_B200_OK = 0,
  SCANN_B200_INVALID_ARGUMENT = 3,
  SCANN_B200_FAILED_PRECONDITION = 9,
  SCANN_B200_UNIMPLEMENTED = 12,
  SCANN_B200_INTERNAL = 13,
};

typedef struct scann_b200_index scann_b200_index;

/*
 * Everything a searcher is made of, as host pointers in the serialized-asset layouts
 * (scann_ops/cc/scann.cc:105-233 LoadArtifacts).  The library copies
 * what it needs to the device during scann_b200_index_create; the caller keeps ownership.
 *   tree-AH      : centers + tokens + codes (+ soar_codes) + codebook
 *                  (+ dataset for f32 reordering, or bf16_dataset alone for bfloat16 reordering:
 *                   Bfloat16ReorderingHelper, utils/reordering_helper.cc:720-757, or int8_dataset +
 *                   int8_multipliers (+ dp_norms) for fixed-point reordering)
 *   brute force  : n_leaves = n_blocks = 0 and bf16_dataset (Bfloat16BruteForceSearcher,
 *                  brute_force/bfloat16_brute_force.cc:101-152) or dataset (BruteForceSearcher<float>,
 *                  brute_force/brute_force.cc:376-393); bf16: dot product only (as the reference), float rows:
 *                  dot product or squared L2
 *   n_blocks     : up to 256 (the reference's int16-accumulator range, asymmetric_hashing_impl.cc:656-688)
 */
typedef struct {
  int32_t distance;            /* SCANN_B200_DOT_PRODUCT | SCANN_B200_SQUARED_L2 */
  uint32_t n;                  /* datapoints in the whole (unsharded) database */
  uint32_t d;                  /* dimensionality */
  uint32_t n_leaves;           /* L partitions (serialized_partitioner.pb) */
  uint32_t n_blocks;           /* B AH blocks (ah_codebook.pb) */
  uint32_t dims_per_block;     /* row stride of `codebook` (short blocks are zero padded) */
  const int32_t* block_dims;   /* [B] real dims per block, NULL => all dims_per_block */
  const float* centers;        /* [L][D] */
  const int32_t* tokens;       /* datapoint_to_token.npy: [N], or [2N] with SOAR (-1 = none) */
  int32_t soar;                /* tokens has 2N entries, soar_codes present */
  const uint8_t* codes;        /* hashed_dataset.npy [N][B], one 4-bit code per byte */
  const uint8_t* soar_codes;   /* hashed_dataset_soar.npy [N][B] or NULL */
  const float* codebook;       /* [B][16][dims_per_block] */
  const float* dataset;        /* dataset.npy [N][D] or NULL (no exact reordering) */
  const int16_t* bf16_dataset; /* bfloat16_dataset.npy [N][D] or NULL */
  float overretrieve;          /* DatabaseSpillingConfig.overretrieve_factor */
  int32_t default_leaves;      /* QuerySpillingConfig.max_spill_centers */
  int32_t default_pre_nn;      /* ExactReordering.approx_num_neighbors */
  int32_t default_final_nn;    /* ScannConfig.num_neighbors */
  int32_t device;              /* CUDA device ordinal */
  int32_t shard_rank;          /* this index holds datapoints with id % shard_world == shard_rank */
  int32_t shard_world;         /* 1 = unsharded */
  /* int8 (fixed point) reordering, used when dataset and bf16_dataset are NULL
   * (FixedPointFloatDense{DotProduct,SquaredL2}ReorderingHelper, utils/reordering_helper.cc:384-441,581-618) */
  const int8_t* int8_dataset;      /* int8_dataset.npy [N][D] */
  const float* int8_multipliers;   /* int8_multipliers.npy [D]: multiplier_by_dimension */
  const float* dp_norms;           /* dp_norms.npy [N]: squared L2 norms of the original rows (squared L2 only) */
  /* how a tree-AH database is split over shard_world ranks:
   *   SCANN_B200_SHARD_BY_ID   every leaf is split, this rank holds datapoints with id % shard_world == shard_rank
   *   SCANN_B200_SHARD_BY_LEAF whole leaves (incl. their SOAR copies) are dealt out, this rank holds the leaves with
   *                            leaf % shard_world == shard_rank and the rows of the datapoints stored in them */
  int32_t shard_mode;
  /* PartitioningConfig.query_tokenization_type (tree(quantize_centroids=True) in scann_builder.py:231):
   *   SCANN_B200_TOKENIZE_FLOAT             the batched float path (kmeans_tree_partitioner.cc:642-730)
   *   SCANN_B200_TOKENIZE_FIXED_POINT_INT8  KMeansTreeNode::GetAllDistancesInt8 (trees/kmeans_tree/kmeans_tree_node.h:
   *       222-256) on the centres KMeansTreeNode::CreateFixedPointCenters (kmeans_tree_node.cc:267-281) derives from
   *       `centers` at load time -- the library derives them the same way, nothing extra is serialized.  The leaf
   *       bias of the AH scores is then the int8 distance, as in the reference.  Tree-AH only. */
  int32_t query_tokenization_type;
} scann_b200_index_desc;

enum { SCANN_B200_SHARD_BY_ID = 0, SCANN_B200_SHARD_BY_LEAF = 1 };
enum { SCANN_B200_TOKENIZE_FLOAT = 0, SCANN_B200_TOKENIZE_FIXED_POINT_INT8 = 1 };

/* Replaces ScannInterface::Initialize(ScannArtifacts) (scann_ops/cc/scann.cc:355-381) and the
 * per-leaf asymmetric_hashing2::Searcher construction incl. CreatePackedDataset
 * (hashes/internal/asymmetric_hashing_impl.cc:690-737). */
int scann_b200_index_create(const scann_b200_index_desc* desc, scann_b200_index** out);
void scann_b200_index_destroy(scann_b200_index* index);

/* Replaces ScannInterface::SearchBatched (scann_ops/cc/scann.cc:463-475) +
 * ReshapeBatchedNNResult (scann.h:165-180): host buffers in, host buffers out.
 * final_nn / pre_reorder_nn / leaves: -1 = index default (scann_ops_pybind.py:75-78).
 * out_idx / out_dist are [nq][out_k]; short rows are padded with (0, NaN). */
int scann_b200_search_batched(scann_b200_index* index, const float* queries, uint32_t nq,
                              int32_t final_nn, int32_t pre_reorder_nn, int32_t leaves,
                              uint32_t* out_idx, float* out_dist, int32_t out_k);

/* Same computation with every buffer already resident on the index's device
 * (queries [nq][D] f32, outputs [nq][out_k]); enqueued on the index's stream and
 * synchronised before returning.  This is the kernel-side throughput leg of bench.py.
 * Stream ordering: the index's streams are non-blocking and do NOT wait for the caller's streams.  The
 * producer of d_queries (and of any other device input of the *_device entry points) must have COMPLETED
 * -- cudaStreamSynchronize / cudaEventSynchronize on the caller's side -- before the call; the outputs are
 * complete when the call returns. */
int scann_b200_search_batched_device(scann_b200_index* index, const float* d_queries, uint32_t nq,
                                     int32_t final_nn, int32_t pre_reorder_nn, int32_t leaves,
                                     uint32_t* d_out_idx, float* d_out_dist, int32_t out_k);

/* Sharded search: this rank's local pre-reorder candidates with their
 * exact distance already computed.  Records are (global datapoint id, AH score, exact distance),
 * rows sorted by (AH score, tie-break key), padded with id 0xFFFFFFFF.  Device buffers
 * [nq][n_cand]; n_cand must be >= the over-retrieved pre-reorder count. */
int scann_b200_search_partial_device(scann_b200_index* index, const float* d_queries, uint32_t nq,
                                     int32_t pre_reorder_nn, int32_t leaves, uint32_t* d_ids,
                                     uint64_t* d_tiebreak, float* d_ah_score, float* d_exact,
                                     int32_t n_cand);
/* Merge `world` gathered partial lists ([world][nq][n_cand], as all-gathered) into the final
 * top-k, reproducing "global top-N' by AH score -> exact distance -> top-k by (distance, id)". */
int scann_b200_merge_partials_device(scann_b200_index* index, uint32_t nq, int32_t world,
                                     int32_t n_cand, const uint32_t* d_ids,
                                     const uint64_t* d_tiebreak, const float* d_ah_score,
                                     const float* d_exact, int32_t pre_reorder_nn, int32_t final_nn,
                                     uint32_t* d_out_idx, float* d_out_dist, int32_t out_k);

/*
 * Database-sharded search, one process per GPU (the reference has no multi-device searcher:
 * the single-device semantics are those of scann_b200_search_batched_device, and the result is bit-identical to it).
 *   scann_b200_comm_unique_id   rank 0 creates the 128-byte NCCL id; the caller distributes it (any side channel)
 *   scann_b200_comm_init        every rank joins; `index` must have been created with the same shard_rank / shard_world
 *   scann_b200_search_sharded_device   every rank calls it with the SAME queries (device buffers, [nq][D]); every rank
 *                               receives the full result ([nq][out_k] device buffers).
 * Per batch, on the index's stream, with NCCL over NVLink as the only exchange:
 *   1. each rank tokenizes nq / world queries; all-gather of the (leaf, centre distance) lists          8 B x P per query
 *   2. the rank that owns a query's nearest leaf fixes the pruning threshold tau; all-reduce(min)         8 B per query
 *   3. LUT16 scan of the rank's own leaves / datapoints, local top-N', exact distances of those candidates
 *   4. all-to-all: the 16-byte records (tie-break key, id, exact distance) of query q go to rank q / ceil(nq / world)
 *   5. that rank merges world sorted lists: global top-N' -> SOAR dedup -> top-k by (distance, id)
 *   6. all-gather of the k results                                                                       8 B x k per query
 * light != 0 selects the north star's light protocol instead of 4-6: every rank finishes its local top-k and one
 * all-gather of (id, distance) x k per query per rank is merged (duplicates by id removed).  The light result
 * reorders a superset of the single-GPU candidates, so it is NOT bit-identical (recall >= the parity mode's).
 */
int scann_b200_comm_unique_id(void* out_id128);
int scann_b200_comm_init(scann_b200_index* index, int32_t rank, int32_t world, const void* id128);
int scann_b200_search_sharded_device(scann_b200_index* index, const float* d_queries, uint32_t nq, int32_t final_nn,
                                     int32_t pre_reorder_nn, int32_t leaves, int32_t light, uint32_t* d_out_idx,
                                     float* d_out_dist, int32_t out_k);
/* The same protocol with all `world` shards living in THIS process (indexes created with shard_rank 0..world-1, on
 * one device or several): the exchanges are device copies.  Used by the tests to run world-size 2/3/4/8 searches on
 * one GPU; the result is written once. */
int scann_b200_search_sharded_local(scann_b200_index* const* shards, int32_t world, const float* d_queries, uint32_t nq,
                                    int32_t final_nn, int32_t pre_reorder_nn, int32_t leaves, int32_t light,
                                    uint32_t* d_out_idx, float* d_out_dist, int32_t out_k);

/* Row-sharded brute force (the reference has no multi-device searcher, the single
 * device semantics are Bfloat16BruteForceSearcher::FindNeighborsImpl, brute_force/bfloat16_brute_force.cc:101-152).
 * A brute-force index created with shard_world > 1 keeps the contiguous rows
 * [rank * ceil(N / world), ...) and its search calls return GLOBAL ids.  This entry point merges the
 * `world` all-gathered local results ([world][nq][k_in] ids + API-signed distances) into the global
 * top-k by (distance, id): identical to the unsharded result. */
int scann_b200_merge_topk_device(scann_b200_index* index, uint32_t nq, int32_t world, int32_t k_in,
                                 const uint32_t* d_ids, const float* d_dists, int32_t final_nn,
                                 uint32_t* d_out_idx, float* d_out_dist, int32_t out_k);

const char* scann_b200_last_error(void);
int scann_b200_abi_version(void);

/* ---- index construction, deterministic part ---- */

/*
 * What builder(...).build() does to every datapoint once the partitioner and the AH codebook are trained:
 *   database tokenization   KMeansTreePartitioner::TokenizeDatabase (partitioning/kmeans_tree_partitioner.cc:475-560)
 *                           -> DenseDistanceManyToManyTop1 under SquaredL2Distance (scann_builder.py:213-238)
 *   SOAR secondary leaf     OrthogonalityAmplifiedTokenForDatapointBatched (:925-997),
 *                           DenseManyToManyOrthogonalityAmplified (distance_measures/many_to_many/many_to_many_impl.inc:729-781)
 *   residual + AH codes     TreeAHHybridResidual::BuildLeafSearchers (tree_x_hybrid/tree_ah_hybrid_residual.cc:395-428)
 *                           -> Indexer::Hash / HashWithNoiseShaping (hashes/asymmetric_hashing2/indexing.cc:87-246,
 *                           hashes/internal/asymmetric_hashing_impl.cc:199-244,434-503)
 * Host pointers in, host pointers out, in the serialized-asset layouts of scann_b200_index_desc.
 */
typedef struct {
  uint32_t n, d;                   /* datapoints, dimensionality */
  uint32_t n_leaves;               /* L */
  uint32_t n_blocks;               /* B <= 256 */
  uint32_t dims_per_block;         /* row stride of `codebook` */
  const int32_t* block_dims;       /* [B] real dims per block (must sum to d), NULL => all dims_per_block */
  const float* dataset;            /* [N][D] */
  const float* centers;            /* [L][D] trained partition centres */
  const float* codebook;           /* [B][16][dims_per_block] trained AH centres */
  int32_t residual;                /* 1: hash x - centre[token] (use_residual_quantization, dot product); 0: hash x */
  float soar_lambda;               /* orthogonality_amplification_lambda; NaN = no database spilling */
  double noise_shaping_threshold;  /* AsymmetricHasherConfig.noise_shaping_threshold; NaN = Indexer::Hash */
  int32_t device;
} scann_b200_encode_desc;

typedef struct {
  float ms_tokenize, ms_soar, ms_encode, ms_total; /* CUDA events; ms_total includes the host <-> device copies */
  uint64_t soar_evaluated;      /* exact SOAR cost evaluations (the reference evaluates N * L) */
  uint64_t spilled;             /* datapoints whose secondary leaf differs from the primary */
  uint64_t norm_ties;           /* noise shaping: datapoints whose initial block norms tie (visiting order unspecified in the reference) */
  uint64_t tokenize_fallbacks;  /* rows whose tensor-core tokenization pre-filter fell back to exact distances */
  uint32_t chunk_rows;          /* rows processed per pass */
} scann_b200_encode_stats;

/* tokens_out: [N] i32, or [2N] with SOAR (slot 2i = lower-numbered leaf, 2i + 1 = the other leaf or -1:
 * scann_ops/cc/scann.cc:533-551); codes_out [N][B] (the code of slot 2i's leaf); soar_codes_out [N][B] (slot 2i + 1,
 * zero when not spilled; required with SOAR, else ignored).  stats may be NULL. */
int scann_b200_encode_database(const scann_b200_encode_desc* desc, int32_t* tokens_out, uint8_t* codes_out,
                               uint8_t* soar_codes_out, scann_b200_encode_stats* stats);

/* ---- index construction, training ---- */

/*
 * Lloyd iterations of the k-means trainers: KMeansTreePartitioner::TrainKMeans (partitioning/kmeans_tree_partitioner.cc:
 * 424-441) -> GmmUtils::GenericKmeans (utils/gmm_utils.cc:846-915: assignment with UnbalancedFloat32PartitionAssignment,
 * RecomputeCentroidsSimple :1052-1132), and the 16-centre AH codebooks per block (hashes/internal/
 * asymmetric_hashing_impl.cc:41-197).  Deterministic: the result is a function of (data, init_centers, iterations).
 * The caller supplies the initial centres (the reference draws them from an unseeded generator); a cluster that ends an
 * iteration empty keeps its centre (the reference re-initialises small clusters randomly).  Host pointers in / out.
 */
typedef struct {
  uint32_t n, d, k;            /* training points, dimensionality, clusters (k <= n) */
  const float* data;           /* [n][d] */
  const float* init_centers;   /* [k][d] */
  int32_t iterations;          /* centroid recomputations (GmmUtils::Options::max_iterations) */
  int32_t device;
} scann_b200_kmeans_desc;

typedef struct {
  float ms_assign, ms_update, ms_total;  /* CUDA events */
  uint32_t iterations;
  uint32_t empty_clusters;     /* clusters that were empty in the last recomputation */
  double mean_sq_distance;     /* mean squared distance of the final assignment (0 when assign_out is NULL) */
} scann_b200_kmeans_stats;

/* centers_out [k][d]; assign_out [n] (the final partition: nearest centre of every point) or NULL; stats may be NULL. */
int scann_b200_train_kmeans(const scann_b200_kmeans_desc* desc, float* centers_out, int32_t* assign_out,
                            scann_b200_kmeans_stats* stats);

/* ---- serialized assets (the reference's on-disk format) ---- */

typedef struct scann_b200_assets scann_b200_assets;

/* Replaces ScannInterface::LoadArtifacts (scann_ops/cc/scann.cc:105-264): reads binary
 * scann_config.pb from `artifacts_dir` and every asset listed in the text-format ScannAssets
 * manifest (`assets_pbtxt` = contents of scann_assets.pbtxt; relative paths are re-rooted at
 * artifacts_dir): serialized_partitioner.pb, ah_codebook.pb, datapoint_to_token.npy,
 * hashed_dataset[_soar].npy, dataset.npy, bfloat16_dataset.npy, int8_dataset.npy + int8_multipliers.npy +
 * dp_norms.npy.  The handle owns the host copies. */
int scann_b200_assets_load(const char* artifacts_dir, const char* assets_pbtxt, scann_b200_assets** out);
void scann_b200_assets_free(scann_b200_assets* assets);
/* Fills an index descriptor whose pointers alias the handle's memory (valid until _free) and whose
 * defaults (num_neighbors, approx_num_neighbors, max_spill_centers, SOAR, overretrieve, distance)
 * come from the loaded ScannConfig.  device / shard fields are left 0 / 0 / 1. */
int scann_b200_assets_describe(const scann_b200_assets* assets, scann_b200_index_desc* out);
/* The loaded ScannConfig in protobuf text format (ScannNumpy::config, scann_npy.cc). */
const char* scann_b200_assets_config(const scann_b200_assets* assets);
/* Replaces ScannInterface::Serialize (scann_ops/cc/scann.cc:504-601) + the manifest of
 * ScannNumpy::Serialize (scann_npy.cc:272-282): writes scann_config.pb (binary ScannConfig encoded
 * from `config_text`) and one file per non-null array of `desc`; returns the text-format
 * ScannAssets manifest in `assets_pbtxt_out` (the caller writes scann_assets.pbtxt). */
int scann_b200_assets_save(const char* artifacts_dir, const scann_b200_index_desc* desc,
                           const char* config_text, int relative_path, char* assets_pbtxt_out,
                           size_t assets_pbtxt_cap);
/* ScannConfig text <-> binary (ParseTextProto scann.h:185-188; WriteProtobufToFile). */
int scann_b200_config_text_to_binary(const char* text, void* out, size_t cap, size_t* out_len);
int scann_b200_config_binary_to_text(const void* bin, size_t len, char* out, size_t cap);

/* ---- parity / measurement hooks (used by tests/ and bench.py) ---- */

/* KMeansTreePartitioner::TokensForDatapointWithSpillingBatched
 * (partitioning/kmeans_tree_partitioner.cc:642-730): the P nearest leaves per query and the
 * distance to their centres, sorted by (distance, leaf).  Host buffers [nq][P]. */
int scann_b200_debug_tokenize(scann_b200_index* index, const float* queries, uint32_t nq,
                              int32_t leaves, int32_t* out_leaf, float* out_dist);
/* AsymmetricQueryer::CreateLookupTable (hashes/asymmetric_hashing2/querying.h:284-329):
 * u8 LUT [nq][B][16] and fixed point multiplier [nq]. */
int scann_b200_debug_lut(scann_b200_index* index, const float* queries, uint32_t nq,
                         uint8_t* out_lut, float* out_mult);
/* LUT16Interface::GetDistances (hashes/internal/lut16_interface.h:40-135): int16 scores of
 * every slot of `leaf` under the given u8 LUT [B][16].  out has leaf-size entries. */
int scann_b200_debug_leaf_scores(scann_b200_index* index, const uint8_t* lut, uint32_t leaf,
                                 int16_t* out, uint32_t out_len);
/* TreeAHHybridResidual::FindNeighborsBatchedImpl (tree_x_hybrid/tree_ah_hybrid_residual.cc:631-786)
 * up to FinishUnsorted: the over-retrieved pre-reorder candidates per query, sorted by
 * (score, leaf, slot).  Host buffers [nq][cap]; out_count [nq]. */
int scann_b200_debug_candidates(scann_b200_index* index, const float* queries, uint32_t nq,
                                int32_t pre_reorder_nn, int32_t leaves, int32_t cap,
                                uint32_t* out_leaf, uint32_t* out_slot, uint32_t* out_dp,
                                float* out_score, uint32_t* out_count);
uint32_t scann_b200_leaf_size(const scann_b200_index* index, uint32_t leaf);

typedef struct {
  uint64_t scan_bytes_alg;   /* sum over probed (query, leaf) of ceil(n/32)*16*B */
  uint64_t scan_pairs;       /* probed (query, leaf) pairs */
  uint64_t scan_lookups;     /* (query, slot, block) lookups issued by the scan kernels */
  uint32_t kernel_launches;  /* kernels launched by the last search call */
  uint32_t overflow_retries; /* candidate-buffer overflow re-scans (0 in the common case) */
  float ms_tokenize, ms_lut, ms_pilot, ms_worklist, ms_scan, ms_compact, ms_finalize, ms_total;
  uint32_t scan_kernel_count; /* launches of the dominant LUT16 scan kernel in the last call */
  uint64_t cand_sum;          /* candidates buffered per query after the main scan, summed */
  uint64_t cand_max;          /* ... and the maximum over queries */
  uint64_t tokenize_fallbacks; /* queries whose tensor-core tokenization pre-filter fell back to exact distances */
  /* sharded search: collectives (all of them, CUDA events around the NCCL calls) and the owner-side merge */
  float ms_exchange, ms_merge;
  uint64_t exchange_bytes;     /* bytes this rank sent in the last call */
  uint32_t bf_widenings;       /* brute force: re-runs with a wider candidate window / safe rounds (0 in the common case) */
  uint32_t bf_exact_fallbacks; /* brute force: queries finished by the exact all-rows kernel */
  /* main-scan launches by kernel: octs (dense work lists), wide quads (sparse ones), tensor cores (opt-in) */
  uint32_t scan_oct_launches, scan_wide_launches, scan_tc_launches;
  uint32_t reserved0;
} scann_b200_stats;
/* Timing (CUDA events on the index's stream) and traffic figures of the last search call. */
int scann_b200_last_stats(scann_b200_index* index, scann_b200_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* SCANN_B200_H_ */
