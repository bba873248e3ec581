// index.cu -- the C ABI (include/scann_b200.h) and the host-side search driver.
//
// Host-side mirror of, in the reference:
//   ScannInterface::Initialize / SearchBatched        scann_ops/cc/scann.cc:355-381,463-475
//   SingleMachineSearcherBase::FindNeighborsBatched   base/single_machine_base.cc:569-587
//   TreeAHHybridResidual::BuildLeafSearchers           tree_x_hybrid/tree_ah_hybrid_residual.cc:325-495
//   TreeAHHybridResidual::FindNeighborsBatchedImpl     tree_x_hybrid/tree_ah_hybrid_residual.cc:631-846
// There is no CPU path: every stage is a CUDA kernel; without a device creation fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "index_internal.h"

namespace {

thread_local std::string g_err;

size_t word_pos(int W, int j, int m) {
  const int N4 = W / 4, R = W % 4;
  if (j < 4 * N4) return (size_t)(j / 4) * 128 + (size_t)m * 4 + (j % 4);
  const int jj = j - 4 * N4;
  if (R >= 2 && jj < 2) return (size_t)N4 * 128 + (size_t)m * 2 + jj;
  return (size_t)N4 * 128 + (R >= 2 ? 64 : 0) + m;
}

}  // namespace

namespace sbi {
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace sbi
using namespace sbi;

namespace sb {
void set_last_error(const char* msg) { g_err = msg ? msg : ""; }  // used by assets.cc
}  // namespace sb

namespace {

int build_index(scann_b200_index* ix, const scann_b200_index_desc* d) {
  const uint32_t N = d->n, D = d->d, L = d->n_leaves, B = d->n_blocks;
  if (d->distance != SCANN_B200_DOT_PRODUCT && d->distance != SCANN_B200_SQUARED_L2)
    return fail(SCANN_B200_INVALID_ARGUMENT, "unknown distance measure %d", d->distance);
  if (!L && !B) {
    // Bfloat16BruteForceSearcher (brute_force/bfloat16_brute_force.cc) or BruteForceSearcher<float>
    // (brute_force/brute_force.cc:376-393): MIPS only
    if (!d->bf16_dataset && !d->dataset) return fail(SCANN_B200_INVALID_ARGUMENT, "brute force needs dataset or bf16_dataset");
    // squared L2: BruteForceSearcher<float> only (Bfloat16BruteForceSearcher is MIPS-only, bfloat16_brute_force.cc:60-75)
    if (d->distance != SCANN_B200_DOT_PRODUCT && d->bf16_dataset)
      return fail(SCANN_B200_UNIMPLEMENTED, "bfloat16 brute force supports dot product distance only");
    // row-sharded brute force: rank r keeps the contiguous rows [r * ceil(N / world), ...) and reports global ids
    const int world_bf = d->shard_world > 0 ? d->shard_world : 1;
    const int rank_bf = d->shard_rank;
    if (rank_bf < 0 || rank_bf >= world_bf) return fail(SCANN_B200_INVALID_ARGUMENT, "shard_rank %d outside [0, %d)", rank_bf, world_bf);
    const uint32_t per = (N + (uint32_t)world_bf - 1) / (uint32_t)world_bf;
    const uint32_t row0 = std::min<uint64_t>((uint64_t)per * rank_bf, N), row1 = std::min<uint64_t>((uint64_t)row0 + per, N);
    const uint32_t nloc = row1 - row0;
    sb::DevIndex& vb = ix->dev;
    vb = sb::DevIndex{};
    vb.distance = d->distance; vb.n = nloc; vb.d = D; vb.disjoint = 1;
    ix->brute = true;
    ix->bf_row0 = row0;
    if (!d->bf16_dataset) {
      // float rows: kept as they are for the exact re-scoring, plus the bf16 GEMM operand [hi | hi | lo]
      // (K = 3D padded to 64; the queries are [hi | lo | hi], so one GEMM gives qh.xh + ql.xh + qh.xl)
      ix->bf_f32 = true;
      CU(ix->dataset.ensure(sizeof(float) * (size_t)std::max<uint32_t>(nloc, 1) * D));
      if (nloc) CU(cudaMemcpy(ix->dataset.p, d->dataset + (size_t)row0 * D, sizeof(float) * (size_t)nloc * D, cudaMemcpyHostToDevice));
      vb.dataset = ix->dataset.as<float>();
      if (d->distance == SCANN_B200_SQUARED_L2) {
        // Squared L2 (brute_force.cc:376-393 with SquaredL2Distance): ||q - x||^2 orders like -(<q, x> - ||x||^2 / 2), the
        // dot product of the AUGMENTED vectors x' = [x, -||x||^2 / 2], q' = [q, 1].  The tensor-core pre-filter runs on
        // the augmented operands (D + 1 dimensions); the re-scoring evaluates the reference's squared-L2 chain on the
        // original rows with the row norms kept here.
        ix->bf_l2 = true;
        ix->bf_dpitch = sb::tokenize_kpitch(D + 1);
        CU(ix->bf_xnorm.ensure(sizeof(float) * std::max<uint32_t>(nloc, 1)));
        CU(sb::bf_row_sqnorms(ix->dataset.as<float>(), nloc, D, ix->bf_xnorm.as<float>(), 0));
        sbi::DevBuf aug;
        CU(aug.ensure(sizeof(float) * (size_t)std::max<uint32_t>(nloc, 1) * (D + 1)));
        CU(sb::bf_augment_rows(ix->dataset.as<float>(), ix->bf_xnorm.as<float>(), 0.f, nloc, D, aug.as<float>(), 0));
        CU(ix->bf_db.ensure(sb::tokenize_operand_bytes(std::max<uint32_t>(nloc, 1), D + 1)));
        CU(sb::build_tokenize_operand(aug.as<float>(), nloc, D + 1, 2, ix->bf_db.p, 0));
        CU(cudaStreamSynchronize(0));
        CU(sb::bf_max_row_norm(aug.p, true, nloc, D + 1, D + 1, &ix->bf_max_row_norm, 0));
        return 0;
      }
      ix->bf_dpitch = sb::tokenize_kpitch(D);
      CU(ix->bf_db.ensure(sb::tokenize_operand_bytes(std::max<uint32_t>(nloc, 1), D)));
      CU(sb::build_tokenize_operand(ix->dataset.as<float>(), nloc, D, 2, ix->bf_db.p, 0));
      CU(cudaStreamSynchronize(0));
      CU(sb::bf_max_row_norm(ix->dataset.p, true, nloc, D, D, &ix->bf_max_row_norm, 0));
      return 0;
    }
    ix->bf_dpitch = (D + 7) / 8 * 8;
    CU(ix->bf_db.ensure((size_t)std::max<uint32_t>(nloc, 1) * ix->bf_dpitch * 2));
    CU(cudaMemset(ix->bf_db.p, 0, (size_t)std::max<uint32_t>(nloc, 1) * ix->bf_dpitch * 2));
    if (nloc) CU(cudaMemcpy2D(ix->bf_db.p, (size_t)ix->bf_dpitch * 2, d->bf16_dataset + (size_t)row0 * D, (size_t)D * 2,
                              (size_t)D * 2, nloc, cudaMemcpyHostToDevice));
    CU(sb::bf_max_row_norm(ix->bf_db.p, false, nloc, D, ix->bf_dpitch, &ix->bf_max_row_norm, 0));
    return 0;
  }
  if (!L || !B) return fail(SCANN_B200_UNIMPLEMENTED, "only tree-AH and bf16 brute-force indexes are implemented (n_leaves=%u, n_blocks=%u)", L, B);
  if (!d->centers || !d->tokens || !d->codes || !d->codebook)
    return fail(SCANN_B200_INVALID_ARGUMENT, "tree-AH index needs centers, tokens, codes and codebook");
  // B <= 256: the reference's int16 accumulator covers it (CanUseInt16Accumulator, asymmetric_hashing_impl.cc:656-688);
  // above that it switches to 32-bit sums, which this path does not restate
  if (B > 256) return fail(SCANN_B200_UNIMPLEMENTED, "n_blocks=%u > 256 not supported", B);
  if (d->query_tokenization_type != SCANN_B200_TOKENIZE_FLOAT && d->query_tokenization_type != SCANN_B200_TOKENIZE_FIXED_POINT_INT8)
    return fail(SCANN_B200_INVALID_ARGUMENT, "unknown query_tokenization_type %d", d->query_tokenization_type);
  if (d->soar && !d->soar_codes) return fail(SCANN_B200_INVALID_ARGUMENT, "SOAR index without soar_codes");
  const int world = d->shard_world > 0 ? d->shard_world : 1, rank = d->shard_rank;
  if (rank < 0 || rank >= world) return fail(SCANN_B200_INVALID_ARGUMENT, "bad shard rank %d/%d", rank, world);
  if (d->shard_mode != SCANN_B200_SHARD_BY_ID && d->shard_mode != SCANN_B200_SHARD_BY_LEAF)
    return fail(SCANN_B200_INVALID_ARGUMENT, "unknown shard_mode %d", d->shard_mode);
  const bool by_leaf = world > 1 && d->shard_mode == SCANN_B200_SHARD_BY_LEAF;
  ix->shard_rank = rank; ix->shard_world = world; ix->shard_mode = d->shard_mode;
  // which (datapoint, leaf) pairs this rank stores: a residue class of the ids inside every leaf, or whole leaves
  auto mine = [&](uint32_t i, int32_t t) -> bool {
    return by_leaf ? ((uint32_t)t % (uint32_t)world == (uint32_t)rank) : (i % (uint32_t)world == (uint32_t)rank);
  };
  sb::DevIndex& v = ix->dev;
  v.distance = d->distance; v.n = N; v.d = D; v.L = L; v.B = B; v.W = (B + 7) / 8; v.dpb = d->dims_per_block;
  const int W = (int)v.W;

  // datapoints_by_token in file order (scann_ops/cc/scann.cc:88-98), filtered to this shard
  const uint32_t mult = d->soar ? 2 : 1;
  const size_t len = (size_t)N * mult;
  std::vector<uint32_t> lsize(L, 0), lsize_full(L, 0);
  for (size_t j = 0; j < len; ++j) {
    const int32_t t = d->tokens[j];
    if (t < 0) continue;
    if ((uint32_t)t >= L) return fail(SCANN_B200_INVALID_ARGUMENT, "token %d out of range [0,%u)", t, L);
    lsize_full[t]++;
    if (!mine((uint32_t)(j / mult), t)) continue;
    lsize[t]++;
  }
  // unsharded group offsets: the tie-break half of a candidate key must be the slot the
  // datapoint has in the UNSHARDED index, so that sharded + merged == single GPU
  std::vector<uint32_t> goff_full(L + 1, 0);
  for (uint32_t l = 0; l < L; ++l) goff_full[l + 1] = goff_full[l] + (lsize_full[l] + 31) / 32;
  std::vector<uint32_t> goff(L + 1, 0), ntiles(L, 0), gpt(L, 0);
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t ng = (lsize[l] + 31) / 32;
    goff[l + 1] = goff[l] + ng;
    if (ng) {
      ntiles[l] = (ng + sb::kMaxGroupsPerTile - 1) / sb::kMaxGroupsPerTile;
      gpt[l] = (ng + ntiles[l] - 1) / ntiles[l];
    }
  }
  const size_t ngroups = goff[L];
  {
    uint32_t nonempty = 0;  // a leaf-sharded rank holds L / world leaves; the rest are empty here
    for (uint32_t l = 0; l < L; ++l) nonempty += lsize[l] ? 1u : 0u;
    ix->avg_leaf_slots = (uint32_t)(ngroups * 32 / std::max<uint32_t>(by_leaf ? nonempty : L, 1));
    ix->nonempty_leaves = nonempty;
  }
  if (ngroups * 32 > 0xFFFFFFF0ull) return fail(SCANN_B200_UNIMPLEMENTED, "more than 2^32 slots");
  std::vector<uint32_t> slot_dp(ngroups * 32, 0xFFFFFFFFu);
  std::vector<uint32_t> slot_tie(world > 1 ? ngroups * 32 : 0, 0xFFFFFFFFu);
  std::vector<uint32_t> codes((size_t)ngroups * W * 32, 0u);
  int disjoint = 1;
  {
    std::vector<uint32_t> cur(L, 0), cur_full(L, 0);
    for (size_t j = 0; j < len; ++j) {
      const int32_t t = d->tokens[j];
      if (t < 0) continue;
      const uint32_t i = (uint32_t)(j / mult);
      if (d->soar && (j & 1) && d->tokens[j - 1] >= 0) disjoint = 0;
      const uint32_t s_full = cur_full[t]++;
      if (!mine(i, t)) continue;
      const uint32_t s = cur[t]++;
      const size_t g = goff[t] + s / 32;
      const int m = (int)(s % 32);
      slot_dp[g * 32 + m] = i;
      if (world > 1) slot_tie[g * 32 + m] = goff_full[t] * 32 + s_full;
      // tree_ah_hybrid_residual.cc:385-396: the SOAR code row iff tok[2i+1] == leaf
      const uint8_t* row = d->codes + (size_t)i * B;
      if (d->soar && d->tokens[2 * (size_t)i + 1] == t) row = d->soar_codes + (size_t)i * B;
      uint32_t* gw = codes.data() + g * W * 32;
      for (int jw = 0; jw < W; ++jw) {
        uint32_t wv = 0;
        for (int k = 0; k < 8; ++k) {
          const uint32_t b = 8 * jw + k;
          if (b < B) {
            if (row[b] > 15) return fail(SCANN_B200_INVALID_ARGUMENT, "AH code %u > 15 at datapoint %u block %u", row[b], i, b);
            wv |= (uint32_t)row[b] << (4 * k);
          }
        }
        gw[word_pos(W, jw, m)] = wv;
      }
    }
  }
  v.disjoint = disjoint;
  v.key_by_dp = d->distance == SCANN_B200_SQUARED_L2 ? 1 : 0;
  if (v.key_by_dp && d->soar) return fail(SCANN_B200_INVALID_ARGUMENT, "SOAR requires dot product distance.");
  ix->h_leaf_size = lsize;

  std::vector<int32_t> bdims(B);
  std::vector<uint32_t> boff(B + 1, 0);
  for (uint32_t b = 0; b < B; ++b) {
    bdims[b] = d->block_dims ? d->block_dims[b] : (int32_t)d->dims_per_block;
    if (bdims[b] <= 0 || (uint32_t)bdims[b] > d->dims_per_block)
      return fail(SCANN_B200_INVALID_ARGUMENT, "block %u has %d dims (stride %u)", b, bdims[b], d->dims_per_block);
    boff[b + 1] = boff[b] + (uint32_t)bdims[b];
  }
  if (boff[B] != D) return fail(SCANN_B200_INVALID_ARGUMENT, "AH blocks cover %u dims, dimensionality is %u", boff[B], D);

#define UP(buf, ptr, bytes_)                                                        \
  do {                                                                              \
    CU(buf.ensure((bytes_) ? (bytes_) : 16));                                       \
    if (bytes_) CU(cudaMemcpy(buf.p, ptr, (bytes_), cudaMemcpyHostToDevice));       \
  } while (0)
  UP(ix->centers, d->centers, sizeof(float) * (size_t)L * D);
  if (d->distance == SCANN_B200_SQUARED_L2) {
    // many_to_many_impl.inc:236-257: ||c||^2 = -(fnmadd chain over dims)
    std::vector<float> cn(L);
    for (uint32_t l = 0; l < L; ++l) {
      float a = 0.f;
      for (uint32_t k = 0; k < D; ++k) a = fmaf(-d->centers[(size_t)l * D + k], d->centers[(size_t)l * D + k], a);
      cn[l] = a * -1.0f;
    }
    UP(ix->cnorm, cn.data(), sizeof(float) * L);
  }
  UP(ix->codebook, d->codebook, sizeof(float) * (size_t)B * 16 * d->dims_per_block);
  UP(ix->block_dims, bdims.data(), sizeof(int32_t) * B);
  UP(ix->block_off, boff.data(), sizeof(uint32_t) * (B + 1));
  UP(ix->leaf_size, lsize.data(), sizeof(uint32_t) * L);
  UP(ix->leaf_goff, goff.data(), sizeof(uint32_t) * (L + 1));
  UP(ix->leaf_ntiles, ntiles.data(), sizeof(uint32_t) * L);
  UP(ix->leaf_gpt, gpt.data(), sizeof(uint32_t) * L);
  UP(ix->codes, codes.data(), sizeof(uint32_t) * codes.size());
  UP(ix->slot_dp, slot_dp.data(), sizeof(uint32_t) * slot_dp.size());
  v.slot_tie = nullptr;
  if (world > 1) {
    UP(ix->slot_tie, slot_tie.data(), sizeof(uint32_t) * slot_tie.size());
    v.slot_tie = ix->slot_tie.as<uint32_t>();
  }
  v.dataset = nullptr; v.dp_row = nullptr; v.dataset_bf16 = nullptr;
  v.dataset_i8 = nullptr; v.i8_inv_mult = nullptr; v.i8_dp_norm = nullptr;
  // Reordering rows of this shard (f32, bf16 or int8): all rows when unsharded; the rows of the rank's residue class
  // (id sharding: a pitched copy of the strided host rows); or the rows of the datapoints stored in the rank's leaves
  // (leaf sharding: gathered through a staging buffer).  dp_row maps a datapoint id to its row on this rank.
  auto upload_rows = [&](const void* src, size_t row_bytes) -> int {
    const char* base = static_cast<const char*>(src);
    if (world == 1) {
      UP(ix->dataset, src, row_bytes * (size_t)N);
      return 0;
    }
    std::vector<uint32_t> rowmap(N, 0xFFFFFFFFu);
    size_t rows = 0;
    if (!by_leaf) {
      for (uint32_t i = rank; i < N; i += world) rowmap[i] = (uint32_t)rows++;
      CU(ix->dataset.ensure(std::max<size_t>(rows, 1) * row_bytes));
      if (rows)
        CU(cudaMemcpy2D(ix->dataset.p, row_bytes, base + (size_t)rank * row_bytes, row_bytes * world, row_bytes, rows,
                        cudaMemcpyHostToDevice));
    } else {
      for (uint32_t i = 0; i < N; ++i) {
        bool here = false;
        for (uint32_t c = 0; c < mult && !here; ++c) {
          const int32_t t = d->tokens[(size_t)i * mult + c];
          here = t >= 0 && mine(i, t);
        }
        if (here) rowmap[i] = (uint32_t)rows++;
      }
      CU(ix->dataset.ensure(std::max<size_t>(rows, 1) * row_bytes));
      const size_t stage_rows = std::max<size_t>(1, (64u << 20) / row_bytes);
      std::vector<char> stage(stage_rows * row_bytes);
      size_t filled = 0, done = 0;
      for (uint32_t i = 0; i < N; ++i) {
        if (rowmap[i] == 0xFFFFFFFFu) continue;
        memcpy(stage.data() + filled * row_bytes, base + (size_t)i * row_bytes, row_bytes);
        if (++filled == stage_rows) {
          CU(cudaMemcpy(static_cast<char*>(ix->dataset.p) + done * row_bytes, stage.data(), filled * row_bytes, cudaMemcpyHostToDevice));
          done += filled; filled = 0;
        }
      }
      if (filled) CU(cudaMemcpy(static_cast<char*>(ix->dataset.p) + done * row_bytes, stage.data(), filled * row_bytes, cudaMemcpyHostToDevice));
    }
    UP(ix->dp_row, rowmap.data(), sizeof(uint32_t) * (size_t)N);
    v.dp_row = ix->dp_row.as<uint32_t>();
    return 0;
  };
  if (!d->dataset && !d->bf16_dataset && d->int8_dataset) {
    // int8 reordering (exact_reordering { fixed_point { enabled: true } }; int8_dataset.npy + int8_multipliers.npy
    // + dp_norms.npy): a quarter of the reorder gather bytes and device memory of the f32 rows
    if (!d->int8_multipliers) return fail(SCANN_B200_INVALID_ARGUMENT, "int8_dataset needs int8_multipliers");
    if (d->distance == SCANN_B200_SQUARED_L2 && !d->dp_norms)
      return fail(SCANN_B200_INVALID_ARGUMENT, "int8 reordering under squared L2 needs dp_norms");
    if (int rc = upload_rows(d->int8_dataset, (size_t)D)) return rc;
    v.dataset_i8 = ix->dataset.as<int8_t>();
    std::vector<float> inv(D);
    for (uint32_t j = 0; j < D; ++j) inv[j] = 1.0f / d->int8_multipliers[j];  // reordering_helper.cc:407-412
    UP(ix->i8_inv, inv.data(), sizeof(float) * D);
    v.i8_inv_mult = ix->i8_inv.as<float>();
    if (d->dp_norms) {
      UP(ix->i8_norm, d->dp_norms, sizeof(float) * (size_t)N);
      v.i8_dp_norm = ix->i8_norm.as<float>();
    }
  }
  if (!d->dataset && d->bf16_dataset) {
    // bfloat16 reordering (exact_reordering { bfloat16 { enabled: true } }, bfloat16_dataset.npy): half the
    // reorder gather bytes and device memory of the f32 rows
    if (int rc = upload_rows(d->bf16_dataset, sizeof(uint16_t) * (size_t)D)) return rc;
    v.dataset_bf16 = ix->dataset.as<uint16_t>();
  }
  if (d->dataset) {
    if (int rc = upload_rows(d->dataset, sizeof(float) * (size_t)D)) return rc;
    v.dataset = ix->dataset.as<float>();
  }
#undef UP
  v.centers = ix->centers.as<float>();
  {
    // bf16 centre operand of the tensor-core tokenization (prep.cu) and the norm bound of its error term
    double cmax2 = 0.0;
    for (uint32_t l = 0; l < L; ++l) {
      double a = 0.0;
      for (uint32_t k = 0; k < D; ++k) a += (double)d->centers[(size_t)l * D + k] * (double)d->centers[(size_t)l * D + k];
      cmax2 = std::max(cmax2, a);
    }
    v.center_max_norm = (float)(std::sqrt(cmax2) * 1.0001);
    v.tok_kp = sb::tokenize_kpitch(D);
    CU(ix->tok_b.ensure(sb::tokenize_operand_bytes(L, D)));
    CU(sb::build_tokenize_operand(v.centers, L, D, 2, ix->tok_b.p, 0));
    CU(cudaStreamSynchronize(0));
    v.tok_b = ix->tok_b.p;
  }
  if (d->query_tokenization_type == SCANN_B200_TOKENIZE_FIXED_POINT_INT8) {
    // KMeansTreeNode::CreateFixedPointCenters (trees/kmeans_tree/kmeans_tree_node.cc:267-281), run when the tree is
    // loaded: ScalarQuantizeFloatDataset(float_centers, 1.0, NaN) (utils/scalar_quantization_helpers.cc:39-63,94-145)
    // -- multiplier[k] = 127 / max_l |c[l][k]| (1 for an all-zero column), Int8Quantize = clamp(std::round(c * m)) --
    // its inverse multipliers 1.0f / m, and float(SquaredL2Norm(float centre)) with DenseSingleAccumulate's four
    // strided double accumulators (utils/reduction.h:357-390)
    std::vector<float> mult(D, 0.0f), inv(D), sqn(L);
    std::vector<int8_t> ci8((size_t)L * D);
    for (uint32_t l = 0; l < L; ++l)
      for (uint32_t k = 0; k < D; ++k) mult[k] = std::max(mult[k], std::fabs(d->centers[(size_t)l * D + k]));
    const bool l2 = d->distance == SCANN_B200_SQUARED_L2;
    for (uint32_t k = 0; k < D; ++k) {
      mult[k] = mult[k] == 0.0f ? 1.0f : 127.0f / mult[k];
      inv[k] = 1.0f / mult[k];
      // GetAllDistancesInt8 scales the query by inv_mult, `inv_mult * 2` for squared L2 (kmeans_tree_node.h:239-244);
      // the doubling is exact, so the kernels read the ready-made scale
      if (l2) inv[k] = inv[k] * 2.0f;
    }
    for (uint32_t l = 0; l < L; ++l) {
      const float* c = d->centers + (size_t)l * D;
      for (uint32_t k = 0; k < D; ++k) {
        const float r = std::round(c[k] * mult[k]);
        ci8[(size_t)l * D + k] = (int8_t)(r > 127.0f ? 127.0f : (r < -128.0f ? -128.0f : r));
      }
      double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
      uint32_t k = 0;
      for (; k + 4 <= D; k += 4) {
        r0 += (double)c[k] * (double)c[k];
        r1 += (double)c[k + 1] * (double)c[k + 1];
        r2 += (double)c[k + 2] * (double)c[k + 2];
        r3 += (double)c[k + 3] * (double)c[k + 3];
      }
      r2 += r3;
      if (k + 2 <= D) {
        r0 += (double)c[k] * (double)c[k];
        r1 += (double)c[k + 1] * (double)c[k + 1];
        k += 2;
      }
      r1 += r2;
      if (k < D) r0 += (double)c[k] * (double)c[k];
      sqn[l] = (float)(r0 + r1);
    }
    CU(ix->cen_i8.ensure(ci8.size() ? ci8.size() : 16));
    CU(cudaMemcpy(ix->cen_i8.p, ci8.data(), ci8.size(), cudaMemcpyHostToDevice));
    CU(ix->cen_inv.ensure(sizeof(float) * D));
    CU(cudaMemcpy(ix->cen_inv.p, inv.data(), sizeof(float) * D, cudaMemcpyHostToDevice));
    CU(ix->cen_sqn.ensure(sizeof(float) * L));
    CU(cudaMemcpy(ix->cen_sqn.p, sqn.data(), sizeof(float) * L, cudaMemcpyHostToDevice));
    v.centers_i8 = ix->cen_i8.as<int8_t>();
    v.cen_qscale = ix->cen_inv.as<float>();
    v.cen_sqnorm = ix->cen_sqn.as<float>();
    {
      std::vector<float> sqn2(L);
      for (uint32_t l = 0; l < L; ++l) sqn2[l] = 2.0f * sqn[l];
      CU(ix->cen_sqn2.ensure(sizeof(float) * L));
      CU(cudaMemcpy(ix->cen_sqn2.p, sqn2.data(), sizeof(float) * L, cudaMemcpyHostToDevice));
      v.cen_sqnorm2 = ix->cen_sqn2.as<float>();
    }
    // tensor-core pre-filter of the int8 tokenization (prep.cu): an int8 value is exact in bf16, so the centre operand
    // is [c | c | 0] and the only dropped term of the split GEMM is the query's third bf16 term
    {
      std::vector<float> cf(ci8.size());
      double m2 = 0.0, s2 = 0.0;
      for (uint32_t l = 0; l < L; ++l) {
        double a = 0.0;
        for (uint32_t k = 0; k < D; ++k) {
          const float f = (float)ci8[(size_t)l * D + k];
          cf[(size_t)l * D + k] = f;
          a += (double)f * (double)f;
        }
        m2 = std::max(m2, a);
        s2 = std::max(s2, (double)sqn[l]);
      }
      v.cen_i8_max_norm = (float)(std::sqrt(m2) * 1.0001);
      v.cen_sqnorm_max = (float)(s2 * 1.0001);
      sbi::DevBuf tmp;
      CU(tmp.ensure(sizeof(float) * std::max<size_t>(cf.size(), 4)));
      CU(cudaMemcpy(tmp.p, cf.data(), sizeof(float) * cf.size(), cudaMemcpyHostToDevice));
      CU(sb::build_tokenize_operand(tmp.as<float>(), L, D, 2, ix->tok_b.p, 0));
      CU(cudaStreamSynchronize(0));
    }
  }
  v.centers_t = nullptr;
  v.center_sqnorm = ix->cnorm.as<float>();
  v.codebook = ix->codebook.as<float>();
  v.block_dims = ix->block_dims.as<int32_t>();
  v.block_off = ix->block_off.as<uint32_t>();
  v.leaf_size = ix->leaf_size.as<uint32_t>();
  v.leaf_goff = ix->leaf_goff.as<uint32_t>();
  v.leaf_ntiles = ix->leaf_ntiles.as<uint32_t>();
  v.leaf_gpt = ix->leaf_gpt.as<uint32_t>();
  v.codes = ix->codes.as<uint32_t>();
  v.slot_dp = ix->slot_dp.as<uint32_t>();
  return 0;
}

}  // namespace

namespace sbi {

int resolve(const scann_b200_index* ix, int final_nn, int pre_nn, int leaves, Params* p) {
  // scann_ops/cc/scann.cc:406-430 + SetUnspecifiedParametersToDefaults
  if (ix->brute) {
    const int kb = final_nn > 0 ? final_nn : ix->desc.default_final_nn;
    if (kb <= 0) return fail(SCANN_B200_INVALID_ARGUMENT, "final_num_neighbors must be positive");
    p->k = (uint32_t)kb; p->npre = p->k; p->nover = p->k; p->P = 1;
    return 0;
  }
  const bool has_reorder = ix->dev.dataset != nullptr || ix->dev.dataset_bf16 != nullptr || ix->dev.dataset_i8 != nullptr;
  const int k = final_nn > 0 ? final_nn : ix->desc.default_final_nn;
  int npre = has_reorder ? (pre_nn > 0 ? pre_nn : ix->desc.default_pre_nn) : k;
  int P = leaves > 0 ? leaves : ix->desc.default_leaves;
  if (k <= 0 || npre <= 0 || P <= 0) return fail(SCANN_B200_INVALID_ARGUMENT, "search parameters must be positive (k=%d pre=%d leaves=%d)", k, npre, P);
  if ((uint32_t)P > ix->dev.L) P = (int)ix->dev.L;
  // tree_ah_hybrid_residual.h:263-267 + internal/utils.h:146-157
  long long nover = npre;
  if (!ix->dev.disjoint) {
    const double r = (double)npre * (double)ix->desc.overretrieve;
    nover = r > 2147483647.0 ? 2147483647LL : (long long)(int)r;
  }
  if (nover > 8192 || npre > 8192) return fail(SCANN_B200_UNIMPLEMENTED, "pre-reorder neighbours %lld > 8192 not supported", nover);
  if (nover < 1) nover = 1;
  if (P > 4096) return fail(SCANN_B200_UNIMPLEMENTED, "leaves_to_search %d > 4096 not supported", P);
  p->k = (uint32_t)k; p->npre = (uint32_t)npre; p->nover = (uint32_t)nover; p->P = (uint32_t)P;
  return 0;
}

uint32_t pick_cap(uint32_t nover) {
  // Candidate buffer entries per query.  Big enough that the heavy tail of the candidate inflow
  // (C2: mean 290, max 13k per query) never needs a re-scan; SCANN_B200_CAND_CAP shrinks it so
  // that tests can exercise the overflow path.
  uint32_t cap = 16384;
  if (const char* e = getenv("SCANN_B200_CAND_CAP")) {
    const long v = strtol(e, nullptr, 10);
    if (v >= 64 && v <= (1 << 20)) { cap = 64; while (cap < (uint32_t)v) cap <<= 1; }
  }
  while (cap < 2 * nover) cap <<= 1;  // nover <= 8192 => cap <= 16384 keys = 128 KB of shared memory in compact_big_kernel
  return cap;
}

int ensure_workspace(scann_b200_index* ix, uint32_t nq, const Params& p, uint32_t out_k, uint32_t cap) {
  const sb::DevIndex& v = ix->dev;
  CU(ix->q.ensure(sizeof(float) * (size_t)nq * v.d));
  CU(ix->dist.ensure(sizeof(float) * (size_t)nq * v.L));
  CU(ix->tok_a.ensure(sb::tokenize_operand_bytes(nq, v.d)));
  CU(ix->tok_cmax.ensure(sizeof(float) * (size_t)nq * ((v.L + 31) / 32)));
  ix->dev.tok_cmax_ws = ix->tok_cmax.as<float>();
  CU(ix->leaves.ensure(sizeof(int32_t) * (size_t)nq * p.P));
  CU(ix->bias.ensure(sizeof(float) * (size_t)nq * p.P));
  CU(ix->lut.ensure((size_t)nq * v.W * 128));
  if (sb::scan_tc_supported(v)) CU(ix->lut_e4m3.ensure((size_t)nq * v.W * 128 * 2));
  CU(ix->mult.ensure(sizeof(float) * nq));
  CU(ix->inv.ensure(sizeof(float) * nq));
  CU(ix->pilot_end.ensure(sizeof(int32_t) * nq));
  CU(ix->buf.ensure(sizeof(uint64_t) * (size_t)nq * cap));
  CU(ix->cnt.ensure(sizeof(uint32_t) * nq));
  CU(ix->tau.ensure(sizeof(uint64_t) * nq));
  CU(ix->ovf.ensure(sizeof(uint32_t) * nq));
  CU(ix->leaf_cnt.ensure(sizeof(uint32_t) * (v.L + 1)));
  CU(ix->leaf_eoff.ensure(sizeof(uint32_t) * (v.L + 1)));
  CU(ix->leaf_cur.ensure(sizeof(uint32_t) * (v.L + 1)));
  CU(ix->item_off.ensure(sizeof(uint32_t) * (v.L + 1)));
  CU(ix->item_leaf.ensure(sizeof(uint32_t) * std::min<size_t>((size_t)nq * p.P * 4 + v.L + 1024, (size_t)8 << 20)));
  CU(ix->entry_q.ensure(sizeof(uint32_t) * (size_t)nq * p.P));
  CU(ix->pair_pos.ensure(sizeof(uint32_t) * (size_t)nq * p.P));
  CU(ix->entry_bias.ensure(sizeof(float) * (size_t)nq * p.P));
  CU(ix->counters.ensure(sizeof(uint32_t) * 8));
  CU(ix->stats.ensure(sizeof(unsigned long long) * 4));
  CU(ix->out_idx.ensure(sizeof(uint32_t) * (size_t)nq * out_k));
  CU(ix->out_dist.ensure(sizeof(float) * (size_t)nq * out_k));
  CU(ix->h_counters.ensure(64));
  return 0;
}

void fill_scan_work(scann_b200_index* ix, uint32_t nq, const Params& p, uint32_t cap, sb::ScanWork* wp) {
  sb::ScanWork& w = *wp;
  w = sb::ScanWork{};
  w.leaves = ix->leaves.as<int32_t>(); w.bias = ix->bias.as<float>(); w.lut = ix->lut.as<uint8_t>();
  w.mult = ix->mult.as<float>(); w.inv_mult = ix->inv.as<float>(); w.pilot_end = ix->pilot_end.as<int32_t>();
  w.buf = ix->buf.as<uint64_t>(); w.cnt = ix->cnt.as<uint32_t>(); w.tau = ix->tau.as<uint64_t>();
  w.ovf = ix->ovf.as<uint32_t>(); w.leaf_cnt = ix->leaf_cnt.as<uint32_t>(); w.leaf_eoff = ix->leaf_eoff.as<uint32_t>();
  w.leaf_cur = ix->leaf_cur.as<uint32_t>(); w.item_off = ix->item_off.as<uint32_t>();
  w.entry_q = ix->entry_q.as<uint32_t>(); w.entry_bias = ix->entry_bias.as<float>(); w.pair_pos = ix->pair_pos.as<uint32_t>();
  w.counters = ix->counters.as<uint32_t>(); w.stats = ix->stats.as<unsigned long long>();
  w.nq = nq; w.P = p.P; w.cap = cap; w.nover = p.nover; w.quads_per_item = 2; w.one = 1;  // 2 octs = 16 queries per work item
  w.item_leaf = ix->item_leaf.as<uint32_t>();
  w.lut_e4m3 = ix->lut_e4m3.as<uint8_t>();
  w.item_leaf_cap = (uint32_t)(ix->item_leaf.bytes / sizeof(uint32_t));
  // Groups per work item.  An item costs a fixed set-up (oct tables, thresholds, a few dependent loads); with many more
  // items than resident CTAs (740) larger tiles amortise it, with few items small tiles balance the load.  Expected
  // items at g groups per tile: (leaves hit) x ceil(queries per leaf / 16) x ceil(groups per leaf / g).
  {
    const double leaves = std::max<double>(1.0, ix->nonempty_leaves);
    const double qpl = (double)nq * p.P * (ix->shard_mode == SCANN_B200_SHARD_BY_LEAF ? 1.0 / ix->shard_world : 1.0) / leaves;
    w.qpl_per_rank = (float)(qpl / std::max<uint32_t>(1u, p.P));
    const double hit = leaves * (1.0 - exp(-qpl));
    const double chunks = std::max(1.0, ceil(qpl / 16.0));
    const double groups = std::max(1.0, ix->avg_leaf_slots / 32.0);
    uint32_t g = (uint32_t)sb::kMaxGroupsPerTile;
    for (uint32_t cand = 256; cand > (uint32_t)sb::kMaxGroupsPerTile; cand >>= 1)
      if (hit * chunks * ceil(groups / cand) >= 16.0 * 740.0) { g = cand; break; }
    w.max_gpt = g;
    if (const char* e = getenv("SCANN_B200_SCAN_GPT")) {
      const int t = atoi(e);
      if (t >= 1 && t <= 4096) w.max_gpt = (uint32_t)t;
    }
  }
  // large leaves (C5 shape: ~20 candidates per (query, item)): stage candidates in shared memory, one global atomic
  // per (query, item); small leaves (C2: ~1.4) append directly
  w.stage = ix->avg_leaf_slots >= 2048 ? 1u : 0u;
  if (const char* e = getenv("SCANN_B200_SCAN_STAGE")) w.stage = e[0] == '1' ? 1u : 0u;
  // pilot sample: 4 N' slots; whole leaves for large leaves (a looser tau costs pushes there), for small leaves the
  // pilot stops inside the leaf once the target is reached (SCANN_B200_PILOT="<factor>[p]" overrides, tests / tuning)
  w.pilot_target = 4 * p.nover;
  w.pilot_partial = 0;
  if (const char* e = getenv("SCANN_B200_PILOT")) {
    const int f = atoi(e);
    if (f > 0) w.pilot_target = (uint32_t)f * p.nover;
    w.pilot_partial = strchr(e, 'p') ? 1u : 0u;
  }
  // pilot buffer: the default (1024 keys or 2 N' + 128) selects every ~900 buffered keys, which also tightens the
  // pilot's own threshold early; buffering a whole large leaf and selecting once was measured and rejected
  // (20M x 96: pilot 0.72 -> 1.47 ms).  SCANN_B200_PILOT_CAP overrides (tests).
  w.pilot_cap = 0;
  if (const char* e = getenv("SCANN_B200_PILOT_CAP")) w.pilot_cap = (uint32_t)atoi(e);
}

}  // namespace sbi

namespace {

struct PartialOut { uint32_t* ids; uint64_t* tie; float* ah; float* exact; uint32_t cap; };

// One chunk of queries, everything on the device.  d_q [nq][D]; outputs may be null when
// only partial (sharded) records are wanted.
int search_chunk(scann_b200_index* ix, const float* d_q, uint32_t nq, const Params& p,
                 uint32_t* d_out_idx, float* d_out_dist, uint32_t out_k, const PartialOut* part,
                 bool stop_after_candidates) {
  const sb::DevIndex& v = ix->dev;
  cudaStream_t s = ix->stream;
  const uint32_t cap = pick_cap(p.nover);
  if (int rc = ensure_workspace(ix, nq, p, out_k ? out_k : 1, cap)) return rc;
  sb::ScanWork w{};
  fill_scan_work(ix, nq, p, cap, &w);
  int launches = 0;
  uint32_t scan_launches = 0, retries = 0;
  uint32_t mode_launches[3] = {0, 0, 0};

  CU(cudaMemsetAsync(w.counters, 0, sizeof(uint32_t) * 8, s));
  CU(cudaMemsetAsync(w.stats, 0, sizeof(unsigned long long) * 4, s));
  CU(cudaEventRecord(ix->ev[EV_START], s));
  CU(sb::launch_tokenize_topp(v, d_q, nq, p.P, ix->dist.as<float>(), ix->tok_a.p, ix->leaves.as<int32_t>(),
                              ix->bias.as<float>(), w.counters + 5, s, &launches));
  CU(cudaEventRecord(ix->ev[EV_TOK], s));
  // LUT build: fused into the pilot kernel when the raw table fits its candidate buffer (every BASELINE.json
  // configuration; SCANN_B200_FUSE_LUT=0 keeps the separate lut_kernel)
  bool fuse_lut = sb::pilot_can_build_lut(v, w);
  if (const char* e = getenv("SCANN_B200_FUSE_LUT")) fuse_lut = fuse_lut && e[0] != '0';
  w.q_for_lut = fuse_lut ? d_q : nullptr;
  if (!fuse_lut) {
    sb::launch_lut(v, d_q, nq, ix->lut.as<uint8_t>(), ix->mult.as<float>(), ix->inv.as<float>(), s);
    launches += 1;
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(ix->ev[EV_LUT], s));
  // Two scan phases when many slots are probed per query (C5-size leaves): the nearest eighth of the
  // leaves first, a compaction that tightens tau from "N-th best of the pilot's sample" to "N-th best of
  // those leaves", then the rest.  The pushes of the second phase shrink several-fold (100M x 96, P = 80:
  // 10.3k -> candidates per query); for C2-size work the extra launches cost more than they save.
  bool two_phase = p.P >= 16 && (uint64_t)p.P * ix->avg_leaf_slots >= 98304;
  if (const char* e = getenv("SCANN_B200_TWO_PHASE")) two_phase = e[0] == '1' && p.P >= 2;
  uint32_t r1 = two_phase ? std::max<uint32_t>(1, p.P / 8) : p.P;
  if (const char* e = getenv("SCANN_B200_PHASE1_RANKS")) { const int t = atoi(e); if (two_phase && t >= 1 && (uint32_t)t < p.P) r1 = (uint32_t)t; }
  w.rank_lo = 0; w.rank_hi = r1;
  sb::scan_prepare_phase(v, &w);
  CU(sb::launch_pilot(v, w, s));
  launches += 1;
  CU(cudaEventRecord(ix->ev[EV_PILOT], s));
  sb::launch_worklist(v, w, false, true, s, &launches);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ix->ev[EV_WORK], s));
  CU(sb::launch_scan(v, w, 0, s));
  launches += 1; scan_launches += 1; mode_launches[w.scan_mode < 3 ? w.scan_mode : 0] += 1;
  CU(cudaEventRecord(ix->ev[EV_SCAN], s));
  int ncl = 0;
  CU(sb::launch_compact(v, w, false, s, &ncl));
  launches += ncl;
  CU(cudaEventRecord(ix->ev[EV_COMPACT], s));
  if (two_phase) {
    w.rank_lo = r1; w.rank_hi = p.P;
    sb::scan_prepare_phase(v, &w);
    sb::launch_worklist(v, w, false, false, s, &launches);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ix->ev[EV2_WORK], s));
    CU(sb::launch_scan(v, w, 0, s));
    launches += 1; scan_launches += 1; mode_launches[w.scan_mode < 3 ? w.scan_mode : 0] += 1;
    CU(cudaEventRecord(ix->ev[EV2_SCAN], s));
    CU(sb::launch_compact(v, w, false, s, &ncl));
    launches += ncl;
    CU(cudaEventRecord(ix->ev[EV2_COMPACT], s));
  }
  w.rank_lo = 0; w.rank_hi = p.P;  // re-scans of overflowed queries cover every probed leaf
  w.rescan = 1;
  sb::scan_prepare_phase(v, &w);
  // Overflow flag, statistics and the finalize kernel share ONE host round trip: finalize is launched
  // optimistically; if some buffer overflowed (rare) its output is discarded, the flagged queries are
  // re-scanned and finalize runs again.
  uint32_t* hc = ix->h_counters.as<uint32_t>();                        // pinned: [0..7] counters, then 4 x u64 stats
  unsigned long long* hs = reinterpret_cast<unsigned long long*>(hc + 8);
  auto finalize = [&]() -> int {
    if (stop_after_candidates) return 0;
    sb::FinalizeArgs a{};
    a.q = d_q; a.nq = nq; a.npre = p.npre; a.k = p.k; a.out_k = out_k;
    a.out_idx = d_out_idx; a.out_dist = d_out_dist;
    if (part) { a.part_ids = part->ids; a.part_tie = part->tie; a.part_ah = part->ah; a.part_exact = part->exact; a.part_cap = part->cap; }
    CU(sb::launch_finalize(v, w, a, s));
    launches += 1;
    return 0;
  };
  if (int rc = finalize()) return rc;
  CU(cudaEventRecord(ix->ev[EV_FIN], s));
  CU(cudaMemcpyAsync(hc, w.counters, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(hs, w.stats, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  const uint32_t tok_fallbacks = hc[5];
  if (hc[7] != 0) return fail(SCANN_B200_INTERNAL, "tensor-core scan: a pipeline barrier timed out (watchdog)");
  if (hc[2] != 0) {
    while (hc[2] != 0) {
      if (++retries > 256) return fail(SCANN_B200_INTERNAL, "candidate buffer overflow did not converge");
      CU(cudaMemsetAsync(w.counters + 2, 0, sizeof(uint32_t), s));
      sb::launch_worklist(v, w, true, false, s, &launches);
      CU(sb::launch_scan(v, w, 0, s));
      CU(sb::launch_compact(v, w, true, s, &ncl));
      launches += 1 + ncl; scan_launches += 1; mode_launches[w.scan_mode < 3 ? w.scan_mode : 0] += 1;
      CU(cudaMemcpyAsync(hc, w.counters, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
    }
    if (int rc = finalize()) return rc;
    CU(cudaEventRecord(ix->ev[EV_FIN], s));
    CU(cudaStreamSynchronize(s));
  }
  float ms[EV_COUNT] = {0};
  for (int i = 1; i < EV_COUNT; ++i) {
    const int prev = (i == EV_FIN && two_phase) ? EV2_COMPACT : i - 1;
    CU(cudaEventElapsedTime(&ms[i], ix->ev[prev], ix->ev[i]));
  }
  if (two_phase) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, ix->ev[EV_COMPACT], ix->ev[EV2_WORK])); ms[EV_WORK] += t;
    CU(cudaEventElapsedTime(&t, ix->ev[EV2_WORK], ix->ev[EV2_SCAN])); ms[EV_SCAN] += t;
    CU(cudaEventElapsedTime(&t, ix->ev[EV2_SCAN], ix->ev[EV2_COMPACT])); ms[EV_COMPACT] += t;
  }
  scann_b200_stats& st = ix->last;
  st.scan_bytes_alg += hs[0];
  st.scan_pairs += hs[1];
  st.scan_lookups += hs[0] * 2;
  st.cand_sum += hs[2];
  st.cand_max = std::max<uint64_t>(st.cand_max, hs[3]);
  st.tokenize_fallbacks += tok_fallbacks;
  st.kernel_launches += (uint32_t)launches;
  st.overflow_retries += retries;
  st.scan_kernel_count += scan_launches;
  st.scan_oct_launches += mode_launches[0]; st.scan_wide_launches += mode_launches[1]; st.scan_tc_launches += mode_launches[2];
  st.ms_tokenize += ms[EV_TOK]; st.ms_lut += ms[EV_LUT]; st.ms_pilot += ms[EV_PILOT];
  st.ms_worklist += ms[EV_WORK]; st.ms_scan += ms[EV_SCAN]; st.ms_compact += ms[EV_COMPACT];
  st.ms_finalize += ms[EV_FIN];
  float tot = 0;
  CU(cudaEventElapsedTime(&tot, ix->ev[EV_START], ix->ev[EV_FIN]));
  st.ms_total += tot;
  return 0;
}

// Brute force (bf16): geometric rounds of [tcgen05 GEMM + threshold filter] -> compaction, then an
// exact re-scoring of the k' best.  Mirrors Bfloat16BruteForceSearcher::FindNeighborsImpl
// (brute_force/bfloat16_brute_force.cc:101-152) for a whole batch of queries.
//
// The result is the exact top-k by (exact distance, id) on every input:
//  * the pre-filter keeps the k' = 2k + 64 smallest approximate keys; the re-scoring kernel proves per query that the
//    window was wide enough (check_window, bruteforce.cu) and the batch is re-run with a 4x wider window while any query
//    fails the proof; at k' = 8192 the remaining queries (thousands of near-ties at the k-th score) are finished by the
//    exact all-rows kernel;
//  * the geometric rounds assume rows arrive in no particular order; on an ordered database (rows sorted by score for
//    some query) a round can push more keys than a candidate buffer holds.  The overflow is detected and the batch is
//    re-run with rounds of at most cap - k' rows, which cannot overflow by construction.
int search_bf_chunk(scann_b200_index* ix, const float* d_q, uint32_t nq, uint32_t k, uint32_t* d_out_idx,
                    float* d_out_dist, uint32_t out_k) {
  const sb::DevIndex& v = ix->dev;
  cudaStream_t s = ix->stream;
  if (2 * k + 64 > 8192) return fail(SCANN_B200_UNIMPLEMENTED, "brute force with k=%u > 4064 is not supported", k);
  CU(ix->bf_a.ensure(ix->bf_f32 ? (size_t)sb::bf_query_rows_pad(nq) * ix->bf_dpitch * 2 + 256 * (size_t)ix->bf_dpitch * 2
                                : sb::bf_query_operand_bytes(nq, ix->bf_dpitch)));
  CU(ix->cnt.ensure(sizeof(uint32_t) * nq));
  CU(ix->tau.ensure(sizeof(uint64_t) * nq));
  CU(ix->ovf.ensure(sizeof(uint32_t) * nq));
  CU(ix->entry_q.ensure(sizeof(uint32_t) * (size_t)nq));
  CU(ix->bf_flags.ensure(sizeof(uint32_t) * (2 * (size_t)nq + 4)));  // [nq] unsafe flags, [nq] flagged list, counters
  CU(ix->counters.ensure(sizeof(uint32_t) * 8));
  CU(ix->stats.ensure(sizeof(unsigned long long) * 4));
  CU(ix->h_counters.ensure(64));
  uint32_t* d_unsafe = ix->bf_flags.as<uint32_t>();
  uint32_t* d_flagged = d_unsafe + nq;
  uint32_t* d_nflag = d_flagged + nq;   // [0] queries that failed the window proof, [1] length of the flagged list
  uint32_t* hc = ix->h_counters.as<uint32_t>();
  uint32_t launches = 0, gemm_launches = 0, widenings = 0, exact_fallbacks = 0;
  const void* db = ix->bf_db.p;
  CU(cudaEventRecord(ix->ev[EV_START], s));
  if (ix->bf_f32) {
    // the GEMM walks bf_query_rows_pad(nq) rows; the split kernel writes ceil(nq / 128) * 128 of them
    const size_t written = (size_t)((nq + 127) / 128 * 128) * ix->bf_dpitch * 2;
    const size_t walked = (size_t)sb::bf_query_rows_pad(nq) * ix->bf_dpitch * 2;
    if (walked > written) CU(cudaMemsetAsync(static_cast<char*>(ix->bf_a.p) + written, 0, walked - written, s));
    if (ix->bf_l2) {  // q' = [q, 1]
      CU(ix->bf_qaug.ensure(sizeof(float) * (size_t)nq * (v.d + 1)));
      CU(sb::bf_augment_rows(d_q, nullptr, 1.0f, nq, v.d, ix->bf_qaug.as<float>(), s));
      CU(sb::build_tokenize_operand(ix->bf_qaug.as<float>(), nq, v.d + 1, 1, ix->bf_a.p, s));
      launches += 1;
    } else {
      CU(sb::build_tokenize_operand(d_q, nq, v.d, 1, ix->bf_a.p, s));
    }
  }
  else CU(sb::bf_split_queries(d_q, nq, v.d, ix->bf_dpitch, ix->bf_a.p, s));
  launches += 1;
  CU(cudaEventRecord(ix->ev[EV_TOK], s));
  // |approximate - exact| <= eps_rel * ||q|| * max ||x||: K * 2^-21 + 2^-15 with K the accumulated products per score
  const uint32_t kacc = ix->bf_f32 ? ix->bf_dpitch : 2 * ix->bf_dpitch;
  const float eps_rel = (float)kacc * 4.76837158e-7f + 3.05175781e-5f;
  uint32_t kprime = 2 * k + 64;
  if (const char* e = getenv("SCANN_B200_BF_KPRIME")) {  // tests: start from a narrow window
    const long t = strtol(e, nullptr, 10);
    if (t >= (long)k && t <= 8192) kprime = (uint32_t)t;
  }
  bool safe_rounds = false;
  sb::ScanWork w{};
  for (;;) {
    const uint32_t cap = pick_cap(kprime);
    CU(ix->buf.ensure(sizeof(uint64_t) * (size_t)nq * cap));
    w = sb::ScanWork{};
    w.buf = ix->buf.as<uint64_t>(); w.cnt = ix->cnt.as<uint32_t>(); w.tau = ix->tau.as<uint64_t>();
    w.ovf = ix->ovf.as<uint32_t>(); w.entry_q = ix->entry_q.as<uint32_t>();
    w.counters = ix->counters.as<uint32_t>(); w.stats = ix->stats.as<unsigned long long>();
    w.nq = nq; w.cap = cap; w.nover = kprime;
    CU(cudaMemsetAsync(w.counters, 0, sizeof(uint32_t) * 8, s));
    CU(cudaMemsetAsync(w.stats, 0, sizeof(unsigned long long) * 4, s));
    CU(cudaMemsetAsync(d_nflag, 0, sizeof(uint32_t) * 2, s));
    CU(sb::bf_init_state(nq, w.cnt, w.tau, w.ovf, s));
    launches += 1;
    // rounds: the first one must hold >= k' rows (everything passes an infinite threshold) and fit
    // the candidate buffers; afterwards the inflow per round is ~ k' * growth
    // (every row seen so far has a 1-in-rows chance per kept slot, so a round over g x the rows seen
    // pushes ~ g * k' keys).  Rounds are sized so that k' * (1 + g) stays within the 1024 keys the
    // one-CTA-per-query compaction sorts in shared memory.  Safe rounds: at most cap - k' rows each.
    uint32_t row0 = 0;
    uint32_t chunk = std::max<uint32_t>(1024, (kprime + 255) / 256 * 256);
    chunk = std::min(chunk, cap / 2);
    const uint32_t safe_chunk = std::max<uint32_t>(256, (cap - kprime) / 256 * 256);
    if (safe_rounds) chunk = safe_chunk;
    const uint32_t growth = 3 * kprime <= 1000 ? 2 : (2 * kprime <= 1000 ? 1 : 2);
    while (row0 < v.n) {
      const uint32_t row1 = (uint64_t)row0 + chunk >= v.n ? v.n : row0 + chunk;
      CU(sb::bf_gemm_round(ix->bf_a.p, db, nq, v.n, ix->bf_dpitch, row0, row1, w, ix->bf_f32 ? 1 : 2, s));
      int ncl = 0;
      CU(sb::launch_compact(v, w, false, s, &ncl));
      launches += 1 + (uint32_t)ncl;
      gemm_launches += 1;
      row0 = row1;
      chunk = safe_rounds ? safe_chunk : (uint32_t)std::min<uint64_t>((uint64_t)row1 * growth, 0x40000000ull);
    }
    CU(cudaEventRecord(ix->ev[EV_SCAN], s));
    // optimistic: the re-scoring is launched before the overflow counter is known (one host round trip)
    sb::BfSafety safety{eps_rel, ix->bf_max_row_norm, d_unsafe, d_nflag};
    if (ix->bf_f32)
      CU(sb::bf_rescore_f32(d_q, v.dataset, nq, v.d, w, kprime, k, out_k, ix->bf_row0, d_out_idx, d_out_dist, s, &safety,
                            ix->bf_l2 ? ix->bf_xnorm.as<float>() : nullptr));
    else
      CU(sb::bf_rescore(d_q, db, nq, v.d, ix->bf_dpitch, w, kprime, k, out_k, ix->bf_row0, d_out_idx, d_out_dist, s, &safety));
    launches += 1;
    CU(cudaEventRecord(ix->ev[EV_FIN], s));
    CU(cudaMemcpyAsync(hc, w.counters, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hc + 8, d_nflag, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (hc[2] != 0) {
      if (safe_rounds) return fail(SCANN_B200_INTERNAL, "brute force: candidate buffer overflow in safe rounds (%u queries)", hc[2]);
      safe_rounds = true;
      ++widenings;
      continue;
    }
    if (hc[8] == 0) break;
    if (kprime < 8192) {
      kprime = std::min<uint32_t>(8192, kprime * 4);
      ++widenings;
      continue;
    }
    // exact all-rows fallback for the flagged queries: keys are exact, keep k
    exact_fallbacks = hc[8];
    w.nover = k;
    CU(sb::bf_exact_prepare(nq, d_unsafe, w, d_flagged, d_nflag + 1, s));
    const uint32_t rows_per_round = std::max<uint32_t>(256, cap - k);
    const void* rows = ix->bf_f32 ? static_cast<const void*>(v.dataset) : db;
    for (uint32_t r0 = 0; r0 < v.n; r0 += rows_per_round) {
      const uint32_t r1 = (uint32_t)std::min<uint64_t>((uint64_t)r0 + rows_per_round, v.n);
      CU(sb::bf_exact_round(d_q, rows, ix->bf_f32, v.d, ix->bf_dpitch, d_flagged, exact_fallbacks, r0, r1, w, s,
                            ix->bf_l2 ? ix->bf_xnorm.as<float>() : nullptr));
      int ncl = 0;
      CU(sb::launch_compact(v, w, false, s, &ncl));
      launches += 1 + (uint32_t)ncl;
    }
    CU(sb::bf_exact_emit(d_flagged, exact_fallbacks, w, k, out_k, ix->bf_row0, d_out_idx, d_out_dist, s, ix->bf_l2));
    launches += 2;
    CU(cudaEventRecord(ix->ev[EV_FIN], s));
    CU(cudaStreamSynchronize(s));
    break;
  }
  float ms_prep = 0, ms_gemm = 0, ms_fin = 0, tot = 0;
  CU(cudaEventElapsedTime(&ms_prep, ix->ev[EV_START], ix->ev[EV_TOK]));
  CU(cudaEventElapsedTime(&ms_fin, ix->ev[EV_SCAN], ix->ev[EV_FIN]));
  CU(cudaEventElapsedTime(&tot, ix->ev[EV_START], ix->ev[EV_FIN]));
  ms_gemm = tot - ms_prep - ms_fin;  // all GEMM rounds, including those of a re-run
  scann_b200_stats& st = ix->last;
  st.kernel_launches += launches;
  st.scan_kernel_count += gemm_launches;
  st.ms_tokenize += ms_prep; st.ms_scan += ms_gemm; st.ms_finalize += ms_fin; st.ms_total += tot;
  st.scan_pairs += (uint64_t)nq * v.n;
  st.scan_bytes_alg += (uint64_t)v.n * ix->bf_dpitch * 2;  // compulsory database bytes of one pass
  st.bf_widenings += widenings;
  st.bf_exact_fallbacks += exact_fallbacks;
  return 0;
}

int run_chunk(scann_b200_index* ix, const float* d_q, uint32_t nq, const Params& p, uint32_t* d_out_idx,
              float* d_out_dist, uint32_t out_k) {
  if (ix->brute) return search_bf_chunk(ix, d_q, nq, p.k, d_out_idx, d_out_dist, out_k);
  return search_chunk(ix, d_q, nq, p, d_out_idx, d_out_dist, out_k, nullptr, false);
}

int check_query_args(scann_b200_index* ix, const void* q, uint32_t nq) {
  if (!ix) return fail(SCANN_B200_INVALID_ARGUMENT, "null index");
  if (nq && !q) return fail(SCANN_B200_INVALID_ARGUMENT, "null queries");
  return 0;
}

}  // namespace

namespace sbi {

static scann_b200_index* make_lane(scann_b200_index* root) {
  scann_b200_index* c = new scann_b200_index();
  c->dev = root->dev; c->desc = root->desc; c->device = root->device; c->sm_count = root->sm_count;
  c->h_leaf_size = root->h_leaf_size; c->max_chunk = root->max_chunk;
  c->brute = root->brute; c->bf_f32 = root->bf_f32; c->bf_dpitch = root->bf_dpitch; c->bf_row0 = root->bf_row0;
  c->avg_leaf_slots = root->avg_leaf_slots; c->nonempty_leaves = root->nonempty_leaves;
  c->bf_max_row_norm = root->bf_max_row_norm;
  c->bf_l2 = root->bf_l2;
  c->shard_rank = root->shard_rank; c->shard_world = root->shard_world; c->shard_mode = root->shard_mode;
  c->parent = root;
  // the index arrays a search path reaches through DevBuf members rather than through `dev`
  auto alias = [](DevBuf& dst, const DevBuf& src) { dst.p = src.p; dst.bytes = src.bytes; dst.own = false; };
  alias(c->bf_xnorm, root->bf_xnorm);
  alias(c->bf_db, root->bf_db); alias(c->dataset, root->dataset); alias(c->leaf_goff, root->leaf_goff);
  alias(c->slot_dp, root->slot_dp);
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < EV_ALL && ok; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  if (!ok) {
    fail(SCANN_B200_INTERNAL, "could not create a search lane (stream / events)");
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int i = 0; i < EV_ALL; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
    return nullptr;
  }
  return c;
}

LaneGuard::LaneGuard(scann_b200_index* root) : root_(root) {
  std::unique_lock<std::mutex> lk(root->pool_mu);
  if (root->lane_busy.empty()) root->lane_busy.assign(1, 0);
  for (;;) {
    for (size_t i = 0; i < root->lane_busy.size(); ++i)
      if (!root->lane_busy[i]) { slot_ = (int)i; break; }
    if (slot_ >= 0) break;
    if ((int)root->lane_busy.size() < scann_b200_index::kMaxLanes) {
      cudaSetDevice(root->device);
      scann_b200_index* c = make_lane(root);
      if (!c) return;
      root->lanes.push_back(c);
      root->lane_busy.push_back(0);
      slot_ = (int)root->lane_busy.size() - 1;
      break;
    }
    root->pool_cv.wait(lk);
  }
  root->lane_busy[slot_] = 1;
  ix = slot_ == 0 ? root : root->lanes[slot_ - 1];
  lk.unlock();
  ix->mu.lock();
}

LaneGuard::~LaneGuard() {
  if (!ix) return;
  const scann_b200_stats st = ix->last;
  ix->mu.unlock();
  {
    std::lock_guard<std::mutex> lk(root_->pool_mu);
    root_->lane_busy[slot_] = 0;
    root_->last_any = st;  // scann_b200_last_stats: the most recently finished call
    root_->last_any_valid = true;
  }
  root_->pool_cv.notify_one();
}

}  // namespace sbi

extern "C" {

const char* scann_b200_last_error(void) { return g_err.c_str(); }
int scann_b200_abi_version(void) { return SCANN_B200_ABI_VERSION; }

int scann_b200_index_create(const scann_b200_index_desc* desc, scann_b200_index** out) {
  if (!desc || !out) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SCANN_B200_FAILED_PRECONDITION, "no CUDA device available (%s); scann_b200 has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (desc->device < 0 || desc->device >= ndev) return fail(SCANN_B200_INVALID_ARGUMENT, "device %d out of range", desc->device);
  CU(cudaSetDevice(desc->device));
  scann_b200_index* ix = new scann_b200_index();
  ix->desc = *desc;
  ix->device = desc->device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, desc->device) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
  int rc = build_index(ix, desc);
  if (rc == 0) {
    cudaError_t se = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) rc = fail(SCANN_B200_INTERNAL, "cudaStreamCreate: %s", cudaGetErrorString(se));
    for (int i = 0; i < EV_ALL && rc == 0; ++i)
      if (cudaEventCreate(&ix->ev[i]) != cudaSuccess) rc = fail(SCANN_B200_INTERNAL, "cudaEventCreate failed");
  }
  if (rc) { delete ix; return rc; }
  // host pointers must not outlive this call
  ix->desc.centers = nullptr; ix->desc.tokens = nullptr; ix->desc.codes = nullptr; ix->desc.soar_codes = nullptr;
  ix->desc.codebook = nullptr; ix->desc.dataset = nullptr; ix->desc.bf16_dataset = nullptr; ix->desc.int8_dataset = nullptr;
  ix->desc.int8_multipliers = nullptr; ix->desc.dp_norms = nullptr; ix->desc.block_dims = nullptr;
  *out = ix;
  return 0;
}

void scann_b200_index_destroy(scann_b200_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  for (scann_b200_index* c : ix->lanes) {
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (int i = 0; i < EV_ALL; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
  }
  ix->lanes.clear();
  if (ix->stream) { cudaStreamSynchronize(ix->stream); cudaStreamDestroy(ix->stream); }
  for (int i = 0; i < EV_ALL; ++i) if (ix->ev[i]) cudaEventDestroy(ix->ev[i]);
  if (ix->comm) comm_destroy(ix->comm);
  delete ix;
}

uint32_t scann_b200_leaf_size(const scann_b200_index* ix, uint32_t leaf) {
  return (ix && leaf < ix->h_leaf_size.size()) ? ix->h_leaf_size[leaf] : 0;
}

int scann_b200_search_batched_device(scann_b200_index* ix, const float* d_queries, uint32_t nq,
                                     int32_t final_nn, int32_t pre_nn, int32_t leaves,
                                     uint32_t* d_out_idx, float* d_out_dist, int32_t out_k) {
  if (int rc = check_query_args(ix, d_queries, nq)) return rc;
  Params p;
  if (int rc = resolve(ix, final_nn, pre_nn, leaves, &p)) return rc;
  if (out_k <= 0 || !d_out_idx || !d_out_dist) return fail(SCANN_B200_INVALID_ARGUMENT, "bad output buffers");
  LaneGuard lane(ix);
  if (!lane.ix) return SCANN_B200_INTERNAL;
  ix = lane.ix;
  CU(cudaSetDevice(ix->device));
  ix->last = scann_b200_stats{};
  for (uint32_t s0 = 0; s0 < nq; s0 += ix->max_chunk) {
    const uint32_t c = std::min(ix->max_chunk, nq - s0);
    if (int rc = run_chunk(ix, d_queries + (size_t)s0 * ix->dev.d, c, p, d_out_idx + (size_t)s0 * out_k,
                           d_out_dist + (size_t)s0 * out_k, (uint32_t)out_k))
      return rc;
  }
  return 0;
}

int scann_b200_search_batched(scann_b200_index* ix, const float* queries, uint32_t nq, int32_t final_nn,
                              int32_t pre_nn, int32_t leaves, uint32_t* out_idx, float* out_dist,
                              int32_t out_k) {
  if (int rc = check_query_args(ix, queries, nq)) return rc;
  Params p;
  if (int rc = resolve(ix, final_nn, pre_nn, leaves, &p)) return rc;
  if (out_k <= 0 || !out_idx || !out_dist) return fail(SCANN_B200_INVALID_ARGUMENT, "bad output buffers");
  LaneGuard lane(ix);  // concurrent callers run on different lanes: one batch's copies overlap another's kernels
  if (!lane.ix) return SCANN_B200_INTERNAL;
  ix = lane.ix;
  CU(cudaSetDevice(ix->device));
  ix->last = scann_b200_stats{};
  const uint32_t D = ix->dev.d;
  auto pinned = [](const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  const bool q_pinned = pinned(queries), out_pinned = pinned(out_idx) && pinned(out_dist);
  for (uint32_t s0 = 0; s0 < nq; s0 += ix->max_chunk) {
    const uint32_t c = std::min(ix->max_chunk, nq - s0);
    CU(ix->q.ensure(sizeof(float) * (size_t)c * D));
    CU(ix->out_idx.ensure(sizeof(uint32_t) * (size_t)c * out_k));
    CU(ix->out_dist.ensure(sizeof(float) * (size_t)c * out_k));
    CU(ix->h_q.ensure(sizeof(float) * (size_t)c * D));
    CU(ix->h_idx.ensure(sizeof(uint32_t) * (size_t)c * out_k));
    CU(ix->h_dist.ensure(sizeof(float) * (size_t)c * out_k));
    // page-locked caller memory goes to the device directly; pageable memory through the pinned staging buffer
    const float* src = queries + (size_t)s0 * D;
    if (!q_pinned) {
      memcpy(ix->h_q.p, src, sizeof(float) * (size_t)c * D);
      src = ix->h_q.as<float>();
    }
    CU(cudaMemcpyAsync(ix->q.p, src, sizeof(float) * (size_t)c * D, cudaMemcpyHostToDevice, ix->stream));
    if (int rc = run_chunk(ix, ix->q.as<float>(), c, p, ix->out_idx.as<uint32_t>(), ix->out_dist.as<float>(),
                           (uint32_t)out_k))
      return rc;
    uint32_t* dst_i = out_pinned ? out_idx + (size_t)s0 * out_k : ix->h_idx.as<uint32_t>();
    float* dst_d = out_pinned ? out_dist + (size_t)s0 * out_k : ix->h_dist.as<float>();
    CU(cudaMemcpyAsync(dst_i, ix->out_idx.p, sizeof(uint32_t) * (size_t)c * out_k, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaMemcpyAsync(dst_d, ix->out_dist.p, sizeof(float) * (size_t)c * out_k, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    if (!out_pinned) {
      memcpy(out_idx + (size_t)s0 * out_k, ix->h_idx.p, sizeof(uint32_t) * (size_t)c * out_k);
      memcpy(out_dist + (size_t)s0 * out_k, ix->h_dist.p, sizeof(float) * (size_t)c * out_k);
    }
  }
  return 0;
}

int scann_b200_search_partial_device(scann_b200_index* ix, const float* d_queries, uint32_t nq,
                                     int32_t pre_nn, int32_t leaves, uint32_t* d_ids, uint64_t* d_tie,
                                     float* d_ah, float* d_exact, int32_t n_cand) {
  if (int rc = check_query_args(ix, d_queries, nq)) return rc;
  Params p;
  if (int rc = resolve(ix, -1, pre_nn, leaves, &p)) return rc;
  if (!d_ids || !d_tie || !d_ah || !d_exact || n_cand < (int)p.nover)
    return fail(SCANN_B200_INVALID_ARGUMENT, "partial buffers too small: n_cand=%d < %u", n_cand, p.nover);
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  { std::lock_guard<std::mutex> lk(ix->pool_mu); ix->last_any_valid = false; }
  ix->last = scann_b200_stats{};
  for (uint32_t s0 = 0; s0 < nq; s0 += ix->max_chunk) {
    const uint32_t c = std::min(ix->max_chunk, nq - s0);
    PartialOut po{d_ids + (size_t)s0 * n_cand, d_tie + (size_t)s0 * n_cand, d_ah + (size_t)s0 * n_cand,
                  d_exact + (size_t)s0 * n_cand, (uint32_t)n_cand};
    if (int rc = search_chunk(ix, d_queries + (size_t)s0 * ix->dev.d, c, p, nullptr, nullptr, 0, &po, false)) return rc;
  }
  return 0;
}

int scann_b200_merge_partials_device(scann_b200_index* ix, uint32_t nq, int32_t world, int32_t n_cand,
                                     const uint32_t* d_ids, const uint64_t* d_tie, const float* d_ah,
                                     const float* d_exact, int32_t pre_nn, int32_t final_nn,
                                     uint32_t* d_out_idx, float* d_out_dist, int32_t out_k) {
  if (!ix) return fail(SCANN_B200_INVALID_ARGUMENT, "null index");
  Params p;
  if (int rc = resolve(ix, final_nn, pre_nn, -1, &p)) return rc;
  if ((long long)world * n_cand > 8192) return fail(SCANN_B200_UNIMPLEMENTED, "merge of %d x %d candidates too large", world, n_cand);
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  (void)d_ah;
  CU(sb::launch_merge_partials(ix->dev, nq, world, n_cand, d_ids, d_tie, d_exact, p.nover, p.npre, p.k, d_out_idx,
                               d_out_dist, (uint32_t)out_k, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return 0;
}

int scann_b200_merge_topk_device(scann_b200_index* ix, uint32_t nq, int32_t world, int32_t k_in, const uint32_t* d_ids,
                                 const float* d_dists, int32_t final_nn, uint32_t* d_out_idx, float* d_out_dist,
                                 int32_t out_k) {
  if (!ix) return fail(SCANN_B200_INVALID_ARGUMENT, "null index");
  if (!d_ids || !d_dists || !d_out_idx || !d_out_dist || world < 1 || k_in < 1 || out_k < 1)
    return fail(SCANN_B200_INVALID_ARGUMENT, "merge_topk: bad arguments");
  if ((long long)world * k_in > 8192) return fail(SCANN_B200_UNIMPLEMENTED, "merge of %d x %d results too large", world, k_in);
  const int k = final_nn > 0 ? final_nn : k_in;
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  CU(sb::launch_merge_topk(ix->dev.distance, nq, world, k_in, d_ids, d_dists, (uint32_t)k, d_out_idx, d_out_dist,
                           (uint32_t)out_k, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return 0;
}

int scann_b200_last_stats(scann_b200_index* ix, scann_b200_stats* out) {
  if (!ix || !out) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  {
    std::lock_guard<std::mutex> lk(ix->pool_mu);
    if (ix->last_any_valid) { *out = ix->last_any; return 0; }
  }
  std::lock_guard<std::mutex> lock(ix->mu);
  *out = ix->last;
  return 0;
}

// ---- debug / parity hooks ------------------------------------------------------------------

int scann_b200_debug_tokenize(scann_b200_index* ix, const float* queries, uint32_t nq, int32_t leaves,
                              int32_t* out_leaf, float* out_dist) {
  if (int rc = check_query_args(ix, queries, nq)) return rc;
  Params p;
  if (int rc = resolve(ix, -1, -1, leaves, &p)) return rc;
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  const sb::DevIndex& v = ix->dev;
  if (int rc = ensure_workspace(ix, nq, p, 1, pick_cap(p.nover))) return rc;
  CU(cudaMemcpyAsync(ix->q.p, queries, sizeof(float) * (size_t)nq * v.d, cudaMemcpyHostToDevice, ix->stream));
  CU(sb::launch_tokenize_topp(v, ix->q.as<float>(), nq, p.P, ix->dist.as<float>(), ix->tok_a.p,
                              ix->leaves.as<int32_t>(), ix->bias.as<float>(), nullptr, ix->stream, nullptr));
  CU(cudaMemcpyAsync(out_leaf, ix->leaves.p, sizeof(int32_t) * (size_t)nq * p.P, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaMemcpyAsync(out_dist, ix->bias.p, sizeof(float) * (size_t)nq * p.P, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return 0;
}

int scann_b200_debug_lut(scann_b200_index* ix, const float* queries, uint32_t nq, uint8_t* out_lut, float* out_mult) {
  if (int rc = check_query_args(ix, queries, nq)) return rc;
  Params p;
  if (int rc = resolve(ix, -1, -1, -1, &p)) return rc;
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  const sb::DevIndex& v = ix->dev;
  if (int rc = ensure_workspace(ix, nq, p, 1, pick_cap(p.nover))) return rc;
  CU(cudaMemcpyAsync(ix->q.p, queries, sizeof(float) * (size_t)nq * v.d, cudaMemcpyHostToDevice, ix->stream));
  sb::launch_lut(v, ix->q.as<float>(), nq, ix->lut.as<uint8_t>(), ix->mult.as<float>(), ix->inv.as<float>(), ix->stream);
  CU(cudaGetLastError());
  // device rows are padded to 8W blocks; return the first B blocks of each
  CU(cudaMemcpy2DAsync(out_lut, (size_t)v.B * 16, ix->lut.p, (size_t)v.W * 128, (size_t)v.B * 16, nq,
                       cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaMemcpyAsync(out_mult, ix->mult.p, sizeof(float) * nq, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return 0;
}

int scann_b200_debug_leaf_scores(scann_b200_index* ix, const uint8_t* lut, uint32_t leaf, int16_t* out, uint32_t out_len) {
  if (!ix || !lut || !out) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  if (leaf >= ix->dev.L) return fail(SCANN_B200_INVALID_ARGUMENT, "leaf %u out of range", leaf);
  const uint32_t n = ix->h_leaf_size[leaf];
  if (out_len < n) return fail(SCANN_B200_INVALID_ARGUMENT, "output too small");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  const sb::DevIndex& v = ix->dev;
  DevBuf dl, dout;
  CU(dl.ensure((size_t)v.W * 128));
  CU(dout.ensure(sizeof(int16_t) * n));
  CU(cudaMemsetAsync(dl.p, 0, (size_t)v.W * 128, ix->stream));
  CU(cudaMemcpyAsync(dl.p, lut, (size_t)v.B * 16, cudaMemcpyHostToDevice, ix->stream));
  CU(sb::launch_leaf_scores(v, dl.as<uint8_t>(), leaf, dout.as<int16_t>(), ix->stream));
  CU(cudaMemcpyAsync(out, dout.p, sizeof(int16_t) * n, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return 0;
}

int scann_b200_debug_candidates(scann_b200_index* ix, const float* queries, uint32_t nq, int32_t pre_nn,
                                int32_t leaves, int32_t cap, uint32_t* out_leaf, uint32_t* out_slot,
                                uint32_t* out_dp, float* out_score, uint32_t* out_count) {
  if (int rc = check_query_args(ix, queries, nq)) return rc;
  Params p;
  if (int rc = resolve(ix, -1, pre_nn, leaves, &p)) return rc;
  if (nq > ix->max_chunk) return fail(SCANN_B200_INVALID_ARGUMENT, "debug_candidates: at most %u queries", ix->max_chunk);
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  const sb::DevIndex& v = ix->dev;
  { std::lock_guard<std::mutex> lk(ix->pool_mu); ix->last_any_valid = false; }
  ix->last = scann_b200_stats{};
  CU(ix->q.ensure(sizeof(float) * (size_t)nq * v.d));
  CU(cudaMemcpyAsync(ix->q.p, queries, sizeof(float) * (size_t)nq * v.d, cudaMemcpyHostToDevice, ix->stream));
  if (int rc = search_chunk(ix, ix->q.as<float>(), nq, p, nullptr, nullptr, 0, nullptr, true)) return rc;
  const uint32_t dcap = pick_cap(p.nover);
  std::vector<uint64_t> keys((size_t)nq * dcap);
  std::vector<uint32_t> cnt(nq);
  std::vector<uint32_t> goff(v.L + 1), sdp;
  CU(cudaMemcpy(keys.data(), ix->buf.p, sizeof(uint64_t) * keys.size(), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cnt.data(), ix->cnt.p, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(goff.data(), ix->leaf_goff.p, sizeof(uint32_t) * (v.L + 1), cudaMemcpyDeviceToHost));
  sdp.resize((size_t)goff[v.L] * 32);
  CU(cudaMemcpy(sdp.data(), ix->slot_dp.p, sizeof(uint32_t) * sdp.size(), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < nq; ++i) {
    const uint32_t n = std::min<uint32_t>(std::min(cnt[i], p.nover), (uint32_t)cap);
    out_count[i] = n;
    for (uint32_t j = 0; j < n; ++j) {
      const uint64_t k = keys[(size_t)i * dcap + j];
      const size_t o = (size_t)i * cap + j;
      if (v.key_by_dp) {  // squared-L2 keys carry the datapoint id instead of the slot
        out_leaf[o] = 0xFFFFFFFFu;
        out_slot[o] = 0xFFFFFFFFu;
        out_dp[o] = (uint32_t)k;
      } else {
        const uint32_t gs = (uint32_t)k, g = gs / 32;
        const uint32_t leaf = (uint32_t)(std::upper_bound(goff.begin(), goff.end(), g) - goff.begin()) - 1;
        out_leaf[o] = leaf;
        out_slot[o] = gs - goff[leaf] * 32;
        out_dp[o] = sdp[gs];
      }
      uint32_t ord = (uint32_t)(k >> 32);
      uint32_t u = (ord & 0x80000000u) ? (ord & 0x7fffffffu) : ~ord;
      memcpy(&out_score[o], &u, 4);
    }
  }
  return 0;
}

}  // extern "C"
