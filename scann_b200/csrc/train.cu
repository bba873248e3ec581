// train.cu -- k-means training on the GPU: the Lloyd iterations behind
//   KMeansTreePartitioner::TrainKMeans                  partitioning/kmeans_tree_partitioner.cc:424-441
//   GmmUtils::GenericKmeans / KMeansImpl main loop      utils/gmm_utils.cc:846-915
//   GmmUtils::RecomputeCentroidsSimple                  utils/gmm_utils.cc:1052-1132
//   AH codebook training (16 centres per block)         hashes/internal/asymmetric_hashing_impl.cc:41-197
// One iteration = (1) assignment of every training point to its nearest centre under squared L2 -- the query
// tokenizer with P = 1 (tcgen05 GEMM pre-filter + exact fp32 chain for >= 256 centres, the SIMT kernel below that),
// i.e. bit for bit the reference's UnbalancedFloat32PartitionAssignment on float centres -- and (2) the centroid
// update in the reference's arithmetic: double sums over the members IN INDEX ORDER inside kParallelAggregate = 4
// contiguous slices of the training set (one warp per slice, the reference's thread-pool path; one slice when
// n < 8 k), slice sums added in slice order, multiplied by double(1.0 / count), stored as float.  That makes the result
// a function of (data, initial centres, iterations) alone and lets the oracle restate it exactly.
// Deliberately not reproduced: the reference's initialisations (k-means++ / random, both draw from absl's unseeded
// bit generator: the caller passes the initial centres) and ReinitializeCenters for clusters below min_cluster_size
// (random / PCA splits): an EMPTY cluster keeps its previous centre here and is reported in the statistics.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "index_internal.h"

using namespace sbi;

namespace sb {

constexpr int kUpdThreads = 128;  // 4 warps = the 4 aggregation slices of RecomputeCentroidsSimple

// One CTA per centre.  Warp t walks slice t of the assignment array in index order; for every member all lanes add the
// point's coordinates (lane owns dims lane, lane + 32, ...) to double accumulators.
__global__ void __launch_bounds__(kUpdThreads)
kmeans_update_kernel(const float* __restrict__ x, const int32_t* __restrict__ assign, uint32_t n, uint32_t d, uint32_t k,
                     uint32_t slices, const float* __restrict__ old_centers, float* __restrict__ new_centers,
                     uint32_t* __restrict__ counts, uint32_t* __restrict__ n_empty) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* part = reinterpret_cast<double*>(smem);  // [4][d]
  __shared__ uint32_t s_cnt[4];
  const uint32_t c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kMaxPerLane = 8;  // d <= 256 in registers; larger d falls back to shared-memory accumulators
  double acc[kMaxPerLane];
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) acc[j] = 0.0;
  const bool in_regs = d <= 32u * kMaxPerLane;
  if (!in_regs)
    for (uint32_t j = lane; j < d; j += 32) part[(size_t)warp * d + j] = 0.0;
  uint32_t cnt = 0;
  if ((uint32_t)warp < slices) {
    const uint32_t per = (n + slices - 1) / slices;
    const uint32_t lo = (uint32_t)warp * per, hi = min(n, lo + per);
    for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
      const uint32_t i = i0 + lane;
      const bool mine = i < hi && assign[i] == (int32_t)c;
      uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
      cnt += __popc(m);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const float* row = x + (size_t)(i0 + b) * d;
        if (in_regs) {
#pragma unroll
          for (int j = 0; j < kMaxPerLane; ++j) {
            const uint32_t dim = (uint32_t)lane + 32u * j;
            if (dim < d) acc[j] = __dadd_rn(acc[j], (double)row[dim]);
          }
        } else {
          for (uint32_t dim = lane; dim < d; dim += 32) part[(size_t)warp * d + dim] = __dadd_rn(part[(size_t)warp * d + dim], (double)row[dim]);
        }
      }
    }
  }
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      const uint32_t dim = (uint32_t)lane + 32u * j;
      if (dim < d) part[(size_t)warp * d + dim] = acc[j];
    }
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  const uint32_t total = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
  if (threadIdx.x == 0) {
    counts[c] = total;
    if (total == 0) atomicAdd(n_empty, 1u);
  }
  for (uint32_t dim = threadIdx.x; dim < d; dim += kUpdThreads) {
    if (total == 0) { new_centers[(size_t)c * d + dim] = old_centers[(size_t)c * d + dim]; continue; }
    double sum = 0.0;  // out_centroid starts at 0 and the slice sums are added in slice order
    for (uint32_t t = 0; t < slices; ++t) sum = __dadd_rn(sum, part[(size_t)t * d + dim]);
    const double mult = __ddiv_rn(1.0, (double)total);  // NormalizeCentroid: multiplier = 1.0 / divisor
    new_centers[(size_t)c * d + dim] = (float)__dmul_rn(sum, mult);
  }
}

}  // namespace sb

extern "C" int scann_b200_train_kmeans(const scann_b200_kmeans_desc* d, float* centers_out, int32_t* assign_out,
                                       scann_b200_kmeans_stats* stats_out) {
  if (!d || !centers_out) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  const uint32_t N = d->n, D = d->d, K = d->k;
  if (!N || !D || !K || !d->data || !d->init_centers) return fail(SCANN_B200_INVALID_ARGUMENT, "train_kmeans: empty input");
  if (K > N) return fail(SCANN_B200_INVALID_ARGUMENT, "Number of points (%u) is less than the number of clusters (%u).", N, K);
  if (d->iterations < 0) return fail(SCANN_B200_INVALID_ARGUMENT, "train_kmeans: negative iteration count");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SCANN_B200_FAILED_PRECONDITION, "no CUDA device available; scann_b200 has no CPU path");
  if (d->device < 0 || d->device >= ndev) return fail(SCANN_B200_INVALID_ARGUMENT, "device %d out of range", d->device);
  CU(cudaSetDevice(d->device));
  cudaStream_t s = nullptr;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { if (s) cudaStreamDestroy(s); } } sg{s};
  cudaEvent_t ev[4] = {};
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 4; ++i) if (e[i]) cudaEventDestroy(e[i]); } } eg{ev};
  for (auto& e : ev) CU(cudaEventCreate(&e));

  DevBuf x, cen[2], cnorm, tok_b, tok_a, dist, tok_cmax, assign, bias, counts, nempty;
  CU(x.ensure(sizeof(float) * (size_t)N * D));
  CU(cudaMemcpyAsync(x.p, d->data, sizeof(float) * (size_t)N * D, cudaMemcpyHostToDevice, s));
  for (auto& c : cen) CU(c.ensure(sizeof(float) * (size_t)K * D));
  CU(cudaMemcpyAsync(cen[0].p, d->init_centers, sizeof(float) * (size_t)K * D, cudaMemcpyHostToDevice, s));
  CU(cnorm.ensure(sizeof(float) * K));
  CU(tok_b.ensure(sb::tokenize_operand_bytes(K, D)));
  // rows per assignment launch: 16k for a large tree (the [R][K] distance matrix), up to 1M for a 16-centre codebook
  const uint32_t R = std::min<uint32_t>(N, std::max<uint32_t>(16384u, std::min<uint32_t>(1u << 20, (1u << 25) / K)));
  CU(tok_a.ensure(sb::tokenize_operand_bytes(R, D)));
  CU(dist.ensure(sizeof(float) * (size_t)R * K));
  CU(tok_cmax.ensure(sizeof(float) * (size_t)R * ((K + 31) / 32)));
  CU(assign.ensure(sizeof(int32_t) * (size_t)N));
  CU(bias.ensure(sizeof(float) * (size_t)N));
  CU(counts.ensure(sizeof(uint32_t) * K));
  CU(nempty.ensure(sizeof(uint32_t) * 2));
  std::vector<float> hc((size_t)K * D), cn(K);
  memcpy(hc.data(), d->init_centers, sizeof(float) * (size_t)K * D);
  const uint32_t slices = (N >= (uint64_t)K * 4 * 2) ? 4u : 1u;  // gmm_utils.cc:1064-1065 (thread-pool path)
  const size_t upd_smem = sizeof(double) * 4 * (size_t)D;
  if (upd_smem > 200 * 1024) return fail(SCANN_B200_UNIMPLEMENTED, "train_kmeans: dimensionality %u too large", D);
  CU(cudaFuncSetAttribute(sb::kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  float ms_assign = 0.f, ms_update = 0.f;
  uint32_t empty_last = 0;
  int cur = 0;
  CU(cudaEventRecord(ev[0], s));
  const int passes = d->iterations + (assign_out ? 1 : 0);  // T recomputes; one more assignment for the final partition
  for (int it = 0; it < passes; ++it) {
    // centre norms in the many-to-many kernel's arithmetic (many_to_many_impl.inc:236-257) and the error bound's ||c||
    double cmax2 = 0.0;
    for (uint32_t l = 0; l < K; ++l) {
      float acc = 0.f;
      double a64 = 0.0;
      for (uint32_t k = 0; k < D; ++k) {
        const float c = hc[(size_t)l * D + k];
        acc = fmaf(-c, c, acc);
        a64 += (double)c * (double)c;
      }
      cn[l] = acc * -1.0f;
      cmax2 = std::max(cmax2, a64);
    }
    CU(cudaMemcpyAsync(cnorm.p, cn.data(), sizeof(float) * K, cudaMemcpyHostToDevice, s));
    sb::DevIndex v{};
    v.distance = SCANN_B200_SQUARED_L2;
    v.n = N; v.d = D; v.L = K;
    v.centers = cen[cur].as<float>();
    v.center_sqnorm = cnorm.as<float>();
    v.center_max_norm = (float)(std::sqrt(cmax2) * 1.0001);
    v.tok_kp = sb::tokenize_kpitch(D);
    CU(sb::build_tokenize_operand(v.centers, K, D, 2, tok_b.p, s));
    v.tok_b = tok_b.p;
    v.tok_cmax_ws = tok_cmax.as<float>();
    CU(cudaEventRecord(ev[1], s));
    for (uint64_t r0 = 0; r0 < N; r0 += R) {
      const uint32_t nr = (uint32_t)std::min<uint64_t>(R, N - r0);
      CU(sb::launch_tokenize_topp(v, x.as<float>() + r0 * D, nr, 1, dist.as<float>(), tok_a.p, assign.as<int32_t>() + r0,
                                  bias.as<float>() + r0, nullptr, s, nullptr));
    }
    CU(cudaEventRecord(ev[2], s));
    if (it < d->iterations) {
      CU(cudaMemsetAsync(nempty.p, 0, sizeof(uint32_t) * 2, s));
      sb::kmeans_update_kernel<<<K, sb::kUpdThreads, upd_smem, s>>>(x.as<float>(), assign.as<int32_t>(), N, D, K, slices,
                                                                   cen[cur].as<float>(), cen[cur ^ 1].as<float>(),
                                                                   counts.as<uint32_t>(), nempty.as<uint32_t>());
      CU(cudaGetLastError());
      cur ^= 1;
      CU(cudaMemcpyAsync(hc.data(), cen[cur].p, sizeof(float) * (size_t)K * D, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpyAsync(&empty_last, nempty.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaEventRecord(ev[3], s));
    CU(cudaStreamSynchronize(s));
    float t = 0;
    CU(cudaEventElapsedTime(&t, ev[1], ev[2])); ms_assign += t;
    CU(cudaEventElapsedTime(&t, ev[2], ev[3])); ms_update += t;
  }
  memcpy(centers_out, hc.data(), sizeof(float) * (size_t)K * D);
  double mean_d = 0.0;
  if (assign_out) {
    CU(cudaMemcpy(assign_out, assign.p, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost));
    std::vector<float> hb(N);
    CU(cudaMemcpy(hb.data(), bias.p, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < N; ++i) mean_d += hb[i];
    mean_d /= N;
  }
  float tot = 0;
  CU(cudaEventRecord(ev[3], s));
  CU(cudaStreamSynchronize(s));
  CU(cudaEventElapsedTime(&tot, ev[0], ev[3]));
  if (stats_out) {
    stats_out->ms_assign = ms_assign; stats_out->ms_update = ms_update; stats_out->ms_total = tot;
    stats_out->iterations = (uint32_t)d->iterations; stats_out->empty_clusters = empty_last;
    stats_out->mean_sq_distance = mean_d;
  }
  return SCANN_B200_OK;
}
