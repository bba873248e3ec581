// sharded.cu -- database-sharded search over the GPUs of one box (BASELINE.json configs[3], [4]).
//
// The reference has no multi-device searcher; what has to be reproduced is the single-device result of
//   TreeAHHybridResidual::FindNeighborsBatchedImpl   tree_x_hybrid/tree_ah_hybrid_residual.cc:631-846
//   SingleMachineSearcherBase::FindNeighborsBatched  base/single_machine_base.cc:569-587 (reorder + top-k)
// with the packed leaves dealt out over `world` ranks (one process per GPU).  Protocol per batch, all on the
// index's stream, NCCL over NVLink as the only exchange (include/scann_b200.h has the byte counts):
//   tokenization of nq / world queries per rank  -> all-gather (leaf, centre distance)
//   pilot on the owner of each query's nearest leaf -> all-reduce(min) of the pruning thresholds
//   LUT16 scan + local top-N' + exact distances of the local candidates (search_chunk's kernels, unchanged)
//   all-to-all of 16-byte records to the query's owner -> merge of `world` sorted lists -> all-gather of the k results
// NCCL is dlopen'ed (libnccl.so.2, the copy the process already holds if torch loaded one), so a single-GPU
// deployment of libscann_b200.so has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "index_internal.h"

using namespace sbi;

namespace sbi {

struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (tried) return api.so ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("SCANN_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.so) break;
  }
  if (!api.so) return nullptr;
  bool ok = true;
  auto sym = [&](const char* n) { void* p = dlsym(api.so, n); if (!p) ok = false; return p; };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) { dlclose(api.so); api.so = nullptr; return nullptr; }
  return &api;
}

struct ShardComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

void comm_destroy(ShardComm* c) {
  if (!c) return;
  NcclApi* api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
}

}  // namespace sbi

#define NC(expr)                                                                                     \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess)                                                                           \
      return fail(SCANN_B200_INTERNAL, "NCCL error at %s:%d: %s", __FILE__, __LINE__, api->GetErrorString(_r)); \
  } while (0)

namespace {

// The exchanges of the protocol.  NCCL: one rank in this process (R.size() == 1).  Local: all ranks in this process,
// device copies through the host-synchronised streams (a test vehicle, not a fast path).
struct Coll {
  std::vector<scann_b200_index*>& R;
  int world;
  bool local;
  NcclApi* api = nullptr;

  int sync_all() {
    for (auto* ix : R) { CU(cudaSetDevice(ix->device)); CU(cudaStreamSynchronize(ix->stream)); }
    return 0;
  }
  // every rank holds `world` slices of `bytes`; rank r's own slice (at r * bytes) is distributed to all
  int allgather(std::vector<char*> base, size_t bytes) {
    if (world == 1 || bytes == 0) return 0;
    if (!local) {
      scann_b200_index* ix = R[0];
      NC(api->AllGather(base[0] + (size_t)ix->shard_rank * bytes, base[0], bytes, ncclUint8, ix->comm->comm, ix->stream));
      return 0;
    }
    if (int rc = sync_all()) return rc;
    for (size_t dst = 0; dst < R.size(); ++dst) {
      CU(cudaSetDevice(R[dst]->device));
      for (size_t src = 0; src < R.size(); ++src)
        if (src != dst)
          CU(cudaMemcpyAsync(base[dst] + src * bytes, base[src] + src * bytes, bytes, cudaMemcpyDefault, R[dst]->stream));
    }
    return sync_all();
  }
  int allreduce_min_u64(std::vector<uint64_t*> buf, size_t count) {
    if (world == 1 || count == 0) return 0;
    if (!local) {
      scann_b200_index* ix = R[0];
      NC(api->AllReduce(buf[0], buf[0], count, ncclUint64, ncclMin, ix->comm->comm, ix->stream));
      return 0;
    }
    if (int rc = sync_all()) return rc;
    std::vector<uint64_t> acc(count, ~0ull), tmp(count);
    for (size_t r = 0; r < R.size(); ++r) {
      CU(cudaMemcpy(tmp.data(), buf[r], count * 8, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < count; ++i) acc[i] = std::min(acc[i], tmp[i]);
    }
    for (size_t r = 0; r < R.size(); ++r) CU(cudaMemcpy(buf[r], acc.data(), count * 8, cudaMemcpyHostToDevice));
    return 0;
  }
  // block p of rank r's send buffer goes to block r of rank p's receive buffer
  int alltoall(std::vector<char*> send, std::vector<char*> recv, size_t bytes) {
    if (bytes == 0) return 0;
    if (world == 1) {
      CU(cudaMemcpyAsync(recv[0], send[0], bytes, cudaMemcpyDeviceToDevice, R[0]->stream));
      return 0;
    }
    if (!local) {
      scann_b200_index* ix = R[0];
      NC(api->GroupStart());
      for (int peer = 0; peer < world; ++peer) {
        NC(api->Send(send[0] + (size_t)peer * bytes, bytes, ncclUint8, peer, ix->comm->comm, ix->stream));
        NC(api->Recv(recv[0] + (size_t)peer * bytes, bytes, ncclUint8, peer, ix->comm->comm, ix->stream));
      }
      NC(api->GroupEnd());
      return 0;
    }
    if (int rc = sync_all()) return rc;
    for (size_t dst = 0; dst < R.size(); ++dst) {
      CU(cudaSetDevice(R[dst]->device));
      for (size_t src = 0; src < R.size(); ++src)
        CU(cudaMemcpyAsync(recv[dst] + src * bytes, send[src] + dst * bytes, bytes, cudaMemcpyDefault, R[dst]->stream));
    }
    return sync_all();
  }
};

struct RankState {
  sb::ScanWork w{};
  int launches = 0;
  uint32_t scan_launches = 0, retries = 0;
  uint32_t mode_launches[3] = {0, 0, 0};
  bool two_phase = false;
};

// One chunk of queries (identical on every rank).  d_q [nq][D] on each rank's device; outputs [nq][out_k] per rank.
int sharded_chunk(Coll& coll, const std::vector<const float*>& d_q, uint32_t nq, const Params& p, bool light,
                  const std::vector<uint32_t*>& d_out_idx, const std::vector<float*>& d_out_dist, uint32_t out_k) {
  std::vector<scann_b200_index*>& R = coll.R;
  const int G = coll.world;
  const uint32_t slice = (nq + G - 1) / G, nq_pad = slice * (uint32_t)G;
  const uint32_t cap = pick_cap(p.nover);
  const uint32_t ncand = p.nover;
  // sampled global threshold: SCANN_B200_SHARD_SAMPLE=0 sends every rank's whole local top-N' instead
  bool sampled = !light && G > 1;
  if (const char* e = getenv("SCANN_B200_SHARD_SAMPLE")) sampled = sampled && e[0] != '0';
  const uint32_t S = (p.nover + 15) / 16;
  std::vector<RankState> RS(R.size());
  auto each = [&](auto&& f) -> int {
    for (size_t r = 0; r < R.size(); ++r) {
      CU(cudaSetDevice(R[r]->device));
      if (int rc = f(R[r], RS[r], r)) return rc;
    }
    return 0;
  };
  auto ptrs = [&](auto&& get) { std::vector<char*> v; for (auto* ix : R) v.push_back(reinterpret_cast<char*>(get(ix))); return v; };
  auto mark = [&](int ev) -> int { return each([&](scann_b200_index* ix, RankState&, size_t) -> int { CU(cudaEventRecord(ix->ev[ev], ix->stream)); return 0; }); };

  // 1. tokenization of this rank's query slice, all-gather of the probed leaves and their centre distances
  if (int rc = each([&](scann_b200_index* ix, RankState& st, size_t r) -> int {
        Params pp = p;
        if (int rc = ensure_workspace(ix, nq_pad, pp, out_k, cap)) return rc;
        CU(ix->sh_send.ensure(sizeof(uint4) * (size_t)nq_pad * ncand));
        CU(ix->sh_recv.ensure(sizeof(uint4) * (size_t)nq_pad * ncand));
        const size_t gk = light ? (size_t)G * nq * out_k : (size_t)nq_pad * out_k;
        CU(ix->sh_idx.ensure(sizeof(uint32_t) * gk));
        CU(ix->sh_dist.ensure(sizeof(float) * gk));
        CU(ix->sh_samples.ensure(sizeof(uint32_t) * ((size_t)G * nq * S + 4)));
        CU(ix->sh_limit.ensure(sizeof(uint32_t) * ((size_t)nq + 4)));
        fill_scan_work(ix, nq, p, cap, &st.w);
        sb::ScanWork& w = st.w;
        w.pilot_world = (G > 1 && ix->shard_mode == SCANN_B200_SHARD_BY_LEAF) ? (uint32_t)G : 1u;
        w.pilot_rank = (uint32_t)ix->shard_rank;
        cudaStream_t s = ix->stream;
        CU(cudaMemsetAsync(w.counters, 0, sizeof(uint32_t) * 8, s));
        CU(cudaMemsetAsync(w.stats, 0, sizeof(unsigned long long) * 4, s));
        CU(cudaEventRecord(ix->ev[EV_START], s));
        const uint32_t q0 = std::min<uint32_t>((uint32_t)ix->shard_rank * slice, nq);
        const uint32_t nloc = std::min<uint32_t>(slice, nq - q0);
        if (nloc)
          CU(sb::launch_tokenize_topp(ix->dev, d_q[r] + (size_t)q0 * ix->dev.d, nloc, p.P, ix->dist.as<float>(), ix->tok_a.p,
                                      ix->leaves.as<int32_t>() + (size_t)q0 * p.P, ix->bias.as<float>() + (size_t)q0 * p.P,
                                      w.counters + 5, s, &st.launches));
        CU(cudaEventRecord(ix->ev[EV_TOK], s));
        return 0;
      })) return rc;
  if (int rc = mark(EV_C0)) return rc;
  if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->leaves.p; }), sizeof(int32_t) * (size_t)slice * p.P)) return rc;
  if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->bias.p; }), sizeof(float) * (size_t)slice * p.P)) return rc;
  if (int rc = mark(EV_C1)) return rc;

  // 2. LUTs (replicated: 16 B per block per query), pilot on the owner of the nearest leaf, all-reduce(min) of tau
  if (int rc = each([&](scann_b200_index* ix, RankState& st, size_t r) -> int {
        sb::ScanWork& w = st.w;
        cudaStream_t s = ix->stream;
        w.q_for_lut = nullptr;
        sb::launch_lut(ix->dev, d_q[r], nq, ix->lut.as<uint8_t>(), ix->mult.as<float>(), ix->inv.as<float>(), s);
        st.launches += 1;
        CU(cudaGetLastError());
        CU(cudaEventRecord(ix->ev[EV_LUT], s));
        // two scan phases only where this RANK probes many slots per query (leaf sharding: P / world leaves)
        const uint64_t volume = (uint64_t)p.P * ix->avg_leaf_slots / (w.pilot_world > 1 ? (uint64_t)G : 1ull);
        st.two_phase = p.P >= 16 && volume >= 49152;
        if (const char* e = getenv("SCANN_B200_TWO_PHASE")) st.two_phase = e[0] == '1' && p.P >= 2;
        uint32_t r1 = st.two_phase ? std::max<uint32_t>(1, p.P / 8) : p.P;
        if (const char* e = getenv("SCANN_B200_PHASE1_RANKS")) { const int t = atoi(e); if (st.two_phase && t >= 1 && (uint32_t)t < p.P) r1 = (uint32_t)t; }
        w.rank_lo = 0; w.rank_hi = r1;
        sb::scan_prepare_phase(ix->dev, &w);
        CU(sb::launch_pilot(ix->dev, w, s));
        st.launches += 1;
        CU(cudaEventRecord(ix->ev[EV_PILOT], s));
        return 0;
      })) return rc;
  if (int rc = mark(EV_C2)) return rc;
  {
    std::vector<uint64_t*> taus;
    for (auto* ix : R) taus.push_back(ix->tau.as<uint64_t>());
    if (int rc = coll.allreduce_min_u64(taus, nq)) return rc;
  }
  if (int rc = mark(EV_C3)) return rc;

  // 3. scan of the rank's own leaves, local top-N', exact distances (or the local top-k in light mode)
  if (int rc = each([&](scann_b200_index* ix, RankState& st, size_t r) -> int {
        sb::ScanWork& w = st.w;
        const sb::DevIndex& v = ix->dev;
        cudaStream_t s = ix->stream;
        int ncl = 0;
        sb::launch_worklist(v, w, false, true, s, &st.launches);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ix->ev[EV_WORK], s));
        CU(sb::launch_scan(v, w, 0, s));
        st.launches += 1; st.scan_launches += 1; st.mode_launches[w.scan_mode < 3 ? w.scan_mode : 0] += 1;
        CU(cudaEventRecord(ix->ev[EV_SCAN], s));
        CU(sb::launch_compact(v, w, false, s, &ncl));
        st.launches += ncl;
        CU(cudaEventRecord(ix->ev[EV_COMPACT], s));
        if (st.two_phase) {
          w.rank_lo = w.rank_hi; w.rank_hi = p.P;
          sb::scan_prepare_phase(v, &w);
          sb::launch_worklist(v, w, false, false, s, &st.launches);
          CU(cudaGetLastError());
          CU(cudaEventRecord(ix->ev[EV2_WORK], s));
          CU(sb::launch_scan(v, w, 0, s));
          st.launches += 1; st.scan_launches += 1; st.mode_launches[w.scan_mode < 3 ? w.scan_mode : 0] += 1;
          CU(cudaEventRecord(ix->ev[EV2_SCAN], s));
          CU(sb::launch_compact(v, w, false, s, &ncl));
          st.launches += ncl;
          CU(cudaEventRecord(ix->ev[EV2_COMPACT], s));
        }
        w.rank_lo = 0; w.rank_hi = p.P;
        w.rescan = 1;
        sb::scan_prepare_phase(v, &w);
        uint32_t* hc = ix->h_counters.as<uint32_t>();
        unsigned long long* hs = reinterpret_cast<unsigned long long*>(hc + 8);
        CU(cudaMemcpyAsync(hc, w.counters, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(hs, w.stats, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        const uint32_t tok_fallbacks = hc[5];
        if (hc[7] != 0) return fail(SCANN_B200_INTERNAL, "tensor-core scan: a pipeline barrier timed out (watchdog)");
        while (hc[2] != 0) {  // candidate-buffer overflow: re-scan the flagged queries (local, no collective involved)
          if (++st.retries > 256) return fail(SCANN_B200_INTERNAL, "candidate buffer overflow did not converge");
          CU(cudaMemsetAsync(w.counters + 2, 0, sizeof(uint32_t), s));
          sb::launch_worklist(v, w, true, false, s, &st.launches);
          CU(sb::launch_scan(v, w, 0, s));
          CU(sb::launch_compact(v, w, true, s, &ncl));
          st.launches += 1 + ncl; st.scan_launches += 1; st.mode_launches[w.scan_mode < 3 ? w.scan_mode : 0] += 1;
          CU(cudaMemcpyAsync(hc, w.counters, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
          CU(cudaStreamSynchronize(s));
        }
        if (sampled) {
          CU(sb::launch_sample_scores(w, S, ix->sh_samples.as<uint32_t>() + (size_t)ix->shard_rank * nq * S, s));
          st.launches += 1;
        }
        CU(cudaEventRecord(ix->ev[EV_S0], s));
        scann_b200_stats& ls = ix->last;
        ls.scan_bytes_alg += hs[0]; ls.scan_pairs += hs[1]; ls.scan_lookups += hs[0] * 2;
        ls.cand_sum += hs[2]; ls.cand_max = std::max<uint64_t>(ls.cand_max, hs[3]);
        ls.tokenize_fallbacks += tok_fallbacks;
        return 0;
      })) return rc;

  // 3b. sampled global threshold (parity mode, world > 1): all-gather every 16th local score, threshold per query
  if (int rc = mark(EV_C8)) return rc;
  if (sampled)
    if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->sh_samples.p; }), sizeof(uint32_t) * (size_t)nq * S)) return rc;
  if (int rc = mark(EV_C9)) return rc;
  if (int rc = each([&](scann_b200_index* ix, RankState& st, size_t r) -> int {
        sb::ScanWork& w = st.w;
        cudaStream_t s = ix->stream;
        sb::FinalizeArgs a{};
        a.q = d_q[r]; a.nq = nq; a.npre = p.npre; a.k = p.k; a.out_k = out_k;
        if (light) {
          a.out_idx = ix->sh_idx.as<uint32_t>() + (size_t)ix->shard_rank * nq * out_k;
          a.out_dist = ix->sh_dist.as<float>() + (size_t)ix->shard_rank * nq * out_k;
        } else {
          a.part_rec = ix->sh_send.as<uint4>();
          a.part_cap = ncand;
          if (sampled) {
            CU(sb::launch_sample_threshold(ix->sh_samples.as<uint32_t>(), G, nq, S, p.nover, ix->sh_limit.as<uint32_t>(), s));
            st.launches += 1;
            a.part_limit = ix->sh_limit.as<uint32_t>();
          }
          if (nq_pad > nq)  // records of the padding queries: invalid ids
            CU(cudaMemsetAsync(ix->sh_send.as<uint4>() + (size_t)nq * ncand, 0xFF, sizeof(uint4) * (size_t)(nq_pad - nq) * ncand, s));
        }
        CU(sb::launch_finalize(ix->dev, w, a, s));
        st.launches += 1;
        CU(cudaEventRecord(ix->ev[EV_FIN], s));
        return 0;
      })) return rc;

  // 4-6. exchange and merge
  if (int rc = mark(EV_C4)) return rc;
  uint64_t sent = 2ull * sizeof(int32_t) * slice * p.P + 8ull * nq;
  if (light) {
    if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->sh_idx.p; }), sizeof(uint32_t) * (size_t)nq * out_k)) return rc;
    if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->sh_dist.p; }), sizeof(float) * (size_t)nq * out_k)) return rc;
    if (int rc = mark(EV_C5)) return rc;
    if (int rc = mark(EV_M0)) return rc;
    if (int rc = each([&](scann_b200_index* ix, RankState& st, size_t r) -> int {
          CU(sb::launch_merge_topk(ix->dev.distance, nq, G, (int)out_k, ix->sh_idx.as<uint32_t>(), ix->sh_dist.as<float>(), p.k,
                                   d_out_idx[r], d_out_dist[r], out_k, ix->stream, true));
          st.launches += 1;
          return 0;
        })) return rc;
    if (int rc = mark(EV_M1)) return rc;
    if (int rc = mark(EV_C6)) return rc;
    if (int rc = mark(EV_C7)) return rc;
    sent += 8ull * nq * out_k;
  } else {
    const size_t blk = sizeof(uint4) * (size_t)slice * ncand;
    if (int rc = coll.alltoall(ptrs([](scann_b200_index* ix) { return ix->sh_send.p; }), ptrs([](scann_b200_index* ix) { return ix->sh_recv.p; }), blk)) return rc;
    if (int rc = mark(EV_C5)) return rc;
    if (int rc = mark(EV_M0)) return rc;
    if (int rc = each([&](scann_b200_index* ix, RankState& st, size_t) -> int {
          const uint32_t q0 = std::min<uint32_t>((uint32_t)ix->shard_rank * slice, nq);
          const uint32_t nloc = std::min<uint32_t>(slice, nq - q0);
          CU(sb::launch_merge_records(ix->dev, nloc, slice, G, (int)ncand, ix->sh_recv.as<uint4>(), p.nover, p.npre, p.k,
                                      ix->sh_idx.as<uint32_t>() + (size_t)ix->shard_rank * slice * out_k,
                                      ix->sh_dist.as<float>() + (size_t)ix->shard_rank * slice * out_k, out_k, ix->stream));
          st.launches += 1;
          return 0;
        })) return rc;
    if (int rc = mark(EV_M1)) return rc;
    if (int rc = mark(EV_C6)) return rc;
    if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->sh_idx.p; }), sizeof(uint32_t) * (size_t)slice * out_k)) return rc;
    if (int rc = coll.allgather(ptrs([](scann_b200_index* ix) { return ix->sh_dist.p; }), sizeof(float) * (size_t)slice * out_k)) return rc;
    if (int rc = mark(EV_C7)) return rc;
    if (int rc = each([&](scann_b200_index* ix, RankState&, size_t r) -> int {
          CU(cudaMemcpyAsync(d_out_idx[r], ix->sh_idx.p, sizeof(uint32_t) * (size_t)nq * out_k, cudaMemcpyDeviceToDevice, ix->stream));
          CU(cudaMemcpyAsync(d_out_dist[r], ix->sh_dist.p, sizeof(float) * (size_t)nq * out_k, cudaMemcpyDeviceToDevice, ix->stream));
          return 0;
        })) return rc;
    sent += (uint64_t)blk * (G - 1) + 8ull * slice * out_k;
  }
  if (int rc = mark(EV_END)) return rc;

  // timings (CUDA events on each rank's stream)
  return each([&](scann_b200_index* ix, RankState& st, size_t) -> int {
    CU(cudaStreamSynchronize(ix->stream));
    auto el = [&](int a, int b, float* out) -> int { CU(cudaEventElapsedTime(out, ix->ev[a], ix->ev[b])); return 0; };
    float t = 0, x = 0;
    scann_b200_stats& ls = ix->last;
    if (int rc = el(EV_START, EV_TOK, &t)) return rc; ls.ms_tokenize += t;
    if (int rc = el(EV_C1, EV_LUT, &t)) return rc; ls.ms_lut += t;
    if (int rc = el(EV_LUT, EV_PILOT, &t)) return rc; ls.ms_pilot += t;
    if (int rc = el(EV_C3, EV_WORK, &t)) return rc; ls.ms_worklist += t;
    if (int rc = el(EV_WORK, EV_SCAN, &t)) return rc; ls.ms_scan += t;
    if (int rc = el(EV_SCAN, EV_COMPACT, &t)) return rc; ls.ms_compact += t;
    if (st.two_phase) {
      if (int rc = el(EV_COMPACT, EV2_WORK, &t)) return rc; ls.ms_worklist += t;
      if (int rc = el(EV2_WORK, EV2_SCAN, &t)) return rc; ls.ms_scan += t;
      if (int rc = el(EV2_SCAN, EV2_COMPACT, &t)) return rc; ls.ms_compact += t;
    }
    if (int rc = el(EV_C9, EV_FIN, &t)) return rc; ls.ms_finalize += t;
    for (int c = 0; c < 5; ++c) { if (int rc = el(EV_C0 + 2 * c, EV_C0 + 2 * c + 1, &t)) return rc; x += t; }
    if (int rc = el(EV_M0, EV_M1, &t)) return rc;
    ls.ms_merge += t;
    ls.ms_exchange += x;
    if (int rc = el(EV_START, EV_END, &t)) return rc; ls.ms_total += t;
    ls.exchange_bytes += sent;
    ls.kernel_launches += (uint32_t)st.launches;
    ls.overflow_retries += st.retries;
    ls.scan_kernel_count += st.scan_launches;
    ls.scan_oct_launches += st.mode_launches[0]; ls.scan_wide_launches += st.mode_launches[1]; ls.scan_tc_launches += st.mode_launches[2];
    return 0;
  });
}

int sharded_search(std::vector<scann_b200_index*>& R, bool local, const float* d_queries, uint32_t nq, int final_nn,
                   int pre_nn, int leaves, int light, uint32_t* d_out_idx, float* d_out_dist, int out_k) {
  if (R.empty() || !R[0]) return fail(SCANN_B200_INVALID_ARGUMENT, "null index");
  if (nq && !d_queries) return fail(SCANN_B200_INVALID_ARGUMENT, "null queries");
  if (out_k <= 0 || !d_out_idx || !d_out_dist) return fail(SCANN_B200_INVALID_ARGUMENT, "bad output buffers");
  const int G = R[0]->shard_world;
  for (auto* ix : R) {
    if (!ix || ix->brute) return fail(SCANN_B200_INVALID_ARGUMENT, "sharded search needs tree-AH shards");
    if (ix->shard_world != G || ix->shard_mode != R[0]->shard_mode)
      return fail(SCANN_B200_INVALID_ARGUMENT, "shards of different worlds / modes");
  }
  if (local) {
    if ((int)R.size() != G) return fail(SCANN_B200_INVALID_ARGUMENT, "%zu local shards of a world of %d", R.size(), G);
    for (int r = 0; r < G; ++r)
      if (R[r]->shard_rank != r) return fail(SCANN_B200_INVALID_ARGUMENT, "local shard %d has rank %d", r, R[r]->shard_rank);
  }
  Coll coll{R, G, local, nullptr};
  if (!local && G > 1) {
    if (!R[0]->comm) return fail(SCANN_B200_FAILED_PRECONDITION, "scann_b200_comm_init has not been called on this index");
    coll.api = nccl_api();
    if (!coll.api) return fail(SCANN_B200_FAILED_PRECONDITION, "libnccl.so.2 could not be loaded");
  }
  Params p;
  if (int rc = resolve(R[0], final_nn, pre_nn, leaves, &p)) return rc;
  if ((long long)G * p.nover > 8192 && !light)
    return fail(SCANN_B200_UNIMPLEMENTED, "merge of %d x %u candidates too large", G, p.nover);
  std::vector<std::unique_lock<std::mutex>> locks;
  for (auto* ix : R) {
    locks.emplace_back(ix->mu);
    ix->last = scann_b200_stats{};
    std::lock_guard<std::mutex> lk(ix->pool_mu);
    ix->last_any_valid = false;
  }
  const uint32_t D = R[0]->dev.d;
  std::vector<uint32_t*> oi(R.size(), nullptr);
  std::vector<float*> od(R.size(), nullptr);
  std::vector<const float*> dq(R.size(), nullptr);
  // local mode writes one result (rank 0's); the other ranks' copies go to their gather buffers
  for (uint32_t s0 = 0; s0 < nq; s0 += R[0]->max_chunk) {
    const uint32_t c = std::min(R[0]->max_chunk, nq - s0);
    for (size_t r = 0; r < R.size(); ++r) {
      dq[r] = d_queries + (size_t)s0 * D;
      if (r == 0) { oi[r] = d_out_idx + (size_t)s0 * out_k; od[r] = d_out_dist + (size_t)s0 * out_k; }
      else {
        CU(cudaSetDevice(R[r]->device));
        // sized for the padded batch, so that ensure_workspace (same buffers) does not move them afterwards
        CU(R[r]->out_idx.ensure(sizeof(uint32_t) * ((size_t)c + G) * out_k));
        CU(R[r]->out_dist.ensure(sizeof(float) * ((size_t)c + G) * out_k));
        oi[r] = R[r]->out_idx.as<uint32_t>(); od[r] = R[r]->out_dist.as<float>();
      }
    }
    if (int rc = sharded_chunk(coll, dq, c, p, light != 0, oi, od, (uint32_t)out_k)) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

int scann_b200_comm_unique_id(void* out_id128) {
  if (!out_id128) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(SCANN_B200_FAILED_PRECONDITION, "libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NC(api->GetUniqueId(&id));
  memcpy(out_id128, &id, sizeof id);
  return 0;
}

int scann_b200_comm_init(scann_b200_index* ix, int32_t rank, int32_t world, const void* id128) {
  if (!ix || !id128) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(SCANN_B200_INVALID_ARGUMENT, "bad rank %d / world %d", rank, world);
  if (ix->shard_world != world || ix->shard_rank != rank)
    return fail(SCANN_B200_INVALID_ARGUMENT, "index holds shard %d of %d, communicator rank is %d of %d", ix->shard_rank,
                ix->shard_world, rank, world);
  NcclApi* api = nccl_api();
  if (!api) return fail(SCANN_B200_FAILED_PRECONDITION, "libnccl.so.2 could not be loaded");
  std::lock_guard<std::mutex> lock(ix->mu);
  CU(cudaSetDevice(ix->device));
  if (ix->comm) { comm_destroy(ix->comm); ix->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ShardComm* c = new ShardComm();
  c->rank = rank; c->world = world;
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(SCANN_B200_INTERNAL, "ncclCommInitRank: %s", api->GetErrorString(r));
  }
  ix->comm = c;
  return 0;
}

int scann_b200_search_sharded_device(scann_b200_index* ix, const float* d_queries, uint32_t nq, int32_t final_nn,
                                     int32_t pre_nn, int32_t leaves, int32_t light, uint32_t* d_out_idx,
                                     float* d_out_dist, int32_t out_k) {
  std::vector<scann_b200_index*> R{ix};
  return sharded_search(R, false, d_queries, nq, final_nn, pre_nn, leaves, light, d_out_idx, d_out_dist, out_k);
}

int scann_b200_search_sharded_local(scann_b200_index* const* shards, int32_t world, const float* d_queries, uint32_t nq,
                                    int32_t final_nn, int32_t pre_nn, int32_t leaves, int32_t light,
                                    uint32_t* d_out_idx, float* d_out_dist, int32_t out_k) {
  if (!shards || world < 1) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  std::vector<scann_b200_index*> R(shards, shards + world);
  return sharded_search(R, true, d_queries, nq, final_nn, pre_nn, leaves, light, d_out_idx, d_out_dist, out_k);
}

}  // extern "C"
