// index_internal.h -- host-side state of one searcher handle, shared by index.cu (C ABI, single-GPU driver) and
// sharded.cu (multi-GPU driver).  Not part of the public boundary (include/scann_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/scann_b200.h"
#include "kernels.h"

namespace sbi {

int fail(int code, const char* fmt, ...);

#define CU(expr)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::sbi::fail(SCANN_B200_INTERNAL, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), \
                         __FILE__, __LINE__, cudaGetErrorString(_e));                         \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool own = true;  // false: an alias of another handle's buffer (search lanes share the index arrays)
  ~DevBuf() { if (p && own) cudaFree(p); }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinnedBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMallocHost(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum { EV_START, EV_TOK, EV_LUT, EV_PILOT, EV_WORK, EV_SCAN, EV_COMPACT, EV_FIN, EV_COUNT,
       EV2_WORK = EV_COUNT, EV2_SCAN, EV2_COMPACT,
       EV_C0, EV_C1, EV_C2, EV_C3, EV_C4, EV_C5, EV_C6, EV_C7, EV_C8, EV_C9, EV_M0, EV_M1, EV_S0, EV_END, EV_ALL };

struct Params { uint32_t k, npre, nover, P; };

struct ShardComm;  // sharded.cu: NCCL communicator of this rank (one process per GPU)

}  // namespace sbi

struct scann_b200_index {
  sb::DevIndex dev{};
  scann_b200_index_desc desc{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::vector<uint32_t> h_leaf_size;
  // persistent device arrays
  sbi::DevBuf i8_inv, i8_norm, tok_cmax, pair_pos;
  sbi::DevBuf cen_i8, cen_inv, cen_sqn, cen_sqn2;  // int8 query tokenization: the fixed-point centres (index.cu)
  sbi::DevBuf centers, cnorm, codebook, block_dims, block_off, leaf_size, leaf_goff, leaf_ntiles, leaf_gpt,
      codes, slot_dp, slot_tie, dataset, dp_row, tok_b;
  // workspace
  sbi::DevBuf lut_e4m3;  // tensor-core scan: the e4m3 (hi nibble x 16, lo nibble) planes of the batch's LUTs
  sbi::DevBuf tok_a, q, dist, leaves, bias, lut, mult, inv, pilot_end, buf, cnt, tau, ovf, leaf_cnt, leaf_eoff,
      leaf_cur, item_off, item_leaf, entry_q, entry_bias, counters, stats, out_idx, out_dist;
  sbi::PinnedBuf h_q, h_idx, h_dist, h_counters;
  cudaEvent_t ev[sbi::EV_ALL] = {};
  scann_b200_stats last{};
  scann_b200_stats last_any{};  // root handle: stats of the most recently finished call on any lane
  bool last_any_valid = false;
  uint32_t max_chunk = 16384;
  // brute-force (bf16) searcher: database rows as bf16 with a 16-byte aligned pitch
  bool brute = false;
  bool bf_f32 = false;  // float brute force: f32 rows in `dataset`, concatenated hi/lo bf16 operand in `bf_db`
  uint32_t bf_dpitch = 0;
  uint32_t bf_row0 = 0;  // first database row of this shard (row-sharded brute force)
  uint32_t avg_leaf_slots = 0;  // mean padded slots per leaf (scan phase heuristic)
  uint32_t nonempty_leaves = 0;
  float bf_max_row_norm = 0.f;  // >= max_i ||x_i|| (error bound of the brute-force pre-filter)
  sbi::DevBuf bf_db, bf_a, bf_flags;
  sbi::DevBuf bf_xnorm, bf_qaug;  // squared-L2 float brute force: row norms (reference arithmetic), augmented queries
  bool bf_l2 = false;
  // Search lanes: ScannInterface::SearchBatched may be called concurrently on one searcher (scann_ops/cc/scann.cc:478-501
  // runs batches on a thread pool).  The handle itself is lane 0; further lanes are shallow clones -- the same index
  // arrays on the device, their own stream, events and workspace -- created on demand (at most kMaxLanes), so that
  // concurrent callers overlap one batch's host <-> device copies with another batch's kernels.
  static constexpr int kMaxLanes = 4;
  std::mutex pool_mu;
  std::condition_variable pool_cv;
  std::vector<scann_b200_index*> lanes;  // clones (lane 0 = this)
  std::vector<char> lane_busy;           // [1 + lanes.size()]
  scann_b200_index* parent = nullptr;    // clones: the handle they belong to
  // sharded search (sharded.cu)
  int shard_rank = 0, shard_world = 1, shard_mode = 0;
  sbi::ShardComm* comm = nullptr;
  sbi::DevBuf sh_send, sh_recv, sh_idx, sh_dist, sh_samples, sh_limit;
};

namespace sbi {
int resolve(const scann_b200_index* ix, int final_nn, int pre_nn, int leaves, Params* p);
uint32_t pick_cap(uint32_t nover);
int ensure_workspace(scann_b200_index* ix, uint32_t nq, const Params& p, uint32_t out_k, uint32_t cap);
void fill_scan_work(scann_b200_index* ix, uint32_t nq, const Params& p, uint32_t cap, sb::ScanWork* w);
void comm_destroy(ShardComm* c);

// RAII: a free search lane of `root` (waits when all kMaxLanes are busy).  lane->mu is held for the lifetime.
class LaneGuard {
 public:
  explicit LaneGuard(scann_b200_index* root);
  ~LaneGuard();
  scann_b200_index* ix = nullptr;  // nullptr: a clone could not be created (last error set)
 private:
  scann_b200_index* root_;
  int slot_ = -1;
};
}  // namespace sbi
