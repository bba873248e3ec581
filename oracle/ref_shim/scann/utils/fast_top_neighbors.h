// shim (oracle/_ref build only) for utils/fast_top_neighbors.h.  The real FastTopNeighbors needs highway's vqsort and
// half of scann/utils; the LUT16 kernels only use it through epsilon() / AcquireMutator() / Mutator::Push*, so this
// stand-in records EVERY pushed (index, distance) pair with an epsilon that never tightens: what comes out is the full
// list of scores the reference kernel computes (lut16_avx2.inc:404-527), in the reference's own arithmetic.
#pragma once
#include "scann/utils/common.h"
#include "scann/utils/types.h"
namespace research_scann {
template <typename DistT, typename DatapointIndexT = DatapointIndex>
class FastTopNeighbors {
 public:
  explicit FastTopNeighbors(DistT epsilon = std::numeric_limits<DistT>::max()) : epsilon_(epsilon) {}
  DistT epsilon() const { return epsilon_; }
  class Mutator {
   public:
    void Init(FastTopNeighbors* p) { p_ = p; }
    bool Push(DatapointIndexT i, DistT d) { if (d < p_->epsilon_) p_->results.emplace_back(i, d); return false; }
    bool PushNoEpsilonCheck(DatapointIndexT i, DistT d) { p_->results.emplace_back(i, d); return false; }
    void GarbageCollect() {}
    DistT epsilon() const { return p_->epsilon_; }
   private:
    FastTopNeighbors* p_ = nullptr;
  };
  void AcquireMutator(Mutator* m) { m->Init(this); }
  std::vector<std::pair<DatapointIndexT, DistT>> results;
 private:
  DistT epsilon_;
};
}  // namespace research_scann
