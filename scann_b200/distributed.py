"""Sharded search over the GPUs of one box.

Two generations live here.  `ShardedIndex` is the one to use: the whole protocol (sliced tokenization, threshold
all-reduce, scan of the rank's own leaves, all-to-all of 16-byte records to the query's owner, merge, all-gather of
the k results) runs inside the C++ library on the index's stream, NCCL called from C++ (`csrc/sharded.cu`);
torch.distributed only carries the 128-byte NCCL id to the ranks.  `ShardedSearcher` below is the round-1 form
(three torch all-gathers of every rank's candidates, every rank merges every query); it is kept because the tests
state the merge rule on its records.

Round-1 form: the database shards by datapoint id (`id % world == rank`) inside every leaf; centres, AH
codebook, tokenization and LUT build are replicated, so no query-side exchange is needed.
Each rank produces its local over-retrieved pre-reorder candidates with exact distances
(`scann_b200_search_partial_device`), ONE all-gather exchanges them (records of
(id u32, tie-break key u64, AH score f32, exact distance f32)), and every rank merges to the
global result (`scann_b200_merge_partials_device`).  The merged result is bit-identical to the
single-GPU result because the tie-break key carries the slot of the datapoint in the
UNSHARDED index.

`torch.distributed` (NCCL on GPUs) is plumbing only.  `merge_partials_reference` is the
host-side statement of the merge rule; tests use it with the gloo backend on CPU.
"""
import ctypes as C

import numpy as np

from . import _lib

INVALID_ID = 0xFFFFFFFF
KEY_MAX = 0xFFFFFFFFFFFFFFFF


def nover_for(pre_nn, disjoint, overretrieve):
  """tree_ah_hybrid_residual.h:263-267 (NumNeighborsWithSpillingMultiplier)."""
  if disjoint:
    return int(pre_nn)
  return int(float(pre_nn) * float(np.float32(overretrieve)))


def shard_tokens(tokens, soar, rank, world):
  """Token array of one shard: datapoints of other ranks are marked absent (-1)."""
  t = np.array(tokens, dtype=np.int32, copy=True)
  mult = 2 if soar else 1
  ids = np.arange(t.shape[0]) // mult
  t[ids % world != rank] = -1
  return t


def f2ord(x):
  """Order-preserving u32 image of float32 (same mapping as the kernels)."""
  u = (np.asarray(x, dtype=np.float32) + np.float32(0.0)).view(np.uint32)
  return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def merge_partials_reference(ids, tie, exact, nover, npre, k, disjoint, dot_product=True):
  """Host statement of the merge rule for ONE query.

  ids/tie/exact: 1-D arrays over all ranks' records (padding: id == INVALID_ID).
  Returns (ids[k], distances[k]) padded with (0, NaN).
  """
  valid = ids != INVALID_ID
  ids, tie, exact = ids[valid], tie[valid], exact[valid]
  order = np.argsort(tie, kind="stable")[:nover]
  ids, tie, exact = ids[order], tie[order], exact[order]
  score_ord = (tie >> np.uint64(32)).astype(np.uint32)
  u = np.where(score_ord & 0x80000000, score_ord & 0x7FFFFFFF, ~score_ord).astype(np.uint32)
  score = u.view(np.float32)
  if disjoint:
    sel = np.arange(min(len(ids), npre))
    sel_ids, sel_exact = ids[sel], exact[sel]
  else:
    best = {}
    for i in range(len(ids)):
      dp = int(ids[i])
      if dp in best:
        a, e = best[dp]
        lo, hi = (a, score[i]) if a <= score[i] else (score[i], a)
        best[dp] = (np.float32(np.float32(0.5) * lo + np.float32(0.5) * hi), e)
      else:
        best[dp] = (score[i], exact[i])
    items = sorted(best.items(), key=lambda kv: (int(f2ord(kv[1][0])), kv[0]))[:npre]
    sel_ids = np.asarray([kv[0] for kv in items], dtype=np.uint32)
    sel_exact = np.asarray([kv[1][1] for kv in items], dtype=np.float32)
  keys = sorted(zip(f2ord(sel_exact).tolist(), sel_ids.tolist(), sel_exact.tolist()))[:k]
  out_i = np.zeros(k, np.uint32)
  out_d = np.full(k, np.nan, np.float32)
  for j, (_, dp, ex) in enumerate(keys):
    out_i[j] = dp
    out_d[j] = -ex if dot_product else ex
  return out_i, out_d


SHARD_BY_ID, SHARD_BY_LEAF = 0, 1


def sampled_threshold_reference(rank_scores, nover):
  """Host statement of the sampled global threshold (csrc/finalize.cu sample_scores_kernel / sample_threshold_kernel).

  rank_scores: one ascending array of score words (f2ord of the AH score) per rank = that rank's local candidate list.
  Every rank samples positions 15, 31, ... of its list (at most nover entries); the ceil(nover / 16)-th smallest sample
  over all ranks bounds the global nover-th best score from above.  Returns that bound (0xFFFFFFFF: no pruning)."""
  need = (nover + 15) // 16
  samples = []
  for s in rank_scores:
    s = np.asarray(s, dtype=np.uint32)[:nover]
    samples.extend(s[15::16].tolist())
  samples.sort()
  return samples[need - 1] if len(samples) >= need else 0xFFFFFFFF


def owner_of_query(q, nq, world):
  """Rank that tokenizes and merges query q of a batch of nq (contiguous slices of ceil(nq / world) queries)."""
  return q // (-(-nq // world))


def leaf_owner(leaf, world):
  """Leaf sharding: the rank that stores (and scans) a leaf."""
  return leaf % world


class ShardedIndex:
  """One rank (one process, one GPU) of a database-sharded tree-AH searcher.

  `group` is a torch.distributed process group; it is used ONCE, to broadcast the NCCL unique id that rank 0 gets
  from the library.  All exchanges of a search are NCCL calls issued by the library itself.
  """

  def __init__(self, arrays, leaves_to_search, pre_reorder_nn, final_nn, rank, world, device, group=None,
               shard_mode=SHARD_BY_LEAF):
    import torch
    import torch.distributed as dist
    self.torch = torch
    self.rank, self.world = rank, world
    self.dev = torch.device("cuda", device)
    self.k = final_nn
    self.index = _lib.NativeIndex(arrays, leaves_to_search, pre_reorder_nn, final_nn, device=device,
                                  shard_rank=rank, shard_world=world, shard_mode=shard_mode)
    if world > 1:
      L = _lib.lib()
      uid = torch.zeros(128, dtype=torch.uint8)
      if rank == 0:
        buf = (C.c_uint8 * 128)()
        _lib.check(L.scann_b200_comm_unique_id(buf))
        uid = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
      backend = dist.get_backend(group)
      t = uid.to(self.dev) if backend == "nccl" else uid
      dist.broadcast(t, src=0, group=group)
      raw = bytes(t.cpu().numpy().tobytes())
      _lib.check(L.scann_b200_comm_init(self.index._h, rank, world, raw))

  def search_batched_device(self, d_q, d_idx, d_dist, light=False, leaves=-1):
    """d_q [nq, D] f32 cuda tensor (the same on every rank); d_idx [nq, k] int32, d_dist [nq, k] f32 outputs, filled
    on every rank.  Returns this rank's stats (CUDA-event times incl. ms_exchange / ms_merge)."""
    vp = C.c_void_p
    # the library's stream does not wait for torch's: the caller's tensors must be complete
    _lib.check(_lib.lib().scann_b200_search_sharded_device(self.index._h, vp(d_q.data_ptr()), d_q.shape[0], -1, -1,
                                                           leaves, 1 if light else 0, vp(d_idx.data_ptr()),
                                                           vp(d_dist.data_ptr()), d_idx.shape[1]))
    return self.index.stats()

  def search_batched(self, q, light=False):
    t = self.torch
    d_q = t.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(self.dev)
    d_idx = t.empty((q.shape[0], self.k), dtype=t.int32, device=self.dev)
    d_dist = t.empty((q.shape[0], self.k), dtype=t.float32, device=self.dev)
    t.cuda.synchronize()
    self.search_batched_device(d_q, d_idx, d_dist, light=light)
    return d_idx.cpu().numpy().view(np.uint32), d_dist.cpu().numpy()


def search_sharded_local(shards, q, k, light=False, leaves=-1):
  """All `world` shards (NativeIndex objects with shard_rank 0..world-1) in this process on one device: the library
  runs the same protocol with device copies instead of NCCL.  Test vehicle; returns (ids, distances, stats of rank 0)."""
  import torch
  dev = torch.device("cuda", shards[0].device)
  d_q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(dev)
  d_idx = torch.empty((q.shape[0], k), dtype=torch.int32, device=dev)
  d_dist = torch.empty((q.shape[0], k), dtype=torch.float32, device=dev)
  torch.cuda.synchronize()
  hs = (C.c_void_p * len(shards))(*[s._h for s in shards])
  vp = C.c_void_p
  _lib.check(_lib.lib().scann_b200_search_sharded_local(hs, len(shards), vp(d_q.data_ptr()), q.shape[0], -1, -1, leaves,
                                                        1 if light else 0, vp(d_idx.data_ptr()), vp(d_dist.data_ptr()), k))
  return d_idx.cpu().numpy().view(np.uint32), d_dist.cpu().numpy(), [s.stats() for s in shards]


class ShardedSearcher:
  """One rank of a sharded searcher.  `group` is a torch.distributed process group (NCCL)."""

  def __init__(self, arrays, leaves_to_search, pre_reorder_nn, final_nn, rank, world, device, group=None):
    import torch
    self.torch = torch
    self.rank, self.world, self.group = rank, world, group
    self.dev = torch.device("cuda", device)
    self.index = _lib.NativeIndex(arrays, leaves_to_search, pre_reorder_nn, final_nn, device=device,
                                  shard_rank=rank, shard_world=world)
    self.pre, self.k = pre_reorder_nn, final_nn
    self.ncand = nover_for(pre_reorder_nn, not arrays.soar, arrays.overretrieve)
    self._bufs = None

  def _buffers(self, nq):
    t = self.torch
    if self._bufs is None or self._bufs[0].shape[0] != nq:
      n, w, dev = self.ncand, self.world, self.dev
      local = (t.empty((nq, n), dtype=t.int32, device=dev), t.empty((nq, n), dtype=t.int64, device=dev),
               t.empty((nq, n), dtype=t.float32, device=dev), t.empty((nq, n), dtype=t.float32, device=dev))
      # concatenated along dim 0 = [world][nq][n] in memory (accepted by both NCCL and gloo)
      gathered = tuple(t.empty((w * nq, n), dtype=x.dtype, device=dev) for x in (local[0], local[1], local[3]))
      self._bufs = local + gathered
    return self._bufs

  def search_batched_device(self, d_q, d_idx, d_dist):
    """d_q [nq, D] f32 cuda tensor; d_idx [nq, k] int32, d_dist [nq, k] f32 outputs. Returns stats."""
    import torch.distributed as dist
    t = self.torch
    nq = d_q.shape[0]
    ids, tie, ah, ex, g_ids, g_tie, g_ex = self._buffers(nq)
    L = _lib.lib()
    vp = C.c_void_p
    h = self.index._h
    _lib.check(L.scann_b200_search_partial_device(h, vp(d_q.data_ptr()), nq, -1, -1, vp(ids.data_ptr()),
                                                  vp(tie.data_ptr()), vp(ah.data_ptr()), vp(ex.data_ptr()),
                                                  self.ncand))
    st = self.index.stats()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    dist.all_gather_into_tensor(g_ids, ids, group=self.group)
    dist.all_gather_into_tensor(g_tie, tie, group=self.group)
    dist.all_gather_into_tensor(g_ex, ex, group=self.group)
    e1.record()
    t.cuda.synchronize()
    import time
    t_merge = time.perf_counter()
    _lib.check(L.scann_b200_merge_partials_device(h, nq, self.world, self.ncand, vp(g_ids.data_ptr()),
                                                  vp(g_tie.data_ptr()), None, vp(g_ex.data_ptr()), -1, -1,
                                                  vp(d_idx.data_ptr()), vp(d_dist.data_ptr()), d_idx.shape[1]))
    st["ms_merge"] = (time.perf_counter() - t_merge) * 1e3  # the call synchronises its stream
    st["ms_allgather"] = e0.elapsed_time(e1)
    st["allgather_bytes_per_rank"] = int(nq * self.ncand * 16)
    return st

  def search_batched(self, q):
    """Host buffers in / out (numpy)."""
    t = self.torch
    d_q = t.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(self.dev)
    d_idx = t.empty((q.shape[0], self.k), dtype=t.int32, device=self.dev)
    d_dist = t.empty((q.shape[0], self.k), dtype=t.float32, device=self.dev)
    self.search_batched_device(d_q, d_idx, d_dist)
    return d_idx.cpu().numpy().view(np.uint32), d_dist.cpu().numpy()


def merge_topk_reference(ids, dists, k, dot_product=True):
  """numpy restatement of scann_b200_merge_topk_device: ids/dists [world][nq][k_in] (API-signed distances, NaN =
  padding) -> global top-k by (internal distance, id)."""
  world, nq, k_in = ids.shape
  out_i = np.zeros((nq, k), dtype=np.uint32)
  out_d = np.full((nq, k), np.nan, dtype=np.float32)
  sign = np.float32(-1.0 if dot_product else 1.0)
  for q in range(nq):
    i = ids[:, q, :].reshape(-1).astype(np.uint32)
    d = dists[:, q, :].reshape(-1).astype(np.float32)
    ok = ~np.isnan(d)
    i, d = i[ok], d[ok]
    order = np.lexsort((i, sign * d))[:k]
    out_i[q, :len(order)] = i[order]
    out_d[q, :len(order)] = d[order]
  return out_i, out_d


class ShardedBruteForce:
  """One rank of a row-sharded bf16 brute-force searcher (BASELINE.json configs[2] at N > 1 GPUs): local top-k over
  this rank's rows, one all-gather of (id, distance), merge on the device.  Equal to the unsharded result."""

  def __init__(self, arrays, final_nn, rank, world, device, group=None):
    import torch
    self.torch = torch
    self.rank, self.world, self.group = rank, world, group
    self.dev = torch.device("cuda", device)
    self.index = _lib.NativeIndex(arrays, 1, final_nn, final_nn, device=device, shard_rank=rank, shard_world=world)
    self.k = final_nn
    self._bufs = None

  def _buffers(self, nq):
    t = self.torch
    if self._bufs is None or self._bufs[0].shape[0] != nq:
      k, w, dev = self.k, self.world, self.dev
      self._bufs = (t.empty((nq, k), dtype=t.int32, device=dev), t.empty((nq, k), dtype=t.float32, device=dev),
                    t.empty((w * nq, k), dtype=t.int32, device=dev), t.empty((w * nq, k), dtype=t.float32, device=dev))
    return self._bufs

  def search_batched_device(self, d_q, d_idx, d_dist):
    import torch.distributed as dist
    t = self.torch
    nq = d_q.shape[0]
    ids, ds, g_ids, g_ds = self._buffers(nq)
    L = _lib.lib()
    vp = C.c_void_p
    h = self.index._h
    _lib.check(L.scann_b200_search_batched_device(h, vp(d_q.data_ptr()), nq, self.k, -1, -1, vp(ids.data_ptr()),
                                                  vp(ds.data_ptr()), self.k))
    st = self.index.stats()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    if self.world > 1:
      dist.all_gather_into_tensor(g_ids, ids, group=self.group)
      dist.all_gather_into_tensor(g_ds, ds, group=self.group)
    else:
      g_ids.copy_(ids)
      g_ds.copy_(ds)
    e1.record()
    t.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    _lib.check(L.scann_b200_merge_topk_device(h, nq, self.world, self.k, vp(g_ids.data_ptr()), vp(g_ds.data_ptr()),
                                              self.k, vp(d_idx.data_ptr()), vp(d_dist.data_ptr()), d_idx.shape[1]))
    st["ms_merge"] = (time.perf_counter() - t0) * 1e3
    st["ms_allgather"] = e0.elapsed_time(e1)
    st["allgather_bytes_per_rank"] = int(nq * self.k * 8)
    return st

  def search_batched(self, q):
    t = self.torch
    d_q = t.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(self.dev)
    d_idx = t.empty((q.shape[0], self.k), dtype=t.int32, device=self.dev)
    d_dist = t.empty((q.shape[0], self.k), dtype=t.float32, device=self.dev)
    self.search_batched_device(d_q, d_idx, d_dist)
    return d_idx.cpu().numpy().view(np.uint32), d_dist.cpu().numpy()
