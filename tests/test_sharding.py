"""N > 1 path on CPU: two gloo ranks, each holding one datapoint-id shard.

What is under test is the sharding contract of SURVEY.md section 8e (shard by id inside every
leaf, one all-gather of (id, tie-break key, exact distance) records, merge rule) and the
torch.distributed plumbing; per-shard candidates come from the CPU oracle here, from the
CUDA kernels in tests/test_gpu_sharding.py.
"""
import copy
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import load_golden, np_slots
from scann_b200 import distributed as sd


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def shard_partials(arrays, q, rank, world, probe, pre, k):
  """(ids, tie, exact) [nq, ncand] for one shard, computed with the CPU oracle."""
  a = copy.copy(arrays)
  a.tokens = sd.shard_tokens(arrays.tokens, arrays.soar, rank, world)
  oi = oracle.OracleIndex(a, probe, pre, k)
  ncand = sd.nover_for(pre, not arrays.soar, arrays.overretrieve)
  c = oi.candidates(q, cap=ncand)
  L = arrays.centers.shape[0]
  full = np_slots(arrays.tokens, L, arrays.soar)
  base = np.concatenate([[0], np.cumsum([len(x) for x in full])])
  nq = q.shape[0]
  ids = np.full((nq, ncand), sd.INVALID_ID, np.uint32)
  tie = np.full((nq, ncand), sd.KEY_MAX, np.uint64)
  exact = np.full((nq, ncand), np.inf, np.float32)
  for i in range(nq):
    n = int(c["count"][i])
    dps, leaves = c["dp"][i, :n], c["leaf"][i, :n]
    gslot = np.asarray([base[l] + np.searchsorted(full[l], dp) for l, dp in zip(leaves.tolist(), dps.tolist())],
                       dtype=np.uint64)
    ids[i, :n] = dps
    tie[i, :n] = (sd.f2ord(c["score"][i, :n]).astype(np.uint64) << np.uint64(32)) | gslot
    exact[i, :n] = oi.exact_distances(q[i], dps)
  return ids, tie, exact


def _worker(rank, world, port, name, out_dir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    a, z = load_golden(name)
    probe, pre, k = int(z["probe"]), int(z["pre"]), int(z["k"])
    q = z["queries"]
    ids, tie, exact = shard_partials(a, q, rank, world, probe, pre, k)
    gathered = []
    for arr, dt in ((ids.view(np.int32), torch.int32), (tie.view(np.int64), torch.int64), (exact, torch.float32)):
      t = torch.from_numpy(np.ascontiguousarray(arr))
      g = torch.empty((world * t.shape[0], t.shape[1]), dtype=dt)
      dist.all_gather_into_tensor(g, t)
      gathered.append(g.numpy().reshape((world,) + tuple(t.shape)))
    g_ids = gathered[0].view(np.uint32)
    g_tie = gathered[1].view(np.uint64)
    g_ex = gathered[2]
    nover = sd.nover_for(pre, not a.soar, a.overretrieve)
    out_i = np.zeros((q.shape[0], k), np.uint32)
    out_d = np.zeros((q.shape[0], k), np.float32)
    for i in range(q.shape[0]):
      out_i[i], out_d[i] = sd.merge_partials_reference(g_ids[:, i].reshape(-1), g_tie[:, i].reshape(-1),
                                                       g_ex[:, i].reshape(-1), nover, pre, k, not a.soar)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=out_i, dist=out_d)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["dot_b16", "dot_soar_b25"])
def test_two_gloo_ranks_reproduce_the_unsharded_result(name, tmp_path):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
  a, z = load_golden(name)
  for r in range(world):
    got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
    np.testing.assert_array_equal(got["idx"], z["exp_idx"])
    np.testing.assert_array_equal(got["dist"].view(np.uint32), z["exp_dist"].view(np.uint32))


def test_shard_tokens_partitions_every_leaf():
  a, z = load_golden("dot_soar_b25")
  full = np_slots(a.tokens, a.centers.shape[0], True)
  parts = [np_slots(sd.shard_tokens(a.tokens, True, r, 3), a.centers.shape[0], True) for r in range(3)]
  for leaf, dps in enumerate(full):
    merged = np.sort(np.concatenate([p[leaf] for p in parts]))
    np.testing.assert_array_equal(merged, dps)
    for r, p in enumerate(parts):
      assert (p[leaf] % 3 == r).all()


# ---- row-sharded bf16 brute force (BASELINE.json configs[2] at N > 1) --------------------------------------
def _bf_worker(rank, world, port, out_dir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from scann_b200 import index_build
    rng = np.random.default_rng(31)
    n, d, nq, k = 3001, 24, 37, 10
    db = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    bits = index_build.bfloat16_quantize(db)
    per = -(-n // world)                                    # the split of scann_b200_index_create
    r0, r1 = min(rank * per, n), min((rank + 1) * per, n)
    li, ld = oracle.bruteforce_bf16(np.ascontiguousarray(bits[r0:r1]), q, k, threads=1)
    li = (li + np.uint32(r0)).astype(np.uint32)             # global ids
    g = []
    for arr, dt in ((li.view(np.int32), torch.int32), (ld, torch.float32)):
      t = torch.from_numpy(np.ascontiguousarray(arr))
      o = torch.empty((world * nq, k), dtype=dt)
      dist.all_gather_into_tensor(o, t)
      g.append(o.numpy().reshape(world, nq, k))
    mi, md = sd.merge_topk_reference(g[0].view(np.uint32), g[1], k)
    fi, fd = oracle.bruteforce_bf16(bits, q, k, threads=1)
    np.savez(os.path.join(out_dir, f"bf{rank}.npz"), idx=mi, dist=md, full_idx=fi, full_dist=fd)
  finally:
    dist.destroy_process_group()


def test_two_gloo_ranks_row_sharded_bruteforce(tmp_path):
  world = 2
  mp.spawn(_bf_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
  for r in range(world):
    got = np.load(os.path.join(str(tmp_path), f"bf{r}.npz"))
    np.testing.assert_array_equal(got["idx"], got["full_idx"])
    np.testing.assert_array_equal(got["dist"].view(np.uint32), got["full_dist"].view(np.uint32))


# ---- host logic of the C++ sharded protocol (csrc/sharded.cu), no GPU needed ----------------------------------------------

@pytest.mark.parametrize("world,nover", [(2, 400), (8, 400), (3, 100), (8, 17), (4, 16), (5, 1)])
def test_sampled_threshold_is_a_valid_and_tight_bound(world, nover):
  """The ceil(N'/16)-th smallest of the ranks' every-16th scores is >= the global N'-th best score (nothing needed is
  dropped) and at most N' + 15 world candidates (+ ties) lie at or below it (little is sent in vain)."""
  rng = np.random.default_rng(world * 1000 + nover)
  for trial in range(50):
    sizes = rng.integers(0, 3 * nover, world)
    if trial % 7 == 0:
      sizes[rng.integers(0, world)] = 4 * nover      # everything on one rank
    if trial % 11 == 0:
      sizes[:] = rng.integers(0, 8, world)           # fewer than N' candidates in total
    lists = [np.sort(rng.integers(0, 1 << (10 if trial % 3 else 28), int(n)).astype(np.uint32)) for n in sizes]
    t = sd.sampled_threshold_reference(lists, nover)
    allv = np.sort(np.concatenate(lists)) if sum(sizes) else np.empty(0, np.uint32)
    if len(allv) >= nover and t != 0xFFFFFFFF:
      assert allv[nover - 1] <= t                    # the global N'-th best survives
      kept = sum(int(np.searchsorted(l[:nover], t, side="right")) for l in lists)
      ties = int((allv == t).sum())
      assert kept <= nover + 15 * world + ties
    if t == 0xFFFFFFFF:
      # no pruning only when the samples cannot prove N' candidates exist
      assert sum(len(l[:nover]) // 16 for l in lists) < (nover + 15) // 16


def test_query_and_leaf_ownership_cover_everything_once():
  for world in (1, 2, 3, 8):
    for nq in (1, 2, 7, 64, 10000):
      owners = [sd.owner_of_query(q, nq, world) for q in range(nq)]
      assert min(owners) >= 0 and max(owners) < world and owners == sorted(owners)
      sl = -(-nq // world)
      assert all(owners.count(r) <= sl for r in range(world))
    assert sorted(set(sd.leaf_owner(l, world) for l in range(100))) == list(range(min(world, 100)))


def test_sharded_protocol_symbols_and_nccl_id():
  """The C ABI exports the sharded entry points; the NCCL id call works without a GPU (it only loads libnccl)."""
  import ctypes as C
  from scann_b200 import _lib
  L = _lib.lib()
  for name in ("scann_b200_comm_unique_id", "scann_b200_comm_init", "scann_b200_search_sharded_device",
               "scann_b200_search_sharded_local"):
    assert hasattr(L, name)
  buf = (C.c_uint8 * 128)()
  rc = L.scann_b200_comm_unique_id(buf)
  assert rc in (0, 9)                                # 9: libnccl.so.2 not loadable on this host
  if rc == 0:
    assert any(buf)
  assert L.scann_b200_comm_init(None, 0, 1, None) == 3
