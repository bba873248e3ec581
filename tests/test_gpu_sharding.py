"""GPU: sharded search == unsharded search, bit for bit.

Shards are emulated on ONE device (two index handles, records concatenated the way an
all-gather would lay them out); the NCCL exchange itself is exercised by `bench.py --gpus N`.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import get_case

pytestmark = pytest.mark.gpu


def run_sharded(case, world):
  import torch
  from scann_b200 import _lib, distributed as sd
  a = case.arrays
  q = np.ascontiguousarray(case.q, dtype=np.float32)
  nq = q.shape[0]
  ncand = sd.nover_for(case.pre, not a.soar, a.overretrieve)
  dev = torch.device("cuda", 0)
  d_q = torch.from_numpy(q).to(dev)
  g_ids = torch.empty((world, nq, ncand), dtype=torch.int32, device=dev)
  g_tie = torch.empty((world, nq, ncand), dtype=torch.int64, device=dev)
  g_ah = torch.empty((world, nq, ncand), dtype=torch.float32, device=dev)
  g_ex = torch.empty((world, nq, ncand), dtype=torch.float32, device=dev)
  torch.cuda.synchronize()
  L = _lib.lib()
  vp = C.c_void_p
  shards = []
  for r in range(world):
    ix = _lib.NativeIndex(a, case.probe, case.pre, case.k, device=0, shard_rank=r, shard_world=world)
    shards.append(ix)
    _lib.check(L.scann_b200_search_partial_device(ix._h, vp(d_q.data_ptr()), nq, -1, -1, vp(g_ids[r].data_ptr()),
                                                  vp(g_tie[r].data_ptr()), vp(g_ah[r].data_ptr()),
                                                  vp(g_ex[r].data_ptr()), ncand))
  d_idx = torch.empty((nq, case.k), dtype=torch.int32, device=dev)
  d_dist = torch.empty((nq, case.k), dtype=torch.float32, device=dev)
  _lib.check(L.scann_b200_merge_partials_device(shards[0]._h, nq, world, ncand, vp(g_ids.data_ptr()),
                                                vp(g_tie.data_ptr()), None, vp(g_ex.data_ptr()), -1, -1,
                                                vp(d_idx.data_ptr()), vp(d_dist.data_ptr()), case.k))
  torch.cuda.synchronize()
  # host statement of the merge rule on the same records
  ids_h = g_ids.cpu().numpy().view(np.uint32)
  tie_h = g_tie.cpu().numpy().view(np.uint64)
  ex_h = g_ex.cpu().numpy()
  ref_i = np.zeros((nq, case.k), np.uint32)
  for i in range(min(nq, 8)):
    ref_i[i], _ = sd.merge_partials_reference(ids_h[:, i].reshape(-1), tie_h[:, i].reshape(-1), ex_h[:, i].reshape(-1),
                                              ncand, case.pre, case.k, not a.soar)
  for ix in shards:
    ix.close()
  return d_idx.cpu().numpy().view(np.uint32), d_dist.cpu().numpy(), ref_i


@pytest.mark.parametrize("kw,world", [(dict(), 2), (dict(), 3), (dict(soar=1.5), 2), (dict(soar=1.5), 4),
                                      (dict(n=3000, leaves=300, probe=40, pre=150), 8)])
def test_sharded_equals_unsharded(kw, world):
  c = get_case(**kw)
  i1, d1 = c.native.search_batched(c.q)
  i2, d2, ref_i = run_sharded(c, world)
  np.testing.assert_array_equal(i1, i2)
  np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))
  np.testing.assert_array_equal(ref_i[:8], i2[:8])
  i0, d0 = c.oracle.search_batched(c.q)
  np.testing.assert_array_equal(i0, i2)


@pytest.mark.parametrize("kw,world", [(dict(), 2), (dict(soar=1.5), 3)])
def test_sharded_bf16_reordering_equals_unsharded(kw, world):
  """Shards hold the bf16 rows of their own datapoints (id mod world) for the reordering."""
  import copy
  import types
  from scann_b200 import _lib, index_build
  c = get_case(**kw)
  a = copy.copy(c.arrays)
  a.bf16_dataset = index_build.bfloat16_quantize(c.db)
  a.dataset = None
  full = _lib.NativeIndex(a, c.probe, c.pre, c.k)
  i1, d1 = full.search_batched(c.q)
  case = types.SimpleNamespace(arrays=a, q=c.q, probe=c.probe, pre=c.pre, k=c.k)
  i2, d2, _ = run_sharded(case, world)
  np.testing.assert_array_equal(i1, i2)
  np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))


@pytest.mark.parametrize("kw,world", [(dict(), 2), (dict(distance="squared_l2", d=64, leaves=50, n=10000), 3)])
def test_sharded_int8_reordering_equals_unsharded(kw, world):
  """Shards hold the int8 rows of their own datapoints; multipliers and row norms are replicated."""
  import copy
  import types
  from scann_b200 import _lib, index_build
  c = get_case(**kw)
  a = copy.copy(c.arrays)
  a.int8_dataset, a.int8_multipliers = index_build.int8_quantize(c.db)
  if a.distance == "squared_l2":
    a.dp_norms = index_build.squared_l2_norms(c.db)
  a.dataset = None
  full = _lib.NativeIndex(a, c.probe, c.pre, c.k)
  i1, d1 = full.search_batched(c.q)
  case = types.SimpleNamespace(arrays=a, q=c.q, probe=c.probe, pre=c.pre, k=c.k)
  i2, d2, _ = run_sharded(case, world)
  np.testing.assert_array_equal(i1, i2)
  np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))


# ---- the C++ protocol (csrc/sharded.cu): sliced tokenization, threshold all-reduce, owner-side merge -----------------

def _local_world(a, probe, pre, k, world, mode):
  from scann_b200 import _lib
  return [_lib.NativeIndex(a, probe, pre, k, device=0, shard_rank=r, shard_world=world, shard_mode=mode)
          for r in range(world)]


@pytest.mark.parametrize("mode", [0, 1], ids=["by_id", "by_leaf"])
@pytest.mark.parametrize("kw,world", [(dict(), 2), (dict(), 3), (dict(soar=1.5), 2), (dict(soar=1.5), 4),
                                      (dict(soar=1.5), 8), (dict(n=3000, leaves=300, probe=40, pre=150), 8),
                                      (dict(distance="squared_l2", d=64, leaves=50, n=10000), 3)])
def test_sharded_protocol_equals_unsharded(kw, world, mode):
  """world shards on one device, exchanges as device copies: ids and distance bits of the unsharded searcher."""
  from scann_b200 import distributed as sd
  c = get_case(**kw)
  i1, d1 = c.native.search_batched(c.q)
  shards = _local_world(c.arrays, c.probe, c.pre, c.k, world, mode)
  i2, d2, st = sd.search_sharded_local(shards, c.q, c.k)
  np.testing.assert_array_equal(i1, i2)
  np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))
  i0, _ = c.oracle.search_batched(c.q)
  np.testing.assert_array_equal(i0, i2)
  assert all(s["exchange_bytes"] > 0 and s["kernel_launches"] >= 8 for s in st)
  if mode == 1:  # whole leaves are dealt out: the ranks' scan work adds up to the unsharded scan work
    c.native.search_batched(c.q)
    assert sum(s["scan_bytes_alg"] for s in st) == c.native.stats()["scan_bytes_alg"]
  for s in shards:
    s.close()


@pytest.mark.parametrize("mode", [0, 1], ids=["by_id", "by_leaf"])
def test_sharded_protocol_ragged_batches_and_overflow(mode, monkeypatch):
  """nq not divisible by the world size, fewer queries than ranks, forced candidate-buffer overflows, two scan phases."""
  from scann_b200 import distributed as sd
  c = get_case(soar=1.5)
  shards = _local_world(c.arrays, c.probe, c.pre, c.k, 4, mode)
  for nq in (1, 3, 37):
    i1, d1 = c.native.search_batched(c.q[:nq])
    i2, d2, _ = sd.search_sharded_local(shards, c.q[:nq], c.k)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))
  i1, d1 = c.native.search_batched(c.q)
  monkeypatch.setenv("SCANN_B200_CAND_CAP", "256")
  monkeypatch.setenv("SCANN_B200_TWO_PHASE", "1")
  i2, d2, st = sd.search_sharded_local(shards, c.q, c.k)
  np.testing.assert_array_equal(i1, i2)
  np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))
  for s in shards:
    s.close()


def test_sharded_protocol_reordering_variants_by_leaf():
  """Leaf-sharded ranks hold the bf16 / int8 rows of the datapoints stored in their leaves."""
  import copy
  from scann_b200 import _lib, index_build, distributed as sd
  c = get_case(soar=1.5)
  for kind in ("bf16", "int8"):
    a = copy.copy(c.arrays)
    if kind == "bf16":
      a.bf16_dataset = index_build.bfloat16_quantize(c.db)
    else:
      a.int8_dataset, a.int8_multipliers = index_build.int8_quantize(c.db)
    a.dataset = None
    full = _lib.NativeIndex(a, c.probe, c.pre, c.k)
    i1, d1 = full.search_batched(c.q)
    shards = _local_world(a, c.probe, c.pre, c.k, 3, 1)
    i2, d2, _ = sd.search_sharded_local(shards, c.q, c.k)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))
    for s in shards:
      s.close()
    full.close()


@pytest.mark.parametrize("mode", [0, 1], ids=["by_id", "by_leaf"])
def test_sharded_light_protocol(mode):
  """Light protocol (local top-k + one all-gather): not bit-identical by design; every returned neighbour carries its
  exact distance, rows are sorted, no id repeats, and the result is at least as good as the parity mode's."""
  from scann_b200 import distributed as sd
  c = get_case(soar=1.5)
  shards = _local_world(c.arrays, c.probe, c.pre, c.k, 4, mode)
  i1, d1 = c.native.search_batched(c.q)
  i2, d2, _ = sd.search_sharded_local(shards, c.q, c.k, light=True)
  for r in range(c.q.shape[0]):
    assert len(set(i2[r].tolist())) == c.k
    exact = c.db[i2[r]].astype(np.float64) @ c.q[r].astype(np.float64)
    np.testing.assert_allclose(d2[r], exact, rtol=1e-5, atol=1e-6)
    assert np.all(np.diff(d2[r]) <= 0)          # dot product: descending similarity
    assert d2[r][-1] >= d1[r][-1] - 1e-6         # k-th neighbour no worse than the parity result's
  for s in shards:
    s.close()
