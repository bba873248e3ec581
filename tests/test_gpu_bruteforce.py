"""GPU: bf16 brute force (tcgen05 GEMM + fused pre-filter + exact re-scoring) against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make(n, d, nq, seed):
  from scann_b200 import index_build
  rng = np.random.default_rng(seed)
  db = rng.standard_normal((n, d), dtype=np.float32)
  q = rng.standard_normal((nq, d), dtype=np.float32)
  bits = index_build.bfloat16_quantize(db)
  a = index_build.IndexArrays(distance="dot_product", dataset=None, n=n, d=d)
  a.bf16_dataset = bits
  return a, bits, q


@pytest.mark.parametrize("n,d,nq,k", [(20000, 96, 300, 20), (5000, 768, 130, 100), (700, 40, 5, 10),
                                      (70000, 128, 257, 100)])
def test_bf16_bruteforce_matches_oracle(n, d, nq, k):
  import oracle
  from scann_b200 import _lib
  a, bits, q = make(n, d, nq, seed=n + d)
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, dist = ix.search_batched(q)
  oi, od = oracle.bruteforce_bf16(bits, q, k, threads=8)
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(dist.view(np.uint32), od.view(np.uint32))
  # f32-query x bf16-row ground truth in float64
  x = (bits.view(np.uint16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)
  truth = np.take_along_axis(q.astype(np.float64) @ x.T, idx.astype(np.int64), axis=1)
  np.testing.assert_allclose(dist, truth, rtol=1e-5, atol=1e-4)
  st = ix.stats()
  assert st["kernel_launches"] >= 4 and st["scan_kernel_count"] >= 1


def test_bf16_bruteforce_recall_equals_oracle_recall():
  """North star: bf16 brute force is checked by recall@k equality against the f32-query restatement."""
  import oracle
  from scann_b200 import _lib
  a, bits, q = make(30000, 64, 200, seed=9)
  k = 50
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, _ = ix.search_batched(q)
  oi, _ = oracle.bruteforce_bf16(bits, q, k, threads=8)
  x = (bits.view(np.uint16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)
  gt = np.argsort(-(q.astype(np.float64) @ x.T), axis=1)[:, :k]
  rec = lambda r: np.mean([len(set(r[i].tolist()) & set(gt[i].tolist())) / k for i in range(len(q))])
  assert rec(idx) == rec(oi) and rec(idx) > 0.999


@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_sharded_bruteforce_equals_unsharded(world):
  """Every shard on this one GPU, the all-gather replaced by a concatenation: local top-k with global ids,
  scann_b200_merge_topk_device, against the unsharded index and the numpy restatement of the merge."""
  import ctypes as C
  import torch
  from scann_b200 import _lib, distributed
  n, d, nq, k = 30011, 96, 140, 50
  a, bits, q = make(n, d, nq, seed=17)
  full = _lib.NativeIndex(a, 1, k, k)
  fi, fd = full.search_batched(q)
  ids = np.zeros((world, nq, k), dtype=np.uint32)
  ds = np.zeros((world, nq, k), dtype=np.float32)
  shards = []
  for r in range(world):
    ix = _lib.NativeIndex(a, 1, k, k, shard_rank=r, shard_world=world)
    shards.append(ix)
    ids[r], ds[r] = ix.search_batched(q)
    per = -(-n // world)
    assert ids[r].min() >= r * per and ids[r].max() < min(n, (r + 1) * per)
  ri, rd = distributed.merge_topk_reference(ids, ds, k)
  np.testing.assert_array_equal(ri, fi)
  np.testing.assert_array_equal(rd.view(np.uint32), fd.view(np.uint32))
  dev = torch.device("cuda", 0)
  g_ids = torch.from_numpy(ids.view(np.int32).reshape(world * nq, k)).to(dev)
  g_ds = torch.from_numpy(ds.reshape(world * nq, k)).to(dev)
  o_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
  o_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
  vp = C.c_void_p
  _lib.check(_lib.lib().scann_b200_merge_topk_device(shards[0]._h, nq, world, k, vp(g_ids.data_ptr()),
                                                     vp(g_ds.data_ptr()), k, vp(o_i.data_ptr()), vp(o_d.data_ptr()), k))
  np.testing.assert_array_equal(o_i.cpu().numpy().view(np.uint32), fi)
  np.testing.assert_array_equal(o_d.cpu().numpy().view(np.uint32), fd.view(np.uint32))


# ---- float brute force (BruteForceSearcher<float>, brute_force/brute_force.cc:376-393) ----
@pytest.mark.parametrize("n,d,nq,k", [(20000, 96, 300, 20), (5000, 768, 130, 100), (700, 40, 5, 10), (30000, 100, 257, 50),
                                      (3000, 17, 40, 10)])
def test_float_bruteforce_matches_oracle(n, d, nq, k):
  import oracle
  from scann_b200 import _lib, index_build
  rng = np.random.default_rng(n + d)
  db = rng.standard_normal((n, d), dtype=np.float32)
  q = rng.standard_normal((nq, d), dtype=np.float32)
  a = index_build.IndexArrays(distance="dot_product", dataset=db, n=n, d=d)
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, dist = ix.search_batched(q)
  oi, od = oracle.bruteforce_f32(db, q, k, threads=8)
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(dist.view(np.uint32), od.view(np.uint32))
  truth = np.take_along_axis(q.astype(np.float64) @ db.astype(np.float64).T, idx.astype(np.int64), axis=1)
  np.testing.assert_allclose(dist, truth, rtol=1e-5, atol=1e-4)
  gt = np.argsort(-(q.astype(np.float64) @ db.astype(np.float64).T), axis=1)[:, :k]
  recall = np.mean([len(set(idx[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)])
  assert recall > 0.999


def test_brute_force_through_the_builder(tmp_path):
  """score_brute_force() (float) and score_brute_force(ReorderType.BFLOAT16): the reference's own brute-force test
  (scann_ops_pybind_test.py:79-90,253-264) checks against numpy at rtol 1e-6 / 1e-5."""
  from scann_b200 import scann_ops_pybind, scann_builder
  rng = np.random.default_rng(0)
  db = rng.random((2000, 32), dtype=np.float32)
  q = rng.random((20, 32), dtype=np.float32)
  s = scann_ops_pybind.builder(db, 10, "dot_product").score_brute_force().build()
  idx, dist = s.search_batched(q)
  full = q.astype(np.float64) @ db.astype(np.float64).T
  np.testing.assert_array_equal(idx, np.argsort(-full, axis=1)[:, :10].astype(np.uint32))
  np.testing.assert_allclose(dist, np.take_along_axis(full, idx.astype(np.int64), axis=1), rtol=1e-5)
  (tmp_path / "f32").mkdir()
  s.serialize(str(tmp_path / "f32"))
  assert (tmp_path / "f32" / "dataset.npy").exists()
  l2 = scann_ops_pybind.load_searcher(str(tmp_path / "f32"))
  i2, d2 = l2.search_batched(q)
  np.testing.assert_array_equal(idx, i2)
  np.testing.assert_array_equal(dist.view(np.uint32), d2.view(np.uint32))
  b = scann_ops_pybind.builder(db, 10, "dot_product").score_brute_force(scann_builder.ReorderType.BFLOAT16).build()
  ib, dbf = b.search_batched(q)
  assert np.mean(ib == idx) > 0.9
  np.testing.assert_allclose(dbf, np.take_along_axis(full, ib.astype(np.int64), axis=1), rtol=2e-2)


# ---- adversarial inputs: ordered databases, duplicates, near-ties, large k (VERDICT r1 / ADVICE r1) --------------------

def _check_bf16(bits, q, k, expect=None):
  import oracle
  from scann_b200 import _lib, index_build
  a = index_build.IndexArrays(distance="dot_product", dataset=None, n=bits.shape[0], d=bits.shape[1])
  a.bf16_dataset = bits
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, dist = ix.search_batched(q)
  oi, od = oracle.bruteforce_bf16(bits, q, k, threads=8)
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(dist.view(np.uint32), od.view(np.uint32))
  st = ix.stats()
  if expect:
    for key, lo in expect.items():
      assert st[key] >= lo, (key, st[key])
  ix.close()
  return st


def test_bf16_bruteforce_rows_sorted_by_score():
  """Rows in ASCENDING score order for query 0: every round's rows beat everything seen before, so the geometric
  rounds overflow that query's candidate buffer; the batch is re-run with rounds that cannot overflow."""
  from scann_b200 import index_build
  rng = np.random.default_rng(5)
  n, d, nq, k = 120000, 64, 40, 20
  db = rng.standard_normal((n, d), dtype=np.float32)
  q = rng.standard_normal((nq, d), dtype=np.float32)
  bits = index_build.bfloat16_quantize(db)
  x = (bits.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
  order = np.argsort(x @ q[0], kind="stable")
  _check_bf16(np.ascontiguousarray(bits[order]), q, k, expect={"bf_widenings": 1})


def test_bf16_bruteforce_cluster_sorted_rows():
  """Cluster-sorted database (rows of one cluster adjacent, clusters in order of their centre's score for a query)."""
  from scann_b200 import datasets, index_build
  n, d, nq, k = 150000, 96, 64, 100
  db = datasets.clustered(n, d, 60, seed=31, centers_seed=32)
  q = datasets.clustered(nq, d, 60, seed=33, centers_seed=32)
  bits = index_build.bfloat16_quantize(db)
  x = (bits.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
  order = np.argsort(np.round(x @ q[3], 0), kind="stable")
  _check_bf16(np.ascontiguousarray(bits[order]), q, k)


def test_bf16_bruteforce_duplicates_and_near_ties():
  """10 % exact duplicates (ties broken by id) and a cloud of near-ties around the k-th score: the window proof fails
  for the affected queries, the window widens, and what cannot be separated goes through the exact all-rows kernel."""
  from scann_b200 import index_build
  rng = np.random.default_rng(11)
  n, d, nq, k = 40000, 64, 24, 50
  db = rng.standard_normal((n, d), dtype=np.float32)
  db[rng.integers(0, n, n // 10)] = db[rng.integers(0, n, n // 10)]
  q = rng.standard_normal((nq, d), dtype=np.float32)
  # 12,000 rows within a relative 1e-6 of each other along query 0's direction: more near-ties than any window holds
  u = q[0] / np.linalg.norm(q[0])
  base = 6.0 * u
  db[:12000] = base[None, :] + 1e-6 * rng.standard_normal((12000, d)).astype(np.float32)
  db[12000:12040] = base  # exact ties at the top as well
  bits = index_build.bfloat16_quantize(db)
  st = _check_bf16(bits, q, k, expect={"bf_widenings": 1, "bf_exact_fallbacks": 1})
  assert st["bf_exact_fallbacks"] < nq


def test_bf16_bruteforce_k1000_and_narrow_window(monkeypatch):
  from scann_b200 import index_build
  rng = np.random.default_rng(13)
  n, d, nq = 60000, 48, 33
  bits = index_build.bfloat16_quantize(rng.standard_normal((n, d), dtype=np.float32))
  q = rng.standard_normal((nq, d), dtype=np.float32)
  _check_bf16(bits, q, 1000)
  # a window of exactly k candidates cannot be proven (the k-th exact equals the last kept): it must widen, not guess
  monkeypatch.setenv("SCANN_B200_BF_KPRIME", "10")
  _check_bf16(bits, q, 10, expect={"bf_widenings": 1})


def test_f32_bruteforce_sorted_rows_and_ties():
  import oracle
  from scann_b200 import _lib, index_build
  rng = np.random.default_rng(17)
  n, d, nq, k = 50000, 40, 20, 30
  db = rng.standard_normal((n, d), dtype=np.float32)
  q = rng.standard_normal((nq, d), dtype=np.float32)
  db = np.ascontiguousarray(db[np.argsort(db @ q[1], kind="stable")])
  db[100:140] = db[99]
  a = index_build.IndexArrays(distance="dot_product", dataset=db, n=n, d=d)
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, dist = ix.search_batched(q)
  oi, od = oracle.bruteforce_f32(db, q, k, threads=8)
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(dist.view(np.uint32), od.view(np.uint32))


# ---- squared-L2 float brute force (BruteForceSearcher<float> with SquaredL2Distance, brute_force.cc:376-393) ----
@pytest.mark.parametrize("n,d,nq,k", [(20000, 96, 300, 20), (5000, 768, 130, 100), (700, 40, 5, 10), (30000, 100, 257, 50),
                                      (3000, 17, 40, 10)])
def test_float_bruteforce_squared_l2_matches_oracle(n, d, nq, k):
  import oracle
  from scann_b200 import _lib, index_build
  rng = np.random.default_rng(n + d + 1)
  db = rng.standard_normal((n, d), dtype=np.float32) * rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)   # norms vary 4x
  q = rng.standard_normal((nq, d), dtype=np.float32)
  a = index_build.IndexArrays(distance="squared_l2", dataset=db, n=n, d=d)
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, dist = ix.search_batched(q)
  oi, od = oracle.bruteforce_f32(db, q, k, threads=8, distance="squared_l2")
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(dist.view(np.uint32), od.view(np.uint32))
  full = ((q.astype(np.float64)[:, None, :] - db.astype(np.float64)[None, :, :]) ** 2).sum(-1) if n * nq * d < 4e8 else None
  if full is not None:
    np.testing.assert_allclose(dist, np.take_along_axis(full, idx.astype(np.int64), axis=1), rtol=1e-5, atol=1e-4)
    gt = np.argsort(full, axis=1)[:, :k]
    recall = np.mean([len(set(idx[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)])
    assert recall > 0.999


def test_squared_l2_bruteforce_adversarial_and_sharded(monkeypatch):
  """Rows sorted by distance to one query, a block of exact duplicates, a narrow start window that has to widen, and the
  row-sharded form (local top-k with global ids, merge by (distance, id)) against the unsharded index."""
  import ctypes as C
  import oracle
  import torch
  from scann_b200 import _lib, index_build
  rng = np.random.default_rng(23)
  n, d, nq, k = 40000, 48, 24, 25
  db = rng.standard_normal((n, d), dtype=np.float32)
  q = rng.standard_normal((nq, d), dtype=np.float32)
  db = np.ascontiguousarray(db[np.argsort(((db - q[2]) ** 2).sum(1), kind="stable")[::-1]])   # best rows of query 2 come last
  db[500:560] = db[499]
  a = index_build.IndexArrays(distance="squared_l2", dataset=db, n=n, d=d)
  oi, od = oracle.bruteforce_f32(db, q, k, threads=8, distance="squared_l2")
  ix = _lib.NativeIndex(a, 1, k, k)
  idx, dist = ix.search_batched(q)
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(dist.view(np.uint32), od.view(np.uint32))
  monkeypatch.setenv("SCANN_B200_BF_KPRIME", str(k))
  idx2, dist2 = ix.search_batched(q)
  assert ix.stats()["bf_widenings"] >= 1
  np.testing.assert_array_equal(idx2, oi)
  np.testing.assert_array_equal(dist2.view(np.uint32), od.view(np.uint32))
  monkeypatch.delenv("SCANN_B200_BF_KPRIME")
  # row-sharded: three shards on this GPU, the all-gather replaced by a concatenation
  world = 3
  dev = torch.device("cuda", 0)
  d_q = torch.from_numpy(q).to(dev)
  ids = torch.empty((world, nq, k), dtype=torch.int32, device=dev)
  dists = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
  shards = [_lib.NativeIndex(a, 1, k, k, shard_rank=r, shard_world=world) for r in range(world)]
  for r, sh in enumerate(shards):
    sh.search_batched_device(d_q.data_ptr(), nq, ids[r].data_ptr(), dists[r].data_ptr(), k)
  out_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
  out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
  _lib.check(_lib.lib().scann_b200_merge_topk_device(shards[0]._h, nq, world, k, C.c_void_p(ids.data_ptr()), C.c_void_p(dists.data_ptr()),
                                                     k, C.c_void_p(out_i.data_ptr()), C.c_void_p(out_d.data_ptr()), k))
  np.testing.assert_array_equal(out_i.cpu().numpy().view(np.uint32), oi)
  np.testing.assert_array_equal(out_d.cpu().numpy().view(np.uint32), od.view(np.uint32))


def test_squared_l2_brute_force_through_the_builder():
  """scann_ops_pybind_test.py:245-264 (test_squared_l2): score_brute_force() with squared_l2 against numpy at rtol 1e-5."""
  from scann_b200 import scann_ops_pybind
  rng = np.random.default_rng(3)
  db = rng.random((2000, 32), dtype=np.float32)
  q = rng.random((20, 32), dtype=np.float32)
  s = scann_ops_pybind.builder(db, 10, "squared_l2").score_brute_force().build()
  idx, dist = s.search_batched(q)
  full = ((q.astype(np.float64)[:, None, :] - db.astype(np.float64)[None, :, :]) ** 2).sum(-1)
  np.testing.assert_array_equal(idx, np.argsort(full, axis=1)[:, :10].astype(np.uint32))
  np.testing.assert_allclose(dist, np.take_along_axis(full, idx.astype(np.int64), axis=1), rtol=1e-5)
