"""GPU: the reference-facing Python surface, mirroring scann/scann_ops/py/scann_ops_pybind_test.py
(serialize -> load -> same results `:39-59`, shapes `:210-242`, batching `:93-106`,
tree-AH parameter product `:108-159`) on seeded data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_data(n=1234, d=20, nq=10, seed=0):
  rng = np.random.default_rng(seed)
  return rng.random((n, d), dtype=np.float32), rng.random((nq, d), dtype=np.float32)


def verify_serialization(searcher, queries, tmp_path, relative):
  from scann_b200 import scann_ops_pybind
  idx0, dist0 = searcher.search_batched(queries)
  searcher.serialize(str(tmp_path), relative_path=relative)
  loaded = scann_ops_pybind.load_searcher(str(tmp_path))
  idx1, dist1 = loaded.search_batched(queries)
  np.testing.assert_array_equal(idx0, idx1)
  np.testing.assert_allclose(dist0, dist1, rtol=1e-6)
  assert loaded.config().strip() != ""
  return loaded


@pytest.mark.parametrize("soar", [None, 1.5])
@pytest.mark.parametrize("reorder", [None, 30])
@pytest.mark.parametrize("relative", [False, True])
def test_tree_ah_serialize_load_round_trip(tmp_path, soar, reorder, relative):
  from scann_b200 import scann_ops_pybind
  db, q = make_data()
  b = scann_ops_pybind.builder(db, 10, "dot_product").tree(27, 10, min_partition_size=10, soar_lambda=soar).score_ah(2)
  if reorder is not None:
    b = b.reorder(reorder)
  s = b.build()
  verify_serialization(s, q, tmp_path, relative)


def test_distances_match_numpy_ground_truth_after_reorder():
  from scann_b200 import scann_ops_pybind
  db, q = make_data(n=3000, d=32, nq=50, seed=3)
  s = scann_ops_pybind.builder(db, 10, "dot_product").tree(20, 20).score_ah(2).reorder(400).build()
  idx, dist = s.search_batched(q)
  truth = np.take_along_axis(q @ db.T, idx.astype(np.int64), axis=1)
  np.testing.assert_allclose(dist, truth, rtol=1e-5)
  # all leaves searched + generous reordering: essentially exact search
  gt = np.argsort(-(q.astype(np.float64) @ db.astype(np.float64).T), axis=1)[:, :10]
  rec = np.mean([len(set(idx[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(q))])
  assert rec > 0.95


def test_batching_matches_single_queries():
  from scann_b200 import scann_ops_pybind
  db, q = make_data()
  s = scann_ops_pybind.builder(db, 10, "dot_product").tree(27, 10, min_partition_size=10).score_ah(2).reorder(30).build()
  idx, dist = s.search_batched(q)
  for i in range(len(q)):
    si, sd = s.search(q[i])
    np.testing.assert_array_equal(si, idx[i])
    np.testing.assert_allclose(sd, dist[i], rtol=1e-6)
  pi, pd = s.search_batched_parallel(q)
  np.testing.assert_array_equal(pi, idx)


def test_shapes_and_overrides():
  from scann_b200 import scann_ops_pybind
  db, q = make_data()
  s = scann_ops_pybind.builder(db, 10, "dot_product").tree(27, 10, min_partition_size=10).score_ah(2).reorder(20).build()
  assert s.search(q[0])[0].shape == (10,)
  assert s.search(q[0], final_num_neighbors=15)[0].shape == (15,)
  assert s.search(q[0], final_num_neighbors=20, pre_reorder_num_neighbors=50)[0].shape == (20,)
  assert s.search_batched(q)[0].shape == (10, 10)
  assert s.search_batched(q, final_num_neighbors=15)[0].shape == (10, 15)
  assert s.search_batched(q, final_num_neighbors=20, pre_reorder_num_neighbors=50, leaves_to_search=27)[1].shape == (10, 20)
  assert s.search_batched(q)[0].dtype == np.uint32 and s.search_batched(q)[1].dtype == np.float32
  assert s.size() == 1234


def test_docids_and_error_behaviour(tmp_path):
  from scann_b200 import scann_ops_pybind
  db, q = make_data(n=300, d=16, nq=4)
  docids = [f"doc{i}" for i in range(300)]
  s = scann_ops_pybind.builder(db, 5, "dot_product").tree(8, 8).score_ah(2).reorder(20).build(docids=docids)
  ids, _ = s.search_batched(q)
  assert all(isinstance(x, str) and x.startswith("doc") for row in ids for x in row)
  s.serialize(str(tmp_path))
  assert (tmp_path / "scann_docids.pkl").exists()
  l2 = scann_ops_pybind.load_searcher(str(tmp_path))
  assert l2.search_batched(q)[0] == ids
  with pytest.raises(ValueError):
    s.searcher.search_batched(q[0], -1, -1, -1, False, 0)           # wrong ndim -> invalid_argument
  with pytest.raises(RuntimeError, match="Error during search: Query doesn't match dataset dim"):
    s.search_batched(np.zeros((2, 7), np.float32))
  with pytest.raises(RuntimeError, match="not supported"):
    s.upsert(["x"], db[:1])
  with pytest.raises(ValueError, match="docid and database size mismatch"):
    scann_ops_pybind.create_searcher(db, "", docids=["a"])
  with pytest.raises(RuntimeError, match="Error initializing searcher: UNIMPLEMENTED"):
    # bfloat16 brute force is MIPS-only, like the reference's (bfloat16_brute_force.cc:60-75); float rows take squared L2
    from scann_b200 import scann_builder
    scann_ops_pybind.builder(db, 5, "squared_l2").score_brute_force(scann_builder.ReorderType.BFLOAT16).build()


def test_empty_partitions_are_tolerated():
  from scann_b200 import scann_ops_pybind
  rng = np.random.default_rng(7)
  db = np.repeat(rng.random((20, 16), dtype=np.float32), 30, axis=0)   # 20 distinct points, 600 rows
  q = rng.random((5, 16), dtype=np.float32)
  s = scann_ops_pybind.builder(db, 10, "dot_product").tree(20, 20, min_partition_size=1).score_ah(2).reorder(50).build()
  idx, dist = s.search_batched(q)
  truth = np.take_along_axis(q @ db.T, idx.astype(np.int64), axis=1)
  np.testing.assert_allclose(dist, truth, rtol=1e-5)


def test_bfloat16_reordering_through_the_builder(tmp_path):
  """reorder(N, quantize=ReorderType.BFLOAT16): bfloat16_dataset.npy asset, round trip, distances of the bf16 rows."""
  from scann_b200 import scann_ops_pybind, scann_builder, index_build
  db, q = make_data()
  s = (scann_ops_pybind.builder(db, 10, "dot_product").tree(27, 10, min_partition_size=10).score_ah(2)
       .reorder(60, quantize=scann_builder.ReorderType.BFLOAT16).build())
  idx, dist = s.search_batched(q)
  bits = index_build.bfloat16_quantize(db)
  x = (bits.view(np.uint16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)
  truth = np.take_along_axis(q.astype(np.float64) @ x.T, idx.astype(np.int64), axis=1)
  np.testing.assert_allclose(dist, truth, rtol=1e-5, atol=1e-5)
  s.serialize(str(tmp_path))
  assert (tmp_path / "bfloat16_dataset.npy").exists() and not (tmp_path / "dataset.npy").exists()
  l2 = scann_ops_pybind.load_searcher(str(tmp_path))
  i2, d2 = l2.search_batched(q)
  np.testing.assert_array_equal(idx, i2)
  np.testing.assert_array_equal(dist.view(np.uint32), d2.view(np.uint32))


def test_int8_reordering_through_the_builder(tmp_path):
  """reorder(N, quantize=ReorderType.INT8): int8 assets, round trip, distances of the dequantized rows."""
  from scann_b200 import scann_ops_pybind, scann_builder, index_build
  db, q = make_data()
  for dist in ("dot_product", "squared_l2"):
    out = tmp_path / dist
    out.mkdir()
    s = (scann_ops_pybind.builder(db, 10, dist).tree(27, 10, min_partition_size=10).score_ah(2)
         .reorder(60, quantize=scann_builder.ReorderType.INT8).build())
    idx, dist_v = s.search_batched(q)
    q8, mult = index_build.int8_quantize(db)
    deq = q8.astype(np.float64) / mult.astype(np.float64)[None, :]
    rows = deq[idx.astype(np.int64)]
    if dist == "dot_product":
      truth = np.einsum("qd,qkd->qk", q.astype(np.float64), rows)
    else:
      truth = ((q.astype(np.float64) ** 2).sum(1)[:, None] + (db.astype(np.float64) ** 2).sum(1)[idx.astype(np.int64)]
               - 2 * np.einsum("qd,qkd->qk", q.astype(np.float64), rows))
    np.testing.assert_allclose(dist_v, truth, rtol=1e-4, atol=1e-3)
    s.serialize(str(out))
    assert (out / "int8_dataset.npy").exists() and (out / "int8_multipliers.npy").exists()
    assert (out / "dp_norms.npy").exists() == (dist == "squared_l2") and not (out / "dataset.npy").exists()
    l2 = scann_ops_pybind.load_searcher(str(out))
    i2, d2 = l2.search_batched(q)
    np.testing.assert_array_equal(idx, i2)
    np.testing.assert_array_equal(dist_v.view(np.uint32), d2.view(np.uint32))


def test_concurrent_search_batched_on_one_handle():
  """ScannInterface::SearchBatched may be called from several threads (scann_ops/cc/scann.cc:478-501): concurrent calls on
  one handle run on separate lanes (own stream + workspace) and return exactly what serial calls return."""
  import threading
  from conftest import get_case
  c = get_case(soar=1.5)
  want_i, want_d = c.native.search_batched(c.q)
  results = {}

  def worker(t):
    out = []
    for rep in range(6):
      lo = (t * 7 + rep * 3) % 40
      out.append((lo, c.native.search_batched(c.q[lo:lo + 24])))
    results[t] = out

  threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]   # more threads than lanes: some wait
  for th in threads:
    th.start()
  for th in threads:
    th.join()
  assert len(results) == 6
  for t, out in results.items():
    for lo, (i, d) in out:
      np.testing.assert_array_equal(i, want_i[lo:lo + 24])
      np.testing.assert_array_equal(d.view(np.uint32), want_d[lo:lo + 24].view(np.uint32))
  assert c.native.stats()["kernel_launches"] >= 8
