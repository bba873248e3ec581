"""GPU: edge cases of the batched query path against the CPU oracle (empty / single / chunked batches, short
result rows, k larger than what the probed leaves hold, squared L2 through the builder API)."""
import numpy as np
import pytest

from conftest import get_case

pytestmark = pytest.mark.gpu


def test_empty_and_single_query_batches():
  c = get_case()
  k = c.native.default_final_nn
  i0, d0 = c.native.search_batched(np.zeros((0, c.q.shape[1]), np.float32))
  assert i0.shape == (0, k) and d0.shape == (0, k)
  oi, od = c.oracle.search_batched(c.q[:1])
  i1, d1 = c.native.search_batched(c.q[:1])
  np.testing.assert_array_equal(oi, i1)
  np.testing.assert_array_equal(od.view(np.uint32), d1.view(np.uint32))
  # every prefix length up to a few warps' worth: partially filled octs / work items / CTAs
  for nq in (2, 7, 8, 9, 15, 17, 33):
    oi, od = c.oracle.search_batched(c.q[:nq])
    i1, d1 = c.native.search_batched(c.q[:nq])
    np.testing.assert_array_equal(oi, i1)
    np.testing.assert_array_equal(od.view(np.uint32), d1.view(np.uint32))


def test_batch_larger_than_one_chunk_equals_per_chunk_results():
  """nq > 16384 is processed in chunks by the host layer; results must not depend on the chunking."""
  c = get_case(n=4000, leaves=40, probe=6, pre=40, d=32)
  rng = np.random.default_rng(5)
  q = c.q[rng.integers(0, len(c.q), size=20001)] + rng.standard_normal((20001, c.q.shape[1]), dtype=np.float32) * 0.01
  q = np.ascontiguousarray(q, dtype=np.float32)
  i_all, d_all = c.native.search_batched(q)
  for s0, s1 in ((0, 300), (16384 - 100, 16384 + 100), (19800, 20001)):
    oi, od = c.oracle.search_batched(q[s0:s1])
    np.testing.assert_array_equal(oi, i_all[s0:s1])
    np.testing.assert_array_equal(od.view(np.uint32), d_all[s0:s1].view(np.uint32))


def test_short_rows_are_padded_like_the_reference():
  """k larger than the number of datapoints in the probed leaves: rows end with (id 0, NaN) (scann.h:175-178)."""
  c = get_case(n=600, leaves=60, probe=2, pre=64, d=24)
  k = 48
  oi, od = c.oracle.search_batched(c.q, final_nn=k, pre_nn=64)
  i1, d1 = c.native.search_batched(c.q, final_nn=k, pre_nn=64)
  assert np.isnan(od).any(), "the case must produce short rows"
  np.testing.assert_array_equal(oi, i1)
  np.testing.assert_array_equal(od.view(np.uint32), d1.view(np.uint32))
  short = np.isnan(d1)
  assert (i1[short] == 0).all()


def test_leaves_override_and_pre_reorder_override():
  c = get_case(soar=1.5)
  for leaves, pre, k in ((1, 10, 5), (3, 33, 10), (c.native.L, 200, 25)):
    oi, od = c.oracle.search_batched(c.q[:64], final_nn=k, pre_nn=pre, leaves=leaves)
    i1, d1 = c.native.search_batched(c.q[:64], final_nn=k, pre_nn=pre, leaves=leaves)
    np.testing.assert_array_equal(oi, i1)
    np.testing.assert_array_equal(od.view(np.uint32), d1.view(np.uint32))


def test_squared_l2_through_the_builder_api():
  """tree().score_ah().reorder() with squared_l2 (TreeXHybridSMMD semantics, BASELINE.json configs[3] family)."""
  from scann_b200 import scann_ops_pybind
  rng = np.random.default_rng(11)
  db = np.round(np.clip(np.abs(rng.standard_normal((5000, 32))) * 40, 0, 218)).astype(np.float32)   # SIFT-like
  q = np.round(np.clip(np.abs(rng.standard_normal((50, 32))) * 40, 0, 218)).astype(np.float32)
  s = (scann_ops_pybind.builder(db, 10, "squared_l2").tree(50, 50, training_sample_size=5000)
       .score_ah(2).reorder(200).build())
  idx, dist = s.search_batched(q)
  exact = ((q[:, None, :].astype(np.float64) - db[None].astype(np.float64)) ** 2).sum(-1)
  np.testing.assert_allclose(dist, np.take_along_axis(exact, idx.astype(np.int64), axis=1), rtol=1e-5)
  assert (np.diff(dist, axis=1) >= 0).all()                          # ascending squared L2
  gt = np.argsort(exact, axis=1)[:, :10]
  recall = np.mean([len(set(idx[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(q))])
  assert recall > 0.9
