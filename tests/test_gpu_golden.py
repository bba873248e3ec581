"""GPU: the CUDA path against the committed golden fixtures (no oracle involved at run time)."""
import numpy as np
import pytest

from helpers import golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_cuda_path_reproduces_golden(name):
  from scann_b200 import _lib
  a, z = load_golden(name)
  ix = _lib.NativeIndex(a, int(z["probe"]), int(z["pre"]), int(z["k"]))
  q = z["queries"]
  leaf, cdist = ix.tokenize(q)
  np.testing.assert_array_equal(leaf, z["exp_leaf"])
  np.testing.assert_array_equal(cdist.view(np.uint32), z["exp_center_dist"].view(np.uint32))
  lut, mult = ix.lut(q)
  np.testing.assert_array_equal(lut, z["exp_lut"])
  np.testing.assert_array_equal(mult.view(np.uint32), z["exp_mult"].view(np.uint32))
  np.testing.assert_array_equal(ix.leaf_scores(lut[0], int(leaf[0, 0])), z["exp_scores_q0_leaf0"])
  c = ix.candidates(q)
  np.testing.assert_array_equal(c["count"], z["exp_cand_count"])
  n = z["exp_cand_dp"].shape[1]
  for i in range(len(q)):
    m = int(c["count"][i])
    np.testing.assert_array_equal(c["dp"][i, :m], z["exp_cand_dp"][i, :m])
    np.testing.assert_array_equal(c["leaf"][i, :m], z["exp_cand_leaf"][i, :m])
    np.testing.assert_array_equal(c["slot"][i, :m], z["exp_cand_slot"][i, :m])
    np.testing.assert_array_equal(c["score"][i, :m].view(np.uint32), z["exp_cand_score"][i, :m].view(np.uint32))
  idx, dist = ix.search_batched(q)
  np.testing.assert_array_equal(idx, z["exp_idx"])
  np.testing.assert_array_equal(dist.view(np.uint32), z["exp_dist"].view(np.uint32))
  ix.close()


def test_overflow_rescan_is_exact():
  """Force candidate-buffer overflow (huge N' relative to cap is impossible, so use many leaves
  with an adversarial query far from its nearest leaf) and check parity still holds."""
  import oracle
  from scann_b200 import _lib, datasets, index_build
  db = datasets.clustered(60000, 32, 64, seed=31, centers_seed=131)
  a = index_build.build_tree_ah(db, "dot_product", num_leaves=200, dims_per_block=2, training_sample_size=20000,
                                tree_iters=4, ah_iters=4, device="cpu")
  rng = np.random.default_rng(5)
  q = rng.standard_normal((40, 32)).astype(np.float32) * 3.0   # queries unlike the data
  import os
  ix = _lib.NativeIndex(a, 150, 100, 10)
  oi = oracle.OracleIndex(a, 150, 100, 10)
  i0, d0 = oi.search_batched(q, impl=1)
  os.environ["SCANN_B200_CAND_CAP"] = "256"     # tiny candidate buffers: force the re-scan path
  try:
    i1, d1 = ix.search_batched(q)
    assert ix.stats()["overflow_retries"] >= 1
  finally:
    del os.environ["SCANN_B200_CAND_CAP"]
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  i2, d2 = ix.search_batched(q)                  # default buffers: no re-scan needed
  assert ix.stats()["overflow_retries"] == 0
  np.testing.assert_array_equal(i0, i2)
