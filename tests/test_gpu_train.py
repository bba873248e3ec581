"""GPU parity: scann_b200_train_kmeans (Lloyd iterations of the k-means tree / AH codebook trainers)
against the oracle's restatement of GmmUtils' loop (utils/gmm_utils.cc:846-915, 1052-1132), bit for bit, through the
C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixture(n, d, k, seed, noise=0.6, normalize=False):
  rng = np.random.default_rng(seed)
  means = rng.standard_normal((k, d)).astype(np.float32)
  x = (means[rng.integers(0, k, n)] + noise * rng.standard_normal((n, d))).astype(np.float32)
  if normalize:
    x /= np.linalg.norm(x, axis=1, keepdims=True)
  init = x[np.sort(rng.choice(n, k, replace=False))].copy()
  return np.ascontiguousarray(x), init


CASES = [
    # n, d, k, iterations, normalize
    (4000, 32, 16, 5, False),      # SIMT assignment (k < 256), 4 aggregation slices
    (3000, 2, 16, 8, False),       # an AH block: 16 centres over 2 dims
    (100, 8, 16, 3, False),        # n < 8 k: the single-slice aggregation
    (20000, 100, 300, 4, True),    # tensor-core assignment, glove-like
    (12000, 96, 512, 3, True),     # deep-like
    (5000, 300, 260, 2, False),    # d > 256: shared-memory accumulators
    (2000, 24, 2000, 1, False),    # k = n: every point its own centre
]


@pytest.mark.parametrize("n,d,k,iterations,normalize", CASES)
def test_train_kmeans_matches_oracle(n, d, k, iterations, normalize):
  import oracle
  from scann_b200 import _lib
  x, init = _mixture(n, d, k, n + d + k, normalize=normalize)
  centers, assign, st = _lib.train_kmeans(x, init, iterations)
  o_centers, o_assign, o_empty = oracle.kmeans(x, init, iterations, threads=8)
  assert np.array_equal(centers.view(np.uint32), o_centers.view(np.uint32))
  assert np.array_equal(assign, o_assign)
  assert st["empty_clusters"] == o_empty
  assert st["iterations"] == iterations
  # the statistic is the mean squared distance of the final partition
  ref = np.mean(((x.astype(np.float64) - centers[assign].astype(np.float64)) ** 2).sum(1))
  assert abs(st["mean_sq_distance"] - ref) <= 1e-4 * ref + 1e-5   # float tokenizer distances vs float64


def test_train_kmeans_lowers_the_distortion():
  from scann_b200 import _lib
  x, init = _mixture(30000, 64, 400, 5)
  _, _, s0 = _lib.train_kmeans(x, init, 0)
  _, _, s5 = _lib.train_kmeans(x, init, 5)
  assert s5["mean_sq_distance"] < 0.9 * s0["mean_sq_distance"]


def test_train_kmeans_keeps_empty_clusters_and_reports_them():
  import oracle
  from scann_b200 import _lib
  x, init = _mixture(3000, 16, 32, 9)
  init[7] = 1e3          # nothing is ever assigned to it
  centers, assign, st = _lib.train_kmeans(x, init, 3)
  assert st["empty_clusters"] >= 1 and np.array_equal(centers[7], init[7]) and not (assign == 7).any()
  o_centers, o_assign, o_empty = oracle.kmeans(x, init, 3, threads=4)
  assert np.array_equal(centers.view(np.uint32), o_centers.view(np.uint32)) and o_empty == st["empty_clusters"]


def test_train_kmeans_argument_errors():
  from scann_b200 import _lib
  x, init = _mixture(100, 8, 16, 1)
  with pytest.raises(RuntimeError, match="less than the number of clusters"):
    _lib.train_kmeans(x[:8], init, 1)


def test_builder_trains_through_the_library():
  """build_tree_ah on a device: k-means tree and AH codebooks come from scann_b200_train_kmeans (deterministic)."""
  from scann_b200 import datasets, index_build
  db = datasets.clustered(20000, 64, 400, seed=3, centers_seed=4)
  a = index_build.build_tree_ah(db, "dot_product", num_leaves=300, dims_per_block=2, training_sample_size=20000,
                                tree_iters=4, ah_iters=4)
  b = index_build.build_tree_ah(db, "dot_product", num_leaves=300, dims_per_block=2, training_sample_size=20000,
                                tree_iters=4, ah_iters=4)
  assert a.meta["trainer"] == "scann_b200_train_kmeans"
  assert np.array_equal(a.centers, b.centers) and np.array_equal(a.codebook, b.codebook)
  assert np.array_equal(a.codes, b.codes)
  assert np.bincount(a.tokens, minlength=300).min() > 0
