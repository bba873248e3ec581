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
