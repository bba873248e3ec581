"""GPU tests at BASELINE.json's C2 size (1,183,514 x 100, 2000 leaves, 100 probed, reorder 100, k = 10, 10k queries):
the oracle finishes only a sample there, so the rest is checked through size-independent properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_STATE = {}


def _c2():
  if "c2" not in _STATE:
    import bench
    from scann_b200 import _lib
    wl = dict(bench.WORKLOADS["c2_glove_shape"])
    db, q = bench.make_data(wl)
    arrays = bench.build_arrays(wl, db, "cuda:0")
    ix = _lib.NativeIndex(arrays, wl["probe"], wl["pre"], wl["k"])
    _STATE["c2"] = (wl, db, q, arrays, ix)
  return _STATE["c2"]


def test_c2_search_properties_and_oracle_sample():
  import oracle
  from scann_b200 import _lib, index_build
  wl, db, q, arrays, ix = _c2()
  idx, dist = ix.search_batched(q)
  assert idx.shape == (wl["nq"], wl["k"]) and not np.isnan(dist).any()
  # rows sorted by descending dot product (ascending internal distance), ids unique
  assert (np.diff(dist, axis=1) <= 0).all()
  assert all(len(set(r.tolist())) == wl["k"] for r in idx[:2000])
  # batch splitting and query order do not change a bit (each query's result is a function of the query alone)
  ia, da = ix.search_batched(q[:3333])
  ib, db_ = ix.search_batched(q[3333:])
  np.testing.assert_array_equal(np.vstack([ia, ib]), idx)
  np.testing.assert_array_equal(np.vstack([da, db_]).view(np.uint32), dist.view(np.uint32))
  perm = np.random.default_rng(0).permutation(wl["nq"])
  ip, dp = ix.search_batched(q[perm])
  np.testing.assert_array_equal(ip, idx[perm])
  np.testing.assert_array_equal(dp.view(np.uint32), dist[perm].view(np.uint32))
  # a sample against the oracle, bit for bit
  oi = oracle.OracleIndex(arrays, wl["probe"], wl["pre"], wl["k"])
  oidx, odist = oi.search_batched(q[:96], impl=1, threads=8)
  np.testing.assert_array_equal(idx[:96], oidx)
  np.testing.assert_array_equal(dist[:96].view(np.uint32), odist.view(np.uint32))
  # reported distances are the true dot products; recall against exact float brute force on the same device
  sel = np.arange(0, wl["nq"], 20)
  truth_d = np.einsum("qd,qkd->qk", q[sel].astype(np.float64), db[idx[sel].astype(np.int64)].astype(np.float64))
  np.testing.assert_allclose(dist[sel], truth_d, rtol=1e-5, atol=1e-6)
  bf = _lib.NativeIndex(index_build.IndexArrays(distance="dot_product", dataset=db, n=db.shape[0], d=db.shape[1]),
                        1, wl["k"], wl["k"])
  bidx, _ = bf.search_batched(q[:2000])
  recall = np.mean([len(set(idx[i].tolist()) & set(bidx[i].tolist())) / wl["k"] for i in range(2000)])
  assert recall >= 0.90, recall


def test_c2_index_build_properties_and_oracle_sample(monkeypatch):
  import oracle
  from scann_b200 import _lib
  wl, db, q, arrays, ix = _c2()
  bd = arrays.block_dims
  t0, c0, s0, st0 = _lib.encode_database(db, arrays.centers, arrays.codebook, bd, soar_lambda=1.5, noise_shaping_threshold=0.2)
  # every datapoint is in a valid leaf; spilled copies sit in a different, higher-numbered leaf; codes are 4-bit
  lo, hi = t0[0::2], t0[1::2]
  assert lo.min() >= 0 and lo.max() < wl["leaves"] and ((hi == -1) | (hi > lo)).all() and hi.max() < wl["leaves"]
  assert c0.max() < 16 and s0.max() < 16 and (s0[hi == -1] == 0).all()
  assert st0["spilled"] == int((hi >= 0).sum()) and 0.2 < st0["spilled"] / len(db) < 1.0
  # the pruned SOAR search evaluated a sliver of the N x L costs
  assert st0["soar_evaluated"] < 0.02 * len(db) * wl["leaves"]
  # the chunk size is not observable
  monkeypatch.setenv("SCANN_B200_ENCODE_CHUNK", "50000")
  t1, c1, s1, st1 = _lib.encode_database(db, arrays.centers, arrays.codebook, bd, soar_lambda=1.5, noise_shaping_threshold=0.2)
  assert st1["chunk_rows"] == 50000
  np.testing.assert_array_equal(t0, t1)
  np.testing.assert_array_equal(c0, c1)
  np.testing.assert_array_equal(s0, s1)
  # without spilling the leaf is the primary: the assignment the bench index was built with
  np.testing.assert_array_equal(arrays.tokens, _lib.encode_database(db, arrays.centers, arrays.codebook, bd)[0])
  # a sample against the oracle, bit for bit (rows from the start, the middle and the end)
  rows = np.r_[0:700, 600000:600700, len(db) - 700:len(db)]
  ot, oc, os_, _ = oracle.encode_database(db[rows], arrays.centers, arrays.codebook, bd, soar_lambda=1.5, threshold=0.2, threads=8)
  np.testing.assert_array_equal(t0.reshape(-1, 2)[rows].reshape(-1), ot)
  np.testing.assert_array_equal(c0[rows], oc)
  np.testing.assert_array_equal(s0[rows], os_)
  # noise shaping trades a slightly larger residual for a smaller component parallel to the datapoint
  tp, cp, _, _ = _lib.encode_database(db[:200000], arrays.centers, arrays.codebook, bd)
  ts, cs, _, _ = _lib.encode_database(db[:200000], arrays.centers, arrays.codebook, bd, noise_shaping_threshold=0.2)
  np.testing.assert_array_equal(tp, ts)
  x = db[:200000].astype(np.float64)
  res = x - arrays.centers[tp].astype(np.float64)

  def err(codes):
    cbk = arrays.codebook.astype(np.float64)
    rec = np.concatenate([cbk[b][codes[:, b]][:, :bd[b]] for b in range(len(bd))], axis=1)
    e = res - rec
    par = (e * x).sum(1) / np.linalg.norm(x, axis=1)
    return (e ** 2).sum(1).mean(), (par ** 2).mean()

  n_p, p_p = err(cp)
  n_s, p_s = err(cs)
  assert n_s >= n_p and p_s < p_p


# ---- the other BASELINE.json configurations at their own sizes (only C2 had a full-size oracle check) ------

def _tree_ah_fullsize(name, n_oracle_queries, overrides=None):
  """Builds the bench workload `name` exactly as bench.py does, searches all its queries on the GPU, and compares an
  oracle run on the first n_oracle_queries bit for bit."""
  import bench
  import oracle
  from scann_b200 import _lib
  wl = dict(bench.WORKLOADS[name])
  wl.update(overrides or {})
  db, q = bench.make_data(wl)
  arrays = bench.build_arrays(wl, db, "cuda:0")
  ix = _lib.NativeIndex(arrays, wl["probe"], wl["pre"], wl["k"])
  idx, dist = ix.search_batched(q)
  st = ix.stats()
  oi = oracle.OracleIndex(arrays, wl["probe"], wl["pre"], wl["k"])
  m = min(n_oracle_queries, wl["nq"])
  oidx, odist = oi.search_batched(q[:m], impl=1, threads=16)
  np.testing.assert_array_equal(idx[:m], oidx)
  np.testing.assert_array_equal(dist[:m].view(np.uint32), odist.view(np.uint32))
  # the scalar oracle path (impl 0) on a few queries as well: the AVX2 path is itself only a restatement
  o2, d2 = oi.search_batched(q[:8], impl=0)
  np.testing.assert_array_equal(idx[:8], o2)
  np.testing.assert_array_equal(dist[:8].view(np.uint32), d2.view(np.uint32))
  assert not np.isnan(dist).any() and all(len(set(r.tolist())) == wl["k"] for r in idx[:1000])
  # reported distances against float64 on the rows themselves (tolerance of the north star: 1e-5 relative)
  sel = np.arange(0, wl["nq"], 50)
  rows = db[idx[sel].astype(np.int64)].astype(np.float64)
  if wl.get("distance") == "squared_l2":
    truth = ((rows - q[sel].astype(np.float64)[:, None, :]) ** 2).sum(2)
    assert (np.diff(dist, axis=1) >= 0).all()
  else:
    truth = np.einsum("qd,qkd->qk", q[sel].astype(np.float64), rows)
    assert (np.diff(dist, axis=1) <= 0).all()
  np.testing.assert_allclose(dist[sel], truth, rtol=1e-5, atol=1e-5 * float(np.abs(truth).max()))
  ix.close()
  return wl, st


def test_c1_all_queries_against_the_oracle():
  """BASELINE.json configs[0]: 100k x 100, 100 leaves, reorder 100, k = 10 -- every one of the 10,000 queries."""
  wl, st = _tree_ah_fullsize("c1_synthetic", 10000)
  assert st["overflow_retries"] == 0


def test_c4_sift_shape_against_the_oracle():
  """BASELINE.json configs[3] shape: 10M x 128 squared L2 (TreeXHybridSMMD semantics), 4000 leaves, 64 probed."""
  wl, st = _tree_ah_fullsize("c4_sift_shape", 64)
  assert st["scan_bytes_alg"] > 0


def test_c5_deep_shape_against_the_oracle():
  """BASELINE.json configs[4] shape at 20M rows: 96-d, SOAR, reorder 200 (400 over-retrieved), 24 of 8000 leaves: the
  two-phase scan and the per-item candidate staging are active at this size."""
  wl, st = _tree_ah_fullsize("c5_deep_shape", 64)
  assert st["scan_kernel_count"] >= 2     # two scan phases
  assert st["cand_sum"] / wl["nq"] > 400  # more candidates buffered than kept: the compaction classes are exercised


def test_c3_bruteforce_bf16_against_the_oracle():
  """BASELINE.json configs[2]: 1M x 768 bf16 rows, k = 100: 32 queries against the oracle's exact f32 x bf16 scan, all
  10,000 for the invariants (sorted, unique ids, no widening needed on i.i.d. data)."""
  import bench
  import oracle
  from scann_b200 import _lib
  wl = dict(bench.WORKLOADS["c3_bruteforce_bf16"])
  a, bits, q = bench.bruteforce_data(wl, 16)
  ix = _lib.NativeIndex(a, 1, wl["k"], wl["k"])
  idx, dist = ix.search_batched(q)
  st = ix.stats()
  oi, od = oracle.bruteforce_bf16(bits, q[:32], wl["k"], threads=16)
  np.testing.assert_array_equal(idx[:32], oi)
  np.testing.assert_array_equal(dist[:32].view(np.uint32), od.view(np.uint32))
  assert (np.diff(dist, axis=1) <= 0).all() and all(len(set(r.tolist())) == wl["k"] for r in idx[:500])
  assert st["bf_widenings"] == 0 and st["bf_exact_fallbacks"] == 0
  ix.close()
