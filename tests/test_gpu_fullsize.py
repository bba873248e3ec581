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
