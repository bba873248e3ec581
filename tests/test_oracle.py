"""CPU tests: the oracle against committed golden fixtures and an independent numpy restatement."""
import numpy as np
import pytest

import oracle
from helpers import golden_names, i8_tok_arrays as _i8_tok_arrays, load_golden, np_center_distances, np_lut_dpb2, np_search, np_slots


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(name):
  a, z = load_golden(name)
  oi = oracle.OracleIndex(a, int(z["probe"]), int(z["pre"]), int(z["k"]))
  q = z["queries"]
  leaf, cdist = oi.tokenize(q)
  np.testing.assert_array_equal(leaf, z["exp_leaf"])
  np.testing.assert_array_equal(cdist.view(np.uint32), z["exp_center_dist"].view(np.uint32))
  lut, mult = oi.lut(q)
  np.testing.assert_array_equal(lut, z["exp_lut"])
  np.testing.assert_array_equal(mult.view(np.uint32), z["exp_mult"].view(np.uint32))
  np.testing.assert_array_equal(oi.leaf_scores(lut[0], int(leaf[0, 0])), z["exp_scores_q0_leaf0"])
  c = oi.candidates(q)
  np.testing.assert_array_equal(c["count"], z["exp_cand_count"])
  np.testing.assert_array_equal(c["dp"], z["exp_cand_dp"])
  np.testing.assert_array_equal(c["acc"], z["exp_cand_acc"])
  for impl in (0, 1):
    idx, dist = oi.search_batched(q, impl=impl)
    np.testing.assert_array_equal(idx, z["exp_idx"])
    np.testing.assert_array_equal(dist.view(np.uint32), z["exp_dist"].view(np.uint32))


@pytest.mark.parametrize("name", ["dot_b16", "dot_soar_b25"])
def test_oracle_matches_numpy_restatement(name):
  a, z = load_golden(name)
  probe, pre, k = int(z["probe"]), int(z["pre"]), int(z["k"])
  oi = oracle.OracleIndex(a, probe, pre, k)
  q = z["queries"][:12]
  ids, dists, cd, lut, mult, cands = np_search(a, q, probe, pre, k)
  leaf, cdist = oi.tokenize(q)
  for i in range(len(q)):
    np.testing.assert_array_equal(cdist[i].view(np.uint32), cd[i, leaf[i]].view(np.uint32))
    assert sorted(leaf[i].tolist()) == sorted(np.lexsort((np.arange(cd.shape[1]), cd[i]))[:probe].tolist())
  olut, omult = oi.lut(q)
  np.testing.assert_array_equal(omult.view(np.uint32), mult.view(np.uint32))
  np.testing.assert_array_equal(olut, lut)
  c = oi.candidates(q)
  for i in range(len(q)):
    n = int(c["count"][i])
    assert n == len(cands[i])
    np.testing.assert_array_equal(c["leaf"][i, :n], [r[1] for r in cands[i]])
    np.testing.assert_array_equal(c["slot"][i, :n], [r[2] for r in cands[i]])
    np.testing.assert_array_equal(c["dp"][i, :n], [r[3] for r in cands[i]])
    np.testing.assert_array_equal(c["score"][i, :n].view(np.uint32),
                                  np.asarray([r[4] for r in cands[i]], np.float32).view(np.uint32))
  oidx, odist = oi.search_batched(q)
  # ids: identical except where two float32 reorder distances tie within rounding of the f64 truth
  np.testing.assert_allclose(odist, dists, rtol=1e-5, atol=1e-5)
  assert (oidx == ids).mean() > 0.98


def test_leaf_membership_and_packing():
  a, z = load_golden("dot_soar_b25")
  oi = oracle.OracleIndex(a, int(z["probe"]), int(z["pre"]), int(z["k"]))
  slots = np_slots(a.tokens, a.centers.shape[0], True)
  assert not oi.disjoint
  for leaf, dps in enumerate(slots):
    assert oi.leaf_size(leaf) == len(dps)
    np.testing.assert_array_equal(oi.leaf_datapoints(leaf), dps)


def test_scalar_and_avx2_and_parallel_agree():
  a, z = load_golden("dot_b16")
  oi = oracle.OracleIndex(a, 6, 50, 10)
  q = z["queries"]
  i0, d0 = oi.search_batched(q, impl=0)
  i1, d1 = oi.search_batched(q, impl=1)
  i2, d2 = oi.search_batched(q, impl=1, threads=3, batch=5)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(i0, i2)
  np.testing.assert_array_equal(d0.view(np.uint32), d2.view(np.uint32))


def test_recall_against_brute_force():
  a, z = load_golden("dot_b16")
  oi = oracle.OracleIndex(a, 8, 100, 10)
  q = z["queries"]
  idx, _ = oi.search_batched(q)
  truth = np.argsort(-(q.astype(np.float64) @ a.dataset.astype(np.float64).T), axis=1)[:, :10]
  rec = np.mean([len(set(idx[i].tolist()) & set(truth[i].tolist())) / 10 for i in range(len(q))])
  assert rec > 0.9


def test_squared_l2_recall_and_distances():
  a, z = load_golden("l2_b16")
  oi = oracle.OracleIndex(a, 8, 100, 10)
  q = z["queries"]
  idx, dist = oi.search_batched(q)
  d2 = ((q.astype(np.float64)[:, None, :] - a.dataset.astype(np.float64)[None, :, :]) ** 2).sum(-1)
  truth = np.argsort(d2, axis=1)[:, :10]
  rec = np.mean([len(set(idx[i].tolist()) & set(truth[i].tolist())) / 10 for i in range(len(q))])
  assert rec > 0.9
  np.testing.assert_allclose(dist, np.take_along_axis(d2, idx.astype(np.int64), axis=1), rtol=1e-5)
  assert (np.diff(dist, axis=1) >= 0).all()          # ascending squared distances, not negated


def test_padding_when_fewer_candidates_than_k():
  a, z = load_golden("dot_b16")
  oi = oracle.OracleIndex(a, 1, 5, 10)
  idx, dist = oi.search_batched(z["queries"][:3], final_nn=10, pre_nn=5, leaves=1)
  assert np.isnan(dist[:, 5:]).all() and (idx[:, 5:] == 0).all() and not np.isnan(dist[:, :5]).any()


# ---- int8 (fixed point) reordering: FixedPointFloatDense*ReorderingHelper ----
def _np_int8_distance(q, x8, mult, dot, norms=None):
  """Independent numpy restatement: q' = (1/mult) * q, eight fnmadd lanes, 4-wide step, HorizontalSum3X, scalar tail."""
  from helpers import f32, fma32
  inv = f32(np.float32(1.0) / mult)
  qp = f32(inv * q)
  n = len(q)
  xf = x8.astype(np.float32)                      # [rows, D]
  a = np.zeros((xf.shape[0], 8), np.float32)
  j = 0
  while j + 8 <= n:
    a = fma32(np.broadcast_to(-qp[None, j:j + 8], a.shape), xf[:, j:j + 8], a)
    j += 8
  if j + 4 <= n:
    a[:, :4] = fma32(np.broadcast_to(-qp[None, j:j + 4], (xf.shape[0], 4)), xf[:, j:j + 4], a[:, :4])
    j += 4
  r = f32(f32(f32(a[:, 0] + a[:, 4]) + f32(a[:, 2] + a[:, 6])) + f32(f32(a[:, 1] + a[:, 5]) + f32(a[:, 3] + a[:, 7])))
  while j < n:
    r = fma32(np.full_like(r, -qp[j]), xf[:, j], r)
    j += 1
  if dot:
    return r
  qn = np.float32((q.astype(np.float64) ** 2).sum())
  return f32(f32(qn + norms) + f32(np.float32(2.0) * r))


@pytest.mark.parametrize("name", ["dot_b16", "l2_b16"])
def test_int8_reordering_matches_numpy_and_f64(name):
  import copy
  from scann_b200 import index_build
  a, z = load_golden(name)
  a = copy.copy(a)
  db = a.dataset
  a.int8_dataset, a.int8_multipliers = index_build.int8_quantize(db)
  dot = a.distance == "dot_product"
  if not dot:
    a.dp_norms = index_build.squared_l2_norms(db)
  a.dataset = None
  # quantization: |x * mult - q| <= 0.5, the extreme value of every column maps to +-127
  v = db * a.int8_multipliers[None, :]
  assert np.abs(v - a.int8_dataset).max() <= 0.5 + 1e-4
  assert (np.abs(a.int8_dataset).max(0) == 127).all()
  oi = oracle.OracleIndex(a, int(z["probe"]), int(z["pre"]), int(z["k"]))
  q = z["queries"][:6]
  dps = np.arange(0, a.n, 7, dtype=np.uint32)
  for qi in q:
    got = oi.exact_distances(qi, dps)
    ref = _np_int8_distance(qi, a.int8_dataset[dps], a.int8_multipliers, dot, None if dot else a.dp_norms[dps])
    # the float64 emulation of the f32 fma double-rounds in rare cases: allow isolated 1-ulp differences
    same = got.view(np.uint32) == ref.view(np.uint32)
    assert same.mean() > 0.999
    np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-6)
    deq = a.int8_dataset[dps].astype(np.float64) / a.int8_multipliers.astype(np.float64)[None, :]
    if dot:
      truth = -(deq @ qi.astype(np.float64))
    else:  # the reference's formula: |q|^2 + |x|^2 (original row) - 2 <q, dequantized x>
      truth = (qi.astype(np.float64) ** 2).sum() + a.dp_norms[dps].astype(np.float64) - 2.0 * (deq @ qi.astype(np.float64))
    np.testing.assert_allclose(got, truth, rtol=1e-4, atol=1e-3)
  idx, dist = oi.search_batched(z["queries"])
  f = oracle.OracleIndex(load_golden(name)[0], int(z["probe"]), int(z["pre"]), int(z["k"]))
  fidx, _ = f.search_batched(z["queries"])
  assert np.mean([len(set(idx[i].tolist()) & set(fidx[i].tolist())) / idx.shape[1] for i in range(len(idx))]) > 0.9


def test_float_bruteforce_oracles_against_numpy_restatements():
  """so_bruteforce_f32 / so_bruteforce_f32_l2 against a numpy restatement of the many-to-many chains
  (many_to_many_impl.inc:236-257,417-426,522-567) -- bit-exact -- and against float64."""
  rng = np.random.default_rng(5)
  n, d, nq, k = 400, 37, 9, 7
  db = (rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
  q = rng.standard_normal((nq, d)).astype(np.float32)

  def fma(a, b, c):  # one rounding: exact product in float64 (24 x 24 bits), sum rounded once to float32 ...
    # ... float64 addition of an exact product and a float32 is not always exact; use Python's exact rationals
    from fractions import Fraction
    out = np.empty(a.shape, np.float32)
    for i in range(a.size):
      out.flat[i] = np.float32(float(Fraction(float(a.flat[i])) * Fraction(float(b.flat[i])) + Fraction(float(c.flat[i]))))
    return out

  # dot product: acc = 0; acc = fnmadd(q[j], x[j], acc)
  acc = np.zeros((nq, n), np.float32)
  for j in range(d):
    acc = fma(np.broadcast_to(-q[:, j:j + 1], (nq, n)).copy(), np.broadcast_to(db[None, :, j], (nq, n)).copy(), acc)
  idx, dist = oracle.bruteforce_f32(db, q, k)
  order = np.lexsort((np.broadcast_to(np.arange(n), (nq, n)), acc), axis=1)[:, :k]
  np.testing.assert_array_equal(idx, order.astype(np.uint32))
  np.testing.assert_array_equal(dist.view(np.uint32), (-np.take_along_axis(acc, order, axis=1)).view(np.uint32))

  # squared L2: ||x||^2 by fnmadd chain times -1, ||q||^2 in double, then fnmadd(q[j], 2 x[j], acc)
  xn = np.zeros(n, np.float32)
  for j in range(d):
    xn = fma(-db[:, j], db[:, j], xn)
  xn = xn * np.float32(-1.0)
  # ||q||^2: sequential double accumulation in dimension order
  qnf = np.empty(nq, np.float32)
  for i in range(nq):
    a = 0.0
    for j in range(d):
      a += float(q[i, j]) * float(q[i, j])
    qnf[i] = np.float32(a)
  acc = (xn[None, :] + qnf[:, None]).astype(np.float32)
  for j in range(d):
    acc = fma(np.broadcast_to(-q[:, j:j + 1], (nq, n)).copy(), np.broadcast_to((db[:, j] * np.float32(2.0))[None, :], (nq, n)).copy(), acc)
  idx, dist = oracle.bruteforce_f32(db, q, k, distance="squared_l2")
  order = np.lexsort((np.broadcast_to(np.arange(n), (nq, n)), acc), axis=1)[:, :k]
  np.testing.assert_array_equal(idx, order.astype(np.uint32))
  np.testing.assert_array_equal(dist.view(np.uint32), np.take_along_axis(acc, order, axis=1).view(np.uint32))
  d2 = ((q.astype(np.float64)[:, None, :] - db.astype(np.float64)[None, :, :]) ** 2).sum(-1)
  np.testing.assert_allclose(dist, np.take_along_axis(d2, idx.astype(np.int64), axis=1), rtol=1e-5)


# ---- int8 (FIXED_POINT_INT8) query tokenization: tree(quantize_centroids=True) -------------------------------------

def test_fixed_point_centers_two_restatements_agree():
  """KMeansTreeNode::CreateFixedPointCenters: the oracle's C restatement against the numpy one."""
  from scann_b200 import index_build
  for L, D in [(7, 5), (100, 100), (33, 17), (64, 128), (1, 3)]:
    a, _ = _i8_tok_arrays(L, D, "dot_product", seed=L)
    got = oracle.quantize_centers(a.centers)
    want = index_build.quantize_centers(a.centers)
    for g, w in zip(got, want):
      np.testing.assert_array_equal(g.view(np.uint8 if g.dtype == np.int8 else np.uint32),
                                    w.view(np.uint8 if w.dtype == np.int8 else np.uint32))
    assert got[0].min() >= -127 and got[0].max() <= 127 and np.all(got[1][1] == 1.0)


# L mod 3 = 0 / 1 / 2 (the last L mod 3 centres take the one-to-one kernel); D covers every tail of both kernels:
# 100 = 6 x 16 + 4, 17 = 16 + 1, 23 = 16 + 4 + 3, 31 = 16 + 8 + 4 + 3, 12 = 8 + 4, 7 = 4 + 3, 3, 128, 64
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("L,D", [(99, 100), (100, 100), (101, 17), (32, 23), (40, 31), (20, 12), (11, 7), (5, 3),
                                 (64, 128), (50, 64), (2, 24), (1, 40)])
def test_int8_tokenization_matches_numpy_restatement(L, D, distance):
  from helpers import np_center_distances_i8
  a, q = _i8_tok_arrays(L, D, distance, seed=L * 1000 + D)
  P = min(L, 7)
  oi = oracle.OracleIndex(a, P, 10, 5)
  leaf, cdist = oi.tokenize(q, leaves=L)                     # every centre: the full distance row
  want = np_center_distances_i8(q, a.centers, distance)
  for i in range(len(q)):
    np.testing.assert_array_equal(cdist[i].view(np.uint32), want[i, leaf[i]].view(np.uint32))
    np.testing.assert_array_equal(leaf[i], np.lexsort((np.arange(L), want[i])))
  # and it is not the float tokenization: distances differ (quantization error), the nearest centre mostly agrees
  a.int8_tokenization = False
  of = oracle.OracleIndex(a, P, 10, 5)
  leaf_f, cdist_f = of.tokenize(q, leaves=L)
  if D >= 12:
    assert not np.array_equal(cdist_f, cdist)
    order_f = np.argsort(leaf_f, axis=1)
    order_i = np.argsort(leaf, axis=1)
    df, di = np.take_along_axis(cdist_f, order_f, 1), np.take_along_axis(cdist, order_i, 1)
    np.testing.assert_allclose(di, df, rtol=0.1, atol=0.05 * np.abs(df).max())
