"""The oracle against the REFERENCE'S OWN CODE (oracle/_ref/libscann_ref.so, built by oracle/Makefile from
/root/reference: the AVX2 LUT16 kernel of hashes/internal/lut16_avx2.inc with the reference's SIMD wrappers, the LUT
fixed-point conversion and the code packing of hashes/internal/asymmetric_hashing_impl.cc, utils/bfloat16_helpers.h).

This is what pins oracle/scann_oracle.c to the reference for the integer / byte part of the path: packed codes, u8 LUT,
int16 scores, the float score (two roundings), the candidate contract.  Skipped where the library is absent (a checkout
without /root/reference and without the prebuilt file).
"""
import numpy as np
import pytest

from conftest import get_case
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libscann_ref.so not built (needs /root/reference)")


@pytest.mark.parametrize("n,B", [(1, 1), (31, 2), (32, 3), (33, 16), (77, 25), (500, 48), (1000, 50), (257, 64),
                                 (96, 127), (64, 128), (40, 200), (33, 256)])
def test_reference_int16_scores_equal_the_lookup_sum(n, B):
  """LUT16Avx2<1..3>::GetInt16Distances on CreatePackedDataset's layout == sum_b lut[b][code_b] - 128 B, the statement
  the oracle's score_slot and the CUDA kernels implement (odd B exercises the SSE tail of the bottom loop, n % 32 the
  padded last group)."""
  rng = np.random.default_rng(n * 1000 + B)
  codes = rng.integers(0, 16, (n, B), dtype=np.uint8)
  packed = ref.pack_dataset(codes)
  for nq in (1, 2, 3):
    luts = [rng.integers(0, 256, (B, 16), dtype=np.uint8) for _ in range(nq)]
    if B > 128:  # keep the int16 accumulator exact as CanUseInt16Accumulator demands (asymmetric_hashing_impl.cc:656-688)
      luts = [(l.astype(np.int32) // 2 + 64).astype(np.uint8) for l in luts]
    got = ref.lut16_int16(packed, n, B, luts)[:, :n].astype(np.int32)
    want = np.stack([l[np.arange(B)[None, :], codes].astype(np.int32).sum(1) - 128 * B for l in luts])
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("kw", [dict(), dict(dpb=1, d=64), dict(dpb=3, d=50, leaves=40), dict(distance="squared_l2", d=64, leaves=50, n=10000)],
                         ids=["dot_b50", "dot_b64", "dot_varchunk_b17", "l2_b32"])
def test_oracle_leaf_scores_equal_the_reference_kernel(kw):
  """Whole leaves of a real index: the oracle's int16 scores == the reference kernel's, on the reference's packing of
  the same codes, under the oracle's own u8 LUTs."""
  c = get_case(**kw)
  lut, _ = c.oracle.lut(c.q[:6])
  for leaf in range(0, c.oracle.L, max(1, c.oracle.L // 7)):
    dps = c.oracle.leaf_datapoints(leaf)
    if len(dps) == 0:
      continue
    codes = c.arrays.codes[dps]
    packed = ref.pack_dataset(codes)
    got = ref.lut16_int16(packed, len(dps), codes.shape[1], [lut[0], lut[1], lut[2]])
    for j in range(3):
      np.testing.assert_array_equal(got[j, :len(dps)], c.oracle.leaf_scores(lut[j], leaf))


def test_lut_fixed_point_conversion_equals_the_reference():
  """so_lut_quantize == ConvertLookupToFixedPoint<uint8_t> (quantile 1.0, ROUND): bytes and multiplier bits, incl. the
  sqrt(FLT_EPSILON) floor, exact .5 products and the +-127 ends."""
  import ctypes as C
  import oracle
  L = oracle.lib()
  L.so_lut_quantize.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
  L.so_lut_quantize.restype = None
  rng = np.random.default_rng(3)
  tables = [rng.standard_normal(800).astype(np.float32), (rng.standard_normal(400) * 1e-6).astype(np.float32),
            np.zeros(32, np.float32), np.linspace(-2, 2, 255 * 2 + 1).astype(np.float32),
            (rng.integers(-254, 255, 1600) / 254.0).astype(np.float32), np.array([1e30, -1e30, 3.0], np.float32)]
  for raw in tables:
    want, wm = ref.lut_to_fixed_point(raw)
    got = np.zeros(raw.size, np.uint8)
    gm = np.zeros(1, np.float32)
    L.so_lut_quantize(raw.ctypes.data_as(C.c_void_p), raw.size, got.ctypes.data_as(C.c_void_p), gm.ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(got, want)
    assert gm.view(np.uint32)[0] == np.float32(wm).view(np.uint32)


def test_oracle_luts_equal_the_reference_conversion_of_their_raw_tables():
  """The oracle's u8 LUT of real queries, re-derived through the reference's conversion from a float64-accurate raw
  table: the multiplier differs by at most the raw table's own rounding, and wherever the raw entries agree the bytes do
  (the raw table's arithmetic -- the one-to-many kernel -- is pinned by tests/test_oracle.py)."""
  c = get_case()
  lut, mult = c.oracle.lut(c.q[:8])
  a = c.arrays
  for i in range(8):
    qb = c.q[i].reshape(a.codebook.shape[0], -1)
    raw = -np.einsum("bd,bcd->bc", qb.astype(np.float64), a.codebook.astype(np.float64)).astype(np.float32)
    want, wm = ref.lut_to_fixed_point(raw)
    assert abs(wm - mult[i]) <= 2e-6 * abs(wm)
    assert np.mean(want.reshape(lut[i].shape) == lut[i]) > 0.97
    assert np.max(np.abs(want.reshape(lut[i].shape).astype(int) - lut[i].astype(int))) <= 1


@pytest.mark.parametrize("kw", [dict(), dict(soar=1.5)], ids=["dot", "dot_soar"])
def test_oracle_candidates_equal_topn_of_the_reference_float_scores(kw):
  """The pre-reorder candidate list, end to end on reference arithmetic: for every probed leaf the reference kernel
  (GetTopFloatDistances, epsilon = +inf so that nothing is pruned) gives float(acc) * float(1 / mult) + bias for every
  datapoint; the N' smallest (score, leaf, slot) of those are exactly the oracle's candidates, score bits included."""
  c = get_case(**kw)
  nq = 6
  q = c.q[:nq]
  leaves, bias = c.oracle.tokenize(q)
  lut, mult = c.oracle.lut(q)
  cand = c.oracle.candidates(q)
  tokens = c.arrays.tokens
  for i in range(nq):
    rows = []
    for r in range(leaves.shape[1]):
      leaf = int(leaves[i, r])
      dps = c.oracle.leaf_datapoints(leaf)
      if len(dps) == 0:
        continue
      codes = c.arrays.codes[dps]
      if c.arrays.soar:  # the SOAR code row iff the leaf is the datapoint's second token
        second = tokens[2 * dps.astype(np.int64) + 1] == leaf
        codes = np.where(second[:, None], c.arrays.soar_codes[dps], codes)
      packed = ref.pack_dataset(codes)
      (idx, dist), = ref.lut16_top_float(packed, len(dps), codes.shape[1], [lut[i]], [bias[i, r]], [mult[i]])
      assert len(idx) == len(dps)  # nothing pruned: every real datapoint exactly once (the padded tail is masked)
      order = np.argsort(idx)
      for s, d in zip(idx[order], dist[order]):
        rows.append((np.float32(d), leaf, int(s)))
    # ascending (score, leaf, slot); -0.0 == +0.0 as in DistanceComparator
    rows.sort(key=lambda t: (float(t[0]), t[1], t[2]))
    n = int(cand["count"][i])
    want = rows[:n]
    np.testing.assert_array_equal(np.asarray([w[0] for w in want], np.float32).view(np.uint32),
                                  cand["score"][i, :n].view(np.uint32))
    np.testing.assert_array_equal([w[1] for w in want], cand["leaf"][i, :n])
    np.testing.assert_array_equal([w[2] for w in want], cand["slot"][i, :n])


def test_reference_prefilter_band_is_what_the_contract_says():
  """The reference's own candidate filter `acc < trunc((epsilon - bias) * mult)` (lut16_avx2.inc:432-438), run with the
  epsilon the exact contract ends at: everything it keeps has score <= epsilon + one quantum, and everything strictly
  better than epsilon minus one quantum is kept -- the band DESIGN.md section 2 describes, measured on reference code."""
  c = get_case()
  q = c.q[:4]
  leaves, bias = c.oracle.tokenize(q)
  lut, mult = c.oracle.lut(q)
  cand = c.oracle.candidates(q)
  for i in range(4):
    n = int(cand["count"][i])
    eps = float(cand["score"][i, n - 1])
    quantum = float(np.float32(1.0) / mult[i])
    for r in range(leaves.shape[1]):
      leaf = int(leaves[i, r])
      dps = c.oracle.leaf_datapoints(leaf)
      if len(dps) == 0:
        continue
      packed = ref.pack_dataset(c.arrays.codes[dps])
      (aidx, adist), = ref.lut16_top_float(packed, len(dps), c.arrays.codes.shape[1], [lut[i]], [bias[i, r]], [mult[i]])
      (kidx, kdist), = ref.lut16_top_float(packed, len(dps), c.arrays.codes.shape[1], [lut[i]], [bias[i, r]], [mult[i]], epsilon=eps)
      assert np.all(kdist <= eps + quantum * 1.001)
      kept = set(kidx.tolist())
      must = aidx[adist < eps - quantum * 1.001]
      assert set(must.tolist()) <= kept


def test_bfloat16_helpers_equal_the_reference():
  from scann_b200 import index_build
  rng = np.random.default_rng(5)
  x = np.concatenate([rng.standard_normal(100000).astype(np.float32) * 10.0 ** rng.integers(-30, 30, 100000).astype(np.float32),
                      np.array([0.0, -0.0, np.inf, -np.inf, 3.4028235e38, -3.4028235e38, 3.39e38, 1e-45, -1e-45, 1.00390625,
                                1.01171875, 65280.0, 65536.0], np.float32)])
  np.testing.assert_array_equal(index_build.bfloat16_quantize(x), ref.bf16_quantize(x))
  bits = rng.integers(-32768, 32768, 50000).astype(np.int16)
  want = ref.bf16_decompress(bits)
  got = (bits.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
  np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- the asymmetric one-to-many kernels (ref_glue_asym.cc): int8 tokenization, int8 / bfloat16 reordering -----------

needs_asym = pytest.mark.skipif(not (ref.available() and ref.has_asymmetric()),
                                reason="oracle/_ref/libscann_ref.so without the asymmetric kernels")


@needs_asym
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("L,D", [(99, 100), (100, 100), (101, 17), (32, 23), (40, 31), (20, 12), (11, 7), (5, 3),
                                 (64, 128), (50, 64), (2, 24), (1, 40), (700, 96)])
def test_int8_tokenization_equals_the_reference_kernel(L, D, distance):
  """KMeansTreeNode::GetAllDistancesInt8 (kmeans_tree_node.h:222-256) assembled around the REFERENCE'S OWN
  DenseDotProductDistanceOneToManyInt8Float: the oracle's centre distances, all L of them -- the three-at-a-time order
  and the one-to-one order of the last L mod 3 centres -- bit for bit.  D = 128 / 64 take the fixed-dimension template
  instances (one_to_many_asymmetric_impl.inc:681-695)."""
  import oracle
  from helpers import i8_tok_arrays
  from scann_b200 import index_build
  a, q = i8_tok_arrays(L, D, distance, seed=L * 1000 + D)
  ci8, inv, sqn = oracle.quantize_centers(a.centers)
  oi = oracle.OracleIndex(a, min(L, 7), 10, 5)
  leaf, cdist = oi.tokenize(q, leaves=L)
  l2 = distance == "squared_l2"
  for i in range(len(q)):
    qp = (q[i] * (inv * np.float32(2.0) if l2 else inv)).astype(np.float32)
    want = ref.one_to_many_int8_float(qp, ci8)
    if l2:
      qn = index_build.squared_l2_norms(q[i:i + 1])[0]
      want = (want + (qn + sqn).astype(np.float32)).astype(np.float32)
    np.testing.assert_array_equal(cdist[i].view(np.uint32), want[leaf[i]].view(np.uint32))


@needs_asym
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("D", [100, 64, 128, 23, 7])
def test_int8_reordering_distances_equal_the_reference_kernel(D, distance):
  """FixedPointFloatDense{DotProduct,SquaredL2}ReorderingHelper (utils/reordering_helper.cc:430-441,610-618): the
  oracle's int8 reordering distance against the reference's indexed one-to-many kernel.  The reference sends the last
  n mod 3 entries of a candidate list through its one-to-one kernel (another summation order, list order unspecified);
  the oracle states the three-at-a-time order for every row, so lists of 3 m candidates are compared."""
  import oracle
  from scann_b200 import index_build
  rng = np.random.default_rng(D)
  n = 600
  db = (rng.standard_normal((n, D)) * rng.uniform(0.1, 4.0, D)[None, :]).astype(np.float32)
  a = index_build.IndexArrays(distance=distance, dataset=None, n=n, d=D)
  a.int8_dataset, a.int8_multipliers = index_build.int8_quantize(db)
  if distance == "squared_l2":
    a.dp_norms = index_build.squared_l2_norms(db)
  a.centers = db[:4].copy()
  a.tokens = (np.arange(n) % 4).astype(np.int32)
  a.codes = rng.integers(0, 16, (n, D), dtype=np.uint8)
  a.codebook = rng.standard_normal((D, 16, 1)).astype(np.float32)
  a.block_dims = np.ones(D, np.int32)
  a.soar, a.soar_codes, a.overretrieve, a.residual = False, None, 2.0, distance == "dot_product"
  oi = oracle.OracleIndex(a, 2, 30, 10)
  inv = (np.float32(1.0) / a.int8_multipliers).astype(np.float32)
  for qi in range(5):
    q = rng.standard_normal(D).astype(np.float32)
    dps = rng.permutation(n)[:300].astype(np.uint32)      # 3 m candidates
    got = oi.exact_distances(q, dps)
    val = ref.one_to_many_int8_float((inv * q).astype(np.float32), a.int8_dataset, dps)
    if distance == "squared_l2":
      qn = index_build.squared_l2_norms(q[None, :])[0]
      val = ((qn + a.dp_norms[dps]).astype(np.float32) + (np.float32(2.0) * val).astype(np.float32)).astype(np.float32)
    np.testing.assert_array_equal(got.view(np.uint32), val.view(np.uint32))


@needs_asym
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("D", [100, 64, 128, 23, 7, 768])
def test_bfloat16_reordering_distances_equal_the_reference_kernel(D, distance):
  """Bfloat16ReorderingHelper (utils/reordering_helper.cc:745-757): f32 query x bf16 row, the oracle against
  DenseDotProductDistanceOneToManyBf16Float / OneToManyBf16FloatSquaredL2 over 3 m rows (see the int8 test)."""
  import oracle
  from scann_b200 import index_build
  rng = np.random.default_rng(D + 1)
  n = 300
  db = rng.standard_normal((n, D)).astype(np.float32)
  a = index_build.IndexArrays(distance=distance, dataset=None, n=n, d=D)
  a.bf16_dataset = index_build.bfloat16_quantize(db)
  B = min(D, 64)
  a.centers = db[:4].copy()
  a.tokens = (np.arange(n) % 4).astype(np.int32)
  dims = np.full(B, D // B, np.int32)
  dims[:D % B] += 1
  a.codes = rng.integers(0, 16, (n, B), dtype=np.uint8)
  a.codebook = rng.standard_normal((B, 16, int(dims.max()))).astype(np.float32)
  a.block_dims = dims
  a.soar, a.soar_codes, a.overretrieve, a.residual = False, None, 2.0, distance == "dot_product"
  oi = oracle.OracleIndex(a, 2, 30, 10)
  for qi in range(5):
    q = rng.standard_normal(D).astype(np.float32)
    got = oi.exact_distances(q, np.arange(n, dtype=np.uint32))
    want = ref.one_to_many_bf16_float(q, a.bf16_dataset, squared_l2=distance == "squared_l2")
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- the symmetric float one-to-many kernel (ref_glue_sym.cc): exact f32 reordering, LUT build for dims >= 8 --------

needs_sym = pytest.mark.skipif(not (ref.available() and ref.has_symmetric()),
                               reason="oracle/_ref/libscann_ref.so without the symmetric float kernel")


@needs_sym
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("D", [8, 9, 10, 11, 12, 13, 14, 15, 16, 24, 50, 64, 96, 100, 101, 102, 103, 128, 768])
def test_f32_reordering_distances_equal_the_reference_kernel(D, distance):
  """ExactReorderingHelper (utils/reordering_helper.cc) -> DenseDotProductDistanceOneToMany / DenseSquaredL2Distance
  OneToMany -> DenseAccumulatingDistanceMeasureOneToManyInternalAvx2 (one_to_many_symmetric.h:373-503): the oracle's
  exact f32 reordering distance against the reference's compiled kernel, every tail of the dimension loop (8-wide,
  4-wide, 2-wide, the last scalar dim).  3 m rows: the last n mod 3 rows of a call take a one-to-one kernel whose place
  in a candidate list is unspecified (see the oracle)."""
  import oracle
  from scann_b200 import index_build
  rng = np.random.default_rng(D + 7)
  n = 300
  db = (rng.standard_normal((n, D)) * rng.uniform(0.1, 4.0, D)[None, :]).astype(np.float32)
  a = index_build.IndexArrays(distance=distance, dataset=db, n=n, d=D)
  B = min(D, 64)
  dims = np.full(B, D // B, np.int32)
  dims[:D % B] += 1
  a.centers = db[:4].copy()
  a.tokens = (np.arange(n) % 4).astype(np.int32)
  a.codes = rng.integers(0, 16, (n, B), dtype=np.uint8)
  a.codebook = rng.standard_normal((B, 16, int(dims.max()))).astype(np.float32)
  a.block_dims = dims
  a.soar, a.soar_codes, a.overretrieve, a.residual = False, None, 2.0, distance == "dot_product"
  oi = oracle.OracleIndex(a, 2, 30, 10)
  for qi in range(5):
    q = rng.standard_normal(D).astype(np.float32)
    got = oi.exact_distances(q, np.arange(n, dtype=np.uint32))
    want = ref.one_to_many_f32(q, db, squared_l2=distance == "squared_l2")
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


@needs_sym
def test_squared_l2_norm_equals_the_reference_reduction():
  """SquaredL2Norm = DenseSingleAccumulate(v, Square()) (utils/reduction.h:357-390): what dp_norms.npy holds, the query
  norm of squared-L2 tokenization (many_to_many_impl.inc:417-426) and the centre norms of int8 tokenization.  The numpy
  restatement the index build uses and, through the squared-L2 tokenization distances, the oracle's C restatement."""
  import oracle
  from helpers import i8_tok_arrays
  from scann_b200 import index_build
  rng = np.random.default_rng(3)
  for D in [1, 2, 3, 4, 5, 6, 7, 8, 9, 30, 96, 100, 101, 102, 103, 768]:
    x = (rng.standard_normal((40, D)) * 10.0 ** rng.integers(-3, 4, (40, 1))).astype(np.float32)
    want = np.asarray([np.float32(ref.squared_l2_norm(r)) for r in x], np.float32)
    np.testing.assert_array_equal(index_build.squared_l2_norms(x).view(np.uint32), want.view(np.uint32))
  # the oracle's own C restatement, observed through so_quantize_centers' squared norms
  for L, D in [(33, 17), (50, 100), (9, 7)]:
    a, _ = i8_tok_arrays(L, D, "squared_l2", seed=D)
    _, _, sqn = oracle.quantize_centers(a.centers)
    want = np.asarray([np.float32(ref.squared_l2_norm(r)) for r in a.centers], np.float32)
    np.testing.assert_array_equal(sqn.view(np.uint32), want.view(np.uint32))


# ---- the batched float tokenization chain (ref_glue_m2m.cc) ---------------------------------------------------------

needs_m2m = pytest.mark.skipif(not (ref.available() and ref.has_many_to_many()),
                               reason="oracle/_ref/libscann_ref.so without the many-to-many pieces")


@needs_m2m
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("L,D", [(100, 100), (33, 17), (50, 64), (7, 3), (16, 128), (17, 96), (1, 40), (257, 30)])
def test_float_tokenization_equals_the_reference_many_to_many(L, D, distance):
  """KMeansTreePartitioner::TokensForDatapointWithSpillingBatched -> DenseDistanceManyToManyTopK: the oracle's centre
  distances (all L, ragged last block of 2 x 8 centres included) against the reference's own AugmentWithL2Norms /
  DoAccumulationTransposedTemplate and SquaredL2Norm, bit for bit."""
  import oracle
  from helpers import i8_tok_arrays
  a, q = i8_tok_arrays(L, D, distance, seed=L * 100 + D)
  a.int8_tokenization = False
  oi = oracle.OracleIndex(a, min(L, 7), 10, 5)
  leaf, cdist = oi.tokenize(q, leaves=L)
  want = ref.many_to_many_f32(q, a.centers, squared_l2=distance == "squared_l2")
  for i in range(len(q)):
    np.testing.assert_array_equal(cdist[i].view(np.uint32), want[i, leaf[i]].view(np.uint32))


@needs_m2m
def test_database_tokenization_distances_equal_the_reference_many_to_many():
  """The index build assigns every datapoint to its nearest centre under squared L2 through the same kernel
  (kmeans_tree_partitioner.cc:561-612): the oracle's primary assignment is the argmin of the reference's accumulators."""
  import oracle
  rng = np.random.default_rng(12)
  for n, L, D in [(500, 37, 24), (300, 100, 100), (200, 16, 7)]:
    x = rng.standard_normal((n, D)).astype(np.float32)
    centers = rng.standard_normal((L, D)).astype(np.float32)
    want = ref.many_to_many_f32(x, centers, squared_l2=True)
    tok, dist = oracle.assign_primary(x, centers, threads=1)
    np.testing.assert_array_equal(tok, want.argmin(1))
    np.testing.assert_array_equal(dist.view(np.uint32), want.min(1).view(np.uint32))


@needs_m2m
@pytest.mark.parametrize("n,L,D,lam", [(400, 37, 24, 1.5), (300, 100, 100, 1.0), (200, 16, 7, 2.5), (150, 33, 64, 0.0)])
def test_soar_assignment_equals_the_reference_arithmetic(n, L, D, lam):
  """The index build's SOAR secondary assignment (KMeansTreePartitioner::OrthogonalityAmplifiedTokenForDatapointBatched,
  partitioning/kmeans_tree_partitioner.cc:925-997): the oracle's choice and cost against the argmin / min of the costs the
  reference's own ComputeNormalizedResidual + DenseManyToManyOrthogonalityAmplified accumulation produce (first strict
  minimum over all centres).  A datapoint that coincides with its primary centre has a zero residual."""
  import oracle
  if not ref.has_soar_costs():
    pytest.skip("library without ref_soar_costs")
  rng = np.random.default_rng(n + L)
  x = rng.standard_normal((n, D)).astype(np.float32)
  centers = rng.standard_normal((L, D)).astype(np.float32)
  primary, _ = oracle.assign_primary(x, centers, threads=1)
  x[0] = centers[primary[0]]                                  # sqnorm < 1e-7: rhat = 0
  want_cost, _ = ref.soar_costs(x, centers, primary, lam)
  tok, cost = oracle.assign_soar(x, centers, primary, lam, threads=1)
  np.testing.assert_array_equal(tok, want_cost.argmin(1))
  np.testing.assert_array_equal(cost.view(np.uint32), want_cost.min(1).view(np.uint32))


# ---- the noise-shaped AH encoder (ref_glue_ns.cc) -------------------------------------------------------------------

needs_ns = pytest.mark.skipif(not (ref.available() and ref.has_noise_shaped()),
                              reason="oracle/_ref/libscann_ref.so without the noise-shaped encoder")


@needs_ns
@pytest.mark.parametrize("D,dpb,residual,threshold", [(32, 2, True, 0.2), (100, 2, True, 0.2), (100, 3, True, 0.2),
                                                      (64, 4, False, 0.3), (96, 8, True, 0.5), (40, 1, True, 0.2)])
def test_noise_shaped_codes_equal_the_reference_encoder(D, dpb, residual, threshold):
  """Indexer::HashWithNoiseShaping -> AhImpl<float>::IndexDatapointNoiseShaped (asymmetric_hashing_impl.cc:434-503): the
  oracle's noise-shaped codes against the reference's own residual statistics, cost multiplier, block order
  (utils/zip_sort.h) and coordinate descent, on clustered data with trained-looking codebooks (VARIABLE_CHUNK when dpb
  does not divide D)."""
  import oracle
  rng = np.random.default_rng(D * 10 + dpb)
  n, L = 600, 12
  centers = rng.standard_normal((L, D)).astype(np.float32)
  token = rng.integers(0, L, n).astype(np.int32)
  x = centers[token] + 0.3 * rng.standard_normal((n, D))
  x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)      # unit rows: the threshold is relative to |x|
  centers = (centers / np.linalg.norm(centers, axis=1, keepdims=True)).astype(np.float32)
  B = (D + dpb - 1) // dpb
  bd = np.full(B, dpb, np.int32)
  if D % dpb:
    bd[-1] = D % dpb
  scale = (0.3 if residual else 1.0) / np.sqrt(D)
  codebook = (scale * rng.standard_normal((B, 16, dpb))).astype(np.float32)
  for b in range(B):
    codebook[b, :, bd[b]:] = 0.0
  cen = centers if residual else None
  tok = token if residual else None
  got, ties = oracle.encode(x, codebook, bd, cen, tok, threshold, threads=1)
  want = ref.encode_noise_shaped(x, codebook, bd, cen, tok, threshold)
  np.testing.assert_array_equal(got, want)
  plain, _ = oracle.encode(x, codebook, bd, cen, tok, float("nan"), threads=1)
  assert (plain != want).mean() > 0.01        # the shaping does move codes: the test is not vacuous


# ---- a lookup table computed by reference code only (dims per block >= 8 and even) ----------------------------------

@pytest.mark.skipif(not (ref.available() and ref.has_symmetric() and ref.has_sse4_one_to_one()),
                    reason="oracle/_ref/libscann_ref.so without the float one-to-many / one-to-one kernels")
@pytest.mark.parametrize("kw", [dict(dpb=8, d=128, leaves=32, n=6000), dict(dpb=16, d=64, leaves=20, n=4000),
                                dict(distance="squared_l2", dpb=8, d=64, leaves=20, n=4000)],
                         ids=["dot_dpb8", "dot_dpb16", "l2_dpb8"])
def test_lookup_tables_equal_the_reference_pipeline(kw):
  """AsymmetricQueryer::CreateLookupTable for blocks of >= 8 dims, assembled from REFERENCE code only: centres 0..14 of a
  block through DenseAccumulatingDistanceMeasureOneToManyInternalAvx2 (three at a time), centre 15 (16 mod 3 = 1) through
  the SSE4 one-to-one kernel (negated for dot product, one_to_many_symmetric.h:793-799), the whole float table through
  ConvertLookupToFixedPoint<uint8_t> -- against the oracle's u8 tables and multipliers, bit for bit."""
  c = get_case(**kw)
  l2 = c.arrays.distance == "squared_l2"
  cb = c.arrays.codebook                                       # [B, 16, dpb]
  B, _, dpb = cb.shape
  q = c.q[:8]
  lut, mult = c.oracle.lut(q)
  for i in range(len(q)):
    raw = np.empty((B, 16), np.float32)
    for b in range(B):
      qb = q[i, b * dpb:(b + 1) * dpb]
      raw[b, :15] = ref.one_to_many_f32(qb, cb[b, :15], squared_l2=l2)
      one = ref.one_to_one_sse4(qb, cb[b, 15], squared_l2=l2)
      raw[b, 15] = np.float32(one if l2 else -one)
    want_lut, want_mult = ref.lut_to_fixed_point(raw)
    np.testing.assert_array_equal(np.float32(mult[i]).view(np.uint32), np.float32(want_mult).view(np.uint32))
    np.testing.assert_array_equal(lut[i].reshape(-1), want_lut)


@pytest.mark.skipif(not (ref.available() and ref.has_symmetric() and ref.has_sse4_one_to_one() and ref.has_many_to_many()),
                    reason="oracle/_ref/libscann_ref.so without the float kernels")
@pytest.mark.parametrize("kw", [dict(dpb=8, d=128, leaves=32, n=6000),
                                dict(distance="squared_l2", dpb=8, d=64, leaves=20, n=4000, probe=6, pre=64)],
                         ids=["dot_dpb8", "l2_dpb8"])
def test_whole_search_from_reference_arithmetic_only(kw):
  """Capstone: SearchBatched of a tree-AH index with 8-dim blocks where EVERY number comes from compiled reference code --
  centre distances (many-to-many accumulation), the lookup table (one-to-many AVX2 kernel, SSE4 one-to-one kernel,
  fixed-point conversion), the per-leaf scores (LUT16 AVX2 kernel on the reference's packing), the exact reordering
  distances (one-to-many AVX2 kernel) -- and only the selections (top P leaves, top N' candidates by (score, leaf, slot)
  or (score, id), top k by (distance, id)) are written here.  It equals the oracle's search_batched, ids and distance
  bits: the oracle -- and through the `-m gpu` tests the CUDA path -- computes what the reference's arithmetic computes."""
  from helpers import reference_pipeline_search
  c = get_case(**kw)
  q = c.q[:6]
  want_idx, want_dist = c.oracle.search_batched(q)
  got_idx, got_dist = reference_pipeline_search(c, q)
  np.testing.assert_array_equal(got_idx, want_idx)
  np.testing.assert_array_equal(got_dist.view(np.uint32), want_dist.view(np.uint32))
