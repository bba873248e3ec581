"""GPU parity of int8 (FIXED_POINT_INT8) query tokenization -- tree(quantize_centroids=True), SURVEY 8f rank 4:
KMeansTreeNode::GetAllDistancesInt8 (trees/kmeans_tree/kmeans_tree_node.h:222-256) on the centres
KMeansTreeNode::CreateFixedPointCenters derives (kmeans_tree_node.cc:267-281).  The library quantizes the centres itself
(csrc/index.cu) and tokenizes with tokenize_i8_kernel / tokenize_i8_tail_kernel (csrc/prep.cu); everything is compared
bit for bit with the oracle, which gets its fixed-point centres from its own C restatement."""
import numpy as np
import pytest

from conftest import get_case
from helpers import i8_tok_arrays

pytestmark = pytest.mark.gpu


# L mod 3 = 0 / 1 / 2 (the last L mod 3 centres take the reference's one-to-one kernel); D covers every tail of both
# kernels (see tests/test_oracle.py); L = 700 is several centre tiles and would take the tensor-core path if float
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("L,D", [(99, 100), (100, 100), (101, 17), (32, 23), (40, 31), (20, 12), (11, 7), (5, 3),
                                 (64, 128), (50, 64), (2, 24), (1, 40), (700, 96), (335, 50)])
def test_int8_tokenization_full_rows_bit_exact(L, D, distance):
  import oracle
  from scann_b200 import _lib
  a, q = i8_tok_arrays(L, D, distance, seed=L * 1000 + D)
  rng = np.random.default_rng(5)
  q = np.concatenate([q, rng.standard_normal((70, D)).astype(np.float32), np.zeros((1, D), np.float32)])
  oi = oracle.OracleIndex(a, min(L, 7), 10, 5)
  ni = _lib.NativeIndex(a, min(L, 7), 10, 5)
  for P in sorted({L, min(L, 7), 1}):
    l0, d0 = oi.tokenize(q, leaves=P)
    l1, d1 = ni.tokenize(q, leaves=P)
    np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  ni.close()


# Both routes of the library: the exact SIMT kernels, and the tcgen05 pre-filter (the int8 centres as an exact bf16
# operand) + radix refinement with the int8 chain -- the default from 256 centres.  L = 700: row image in shared
# memory, 128 threads per query; 335: L % 4 != 0, no staged row; 4100: 256 threads; 2000 x 100: the C2 shape.  The zero
# query has every approximate distance equal: more candidates than the buffer, the exact all-centres fallback.
# "chunk": the chunk pre-selection of long rows (default from 4096 centres; forced here wherever L / 32 >= 2 P).
@pytest.mark.parametrize("route", ["simt", "tcgen05", "chunk"])
@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
@pytest.mark.parametrize("L,D", [(700, 96), (335, 50), (4100, 33), (2000, 100), (301, 128), (1000, 7), (8201, 40)])
def test_int8_tokenization_routes_bit_exact(L, D, distance, route, monkeypatch):
  import oracle
  from scann_b200 import _lib
  monkeypatch.setenv("SCANN_B200_TOKENIZE", route)
  a, q = i8_tok_arrays(L, D, distance, seed=L * 1000 + D)
  rng = np.random.default_rng(6)
  q = np.concatenate([q, rng.standard_normal((150, D)).astype(np.float32) * rng.uniform(0.01, 30.0, (150, 1)).astype(np.float32),
                      np.zeros((1, D), np.float32), a.centers[:5], -a.centers[5:8]])
  oi = oracle.OracleIndex(a, 10, 10, 5)
  ni = _lib.NativeIndex(a, 10, 10, 5)
  for P in (1, 10, 100, 250):
    l0, d0 = oi.tokenize(q, leaves=P)
    l1, d1 = ni.tokenize(q, leaves=P)
    np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  ni.close()


def test_int8_tokenization_near_ties_and_duplicate_centres(monkeypatch):
  """Both refinements.  Centres that quantize to the same int8 row (exact ties, broken by index) and centres a hair apart: the window of
  the pre-filter must keep every one of them for the exact chain to order."""
  import oracle
  from scann_b200 import _lib
  for distance, route in (("dot_product", "tcgen05"), ("squared_l2", "tcgen05"), ("dot_product", "chunk"), ("squared_l2", "chunk")):
    monkeypatch.setenv("SCANN_B200_TOKENIZE", route)
    a, q = i8_tok_arrays(600, 64, distance, seed=77)
    rng = np.random.default_rng(8)
    a.centers[300:400] = a.centers[:100]                                     # exact duplicates
    a.centers[400:500] = a.centers[100:200] * np.float32(1.0 + 1e-6)        # same int8 row, other float norm
    a.centers[500:600] = a.centers[200:300] + rng.standard_normal((100, 64)).astype(np.float32) * np.float32(1e-3)
    q = np.concatenate([q, rng.standard_normal((120, 64)).astype(np.float32)])
    oi = oracle.OracleIndex(a, 10, 10, 5)
    ni = _lib.NativeIndex(a, 10, 10, 5)
    for P in (1, 16, 200):
      l0, d0 = oi.tokenize(q, leaves=P)
      l1, d1 = ni.tokenize(q, leaves=P)
      np.testing.assert_array_equal(l0, l1)
      np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
    ni.close()


CASES = [
    dict(int8_tok=True),                                                     # C1-like, L = 100 (L mod 3 = 1)
    dict(int8_tok=True, soar=1.5),                                           # SOAR: the leaf bias is the int8 distance
    dict(int8_tok=True, distance="squared_l2", d=64, leaves=50, n=10000),    # L mod 3 = 2, squared L2
    dict(int8_tok=True, n=6000, leaves=300, probe=300, pre=50, d=32),        # every leaf probed
    dict(int8_tok=True, dpb=4, d=96, leaves=64, n=12000),
    dict(int8_tok=True, n=20000, leaves=400, probe=30, pre=100, soar=1.5),   # >= 256 centres: the tcgen05 route by default
    dict(int8_tok=True, distance="squared_l2", d=64, leaves=320, n=16000, probe=20),
]


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_int8_tokenized_search_bit_exact(kw):
  c = get_case(**kw)
  l0, d0 = c.oracle.tokenize(c.q)
  l1, d1 = c.native.tokenize(c.q)
  np.testing.assert_array_equal(l0, l1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  a = c.oracle.candidates(c.q)
  b = c.native.candidates(c.q)
  np.testing.assert_array_equal(a["count"], b["count"])
  for i in range(len(c.q)):
    n = a["count"][i]
    np.testing.assert_array_equal(a["leaf"][i, :n], b["leaf"][i, :n])
    np.testing.assert_array_equal(a["dp"][i, :n], b["dp"][i, :n])
    np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q)
  i1, d1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  # it IS another tokenization: the float searcher over the same arrays sees other centre distances
  f = get_case(**{k: v for k, v in kw.items() if k != "int8_tok"})
  _, df = f.native.tokenize(c.q)
  assert not np.array_equal(df.view(np.uint32), d1.view(np.uint32))
  st = c.native.stats()
  assert st["kernel_launches"] >= 8


@pytest.mark.parametrize("mode", [0, 1], ids=["by_id", "by_leaf"])
def test_int8_tokenized_sharded_protocol_equals_unsharded(mode):
  """The sharded protocol tokenizes slices of the batch on every rank: the int8 kernels run there too."""
  from scann_b200 import _lib, distributed as sd
  c = get_case(int8_tok=True, soar=1.5)
  i1, d1 = c.native.search_batched(c.q)
  shards = [_lib.NativeIndex(c.arrays, c.probe, c.pre, c.k, device=0, shard_rank=r, shard_world=3, shard_mode=mode)
            for r in range(3)]
  i2, d2, _ = sd.search_sharded_local(shards, c.q, c.k)
  np.testing.assert_array_equal(i1, i2)
  np.testing.assert_array_equal(d1.view(np.uint32), d2.view(np.uint32))
  for s in shards:
    s.close()


@pytest.mark.parametrize("distance", ["dot_product", "squared_l2"])
def test_builder_quantize_centroids_round_trip(tmp_path, distance):
  """scann_builder.tree(quantize_centroids=True) (scann_builder.py:231): build, search, serialize, load -- and the
  searcher equals the oracle over the same assets with int8 tokenization, not the float one."""
  import oracle
  from scann_b200 import scann_ops_pybind
  rng = np.random.default_rng(11)
  db = rng.standard_normal((4000, 24)).astype(np.float32)
  q = rng.standard_normal((33, 24)).astype(np.float32)
  s = (scann_ops_pybind.builder(db, 10, distance).tree(40, 8, min_partition_size=10, quantize_centroids=True)
       .score_ah(2).reorder(60).build())
  assert "FIXED_POINT_INT8" in s.config()
  i0, d0 = s.search_batched(q)
  arrays = s.searcher._arrays
  assert arrays.int8_tokenization
  oi = oracle.OracleIndex(arrays, 8, 60, 10)
  oi_idx, oi_dist = oi.search_batched(q)
  np.testing.assert_array_equal(i0, oi_idx)
  np.testing.assert_array_equal(d0.view(np.uint32), oi_dist.view(np.uint32))
  s.serialize(str(tmp_path))
  loaded = scann_ops_pybind.load_searcher(str(tmp_path))
  assert loaded.searcher._arrays.int8_tokenization
  i1, d1 = loaded.search_batched(q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
