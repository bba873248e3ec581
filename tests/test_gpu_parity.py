"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from conftest import get_case

pytestmark = pytest.mark.gpu

CASES = [
    dict(),                                         # C1-like: D=100, dpb=2 -> B=50 (W=7, partial word)
    dict(dpb=3, d=100),                             # VARIABLE_CHUNK: B=34, last block 1 dim
    dict(dpb=1, d=40, leaves=50, n=8000),           # B=40
    dict(dpb=4, d=96, leaves=64, n=12000),          # B=24 (W=3), Highway 4-lane LUT path
    dict(dpb=8, d=128, leaves=32, n=6000),          # B=16, AVX2-order LUT path
    dict(soar=1.5),                                 # SOAR spilled leaves + dedup
    dict(n=3000, leaves=300, probe=40, pre=150),    # tiny / empty leaves, pilot spans many leaves
    dict(n=6000, leaves=300, probe=300, pre=50, d=32),               # every leaf probed (P > 256: block top-P path)
    dict(distance="squared_l2", d=64, leaves=50, n=10000),           # TreeXHybridSMMD semantics (C4 shape family)
    dict(distance="squared_l2", d=30, dpb=4, leaves=20, n=5000, probe=6, pre=64),
    # 128 < B <= 256 (still the reference's int16 accumulator, asymmetric_hashing_impl.cc:656-688): the generic kernels
    dict(dpb=1, d=130, leaves=40, n=6000),          # B=130 (W=17, two blocks in the last word)
    dict(dpb=2, d=512, leaves=30, n=4000, soar=1.5),  # B=256 (W=32): sums up to 65280 in the u16 lanes
    dict(distance="squared_l2", dpb=1, d=200, leaves=25, n=5000),  # B=200 (W=25)
]


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_tokenize_bit_exact(kw):
  c = get_case(**kw)
  l0, d0 = c.oracle.tokenize(c.q)
  l1, d1 = c.native.tokenize(c.q)
  np.testing.assert_array_equal(l0, l1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_lut_bit_exact(kw):
  c = get_case(**kw)
  t0, m0 = c.oracle.lut(c.q)
  t1, m1 = c.native.lut(c.q)
  np.testing.assert_array_equal(m0.view(np.uint32), m1.view(np.uint32))
  np.testing.assert_array_equal(t0, t1)


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_lut16_int16_scores_bit_exact(kw):
  c = get_case(**kw)
  luts, _ = c.oracle.lut(c.q[:3])
  L = c.arrays.centers.shape[0]
  for qi in range(3):
    for leaf in list(range(0, L, max(1, L // 7))) + [L - 1]:
      assert c.native.leaf_size(leaf) == c.oracle.leaf_size(leaf)
      s0 = c.oracle.leaf_scores(luts[qi], leaf)
      s1 = c.native.leaf_scores(luts[qi], leaf)
      np.testing.assert_array_equal(s0, s1)


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_pre_reorder_candidates_bit_exact(kw):
  c = get_case(**kw)
  a = c.oracle.candidates(c.q)
  b = c.native.candidates(c.q)
  np.testing.assert_array_equal(a["count"], b["count"])
  for i in range(len(c.q)):
    n = a["count"][i]
    np.testing.assert_array_equal(a["leaf"][i, :n], b["leaf"][i, :n])
    np.testing.assert_array_equal(a["slot"][i, :n], b["slot"][i, :n])
    np.testing.assert_array_equal(a["dp"][i, :n], b["dp"][i, :n])
    np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_search_batched_ids_and_distances(kw):
  c = get_case(**kw)
  i0, d0 = c.oracle.search_batched(c.q)
  i1, d1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  # float distances within 1e-5 relative of a float64 recomputation (north star tolerance)
  rows = c.db[i1.astype(np.int64)].astype(np.float64)
  if c.arrays.distance == "dot_product":
    truth = np.einsum("qd,qkd->qk", c.q.astype(np.float64), rows)
  else:
    truth = ((c.q.astype(np.float64)[:, None, :] - rows) ** 2).sum(-1)
  np.testing.assert_allclose(d1, truth, rtol=1e-5, atol=1e-5)


def test_parameter_overrides_and_padding():
  c = get_case()
  i0, d0 = c.oracle.search_batched(c.q, final_nn=20, pre_nn=50, leaves=3)
  i1, d1 = c.native.search_batched(c.q, final_nn=20, pre_nn=50, leaves=3)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  # more neighbours requested than candidates exist -> (0, NaN) padding (scann.h:175-178)
  i2, d2 = c.native.search_batched(c.q[:4], final_nn=30, pre_nn=8, leaves=1)
  assert np.isnan(d2[:, 8:]).all() and (i2[:, 8:] == 0).all()
  i3, d3 = c.oracle.search_batched(c.q[:4], final_nn=30, pre_nn=8, leaves=1)
  np.testing.assert_array_equal(i2, i3)


def test_large_batch_matches_small_batches():
  c = get_case(nq=700)
  i_all, d_all = c.native.search_batched(c.q)
  for s in (0, 333):
    i_p, d_p = c.native.search_batched(c.q[s:s + 100])
    np.testing.assert_array_equal(i_all[s:s + 100], i_p)
    np.testing.assert_array_equal(d_all[s:s + 100].view(np.uint32), d_p.view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q, impl=1)
  np.testing.assert_array_equal(i0, i_all)


def test_stats_scan_bytes_match_oracle():
  c = get_case()
  c.native.search_batched(c.q)
  st = c.native.stats()
  c.oracle.search_batched(c.q)
  assert st["scan_bytes_alg"] == c.oracle.last_scan_bytes()
  assert st["scan_pairs"] == len(c.q) * c.probe
  assert st["kernel_launches"] >= 8


# ---- tokenization: the tcgen05 pre-filter + exact refinement against the SIMT path and the oracle ----
TOK_CASES = [
    dict(),
    dict(soar=1.5),
    dict(n=6000, leaves=300, probe=40, pre=50, d=32),
    dict(n=20000, leaves=1000, probe=100, d=100),          # L > 256: the default picks the tensor path
    dict(distance="squared_l2", d=64, leaves=50, n=10000),
    dict(distance="squared_l2", d=30, dpb=4, leaves=300, n=6000, probe=64, pre=64),
    dict(dpb=3, d=100, leaves=700, n=15000, probe=37),     # D % 4 == 0 but 3D not a multiple of 64; odd P
    dict(dpb=3, d=99, leaves=300, n=8000, probe=33),       # D % 4 != 0: scalar centre loads
    dict(n=12000, leaves=3000, probe=24, pre=60, d=32),    # 94 chunks of 32 centres >= 2 P: the chunk pre-selection applies
    dict(n=9000, leaves=2500, probe=7, pre=40, d=50, dpb=2),  # ... with L % 32 != 0 and D % 4 != 0
    dict(distance="squared_l2", n=9000, leaves=2100, probe=9, pre=40, d=32),  # ... squared L2: chunk minima of ||c||^2 - 2 <q, c>
]


@pytest.mark.parametrize("mode", ["tcgen05", "simt", "stream", "chunk"])
@pytest.mark.parametrize("kw", TOK_CASES, ids=[str(i) for i in range(len(TOK_CASES))])
def test_tokenize_modes_bit_exact(kw, mode, monkeypatch):
  monkeypatch.setenv("SCANN_B200_TOKENIZE", mode)
  c = get_case(**kw)
  l0, d0 = c.oracle.tokenize(c.q)
  l1, d1 = c.native.tokenize(c.q)
  np.testing.assert_array_equal(l0, l1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  i0, e0 = c.oracle.search_batched(c.q)
  i1, e1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(e0.view(np.uint32), e1.view(np.uint32))


@pytest.mark.parametrize("kw", [dict(n=6000, leaves=300, probe=40, pre=50, d=32),
                                dict(distance="squared_l2", d=64, leaves=50, n=10000)], ids=["dot", "l2"])
@pytest.mark.parametrize("mode", ["tcgen05", "stream", "chunk"])
def test_tokenize_degenerate_queries_fall_back_to_exact(kw, mode, monkeypatch):
  """Zero / tiny / huge / duplicated queries: the candidate window of the pre-filter degenerates (every centre ties),
  the kernel must fall back to exact distances and still match the oracle bit for bit."""
  monkeypatch.setenv("SCANN_B200_TOKENIZE", mode)
  c = get_case(**kw)
  q = c.q[:64].copy()
  q[0] = 0.0
  q[1] = 1e-30
  q[2] *= 1e-20
  q[3] *= 1e15
  q[4] = -q[5]
  q[6] = c.arrays.centers[7]
  q[8, 1:] = 0.0
  l0, d0 = c.oracle.tokenize(q)
  l1, d1 = c.native.tokenize(q)
  np.testing.assert_array_equal(l0, l1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


# ---- two scan phases (nearest leaves first, tau tightened in between) must not change anything ----
@pytest.mark.parametrize("kw", [CASES[0], CASES[5], CASES[6], CASES[7], CASES[8]], ids=["dot", "soar", "tiny", "allprobed", "l2"])
@pytest.mark.parametrize("cap", [None, "256"])
def test_two_phase_scan_is_bit_exact(kw, cap, monkeypatch):
  monkeypatch.setenv("SCANN_B200_TWO_PHASE", "1")
  if cap:
    monkeypatch.setenv("SCANN_B200_CAND_CAP", cap)   # overflow in either phase -> re-scan with dedup
  c = get_case(**kw)
  a = c.oracle.candidates(c.q)
  b = c.native.candidates(c.q)
  np.testing.assert_array_equal(a["count"], b["count"])
  for i in range(len(c.q)):
    n = a["count"][i]
    np.testing.assert_array_equal(a["dp"][i, :n], b["dp"][i, :n])
    np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q)
  i1, d1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


# ---- per-item candidate staging in shared memory (the large-leaf mode of the main scan) must not change anything ----
@pytest.mark.parametrize("kw", [CASES[0], CASES[5], CASES[7], CASES[8]], ids=["dot", "soar", "allprobed", "l2"])
@pytest.mark.parametrize("mode", ["single", "two_phase", "overflow"])
def test_staged_candidate_push_is_bit_exact(kw, mode, monkeypatch):
  monkeypatch.setenv("SCANN_B200_SCAN_STAGE", "1")
  monkeypatch.setenv("SCANN_B200_PILOT_CAP", "4096")     # the large-leaf pilot buffer as well
  monkeypatch.setenv("SCANN_B200_TWO_PHASE", "1" if mode != "single" else "0")
  if mode == "overflow":
    monkeypatch.setenv("SCANN_B200_CAND_CAP", "256")
  c = get_case(**kw)
  a = c.oracle.candidates(c.q)
  b = c.native.candidates(c.q)
  np.testing.assert_array_equal(a["count"], b["count"])
  for i in range(len(c.q)):
    n = a["count"][i]
    np.testing.assert_array_equal(a["dp"][i, :n], b["dp"][i, :n])
    np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q)
  i1, d1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


# ---- wide quads (the sparse-batch mode of the main scan: u16-lane tables, four queries per LDS.64) ----
@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
@pytest.mark.parametrize("mode", ["plain", "staged_two_phase", "overflow"])
def test_wide_quad_scan_is_bit_exact(kw, mode, monkeypatch):
  monkeypatch.setenv("SCANN_B200_SCAN_WIDE", "1")
  if mode != "plain":
    monkeypatch.setenv("SCANN_B200_SCAN_STAGE", "1")
    monkeypatch.setenv("SCANN_B200_TWO_PHASE", "1")
  if mode == "overflow":
    monkeypatch.setenv("SCANN_B200_CAND_CAP", "256")
  c = get_case(**kw)
  a = c.oracle.candidates(c.q)
  b = c.native.candidates(c.q)
  np.testing.assert_array_equal(a["count"], b["count"])
  for i in range(len(c.q)):
    n = a["count"][i]
    np.testing.assert_array_equal(a["leaf"][i, :n], b["leaf"][i, :n])
    np.testing.assert_array_equal(a["slot"][i, :n], b["slot"][i, :n])
    np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q)
  i1, d1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


def test_wide_quad_and_oct_scans_agree_on_a_sparse_batch(monkeypatch):
  """Few queries per leaf (the regime the wide quads are for): 48 queries over 1,000 leaves, both modes, same bits."""
  c = get_case(n=20000, leaves=1000, probe=24, pre=80, d=96, dpb=2, nq=48)
  out = {}
  for wide in ("0", "1"):
    monkeypatch.setenv("SCANN_B200_SCAN_WIDE", wide)
    out[wide] = c.native.search_batched(c.q)
  np.testing.assert_array_equal(out["0"][0], out["1"][0])
  np.testing.assert_array_equal(out["0"][1].view(np.uint32), out["1"][1].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q)
  np.testing.assert_array_equal(i0, out["1"][0])


# ---- tensor-core scan (scan_tc.cu: e4m3 nibble planes of the u8 LUT x one-hot codes on tcgen05 kind::f8f6f4, exact f32 sums) ----
@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
@pytest.mark.parametrize("mode", ["plain", "two_phase", "overflow"])
def test_tensor_core_scan_is_bit_exact(kw, mode, monkeypatch):
  monkeypatch.setenv("SCANN_B200_SCAN_TC", "1")
  if mode != "plain":
    monkeypatch.setenv("SCANN_B200_TWO_PHASE", "1")
  if mode == "overflow":
    monkeypatch.setenv("SCANN_B200_CAND_CAP", "256")   # overflowed queries are re-scanned by the SIMT kernel
  c = get_case(**kw)
  a = c.oracle.candidates(c.q)
  b = c.native.candidates(c.q)
  np.testing.assert_array_equal(a["count"], b["count"])
  for i in range(len(c.q)):
    n = a["count"][i]
    np.testing.assert_array_equal(a["leaf"][i, :n], b["leaf"][i, :n])
    np.testing.assert_array_equal(a["slot"][i, :n], b["slot"][i, :n])
    np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q)
  i1, d1 = c.native.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


@pytest.mark.parametrize("kw", [dict(nq=700), dict(nq=700, soar=1.5), dict(nq=500, distance="squared_l2", d=64, leaves=50, n=10000),
                                dict(nq=300, leaves=30)],   # ~670 slots per leaf: three 256-slot tiles, the last one ragged
                         ids=["dot", "soar", "l2", "tiles"])
def test_tensor_core_scan_with_several_query_blocks_per_leaf(kw, monkeypatch):
  """Hundreds of queries per leaf: several 64-query blocks (the last one ragged) per leaf, several items per CTA;
  the tensor-core scan, the octs and the oracle agree bit for bit."""
  c = get_case(**kw)
  out = {}
  for tc in ("0", "1"):
    monkeypatch.setenv("SCANN_B200_SCAN_TC", tc)
    out[tc] = c.native.search_batched(c.q, leaves=40)
    assert c.native.stats()["overflow_retries"] == 0
  np.testing.assert_array_equal(out["0"][0], out["1"][0])
  np.testing.assert_array_equal(out["0"][1].view(np.uint32), out["1"][1].view(np.uint32))
  i0, d0 = c.oracle.search_batched(c.q, leaves=40, impl=1)
  np.testing.assert_array_equal(i0, out["1"][0])
  np.testing.assert_array_equal(d0.view(np.uint32), out["1"][1].view(np.uint32))


def test_host_call_with_page_locked_buffers_matches_pageable():
  """scann_b200_search_batched copies page-locked caller memory to / from the device directly."""
  import torch
  c = get_case()
  i0, d0 = c.native.search_batched(c.q)
  qp = torch.from_numpy(c.q).pin_memory()
  oi = torch.empty(i0.shape, dtype=torch.int32).pin_memory()
  od = torch.empty(d0.shape, dtype=torch.float32).pin_memory()
  i1, d1 = c.native.search_batched(qp.numpy(), out=(oi.numpy().view(np.uint32), od.numpy()))
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  # mixed: pinned queries, pageable outputs
  i2, d2 = c.native.search_batched(qp.numpy())
  np.testing.assert_array_equal(i0, i2)
  np.testing.assert_array_equal(d0.view(np.uint32), d2.view(np.uint32))


# ---- bfloat16 reordering (Bfloat16ReorderingHelper) ----
@pytest.mark.parametrize("kw", [dict(), dict(soar=1.5), dict(dpb=3, d=99, leaves=60, n=9000),
                                dict(distance="squared_l2", d=64, leaves=50, n=10000),
                                dict(n=3000, leaves=30, d=6, dpb=2, probe=5, pre=40)],
                         ids=["dot", "soar", "d99", "l2", "d6"])
def test_bf16_reordering_is_bit_exact(kw):
  import copy
  import oracle
  from scann_b200 import _lib, index_build, distributed
  c = get_case(**kw)
  a = copy.copy(c.arrays)
  a.bf16_dataset = index_build.bfloat16_quantize(c.db)
  a.dataset = None
  o = oracle.OracleIndex(a, c.probe, c.pre, c.k)
  g = _lib.NativeIndex(a, c.probe, c.pre, c.k)
  i0, d0 = o.search_batched(c.q)
  i1, d1 = g.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  # distances are those of the bf16 rows (f32 query), within 1e-5 relative of float64
  x = (a.bf16_dataset.view(np.uint16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)
  rows = x[i1.astype(np.int64)]
  if a.distance == "dot_product":
    truth = np.einsum("qd,qkd->qk", c.q.astype(np.float64), rows)
  else:
    truth = ((c.q.astype(np.float64)[:, None, :] - rows) ** 2).sum(-1)
  np.testing.assert_allclose(d1, truth, rtol=1e-5, atol=1e-5)
  # and it differs from f32 reordering somewhere (the bf16 rows really are used)
  _, df = c.native.search_batched(c.q)
  assert not np.array_equal(df.view(np.uint32), d1.view(np.uint32))


# ---- int8 (fixed point) reordering (FixedPointFloatDense*ReorderingHelper) ----
@pytest.mark.parametrize("kw", [dict(), dict(soar=1.5), dict(dpb=3, d=99, leaves=60, n=9000),
                                dict(distance="squared_l2", d=64, leaves=50, n=10000),
                                dict(n=3000, leaves=30, d=6, dpb=2, probe=5, pre=40)],
                         ids=["dot", "soar", "d99", "l2", "d6"])
def test_int8_reordering_is_bit_exact(kw):
  import copy
  import oracle
  from scann_b200 import _lib, index_build
  c = get_case(**kw)
  a = copy.copy(c.arrays)
  a.int8_dataset, a.int8_multipliers = index_build.int8_quantize(c.db)
  if a.distance == "squared_l2":
    a.dp_norms = index_build.squared_l2_norms(c.db)
  a.dataset = None
  o = oracle.OracleIndex(a, c.probe, c.pre, c.k)
  g = _lib.NativeIndex(a, c.probe, c.pre, c.k)
  i0, d0 = o.search_batched(c.q)
  i1, d1 = g.search_batched(c.q)
  np.testing.assert_array_equal(i0, i1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
  # distances are those of the dequantized rows, close to the f32 reordering
  _, df = c.native.search_batched(c.q)
  assert not np.array_equal(df.view(np.uint32), d1.view(np.uint32))
  deq = a.int8_dataset.astype(np.float64) / a.int8_multipliers.astype(np.float64)[None, :]
  rows = deq[i1.astype(np.int64)]
  qq = c.q.astype(np.float64)
  if a.distance == "dot_product":
    truth = np.einsum("qd,qkd->qk", qq, rows)
  else:
    truth = (qq ** 2).sum(1)[:, None] + a.dp_norms[i1.astype(np.int64)].astype(np.float64) - 2 * np.einsum("qd,qkd->qk", qq, rows)
  np.testing.assert_allclose(d1, truth, rtol=1e-4, atol=1e-3)


# ---- chunk pre-selection without the stored distance matrix (exact chain for whole candidate chunks) ----
@pytest.mark.parametrize("kw", [TOK_CASES[8], TOK_CASES[9], TOK_CASES[10]], ids=["dot", "odd", "l2"])
def test_chunk_preselection_without_stored_rows(kw, monkeypatch):
  monkeypatch.setenv("SCANN_B200_TOKENIZE", "chunk")
  monkeypatch.setenv("SCANN_B200_TOKENIZE_ROWS", "0")
  c = get_case(**kw)
  l0, d0 = c.oracle.tokenize(c.q)
  l1, d1 = c.native.tokenize(c.q)
  np.testing.assert_array_equal(l0, l1)
  np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))


# ---- LUT build fused into the pilot kernel vs the separate lut_kernel: same table, same results ----
@pytest.mark.parametrize("kw", [CASES[0], CASES[5], CASES[8]], ids=["dot", "soar", "l2"])
def test_fused_and_separate_lut_build_agree(kw, monkeypatch):
  c = get_case(**kw)
  i0, d0 = c.oracle.search_batched(c.q)
  for fuse in ("1", "0"):
    monkeypatch.setenv("SCANN_B200_FUSE_LUT", fuse)
    i1, d1 = c.native.search_batched(c.q)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
    a = c.oracle.candidates(c.q)
    b = c.native.candidates(c.q)
    np.testing.assert_array_equal(a["count"], b["count"])
    for i in range(len(c.q)):
      n = a["count"][i]
      np.testing.assert_array_equal(a["score"][i, :n].view(np.uint32), b["score"][i, :n].view(np.uint32))
