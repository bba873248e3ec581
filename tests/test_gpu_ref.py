"""GPU kernels against the REFERENCE'S OWN CODE (oracle/_ref/libscann_ref.so: the reference's AVX2 LUT16 kernel and code
packing, compiled from /root/reference by oracle/Makefile; the prebuilt library travels to the GPU box).

The other -m gpu tests compare the CUDA path with the oracle; these close the loop without it for the integer part of
the path and the float score: the main scan's int16 sums (score_oct_addr, via scann_b200_debug_leaf_scores, every oct
lane) and the pre-reorder candidate lists.
"""
import numpy as np
import pytest

from conftest import get_case
from oracle import ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libscann_ref.so not built (needs /root/reference)")]


def _leaf_codes(c, leaf):
  dps = c.oracle.leaf_datapoints(leaf)
  codes = c.arrays.codes[dps]
  if c.arrays.soar and len(dps):
    second = c.arrays.tokens[2 * dps.astype(np.int64) + 1] == leaf
    codes = np.where(second[:, None], c.arrays.soar_codes[dps], codes)
  return dps, codes


@pytest.mark.parametrize("kw", [dict(), dict(soar=1.5), dict(dpb=1, d=64), dict(dpb=3, d=50, leaves=40), dict(dpb=4, d=100),
                                dict(distance="squared_l2", d=64, leaves=50, n=10000)],
                         ids=["dot_b50", "dot_soar_b50", "dot_b64", "dot_varchunk_b17", "dot_b25", "l2_b32"])
def test_main_scan_int16_sums_equal_the_reference_kernel(kw):
  """scann_b200_debug_leaf_scores runs the main scan's scoring path with the LUT in oct lane (leaf mod 8): 16 leaves
  cover every lane twice.  Compared with LUT16Avx2<>::GetInt16Distances on CreatePackedDataset's packing."""
  c = get_case(**kw)
  rng = np.random.default_rng(7)
  B = c.arrays.codes.shape[1]
  checked = 0
  for leaf in range(min(c.native.L, 16)):
    dps, codes = _leaf_codes(c, leaf)
    if len(dps) == 0:
      continue
    lut = rng.integers(0, 256, (B, 16), dtype=np.uint8)
    want = ref.lut16_int16(ref.pack_dataset(codes), len(dps), B, [lut])[0, :len(dps)]
    np.testing.assert_array_equal(c.native.leaf_scores(lut, leaf), want)
    checked += 1
  assert checked >= 8


@pytest.mark.parametrize("kw", [dict(), dict(soar=1.5)], ids=["dot", "dot_soar"])
def test_candidates_equal_topn_of_the_reference_float_scores(kw):
  """The GPU's pre-reorder candidate lists == the N' smallest (score, leaf, slot) of the float scores the reference
  kernel computes (GetTopFloatDistances, nothing pruned) for the GPU's own probed leaves, LUTs and multipliers."""
  c = get_case(**kw)
  nq = 8
  q = c.q[:nq]
  leaves, bias = c.native.tokenize(q)
  lut, mult = c.native.lut(q)
  cand = c.native.candidates(q)
  B = c.arrays.codes.shape[1]
  for i in range(nq):
    rows = []
    for r in range(leaves.shape[1]):
      leaf = int(leaves[i, r])
      dps, codes = _leaf_codes(c, leaf)
      if len(dps) == 0:
        continue
      (idx, dist), = ref.lut16_top_float(ref.pack_dataset(codes), len(dps), B, [lut[i]], [bias[i, r]], [mult[i]])
      rows += [(float(d), leaf, int(s), np.float32(d)) for s, d in zip(idx, dist)]
    rows.sort(key=lambda t: (t[0], t[1], t[2]))
    n = int(cand["count"][i])
    want = rows[:n]
    np.testing.assert_array_equal(np.asarray([w[3] for w in want], np.float32).view(np.uint32), cand["score"][i, :n].view(np.uint32))
    np.testing.assert_array_equal([w[1] for w in want], cand["leaf"][i, :n])
    np.testing.assert_array_equal([w[2] for w in want], cand["slot"][i, :n])
