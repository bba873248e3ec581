"""GPU against the REFERENCE'S OWN ARITHMETIC, end to end, with no oracle in between: scann_b200_search_batched of a
tree-AH index with 8-dim blocks against a SearchBatched assembled from compiled reference code only
(tests/helpers.py::reference_pipeline_search over oracle/_ref/libscann_ref.so, prebuilt in the container that has
/root/reference) -- ids and distance bits.  (tests/test_oracle_ref.py::test_whole_search_from_reference_arithmetic_only is
the CPU counterpart against the oracle.)"""
import numpy as np
import pytest

from conftest import get_case
from oracle import ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref.available() and ref.has_symmetric() and ref.has_sse4_one_to_one()
                                      and ref.has_many_to_many()),
                                 reason="oracle/_ref/libscann_ref.so not built (needs /root/reference)")]


# the index configuration is one of tests/test_gpu_parity.py's (GPU == oracle there, stage by stage)
@pytest.mark.parametrize("kw", [dict(dpb=8, d=128, leaves=32, n=6000)], ids=["dot_dpb8"])
def test_gpu_search_equals_the_reference_arithmetic(kw):
  from helpers import reference_pipeline_search
  c = get_case(**kw)
  q = c.q[:8]
  want_idx, want_dist = reference_pipeline_search(c, q)
  got_idx, got_dist = c.native.search_batched(q)
  np.testing.assert_array_equal(got_idx, want_idx)
  np.testing.assert_array_equal(got_dist.view(np.uint32), want_dist.view(np.uint32))
