import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:
    return False


def pytest_collection_modifyitems(config, items):
  if _has_gpu():
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


class Case:
  """A small tree-AH index + queries + oracle, shared by the parity tests."""

  def __init__(self, n=20000, d=100, leaves=100, dpb=2, nq=64, soar=None, distance="dot_product",
               probe=10, pre=100, k=10, seed=1, int8_tok=False):
    from scann_b200 import datasets, index_build
    import oracle
    self.db = datasets.clustered(n, d, 4 * leaves, seed=seed, centers_seed=100 + seed)
    self.q = datasets.clustered(nq, d, 4 * leaves, seed=seed + 1, centers_seed=100 + seed)
    self.arrays = index_build.build_tree_ah(self.db, distance, num_leaves=leaves, dims_per_block=dpb,
                                            training_sample_size=min(n, 20000), soar_lambda=soar,
                                            tree_iters=6, ah_iters=5, device="cpu")
    # tree(quantize_centroids=True): queries are tokenized against the fixed-point centres
    self.arrays.int8_tokenization = bool(int8_tok)
    self.probe, self.pre, self.k = probe, pre, k
    self.oracle = oracle.OracleIndex(self.arrays, probe, pre, k)
    self._native = None

  @property
  def native(self):
    if self._native is None:
      from scann_b200 import _lib
      self._native = _lib.NativeIndex(self.arrays, self.probe, self.pre, self.k)
    return self._native


_CASES = {}


def get_case(**kw):
  key = tuple(sorted(kw.items()))
  if key not in _CASES:
    _CASES[key] = Case(**kw)
  return _CASES[key]


@pytest.fixture(scope="session")
def case_default():
  return get_case()
