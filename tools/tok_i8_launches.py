"""One int8-tokenized search of 10k queries on the C2-shaped centre set (2000 x 100, P = 100) and one on 40k x 96 centres
(P = 24), for an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): the kernels of
the int8 tokenization routes (split_rows_kernel, bf::gemm_kernel, topp_refine_kernel / topp_chunk_kernel) by name."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
  from helpers import i8_tok_arrays
  from scann_b200 import _lib
  for L, D, P in [(2000, 100, 100), (40000, 96, 24)]:
    a, _ = i8_tok_arrays(L, D, "dot_product", seed=3)
    a.int8_tokenization = True
    q = np.random.default_rng(1).standard_normal((10000, D)).astype(np.float32)
    ix = _lib.NativeIndex(a, P, 20, 10)
    ix.tokenize(q, leaves=P)
    ix.tokenize(q, leaves=P)
    ix.close()
  print("done")


if __name__ == "__main__":
  main()
