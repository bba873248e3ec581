"""Times the tokenization stage (CUDA events, `ms_tokenize` of the library's stats) of a float and an int8-tokenized
searcher over the same centres: L centres x D dims, nq queries per call.  One JSON line per configuration."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
  from helpers import i8_tok_arrays
  from scann_b200 import _lib
  nq = 10000
  for L, D, P in [(2000, 100, 100), (2000, 128, 100), (10000, 96, 24), (40000, 96, 24)]:
    a, _ = i8_tok_arrays(L, D, "dot_product", seed=3)
    q = np.random.default_rng(1).standard_normal((nq, D)).astype(np.float32)
    row = {"L": L, "D": D, "P": P, "nq": nq}
    for name, flag in (("float", False), ("int8", True)):
      a.int8_tokenization = flag
      ix = _lib.NativeIndex(a, P, 20, 10)
      ms = []
      for _ in range(6):
        ix.tokenize(q[:16], leaves=P)  # warm
        ix.search_batched(q)
        ms.append(ix.stats()["ms_tokenize"])
      row["ms_tokenize_" + name] = float(np.median(ms[2:]))
      ix.close()
    row["int8_gfma_per_s"] = nq * L * D / (row["ms_tokenize_int8"] * 1e-3) / 1e9
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
  main()
