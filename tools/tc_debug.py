"""Debug helper: the multi-block tensor-core case of tests/test_gpu_parity.py, step by step with prints."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import get_case
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 700
leaves = int(sys.argv[2]) if len(sys.argv) > 2 else 40
c = get_case(nq=nq)
print("case built", flush=True)
os.environ["SCANN_B200_SCAN_TC"] = "0"
a = c.native.search_batched(c.q, leaves=leaves)
print("simt done", c.native.stats()["ms_scan"], flush=True)
os.environ["SCANN_B200_SCAN_TC"] = "1"
b = c.native.search_batched(c.q, leaves=leaves)
print("tc done", c.native.stats()["ms_scan"], "equal", np.array_equal(a[0], b[0]), flush=True)
