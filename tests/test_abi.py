"""CPU tests of the boundary: the C-ABI library loads and exports what include/scann_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
  text = open(os.path.join(ROOT, "include", "scann_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(scann_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  from scann_b200 import _lib
  lib = _lib.lib()
  names = declared_symbols()
  assert len(names) >= 14
  for n in names:
    assert hasattr(lib, n), f"{n} declared in include/scann_b200.h but not exported"
  assert sorted(_lib.EXPORTS) == names
  assert lib.scann_b200_abi_version() == 6


def test_struct_layouts_match_header():
  from scann_b200 import _lib
  # 6 u32/i32, 3 pointers, i32 (+pad), 5 pointers, float + 6 i32 (+pad), 3 pointers (int8 reordering), shard_mode (+pad)
  assert ctypes.sizeof(_lib.IndexDesc) == 24 + 24 + 8 + 40 + 28 + 4 + 24 + 8
  # 5 u32 (+pad), 4 pointers, i32, float, double, i32 (+pad)
  assert ctypes.sizeof(_lib.EncodeDesc) == 24 + 32 + 8 + 8 + 8
  # 4 floats, 4 u64, u32 (+pad)
  assert ctypes.sizeof(_lib.EncodeStats) == 16 + 32 + 8
  # ... + ms_exchange, ms_merge, exchange_bytes, bf_widenings, bf_exact_fallbacks, 3 scan-launch counters + reserved
  assert ctypes.sizeof(_lib.Stats) == 24 + 8 + 32 + 4 + 4 + 16 + 8 + 8 + 8 + 8 + 16


def test_create_fails_loudly_without_gpu_or_arguments():
  import torch
  from scann_b200 import _lib
  lib = _lib.lib()
  h = ctypes.c_void_p()
  d = _lib.IndexDesc()
  rc = lib.scann_b200_index_create(ctypes.byref(d), ctypes.byref(h))
  assert rc != 0 and not h.value
  msg = lib.scann_b200_last_error().decode()
  if not torch.cuda.is_available():
    assert rc == 9 and "no CPU path" in msg      # FAILED_PRECONDITION: never falls back to the CPU
  else:
    assert rc in (3, 12)
  assert lib.scann_b200_index_create(None, None) == 3


def test_product_package_does_not_import_oracle():
  import subprocess
  import sys
  code = ("import sys; sys.path.insert(0, %r); import scann_b200, scann_b200._lib, scann_b200.index_build; "
          "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'") % ROOT
  subprocess.check_call([sys.executable, "-c", code])
  for dirpath, _, files in os.walk(os.path.join(ROOT, "scann_b200")):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
        text = open(os.path.join(dirpath, f)).read()
        assert not re.search(r"^\s*(import oracle|from oracle)", text, flags=re.M), f
        assert not re.search(r"#include.*oracle|dlopen.*oracle|CDLL.*oracle", text), f


def test_built_library_is_hand_written_sm100a_code():
  """Evidence hygiene (no GPU needed): the in-tree libscann_b200.so carries sm_100a cubins only, their SASS holds the
  Blackwell tensor-core / TMA / TMEM instructions the kernels are written with (tcgen05.mma -> UTCHMMA, TMA loads ->
  UTMALDG, tcgen05.ld -> LDTM), and the kernels of every stage -- the int8 tokenization ones included -- are in it."""
  import shutil
  import subprocess
  tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
  if not os.path.exists(tool):
    pytest.skip("cuobjdump not available")
  so = os.path.join(ROOT, "scann_b200", "libscann_b200.so")
  elfs = subprocess.run([tool, "-lelf", so], capture_output=True, text=True, check=True).stdout.split()
  cubins = [e for e in elfs if e.endswith(".cubin")]
  assert cubins and all("sm_100a" in e for e in cubins), cubins
  sass = subprocess.run([tool, "-sass", so], capture_output=True, text=True, check=True).stdout
  for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "LDTM.x32", "UTCBAR"):
    assert mnemonic in sass, mnemonic
  for kernel in ("scan_main_kernel", "pilot_kernel", "tokenize_i8_kernel", "tokenize_i8_tail_kernel", "topp_refine_kernel",
                 "topp_chunk_kernel", "gemm_pair_kernel", "gemm_kernel", "finalize_kernel", "scan_tc_kernel"):
    assert kernel in sass, kernel
