"""ctypes binding of the C ABI declared in include/scann_b200.h.

The shared library is built in-tree (scann_b200/libscann_b200.so) by
`__graft_entry__.build()` / `make -C scann_b200/csrc`.  There is no fallback:
if the library is missing, or no CUDA device is usable, the product path
raises instead of computing anything on the CPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscann_b200.so")
_LIB = None


class IndexDesc(C.Structure):
  _fields_ = [
      ("distance", C.c_int32), ("n", C.c_uint32), ("d", C.c_uint32),
      ("n_leaves", C.c_uint32), ("n_blocks", C.c_uint32), ("dims_per_block", C.c_uint32),
      ("block_dims", C.c_void_p), ("centers", C.c_void_p), ("tokens", C.c_void_p),
      ("soar", C.c_int32), ("codes", C.c_void_p), ("soar_codes", C.c_void_p),
      ("codebook", C.c_void_p), ("dataset", C.c_void_p), ("bf16_dataset", C.c_void_p),
      ("overretrieve", C.c_float), ("default_leaves", C.c_int32),
      ("default_pre_nn", C.c_int32), ("default_final_nn", C.c_int32),
      ("device", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
      ("int8_dataset", C.c_void_p), ("int8_multipliers", C.c_void_p), ("dp_norms", C.c_void_p),
      ("shard_mode", C.c_int32),
      ("query_tokenization_type", C.c_int32),
  ]


class Stats(C.Structure):
  _fields_ = [
      ("scan_bytes_alg", C.c_uint64), ("scan_pairs", C.c_uint64), ("scan_lookups", C.c_uint64),
      ("kernel_launches", C.c_uint32), ("overflow_retries", C.c_uint32),
      ("ms_tokenize", C.c_float), ("ms_lut", C.c_float), ("ms_pilot", C.c_float),
      ("ms_worklist", C.c_float), ("ms_scan", C.c_float), ("ms_compact", C.c_float),
      ("ms_finalize", C.c_float), ("ms_total", C.c_float), ("scan_kernel_count", C.c_uint32),
      ("cand_sum", C.c_uint64), ("cand_max", C.c_uint64),
      ("tokenize_fallbacks", C.c_uint64),
      ("ms_exchange", C.c_float), ("ms_merge", C.c_float), ("exchange_bytes", C.c_uint64),
      ("bf_widenings", C.c_uint32), ("bf_exact_fallbacks", C.c_uint32),
      ("scan_oct_launches", C.c_uint32), ("scan_wide_launches", C.c_uint32), ("scan_tc_launches", C.c_uint32),
      ("reserved0", C.c_uint32),
  ]

  def as_dict(self):
    return {k: getattr(self, k) for k, _ in self._fields_}


class EncodeDesc(C.Structure):
  _fields_ = [
      ("n", C.c_uint32), ("d", C.c_uint32), ("n_leaves", C.c_uint32), ("n_blocks", C.c_uint32),
      ("dims_per_block", C.c_uint32), ("block_dims", C.c_void_p), ("dataset", C.c_void_p),
      ("centers", C.c_void_p), ("codebook", C.c_void_p), ("residual", C.c_int32),
      ("soar_lambda", C.c_float), ("noise_shaping_threshold", C.c_double), ("device", C.c_int32),
  ]


class EncodeStats(C.Structure):
  _fields_ = [
      ("ms_tokenize", C.c_float), ("ms_soar", C.c_float), ("ms_encode", C.c_float), ("ms_total", C.c_float),
      ("soar_evaluated", C.c_uint64), ("spilled", C.c_uint64), ("norm_ties", C.c_uint64),
      ("tokenize_fallbacks", C.c_uint64), ("chunk_rows", C.c_uint32),
  ]

  def as_dict(self):
    return {k: getattr(self, k) for k, _ in self._fields_}


class KMeansDesc(C.Structure):
  _fields_ = [("n", C.c_uint32), ("d", C.c_uint32), ("k", C.c_uint32), ("data", C.c_void_p), ("init_centers", C.c_void_p),
              ("iterations", C.c_int32), ("device", C.c_int32)]


class KMeansStats(C.Structure):
  _fields_ = [("ms_assign", C.c_float), ("ms_update", C.c_float), ("ms_total", C.c_float), ("iterations", C.c_uint32),
              ("empty_clusters", C.c_uint32), ("mean_sq_distance", C.c_double)]

  def as_dict(self):
    return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "scann_b200_index_create", "scann_b200_index_destroy", "scann_b200_search_batched",
    "scann_b200_search_batched_device", "scann_b200_search_partial_device",
    "scann_b200_merge_partials_device", "scann_b200_merge_topk_device", "scann_b200_last_error",
    "scann_b200_comm_unique_id", "scann_b200_comm_init", "scann_b200_search_sharded_device",
    "scann_b200_search_sharded_local",
    "scann_b200_abi_version",
    "scann_b200_debug_tokenize", "scann_b200_debug_lut", "scann_b200_debug_leaf_scores",
    "scann_b200_debug_candidates", "scann_b200_leaf_size", "scann_b200_last_stats",
    "scann_b200_assets_load", "scann_b200_assets_free", "scann_b200_assets_describe",
    "scann_b200_assets_config", "scann_b200_assets_save", "scann_b200_config_text_to_binary",
    "scann_b200_config_binary_to_text", "scann_b200_encode_database", "scann_b200_train_kmeans",
]


class NativeLibraryMissing(RuntimeError):
  pass


def lib():
  """Loads libscann_b200.so; raises NativeLibraryMissing if it has not been built."""
  global _LIB
  if _LIB is not None:
    return _LIB
  if not os.path.exists(LIB_PATH):
    raise NativeLibraryMissing(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(scann_b200 has no CPU fallback)")
  L = C.CDLL(LIB_PATH)
  vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
  L.scann_b200_index_create.argtypes = [C.POINTER(IndexDesc), C.POINTER(vp)]
  L.scann_b200_index_destroy.argtypes = [vp]
  L.scann_b200_index_destroy.restype = None
  L.scann_b200_search_batched.argtypes = [vp, vp, u32, i32, i32, i32, vp, vp, i32]
  L.scann_b200_search_batched_device.argtypes = [vp, vp, u32, i32, i32, i32, vp, vp, i32]
  L.scann_b200_search_partial_device.argtypes = [vp, vp, u32, i32, i32, vp, vp, vp, vp, i32]
  L.scann_b200_merge_partials_device.argtypes = [vp, u32, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, i32]
  L.scann_b200_merge_topk_device.argtypes = [vp, u32, i32, i32, vp, vp, i32, vp, vp, i32]
  L.scann_b200_comm_unique_id.argtypes = [vp]
  L.scann_b200_comm_init.argtypes = [vp, i32, i32, vp]
  L.scann_b200_search_sharded_device.argtypes = [vp, vp, u32, i32, i32, i32, i32, vp, vp, i32]
  L.scann_b200_search_sharded_local.argtypes = [C.POINTER(vp), i32, vp, u32, i32, i32, i32, i32, vp, vp, i32]
  L.scann_b200_last_error.restype = C.c_char_p
  L.scann_b200_abi_version.restype = C.c_int
  L.scann_b200_debug_tokenize.argtypes = [vp, vp, u32, i32, vp, vp]
  L.scann_b200_debug_lut.argtypes = [vp, vp, u32, vp, vp]
  L.scann_b200_debug_leaf_scores.argtypes = [vp, vp, u32, vp, u32]
  L.scann_b200_debug_candidates.argtypes = [vp, vp, u32, i32, i32, i32, vp, vp, vp, vp, vp]
  L.scann_b200_leaf_size.argtypes = [vp, u32]
  L.scann_b200_leaf_size.restype = u32
  L.scann_b200_last_stats.argtypes = [vp, C.POINTER(Stats)]
  L.scann_b200_assets_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
  L.scann_b200_assets_free.argtypes = [vp]
  L.scann_b200_assets_free.restype = None
  L.scann_b200_assets_describe.argtypes = [vp, C.POINTER(IndexDesc)]
  L.scann_b200_assets_config.argtypes = [vp]
  L.scann_b200_assets_config.restype = C.c_char_p
  L.scann_b200_assets_save.argtypes = [C.c_char_p, C.POINTER(IndexDesc), C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
  L.scann_b200_config_text_to_binary.argtypes = [C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
  L.scann_b200_config_binary_to_text.argtypes = [vp, C.c_size_t, C.c_char_p, C.c_size_t]
  L.scann_b200_encode_database.argtypes = [C.POINTER(EncodeDesc), vp, vp, vp, C.POINTER(EncodeStats)]
  L.scann_b200_train_kmeans.argtypes = [C.POINTER(KMeansDesc), vp, vp, C.POINTER(KMeansStats)]
  _LIB = L
  return L


_STATUS_NAMES = {3: "INVALID_ARGUMENT", 9: "FAILED_PRECONDITION", 12: "UNIMPLEMENTED", 13: "INTERNAL"}


class ScannB200Error(RuntimeError):

  def __init__(self, code, message):
    super().__init__(f"{_STATUS_NAMES.get(code, code)}: {message}")
    self.code = code
    self.message = message


def check(rc):
  if rc != 0:
    raise ScannB200Error(rc, lib().scann_b200_last_error().decode("utf-8", "replace"))


def ptr(a):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


class NativeIndex:
  """Owns one scann_b200_index handle built from IndexArrays."""

  def __init__(self, arrays, leaves_to_search, pre_reorder_nn, final_nn, device=0, shard_rank=0,
               shard_world=1, shard_mode=0):
    L = lib()
    a = arrays
    keep = []

    def own(x, dt):
      if x is None:
        return None
      y = np.ascontiguousarray(x, dtype=dt)
      keep.append(y)
      return y

    d = IndexDesc()
    d.distance = 0 if a.distance == "dot_product" else 1
    d.n, d.d = a.n, a.d
    d.n_leaves = 0 if a.centers is None else a.centers.shape[0]
    d.n_blocks = 0 if a.codes is None else a.codes.shape[1]
    d.dims_per_block = 0 if a.codebook is None else a.codebook.shape[2]
    d.block_dims = ptr(own(a.block_dims, np.int32))
    d.centers = ptr(own(a.centers, np.float32))
    d.tokens = ptr(own(a.tokens, np.int32))
    d.soar = 1 if a.soar else 0
    d.codes = ptr(own(a.codes, np.uint8))
    d.soar_codes = ptr(own(a.soar_codes, np.uint8))
    d.codebook = ptr(own(a.codebook, np.float32))
    d.dataset = ptr(own(a.dataset, np.float32))
    d.bf16_dataset = ptr(own(a.bf16_dataset, np.int16))
    d.int8_dataset = ptr(own(getattr(a, "int8_dataset", None), np.int8))
    d.int8_multipliers = ptr(own(getattr(a, "int8_multipliers", None), np.float32))
    d.dp_norms = ptr(own(getattr(a, "dp_norms", None), np.float32))
    d.query_tokenization_type = 1 if getattr(a, "int8_tokenization", False) else 0
    d.overretrieve = a.overretrieve
    d.default_leaves = leaves_to_search
    d.default_pre_nn = pre_reorder_nn
    d.default_final_nn = final_nn
    d.device = device
    d.shard_rank = shard_rank
    d.shard_world = shard_world
    d.shard_mode = shard_mode
    h = C.c_void_p()
    check(L.scann_b200_index_create(C.byref(d), C.byref(h)))
    self._h = h
    self.n, self.d = a.n, a.d
    self.L, self.B = d.n_leaves, d.n_blocks
    self.default_leaves, self.default_pre_nn, self.default_final_nn = leaves_to_search, pre_reorder_nn, final_nn
    self.device = device

  def close(self):
    if getattr(self, "_h", None):
      try:
        lib().scann_b200_index_destroy(self._h)
      except Exception:  # interpreter shutdown
        pass
      self._h = None

  __del__ = close

  # ---- hot path ----
  def search_batched(self, q, final_nn=-1, pre_nn=-1, leaves=-1, out=None):
    """Host buffers in / out.  `out` = (uint32 [nq, k], float32 [nq, k]) reuses caller arrays; page-locked
    arrays (queries and/or outputs) are copied to / from the device without the staging copy."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = final_nn if final_nn > 0 else self.default_final_nn
    if out is not None:
      idx, dist = out
      assert idx.dtype == np.uint32 and dist.dtype == np.float32 and idx.shape == dist.shape == (q.shape[0], k)
      assert idx.flags.c_contiguous and dist.flags.c_contiguous
    else:
      idx = np.empty((q.shape[0], k), dtype=np.uint32)
      dist = np.empty((q.shape[0], k), dtype=np.float32)
    check(lib().scann_b200_search_batched(self._h, ptr(q), q.shape[0], final_nn, pre_nn, leaves,
                                          ptr(idx), ptr(dist), k))
    return idx, dist

  def search_batched_device(self, d_q_ptr, nq, d_idx_ptr, d_dist_ptr, out_k, final_nn=-1, pre_nn=-1,
                            leaves=-1):
    check(lib().scann_b200_search_batched_device(self._h, C.c_void_p(d_q_ptr), nq, final_nn, pre_nn, leaves,
                                                 C.c_void_p(d_idx_ptr), C.c_void_p(d_dist_ptr), out_k))

  def stats(self):
    s = Stats()
    check(lib().scann_b200_last_stats(self._h, C.byref(s)))
    return s.as_dict()

  # ---- parity hooks ----
  def tokenize(self, q, leaves=-1):
    q = np.ascontiguousarray(q, dtype=np.float32)
    P = min(leaves if leaves > 0 else self.default_leaves, self.L)
    leaf = np.empty((q.shape[0], P), dtype=np.int32)
    dist = np.empty((q.shape[0], P), dtype=np.float32)
    check(lib().scann_b200_debug_tokenize(self._h, ptr(q), q.shape[0], P, ptr(leaf), ptr(dist)))
    return leaf, dist

  def lut(self, q):
    q = np.ascontiguousarray(q, dtype=np.float32)
    lut = np.empty((q.shape[0], self.B, 16), dtype=np.uint8)
    mult = np.empty(q.shape[0], dtype=np.float32)
    check(lib().scann_b200_debug_lut(self._h, ptr(q), q.shape[0], ptr(lut), ptr(mult)))
    return lut, mult

  def leaf_size(self, leaf):
    return int(lib().scann_b200_leaf_size(self._h, leaf))

  def leaf_scores(self, lut, leaf):
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    out = np.empty(self.leaf_size(leaf), dtype=np.int16)
    check(lib().scann_b200_debug_leaf_scores(self._h, ptr(lut), leaf, ptr(out), out.shape[0]))
    return out

  def candidates(self, q, pre_nn=-1, leaves=-1, cap=None):
    q = np.ascontiguousarray(q, dtype=np.float32)
    nq = q.shape[0]
    if cap is None:
      npre = pre_nn if pre_nn > 0 else self.default_pre_nn
      cap = int(npre * 4 + 8)
    leaf = np.zeros((nq, cap), np.uint32)
    slot = np.zeros((nq, cap), np.uint32)
    dp = np.zeros((nq, cap), np.uint32)
    score = np.zeros((nq, cap), np.float32)
    cnt = np.zeros(nq, np.uint32)
    check(lib().scann_b200_debug_candidates(self._h, ptr(q), nq, pre_nn, leaves, cap, ptr(leaf), ptr(slot),
                                            ptr(dp), ptr(score), ptr(cnt)))
    return dict(leaf=leaf, slot=slot, dp=dp, score=score, count=cnt)


def encode_database(dataset, centers, codebook, block_dims=None, residual=True, soar_lambda=None,
                    noise_shaping_threshold=float("nan"), device=0):
  """scann_b200_encode_database: database tokenization, SOAR secondary assignment and AH encoding on the GPU.

  Returns (tokens [N] or [2N] i32, codes [N, B] u8, soar_codes [N, B] u8 or None, stats dict) in the
  serialized-asset layout (datapoint_to_token.npy, hashed_dataset.npy, hashed_dataset_soar.npy)."""
  x = np.ascontiguousarray(dataset, dtype=np.float32)
  c = np.ascontiguousarray(centers, dtype=np.float32)
  cb = np.ascontiguousarray(codebook, dtype=np.float32)
  bd = None if block_dims is None else np.ascontiguousarray(block_dims, dtype=np.int32)
  n, d = x.shape
  nb = cb.shape[0]
  soar = soar_lambda is not None
  tokens = np.empty(2 * n if soar else n, np.int32)
  codes = np.empty((n, nb), np.uint8)
  soar_codes = np.empty((n, nb), np.uint8) if soar else None
  desc = EncodeDesc()
  desc.n, desc.d, desc.n_leaves, desc.n_blocks, desc.dims_per_block = n, d, c.shape[0], nb, cb.shape[2]
  desc.block_dims, desc.dataset, desc.centers, desc.codebook = ptr(bd), ptr(x), ptr(c), ptr(cb)
  desc.residual = 1 if residual else 0
  desc.soar_lambda = float(soar_lambda) if soar else float("nan")
  desc.noise_shaping_threshold = float(noise_shaping_threshold)
  desc.device = device
  st = EncodeStats()
  check(lib().scann_b200_encode_database(C.byref(desc), ptr(tokens), ptr(codes), ptr(soar_codes), C.byref(st)))
  return tokens, codes, soar_codes, st.as_dict()


def train_kmeans(data, init_centers, iterations, device=0, want_assignment=True):
  """scann_b200_train_kmeans: Lloyd iterations on the GPU (assignment = the tokenizer with P = 1, centroid update in the
  reference's double arithmetic, deterministic).  Returns (centers [k, d] f32, assignment [n] i32 or None, stats)."""
  x = np.ascontiguousarray(data, dtype=np.float32)
  c0 = np.ascontiguousarray(init_centers, dtype=np.float32)
  n, d = x.shape
  k = c0.shape[0]
  desc = KMeansDesc()
  desc.n, desc.d, desc.k, desc.data, desc.init_centers = n, d, k, ptr(x), ptr(c0)
  desc.iterations, desc.device = int(iterations), device
  centers = np.empty((k, d), np.float32)
  assign = np.empty(n, np.int32) if want_assignment else None
  st = KMeansStats()
  check(lib().scann_b200_train_kmeans(C.byref(desc), ptr(centers), ptr(assign), C.byref(st)))
  return centers, assign, st.as_dict()
