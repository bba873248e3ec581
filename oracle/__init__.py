"""ctypes wrapper around the CPU oracle (oracle/scann_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(scann_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Desc(C.Structure):
  _fields_ = [
      ("distance", C.c_int32), ("n", C.c_uint32), ("d", C.c_uint32),
      ("n_leaves", C.c_uint32), ("n_blocks", C.c_uint32), ("dims_per_block", C.c_uint32),
      ("block_dims", C.c_void_p), ("centers", C.c_void_p), ("tokens", C.c_void_p),
      ("soar", C.c_int32), ("codes", C.c_void_p), ("soar_codes", C.c_void_p),
      ("codebook", C.c_void_p), ("dataset", C.c_void_p), ("bf16_dataset", C.c_void_p),
      ("overretrieve", C.c_float), ("default_leaves", C.c_int32),
      ("default_pre_nn", C.c_int32), ("default_final_nn", C.c_int32),
      ("int8_dataset", C.c_void_p), ("int8_multipliers", C.c_void_p), ("dp_norms", C.c_void_p),
      ("centers_i8", C.c_void_p), ("centers_inv_mult", C.c_void_p), ("centers_sqnorm", C.c_void_p),
  ]


def build(force=False):
  so = os.path.join(_HERE, "libscann_oracle.so")
  src = os.path.join(_HERE, "scann_oracle.c")
  if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["make", "-C", _HERE, "-s"])
  return so


def lib():
  global _LIB
  if _LIB is None:
    so = os.path.join(_HERE, "libscann_oracle.so")
    if not os.path.exists(so):
      build()
    L = C.CDLL(so)
    L.so_index_create.restype = C.c_void_p
    L.so_index_create.argtypes = [C.POINTER(_Desc)]
    L.so_index_destroy.argtypes = [C.c_void_p]
    L.so_last_error.restype = C.c_char_p
    L.so_search_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.so_tokenize.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.so_lut.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.so_leaf_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.so_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    L.so_exact_distances.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.so_leaf_size.restype = C.c_uint32
    L.so_leaf_size.argtypes = [C.c_void_p, C.c_uint32]
    L.so_leaf_datapoints.restype = C.POINTER(C.c_uint32)
    L.so_leaf_datapoints.argtypes = [C.c_void_p, C.c_uint32]
    L.so_disjoint.argtypes = [C.c_void_p]
    L.so_quantize_centers.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.so_last_scan_bytes.restype = C.c_uint64
    L.so_last_boundary_band.restype = C.c_uint64
    _LIB = L
  return _LIB


def _p(a):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


def quantize_centers(centers):
  """so_quantize_centers: (int8 centres [L, D], inverse multipliers [D], squared norms of the float centres [L])."""
  c = np.ascontiguousarray(centers, dtype=np.float32)
  L, D = c.shape
  ci8, inv, sqn = np.empty((L, D), np.int8), np.empty(D, np.float32), np.empty(L, np.float32)
  if lib().so_quantize_centers(_p(c), L, D, _p(ci8), _p(inv), _p(sqn)):
    raise RuntimeError(lib().so_last_error().decode())
  return ci8, inv, sqn


class OracleIndex:
  """CPU oracle searcher built from the same arrays the CUDA index is built from."""

  def __init__(self, arrays, leaves_to_search, pre_reorder_nn, final_nn):
    L = lib()
    a = arrays
    self._keep = []

    def own(x, dt):
      if x is None:
        return None
      y = np.ascontiguousarray(x, dtype=dt)
      self._keep.append(y)
      return y

    self.n, self.d = a.n, a.d
    self.B = 0 if a.codes is None else a.codes.shape[1]
    self.L = 0 if a.centers is None else a.centers.shape[0]
    d = _Desc()
    d.distance = 0 if a.distance == "dot_product" else 1
    d.n, d.d, d.n_leaves, d.n_blocks = a.n, a.d, self.L, self.B
    d.dims_per_block = 0 if a.codebook is None else a.codebook.shape[2]
    d.block_dims = _p(own(a.block_dims, np.int32))
    d.centers = _p(own(a.centers, np.float32))
    d.tokens = _p(own(a.tokens, np.int32))
    d.soar = 1 if a.soar else 0
    d.codes = _p(own(a.codes, np.uint8))
    d.soar_codes = _p(own(a.soar_codes, np.uint8))
    d.codebook = _p(own(a.codebook, np.float32))
    d.dataset = _p(own(a.dataset, np.float32))
    d.bf16_dataset = _p(own(a.bf16_dataset, np.int16))
    d.int8_dataset = _p(own(getattr(a, "int8_dataset", None), np.int8))
    d.int8_multipliers = _p(own(getattr(a, "int8_multipliers", None), np.float32))
    d.dp_norms = _p(own(getattr(a, "dp_norms", None), np.float32))
    if getattr(a, "int8_tokenization", False):
      # query_tokenization_type FIXED_POINT_INT8: the fixed-point centres are derived here from the float centres
      ci8, inv, sqn = quantize_centers(a.centers)
      d.centers_i8, d.centers_inv_mult, d.centers_sqnorm = _p(own(ci8, np.int8)), _p(own(inv, np.float32)), _p(own(sqn, np.float32))
    d.overretrieve = a.overretrieve
    d.default_leaves = leaves_to_search
    d.default_pre_nn = pre_reorder_nn
    d.default_final_nn = final_nn
    self._h = L.so_index_create(C.byref(d))
    if not self._h:
      raise RuntimeError(L.so_last_error().decode())
    self.default_leaves = leaves_to_search
    self.default_pre_nn = pre_reorder_nn
    self.default_final_nn = final_nn

  def __del__(self):
    if getattr(self, "_h", None):
      try:
        lib().so_index_destroy(self._h)
      except Exception:  # interpreter shutdown
        pass
      self._h = None

  def search_batched(self, q, final_nn=-1, pre_nn=-1, leaves=-1, impl=0, threads=1, batch=256):
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = final_nn if final_nn > 0 else self.default_final_nn
    idx = np.empty((q.shape[0], k), dtype=np.uint32)
    dist = np.empty((q.shape[0], k), dtype=np.float32)
    rc = lib().so_search_batched(self._h, _p(q), q.shape[0], final_nn, pre_nn, leaves, _p(idx), _p(dist),
                                 k, impl, threads, batch)
    if rc:
      raise RuntimeError(lib().so_last_error().decode())
    return idx, dist

  def tokenize(self, q, leaves=-1):
    q = np.ascontiguousarray(q, dtype=np.float32)
    P = min(leaves if leaves > 0 else self.default_leaves, self.L)
    leaf = np.empty((q.shape[0], P), dtype=np.int32)
    dist = np.empty((q.shape[0], P), dtype=np.float32)
    lib().so_tokenize(self._h, _p(q), q.shape[0], P, _p(leaf), _p(dist))
    return leaf, dist

  def lut(self, q):
    q = np.ascontiguousarray(q, dtype=np.float32)
    lut = np.empty((q.shape[0], self.B, 16), dtype=np.uint8)
    mult = np.empty(q.shape[0], dtype=np.float32)
    lib().so_lut(self._h, _p(q), q.shape[0], _p(lut), _p(mult))
    return lut, mult

  def leaf_size(self, leaf):
    return int(lib().so_leaf_size(self._h, leaf))

  def leaf_datapoints(self, leaf):
    n = self.leaf_size(leaf)
    ptr = lib().so_leaf_datapoints(self._h, leaf)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.empty(0, np.uint32)

  def leaf_scores(self, lut, leaf):
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    out = np.empty(self.leaf_size(leaf), dtype=np.int16)
    lib().so_leaf_scores(self._h, _p(lut), leaf, _p(out))
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
    acc = np.zeros((nq, cap), np.int32)
    cnt = np.zeros(nq, np.uint32)
    lib().so_candidates(self._h, _p(q), nq, pre_nn, leaves, cap, _p(leaf), _p(slot), _p(dp), _p(score),
                        _p(acc), _p(cnt))
    return dict(leaf=leaf, slot=slot, dp=dp, score=score, acc=acc, count=cnt,
                scan_bytes=int(lib().so_last_scan_bytes()), band=int(lib().so_last_boundary_band()))

  def exact_distances(self, q, dps):
    q = np.ascontiguousarray(q, dtype=np.float32)
    dps = np.ascontiguousarray(dps, dtype=np.uint32)
    out = np.empty(dps.shape[0], dtype=np.float32)
    rc = lib().so_exact_distances(self._h, _p(q), _p(dps), dps.shape[0], _p(out))
    if rc:
      raise RuntimeError(lib().so_last_error().decode())
    return out

  @property
  def disjoint(self):
    return bool(lib().so_disjoint(self._h))

  @staticmethod
  def last_scan_bytes():
    return int(lib().so_last_scan_bytes())


def bruteforce_bf16(db_bf16, q, k, threads=1):
  """Exact bf16 brute force on the CPU: db_bf16 [N, D] int16 (bf16 bits), q [nq, D] f32."""
  L = lib()
  L.so_bruteforce_bf16.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int]
  db = np.ascontiguousarray(db_bf16, dtype=np.int16)
  q = np.ascontiguousarray(q, dtype=np.float32)
  idx = np.empty((q.shape[0], k), np.uint32)
  dist = np.empty((q.shape[0], k), np.float32)
  rc = L.so_bruteforce_bf16(_p(db), db.shape[0], db.shape[1], _p(q), q.shape[0], k, _p(idx), _p(dist), threads)
  if rc:
    raise RuntimeError(L.so_last_error().decode())
  return idx, dist


def bruteforce_f32(db, q, k, threads=1, distance="dot_product"):
  """Exact float brute force on the CPU (sequential fnmadd chain): db [N, D] f32, q [nq, D] f32; dot product or
  squared L2 (the many-to-many kernel's ||x||^2 + ||q||^2 - 2 <q, x> chain)."""
  L = lib()
  fn = L.so_bruteforce_f32 if distance == "dot_product" else L.so_bruteforce_f32_l2
  fn.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
  db = np.ascontiguousarray(db, dtype=np.float32)
  q = np.ascontiguousarray(q, dtype=np.float32)
  idx = np.empty((q.shape[0], k), dtype=np.uint32)
  dist = np.empty((q.shape[0], k), dtype=np.float32)
  rc = fn(_p(db), db.shape[0], db.shape[1], _p(q), q.shape[0], k, _p(idx), _p(dist), threads)
  if rc:
    raise RuntimeError(L.so_last_error().decode())
  return idx, dist


# ---- index build, deterministic part ----

def assign_primary(x, centers, threads=1):
  """Database tokenization: nearest centre under squared L2 -> (token [N] i32, distance [N] f32)."""
  L = lib()
  L.so_assign_primary.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
  x = np.ascontiguousarray(x, dtype=np.float32)
  c = np.ascontiguousarray(centers, dtype=np.float32)
  tok = np.empty(x.shape[0], np.int32)
  dist = np.empty(x.shape[0], np.float32)
  rc = L.so_assign_primary(_p(x), x.shape[0], x.shape[1], _p(c), c.shape[0], _p(tok), _p(dist), threads)
  if rc:
    raise RuntimeError(L.so_last_error().decode())
  return tok, dist


def assign_soar(x, centers, primary, lam, threads=1):
  """SOAR secondary assignment -> (token [N] i32, cost [N] f32); the token may equal the primary."""
  L = lib()
  L.so_assign_soar.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_float,
                               C.c_void_p, C.c_void_p, C.c_int]
  x = np.ascontiguousarray(x, dtype=np.float32)
  c = np.ascontiguousarray(centers, dtype=np.float32)
  p = np.ascontiguousarray(primary, dtype=np.int32)
  tok = np.empty(x.shape[0], np.int32)
  cost = np.empty(x.shape[0], np.float32)
  rc = L.so_assign_soar(_p(x), x.shape[0], x.shape[1], _p(c), c.shape[0], _p(p), float(lam), _p(tok), _p(cost), threads)
  if rc:
    raise RuntimeError(L.so_last_error().decode())
  return tok, cost


def encode(x, codebook, block_dims=None, centers=None, token=None, threshold=float("nan"), threads=1):
  """AH codes [N, B] u8 of x (centers None) or of x - centers[token]; threshold NaN = plain nearest-centre hash."""
  L = lib()
  L.so_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                          C.c_void_p, C.c_double, C.c_void_p, C.c_int]
  L.so_last_encode_ties.restype = C.c_uint64
  x = np.ascontiguousarray(x, dtype=np.float32)
  cb = np.ascontiguousarray(codebook, dtype=np.float32)
  bd = None if block_dims is None else np.ascontiguousarray(block_dims, dtype=np.int32)
  c = None if centers is None else np.ascontiguousarray(centers, dtype=np.float32)
  t = None if token is None else np.ascontiguousarray(token, dtype=np.int32)
  out = np.empty((x.shape[0], cb.shape[0]), np.uint8)
  rc = L.so_encode(_p(x), x.shape[0], x.shape[1], _p(c), _p(t), _p(cb), cb.shape[0], cb.shape[2], _p(bd),
                   float(threshold), _p(out), threads)
  if rc:
    raise RuntimeError(L.so_last_error().decode())
  return out, int(L.so_last_encode_ties())


def encode_database(x, centers, codebook, block_dims=None, residual=True, soar_lambda=None,
                    threshold=float("nan"), threads=1):
  """The whole deterministic build stage in the serialized layout (scann_ops/cc/scann.cc:533-566):
  tokens [N] or [2N] (slot 2i = lower-numbered leaf, 2i+1 = the other leaf or -1), codes, soar_codes, and the number
  of (datapoint, leaf) pairs whose initial block norms tie (so_last_encode_ties)."""
  n = x.shape[0]
  prim, _ = assign_primary(x, centers, threads)
  cen = centers if residual else None
  if soar_lambda is None:
    codes, ties = encode(x, codebook, block_dims, cen, prim, threshold, threads)
    return prim, codes, None, ties
  if not residual:
    raise ValueError("SOAR is defined for residual (dot product) tree-AH only (scann_builder.py:170-172)")
  sec, _ = assign_soar(x, centers, prim, soar_lambda, threads)
  spilled = sec != prim
  lo = np.where(spilled, np.minimum(prim, sec), prim).astype(np.int32)
  hi = np.where(spilled, np.maximum(prim, sec), -1).astype(np.int32)
  tokens = np.empty(2 * n, np.int32)
  tokens[0::2] = lo
  tokens[1::2] = hi
  codes, ties0 = encode(x, codebook, block_dims, cen, lo, threshold, threads)
  soar_codes, ties1 = encode(x, codebook, block_dims, cen, hi, threshold, threads)
  return tokens, codes, soar_codes, ties0 + ties1


def kmeans(x, init_centers, iterations, threads=1, want_assignment=True):
  """so_kmeans: (centers [k, d] f32, assignment [n] i32 or None, empty clusters of the last iteration)."""
  L = lib()
  L.so_kmeans.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
  x = np.ascontiguousarray(x, dtype=np.float32)
  c = np.array(init_centers, dtype=np.float32, copy=True, order="C")
  assign = np.empty(x.shape[0], np.int32) if want_assignment else None
  empty = np.zeros(1, np.uint32)
  rc = L.so_kmeans(_p(x), x.shape[0], x.shape[1], _p(c), c.shape[0], int(iterations), _p(assign) if assign is not None else None,
                   _p(empty), threads)
  if rc:
    raise RuntimeError(L.so_last_error().decode())
  return c, assign, int(empty[0])
