"""ctypes binding of oracle/_ref/libscann_ref.so: pieces of the REFERENCE'S OWN code (its AVX2 LUT16 kernel, the LUT
fixed-point conversion, the code packing, bfloat16 helpers, the f32 x int8 / bf16 one-to-many kernels) compiled from
/root/reference by oracle/Makefile.

Test infrastructure: it pins oracle/scann_oracle.c (and, in the -m gpu tests, the CUDA kernels) to reference code
instead of to restatements.  The library is built in the container that has /root/reference and travels to the GPU box
as a prebuilt file; `available()` is False where neither exists.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libscann_ref.so")
_LIB = None


def available():
  return os.path.exists(LIB_PATH)


def lib():
  global _LIB
  if _LIB is None:
    L = C.CDLL(LIB_PATH)
    L.ref_pack_dataset.restype = C.c_uint64
    L.ref_pack_dataset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    L.ref_lut16_int16.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_lut16_top_float.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                      C.c_int]
    L.ref_lut_to_fixed_point.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_bf16_quantize.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.ref_bf16_decompress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    if hasattr(L, "ref_one_to_many_int8_float"):
      L.ref_one_to_many_int8_float.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
      L.ref_one_to_many_int8_float_indexed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
      L.ref_one_to_many_bf16_float.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    if hasattr(L, "ref_one_to_many_f32"):
      L.ref_one_to_many_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
      L.ref_squared_l2_norm.restype = C.c_double
    if hasattr(L, "ref_dot_sse4_f32"):
      for f in (L.ref_dot_sse4_f32, L.ref_sql2_sse4_f32):
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    if hasattr(L, "ref_encode_noise_shaped"):
      L.ref_encode_noise_shaped.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.c_uint64, C.c_void_p, C.c_double, C.c_void_p]
    if hasattr(L, "ref_soar_costs"):
      L.ref_soar_costs.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_void_p]
    if hasattr(L, "ref_many_to_many_f32"):
      L.ref_many_to_many_f32.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
      L.ref_squared_l2_norm.argtypes = [C.c_void_p, C.c_uint64]
    _LIB = L
  return _LIB


def _p(a):
  return a.ctypes.data_as(C.c_void_p)


def pack_dataset(codes):
  """CreatePackedDataset (asymmetric_hashing_impl.cc:690-737): codes [n, B] u8 -> packed bytes."""
  codes = np.ascontiguousarray(codes, dtype=np.uint8)
  n, b = codes.shape
  out = np.zeros(((n + 31) // 32) * b * 16, np.uint8)
  size = lib().ref_pack_dataset(_p(codes), n, b, _p(out), out.size)
  assert size == out.size, (size, out.size)
  return out


def _ptr_array(arrs):
  return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def lut16_int16(packed, n, num_blocks, luts):
  """LUT16Avx2<len(luts)>::GetInt16Distances: int16 scores [len(luts), 32 * ceil(n / 32)]."""
  n32 = (n + 31) // 32
  luts = [np.ascontiguousarray(l, dtype=np.uint8).reshape(-1) for l in luts]
  outs = [np.zeros(32 * n32, np.int16) for _ in luts]
  rc = lib().ref_lut16_int16(_p(packed), n32, num_blocks, _ptr_array(luts), len(luts), _ptr_array(outs))
  assert rc == 0
  return np.stack(outs)


def lut16_top_float(packed, n, num_blocks, luts, biases, mults, epsilon=np.inf, first_dp_index=0, seq_prefetch=False):
  """LUT16Avx2<len(luts)>::GetTopFloatDistances with a top-N that keeps every push: per query (indices, float scores)
  of the datapoints whose int16 sum passes the reference's pre-filter `acc < trunc((epsilon - bias) * mult)`."""
  n32 = (n + 31) // 32
  nq = len(luts)
  luts = [np.ascontiguousarray(l, dtype=np.uint8).reshape(-1) for l in luts]
  idx = [np.zeros(32 * n32, np.uint32) for _ in range(nq)]
  dist = [np.zeros(32 * n32, np.float32) for _ in range(nq)]
  cnt = np.zeros(nq, np.uint32)
  b = np.ascontiguousarray(biases, dtype=np.float32)
  m = np.ascontiguousarray(mults, dtype=np.float32)
  rc = lib().ref_lut16_top_float(_p(packed), n32, num_blocks, n, _ptr_array(luts), nq, _p(b), _p(m),
                                 C.c_float(float(epsilon)), first_dp_index, _ptr_array(idx), _ptr_array(dist), 32 * n32,
                                 _p(cnt), 1 if seq_prefetch else 0)
  assert rc == 0, rc
  return [(idx[j][:cnt[j]].copy(), dist[j][:cnt[j]].copy()) for j in range(nq)]


def lut_to_fixed_point(raw, truncate=False):
  """ConvertLookupToFixedPoint<uint8_t> (asymmetric_hashing_impl.cc:571-645): (u8 table, multiplier)."""
  raw = np.ascontiguousarray(raw, dtype=np.float32).reshape(-1)
  out = np.zeros(raw.size, np.uint8)
  mult = np.zeros(1, np.float32)
  lib().ref_lut_to_fixed_point(_p(raw), raw.size, 1 if truncate else 0, _p(out), _p(mult))
  return out, mult[0]


def bf16_quantize(x):
  x = np.ascontiguousarray(x, dtype=np.float32)
  out = np.zeros(x.shape, np.int16)
  lib().ref_bf16_quantize(_p(x), x.size, _p(out))
  return out


def bf16_decompress(x):
  x = np.ascontiguousarray(x, dtype=np.int16)
  out = np.zeros(x.shape, np.float32)
  lib().ref_bf16_decompress(_p(x), x.size, _p(out))
  return out


def has_asymmetric():
  """True when the library carries the asymmetric one-to-many kernels (ref_glue_asym.cc)."""
  return available() and hasattr(lib(), "ref_one_to_many_int8_float")


def one_to_many_int8_float(query, rows, indices=None):
  """DenseDotProductDistanceOneToManyInt8Float (one_to_many_asymmetric.cc:44-49, :79-86 with an index list):
  -<query, float(row)> per row; the last n mod 3 results come from the reference's one-to-one kernel."""
  query = np.ascontiguousarray(query, dtype=np.float32)
  rows = np.ascontiguousarray(rows, dtype=np.int8)
  if indices is None:
    out = np.zeros(rows.shape[0], np.float32)
    lib().ref_one_to_many_int8_float(_p(query), _p(rows), rows.shape[0], rows.shape[1], _p(out))
    return out
  indices = np.ascontiguousarray(indices, dtype=np.uint32)
  out = np.zeros(indices.shape[0], np.float32)
  lib().ref_one_to_many_int8_float_indexed(_p(query), _p(rows), rows.shape[1], _p(indices), indices.shape[0], _p(out))
  return out


def one_to_many_bf16_float(query, rows, squared_l2=False):
  """DenseDotProductDistanceOneToManyBf16Float / OneToManyBf16FloatSquaredL2 over all rows (bf16 bits as int16)."""
  query = np.ascontiguousarray(query, dtype=np.float32)
  rows = np.ascontiguousarray(rows, dtype=np.int16)
  out = np.zeros(rows.shape[0], np.float32)
  lib().ref_one_to_many_bf16_float(_p(query), _p(rows), rows.shape[0], rows.shape[1], 1 if squared_l2 else 0, _p(out))
  return out


def has_symmetric():
  """True when the library carries the symmetric float one-to-many kernel (ref_glue_sym.cc)."""
  return available() and hasattr(lib(), "ref_one_to_many_f32")


def one_to_many_f32(query, rows, squared_l2=False):
  """DenseAccumulatingDistanceMeasureOneToManyInternalAvx2 (one_to_many_symmetric.h:373-503) with the dot-product or
  squared-L2 lambdas over a row-major f32 matrix of 3 m rows, dims >= 8."""
  query = np.ascontiguousarray(query, dtype=np.float32)
  rows = np.ascontiguousarray(rows, dtype=np.float32)
  out = np.zeros(rows.shape[0], np.float32)
  rc = lib().ref_one_to_many_f32(_p(query), _p(rows), rows.shape[0], rows.shape[1], 1 if squared_l2 else 0, _p(out))
  assert rc == 0, "ref_one_to_many_f32 needs dims >= 8 and 3 m rows"
  return out


def squared_l2_norm(v):
  """SquaredL2Norm(ConstSpan<float>) = DenseSingleAccumulate(v, Square()) (utils/reduction.h:357-390), a double."""
  v = np.ascontiguousarray(v, dtype=np.float32)
  return float(lib().ref_squared_l2_norm(_p(v), v.size))


def has_many_to_many():
  return available() and hasattr(lib(), "ref_many_to_many_f32")


def many_to_many_f32(queries, db, squared_l2=False):
  """DenseDistanceManyToMany's accumulators (many_to_many_impl.inc:236-257,522-560): [nq, n] f32 -- -<q, x> or
  (|x|^2 + |q|^2) - 2 <q, x> in the reference's per-dimension fnmadd order."""
  queries = np.ascontiguousarray(queries, dtype=np.float32)
  db = np.ascontiguousarray(db, dtype=np.float32)
  out = np.zeros((queries.shape[0], db.shape[0]), np.float32)
  lib().ref_many_to_many_f32(_p(queries), queries.shape[0], _p(db), db.shape[0], db.shape[1], 1 if squared_l2 else 0, _p(out))
  return out


def has_soar_costs():
  return available() and hasattr(lib(), "ref_soar_costs")


def soar_costs(x, centers, primary, lam):
  """ComputeNormalizedResidual + DenseManyToManyOrthogonalityAmplified's accumulation: (costs [n, L], rhat [n, D])."""
  x = np.ascontiguousarray(x, dtype=np.float32)
  c = np.ascontiguousarray(centers, dtype=np.float32)
  p = np.ascontiguousarray(primary, dtype=np.int32)
  cost = np.zeros((x.shape[0], c.shape[0]), np.float32)
  rhat = np.zeros(x.shape, np.float32)
  lib().ref_soar_costs(_p(x), x.shape[0], x.shape[1], _p(c), c.shape[0], _p(p), C.c_float(float(lam)), _p(cost), _p(rhat))
  return cost, rhat


def has_noise_shaped():
  return available() and hasattr(lib(), "ref_encode_noise_shaped")


def encode_noise_shaped(x, codebook, block_dims=None, centers=None, token=None, threshold=0.2):
  """AhImpl<float>::IndexDatapointNoiseShaped (asymmetric_hashing_impl.cc:434-503) over rows of x (residual against
  centers[token] when given, the original being x): codes [n, B] u8."""
  x = np.ascontiguousarray(x, dtype=np.float32)
  cb = np.ascontiguousarray(codebook, dtype=np.float32)
  bd = None if block_dims is None else np.ascontiguousarray(block_dims, dtype=np.int32)
  c = None if centers is None else np.ascontiguousarray(centers, dtype=np.float32)
  t = None if token is None else np.ascontiguousarray(token, dtype=np.int32)
  out = np.zeros((x.shape[0], cb.shape[0]), np.uint8)
  lib().ref_encode_noise_shaped(_p(x), x.shape[0], x.shape[1], None if c is None else _p(c), None if t is None else _p(t),
                                _p(cb), cb.shape[0], cb.shape[2], None if bd is None else _p(bd), float(threshold), _p(out))
  return out


def has_sse4_one_to_one():
  return available() and hasattr(lib(), "ref_dot_sse4_f32")


def one_to_one_sse4(a, b, squared_l2=False):
  """DenseDotProductSse4 / DenseSquaredL2DistanceSse4 (float, float), even lengths: the double the reference returns."""
  a = np.ascontiguousarray(a, dtype=np.float32)
  b = np.ascontiguousarray(b, dtype=np.float32)
  f = lib().ref_sql2_sse4_f32 if squared_l2 else lib().ref_dot_sse4_f32
  return float(f(_p(a), _p(b), a.size))
