"""GPU parity: scann_b200_encode_database (database tokenization, SOAR, AH encoding)
against the oracle's restatement of the reference, bit for bit, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixture(n, d, L, seed, noise, normalize):
  rng = np.random.default_rng(seed)
  means = rng.standard_normal((L, d)).astype(np.float32)
  x = (means[rng.integers(0, L, n)] + noise * rng.standard_normal((n, d))).astype(np.float32)
  centers = (means + 0.05 * rng.standard_normal((L, d))).astype(np.float32)
  if normalize:
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
  return np.ascontiguousarray(x), np.ascontiguousarray(centers)


def _codebook(x, centers, dpb, seed):
  """A plausible codebook: 16 residual sub-vectors of the data per block (zero padded short block)."""
  import oracle
  rng = np.random.default_rng(seed)
  n, d = x.shape
  full, part = divmod(d, dpb)
  bd = np.asarray([dpb] * full + ([part] if part else []), np.int32)
  tok, _ = oracle.assign_primary(x[:512], centers, threads=4)
  res = x[:512] - centers[tok]
  cb = np.zeros((len(bd), 16, dpb), np.float32)
  off = np.concatenate([[0], np.cumsum(bd)])
  for b in range(len(bd)):
    pick = rng.choice(len(res), 16, replace=False)
    cb[b, :, :bd[b]] = res[pick, off[b]:off[b + 1]]
  return cb, bd


CASES = [
    # n, d, L, dpb, noise, normalize, soar, threshold
    (3000, 32, 64, 2, 1.2, False, 1.5, float("nan")),      # SIMT tokenization (L < 256), plain hash, SOAR
    (3000, 32, 64, 2, 1.2, True, 1.5, 0.2),                # ... noise-shaped
    (6000, 100, 300, 2, 1.0, True, 1.5, 0.2),              # glove-like: tensor-core tokenization, B = 50
    (6000, 100, 300, 2, 1.0, True, None, 0.2),             # no spilling
    (4000, 96, 512, 2, 0.5, True, 1.5, float("nan")),      # deep-like, plain
    (2500, 33, 260, 2, 1.0, False, 2.0, 0.5),              # VARIABLE_CHUNK (short last block), other lambda / T
    (2500, 24, 40, 3, 1.0, True, 1.5, 0.3),                # dims_per_block 3
    (1500, 128, 256, 8, 1.0, False, None, float("nan")),   # dims_per_block 8: AVX2-order block distances
    (700, 20, 7, 1, 1.0, True, 1.5, 0.4),                  # fewer leaves than SOAR seeds, dims_per_block 1
]


@pytest.mark.parametrize("n,d,L,dpb,noise,normalize,soar,threshold", CASES)
def test_encode_database_matches_oracle(n, d, L, dpb, noise, normalize, soar, threshold):
  import oracle
  from scann_b200 import _lib
  x, centers = _mixture(n, d, L, n + d, noise, normalize)
  cb, bd = _codebook(x, centers, dpb, 7)
  tokens, codes, soar_codes, st = _lib.encode_database(x, centers, cb, bd, residual=True, soar_lambda=soar,
                                                       noise_shaping_threshold=threshold)
  o_tokens, o_codes, o_soar, o_ties = oracle.encode_database(x, centers, cb, bd, residual=True, soar_lambda=soar,
                                                             threshold=threshold, threads=8)
  np.testing.assert_array_equal(tokens, o_tokens)
  np.testing.assert_array_equal(codes, o_codes)
  if soar is not None:
    np.testing.assert_array_equal(soar_codes, o_soar)
    assert st["spilled"] == int((o_tokens[1::2] >= 0).sum())
    # the pruned search evaluates a small fraction of the N * L costs the reference computes
    assert st["soar_evaluated"] < 0.6 * n * L or L < 64
  # datapoints that donated a sub-vector to the codebook have zero residual norm in that block: genuine ties
  assert st["norm_ties"] == o_ties


def test_encode_raw_vectors_for_squared_l2_tree_ah():
  """TreeXHybridSMMD hashes the datapoint itself (no residual), plain and noise-shaped."""
  import oracle
  from scann_b200 import _lib
  x, centers = _mixture(3000, 64, 128, 3, 1.0, False)
  cb, bd = _codebook(x, np.zeros_like(centers), 2, 5)
  for thr in (float("nan"), 0.2):
    tokens, codes, soar_codes, _ = _lib.encode_database(x, centers, cb, bd, residual=False, noise_shaping_threshold=thr)
    o_tok, _ = oracle.assign_primary(x, centers, threads=8)
    o_codes, _ = oracle.encode(x, cb, bd, threshold=thr, threads=8)
    np.testing.assert_array_equal(tokens, o_tok)
    np.testing.assert_array_equal(codes, o_codes)
    assert soar_codes is None


def test_encode_chunking_and_degenerate_rows(monkeypatch):
  """Several passes (chunk smaller than N), datapoints equal to a centre (zero residual: SOAR degenerates to the
  nearest centre, not spilled) and duplicated centres (ties go to the smaller index)."""
  import oracle
  from scann_b200 import _lib
  monkeypatch.setenv("SCANN_B200_ENCODE_CHUNK", "1000")
  x, centers = _mixture(3500, 48, 300, 11, 1.0, True)
  centers[17] = centers[5]
  x[:300] = centers
  cb, bd = _codebook(x[300:], centers, 2, 9)
  tokens, codes, soar_codes, st = _lib.encode_database(x, centers, cb, bd, soar_lambda=1.5, noise_shaping_threshold=0.2)
  assert st["chunk_rows"] == 1000
  o_tokens, o_codes, o_soar, _ = oracle.encode_database(x, centers, cb, bd, soar_lambda=1.5, threshold=0.2, threads=8)
  np.testing.assert_array_equal(tokens, o_tokens)
  np.testing.assert_array_equal(codes, o_codes)
  np.testing.assert_array_equal(soar_codes, o_soar)
  assert (tokens[1:600:2] == -1).all() and not (tokens == 17).any()


@pytest.mark.parametrize("mode", ["stream", "tcgen05"])
def test_encode_with_streaming_tokenizer_and_fallback_rows(mode, monkeypatch):
  """The streaming top-P refinement (opt-in) under the encoder, incl. rows that fall back to exact
  distances (zero vectors, many equal centres): the SOAR pruning must then read that row as distances."""
  import oracle
  from scann_b200 import _lib
  monkeypatch.setenv("SCANN_B200_TOKENIZE", mode)
  x, centers = _mixture(3000, 48, 400, 17, 1.0, True)
  centers[100:140] = centers[99]               # 41 equal centres: a crowded tie window
  x[:50] = 0.0                                 # zero rows: every centre ties in the pre-filter
  x[50:90] = centers[99] * 1.0001
  cb, bd = _codebook(x[100:], centers, 2, 9)
  tokens, codes, soar_codes, st = _lib.encode_database(x, centers, cb, bd, soar_lambda=1.5)
  assert st["tokenize_fallbacks"] > 0
  o_tokens, o_codes, o_soar, _ = oracle.encode_database(x, centers, cb, bd, soar_lambda=1.5, threads=8)
  np.testing.assert_array_equal(tokens, o_tokens)
  np.testing.assert_array_equal(codes, o_codes)
  np.testing.assert_array_equal(soar_codes, o_soar)


@pytest.mark.parametrize("soar", [None, 1.5])
def test_encode_with_chunk_preselection(soar, monkeypatch):
  """The tokenizer's chunk pre-selection under squared L2 (default from 4,096 leaves; forced here on 1,024)."""
  import oracle
  from scann_b200 import _lib
  monkeypatch.setenv("SCANN_B200_TOKENIZE", "chunk")
  x, centers = _mixture(5000, 40, 1024, 23, 0.7, False)
  centers[300:330] = centers[299]              # equal centres inside and across chunks
  x[:20] = 0.0
  cb, bd = _codebook(x[100:], centers, 2, 4)
  tokens, codes, soar_codes, st = _lib.encode_database(x, centers, cb, bd, soar_lambda=soar, noise_shaping_threshold=0.2)
  o_tokens, o_codes, o_soar, _ = oracle.encode_database(x, centers, cb, bd, soar_lambda=soar, threshold=0.2, threads=8)
  np.testing.assert_array_equal(tokens, o_tokens)
  np.testing.assert_array_equal(codes, o_codes)
  if soar is not None:
    np.testing.assert_array_equal(soar_codes, o_soar)


def test_encode_rejects_bad_arguments():
  from scann_b200 import _lib
  x, centers = _mixture(100, 16, 8, 1, 1.0, False)
  cb = np.zeros((8, 16, 2), np.float32)
  with pytest.raises(_lib.ScannB200Error, match="INVALID_ARGUMENT"):
    _lib.encode_database(x, centers, cb[:7])                     # blocks cover 14 of 16 dims
  with pytest.raises(_lib.ScannB200Error, match="SOAR requires residual"):
    _lib.encode_database(x, centers, cb, residual=False, soar_lambda=1.5)


def test_built_index_searches_like_the_oracle():
  """Assets produced by the GPU encoder feed the query path: same results as the oracle on the same assets."""
  import oracle
  from scann_b200 import _lib, index_build
  x, centers = _mixture(20000, 64, 256, 21, 0.6, True)
  cb, bd = _codebook(x, centers, 2, 3)
  tokens, codes, soar_codes, _ = _lib.encode_database(x, centers, cb, bd, soar_lambda=1.5, noise_shaping_threshold=0.2)
  a = index_build.IndexArrays(distance="dot_product", dataset=x, n=x.shape[0], d=x.shape[1])
  a.centers, a.tokens, a.codes, a.soar_codes, a.codebook, a.block_dims = centers, tokens, codes, soar_codes, cb, bd
  a.soar, a.residual, a.overretrieve = True, True, 2.0
  q = x[:200] + 0.01
  ix = _lib.NativeIndex(a, 16, 100, 10)
  oi = oracle.OracleIndex(a, 16, 100, 10)
  idx, dist = ix.search_batched(q)
  o_idx, o_dist = oi.search_batched(q, impl=1, threads=8)
  np.testing.assert_array_equal(idx, o_idx)
  np.testing.assert_array_equal(dist.view(np.uint32), o_dist.view(np.uint32))


def _build_golden_names():
  import glob
  import os
  d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "build")
  return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(d, "*.npz")))


@pytest.mark.parametrize("name", _build_golden_names())
def test_encode_database_matches_committed_golden(name):
  """No oracle in the loop: the GPU encoder against the committed fixtures (oracle/gen_golden_build.py)."""
  import os
  from scann_b200 import _lib
  z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "build", name + ".npz"))
  soar = None if np.isnan(z["soar_lambda"]) else float(z["soar_lambda"])
  tokens, codes, soar_codes, st = _lib.encode_database(z["x"], z["centers"], z["codebook"], z["block_dims"],
                                                       residual=bool(z["residual"]), soar_lambda=soar,
                                                       noise_shaping_threshold=float(z["threshold"]))
  np.testing.assert_array_equal(tokens, z["exp_tokens"])
  np.testing.assert_array_equal(codes, z["exp_codes"])
  if soar is not None:
    np.testing.assert_array_equal(soar_codes, z["exp_soar_codes"])
  assert st["norm_ties"] == int(z["exp_ties"])


@pytest.mark.parametrize("n,d,L,dpb,soar,thr", [
    (1, 8, 5, 2, 1.5, 0.2),          # a single datapoint
    (33, 2, 3, 2, 1.5, float("nan")),  # one AH block
    (40, 6, 1, 2, 1.5, 0.2),         # one leaf: the secondary can only be the primary, nothing is spilled
    (257, 5, 9, 1, None, 0.5),       # dims_per_block 1, odd D
    (64, 16, 300, 4, 2.5, 0.2),      # more leaves than datapoints (tensor-core tokenization with empty leaves)
])
def test_encode_edge_shapes(n, d, L, dpb, soar, thr):
  import oracle
  from scann_b200 import _lib
  rng = np.random.default_rng(n * 7 + d)
  x = rng.standard_normal((n, d)).astype(np.float32)
  centers = rng.standard_normal((L, d)).astype(np.float32)
  full, part = divmod(d, dpb)
  bd = np.asarray([dpb] * full + ([part] if part else []), np.int32)
  cb = (0.5 * rng.standard_normal((len(bd), 16, dpb))).astype(np.float32)
  for b in range(len(bd)):
    cb[b, :, bd[b]:] = 0
  tokens, codes, soar_codes, st = _lib.encode_database(x, centers, cb, bd, soar_lambda=soar, noise_shaping_threshold=thr)
  o_tokens, o_codes, o_soar, _ = oracle.encode_database(x, centers, cb, bd, soar_lambda=soar, threshold=thr, threads=2)
  np.testing.assert_array_equal(tokens, o_tokens)
  np.testing.assert_array_equal(codes, o_codes)
  if soar is not None:
    np.testing.assert_array_equal(soar_codes, o_soar)
    if L == 1:
      assert (tokens[1::2] == -1).all() and st["spilled"] == 0


def test_encode_empty_database():
  from scann_b200 import _lib
  x = np.zeros((0, 8), np.float32)
  centers = np.ones((4, 8), np.float32)
  cb = np.zeros((4, 16, 2), np.float32)
  tokens, codes, soar_codes, st = _lib.encode_database(x, centers, cb)
  assert tokens.shape == (0,) and codes.shape == (0, 4) and soar_codes is None
