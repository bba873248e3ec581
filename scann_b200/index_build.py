"""Index construction (k-means partitioner, residual AH codebooks, encoding, SOAR).

This is the part of `builder(...).build()` that sits *before* the hot path
(SURVEY.md §3.2, §8f rank 1): it only has to produce assets in the reference's
format so that the query path has something to search.  It is written with
torch so that CPU-only test fixtures can be generated here; on a CUDA device the trainers
(`scann_b200_train_kmeans`) and the per-datapoint stage (`scann_b200_encode_database`) run in the library.

What it mirrors in the reference (behaviour, not code):
  * k-means partitioner with SquaredL2 partitioning distance
    (`scann_builder.py:213-238`, `partitioning/kmeans_tree_partitioner.cc:424-441`)
  * database tokenisation = nearest centre (`kmeans_tree_partitioner.cc:475-620`)
  * SOAR secondary assignment with the orthogonality-amplified cost
    ||x-c||^2 + lambda * <x-c, r_hat>^2 (`partitioning/orthogonality_amplification_utils.h:51-68`)
  * residual = x - centre for dot-product tree-AH (`tree_ah_hybrid_residual.cc:145-224`)
  * per-block 16-centre k-means codebooks and nearest-centre encoding
    (`hashes/internal/asymmetric_hashing_impl.cc:41-244`)
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch


@dataclass
class IndexArrays:
  """Host-side arrays that describe one searcher; exactly the serialized assets."""
  distance: str                       # "dot_product" | "squared_l2"
  dataset: Optional[np.ndarray]       # [N, D] f32 (dataset.npy) or None
  centers: Optional[np.ndarray] = None        # [L, D] f32 (serialized_partitioner.pb)
  tokens: Optional[np.ndarray] = None         # [N] i32 or [2N] i32 (datapoint_to_token.npy)
  codes: Optional[np.ndarray] = None          # [N, B] u8 (hashed_dataset.npy)
  soar_codes: Optional[np.ndarray] = None     # [N, B] u8 (hashed_dataset_soar.npy)
  codebook: Optional[np.ndarray] = None       # [B, 16, dpb] f32, zero padded (ah_codebook.pb)
  block_dims: Optional[np.ndarray] = None     # [B] i32 real dims of each block
  bf16_dataset: Optional[np.ndarray] = None   # [N, D] i16 (bfloat16_dataset.npy)
  int8_dataset: Optional[np.ndarray] = None   # [N, D] i8 (int8_dataset.npy): fixed-point reordering
  int8_multipliers: Optional[np.ndarray] = None  # [D] f32 (int8_multipliers.npy)
  dp_norms: Optional[np.ndarray] = None       # [N] f32 (dp_norms.npy), squared L2 only
  # query_tokenization_type FIXED_POINT_INT8 (tree(quantize_centroids=True)): the searcher tokenizes queries against
  # the fixed-point image of `centers` (derived at load time, as KMeansTreeNode::CreateFixedPointCenters does)
  int8_tokenization: bool = False
  n: int = 0
  d: int = 0
  residual: bool = False
  soar: bool = False
  overretrieve: float = 2.0
  meta: dict = field(default_factory=dict)


def _dev(device):
  if device is not None:
    return torch.device(device)
  return torch.device("cuda:0" if torch.cuda.is_available() else "cpu")


def _sqdist_argmin(x, centers, c_norms=None, chunk=65536, exclude=None, return_dist=False):
  """argmin_c ||x - c||^2 for every row of x (chunked GEMM)."""
  if c_norms is None:
    c_norms = (centers * centers).sum(1)
  out = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
  dist = torch.empty(x.shape[0], dtype=torch.float32, device=x.device) if return_dist else None
  for s in range(0, x.shape[0], chunk):
    xb = x[s:s + chunk]
    d = c_norms[None, :] - 2.0 * (xb @ centers.T)
    if exclude is not None:
      d.scatter_(1, exclude[s:s + chunk, None], float("inf"))
    m = d.min(1)
    out[s:s + chunk] = m.indices
    if return_dist:
      dist[s:s + chunk] = m.values + (xb * xb).sum(1)
  return (out, dist) if return_dist else out


def _use_native_trainer(device, spherical=False):
  """The library's trainer (csrc/train.cu) runs whenever a CUDA device is the target; torch only serves CPU-only
  fixtures (device="cpu") and spherical k-means (the library does not normalise centroids)."""
  return (not spherical) and torch.cuda.is_available() and _dev(device).type == "cuda"


def _reseed_empty(x, centers, assign, device_index, extra_iters=2, rounds=4):
  """The library keeps an empty cluster's centre (the reference re-initialises it from an unseeded generator,
  `utils/gmm_utils.cc:1204-1330`); here empty clusters restart from the points farthest from their centre, followed by
  `extra_iters` more Lloyd iterations, so every leaf ends up non-empty when N >= k."""
  from scann_b200 import _lib
  k = centers.shape[0]
  for _ in range(rounds):
    empty = np.flatnonzero(np.bincount(assign, minlength=k) == 0)
    if not len(empty):
      break
    diff = x - centers[assign]
    far = np.argsort(-np.einsum("ij,ij->i", diff, diff), kind="stable")[:len(empty)]
    centers = centers.copy()
    centers[empty] = x[far]
    centers, assign, _ = _lib.train_kmeans(x, centers, extra_iters, device=device_index)
  return centers, assign


def train_kmeans_native(x, k, iters=12, seed=0, device=None):
  """`scann_b200_train_kmeans` (csrc/train.cu): tensor-core assignment + the reference's double-precision centroid
  update; initial centres = a seeded sample of the data (the reference's k-means++ draws from an unseeded generator)."""
  from scann_b200 import _lib
  x = np.ascontiguousarray(x, dtype=np.float32)
  n = x.shape[0]
  if k > n:
    raise ValueError(f"k={k} > n={n}")
  dev_index = _dev(device).index or 0
  g = torch.Generator(device="cpu").manual_seed(seed)
  init = x[torch.randperm(n, generator=g)[:k].numpy()]
  centers, assign, _ = _lib.train_kmeans(x, init, iters, device=dev_index)
  centers, _ = _reseed_empty(x, centers, assign, dev_index)
  return centers


def train_kmeans(x, k, iters=12, seed=0, spherical=False, device=None):
  """Lloyd iterations, random initialisation from the data: the library's trainer on a CUDA device, plain torch
  otherwise (CPU-only test fixtures, spherical k-means).

  Empty clusters are re-seeded from the points currently farthest from their
  centre, so every leaf ends up non-empty when N >= k.
  """
  if _use_native_trainer(device, spherical):
    return train_kmeans_native(x, k, iters=iters, seed=seed, device=device)
  dev = _dev(device)
  x = torch.as_tensor(x, device=dev, dtype=torch.float32)
  n = x.shape[0]
  g = torch.Generator(device="cpu").manual_seed(seed)
  if k > n:
    raise ValueError(f"k={k} > n={n}")
  perm = torch.randperm(n, generator=g)[:k].to(dev)
  centers = x[perm].clone()
  for _ in range(iters):
    assign, dist = _sqdist_argmin(x, centers, return_dist=True)
    sums = torch.zeros_like(centers)
    sums.index_add_(0, assign, x)
    counts = torch.bincount(assign, minlength=k).to(torch.float32)
    empty = counts == 0
    new_centers = sums / counts.clamp(min=1.0)[:, None]
    n_empty = int(empty.sum())
    if n_empty:
      far = torch.topk(dist, n_empty).indices
      new_centers[empty] = x[far]
    if spherical:
      new_centers = new_centers / new_centers.norm(dim=1, keepdim=True).clamp(min=1e-12)
    centers = new_centers
  return centers.cpu().numpy()


def tokenize_database(x, centers, device=None, chunk=65536):
  dev = _dev(device)
  c = torch.as_tensor(centers, device=dev)
  out = np.empty(x.shape[0], dtype=np.int32)
  cn = (c * c).sum(1)
  big = 1 << 20
  for s in range(0, x.shape[0], big):
    xb = torch.as_tensor(x[s:s + big], device=dev)
    out[s:s + big] = _sqdist_argmin(xb, c, cn, chunk).to(torch.int32).cpu().numpy()
  return out


def soar_assign(x, centers, primary, lam=1.5, device=None, chunk=16384):
  """Secondary (spilled) leaf per datapoint, SOAR cost (orthogonality amplification)."""
  dev = _dev(device)
  c = torch.as_tensor(centers, device=dev)
  cn = (c * c).sum(1)
  out = np.empty(x.shape[0], dtype=np.int32)
  for s in range(0, x.shape[0], chunk):
    xb = torch.as_tensor(x[s:s + chunk], device=dev)
    p = torch.as_tensor(primary[s:s + chunk], device=dev, dtype=torch.int64)
    r = xb - c[p]
    rn = r.norm(dim=1, keepdim=True)
    rhat = torch.where(rn * rn < 1e-7, torch.zeros_like(r), r / rn.clamp(min=1e-30))
    # ||x-c||^2 for all c
    t1 = (xb * xb).sum(1, keepdim=True) + cn[None, :] - 2.0 * (xb @ c.T)
    # <x-c, rhat> = <x,rhat> - <c,rhat>
    t2 = (xb * rhat).sum(1, keepdim=True) - rhat @ c.T
    cost = t1 + lam * t2 * t2
    cost.scatter_(1, p[:, None], float("inf"))
    out[s:s + chunk] = cost.argmin(1).to(torch.int32).cpu().numpy()
  return out


def block_layout(d, dims_per_block):
  """CHUNK / VARIABLE_CHUNK projection: contiguous slices, last one may be short
  (`scann_builder.py:275-294`, `projection/chunking_projection.cc:153-213`)."""
  full, part = divmod(d, dims_per_block)
  dims = [dims_per_block] * full + ([part] if part else [])
  return np.asarray(dims, dtype=np.int32)


def _to_blocks(x, block_dims, dpb):
  """[n, D] -> [B, n, dpb] zero padded."""
  n, d = x.shape
  b = len(block_dims)
  if b * dpb != d:
    x = torch.nn.functional.pad(x, (0, b * dpb - d))
  return x.view(n, b, dpb).permute(1, 0, 2).contiguous()


def train_ah_codebook_native(x, dims_per_block, iters=10, seed=0, sample=100000, device=None):
  """One `scann_b200_train_kmeans` call per block (16 centres over the block's real dims,
  `hashes/internal/asymmetric_hashing_impl.cc:41-197`) -> ([B, 16, dpb] f32 zero padded, block_dims)."""
  from scann_b200 import _lib
  x = np.ascontiguousarray(x, dtype=np.float32)
  n, d = x.shape
  block_dims = block_layout(d, dims_per_block)
  g = np.random.default_rng(seed)
  xs = x[np.sort(g.choice(n, size=sample, replace=False))] if n > sample else x
  ns = xs.shape[0]
  init = g.choice(ns, size=16, replace=ns < 16)
  dev_index = _dev(device).index or 0
  cb = np.zeros((len(block_dims), 16, dims_per_block), np.float32)
  off = 0
  for b, bd in enumerate(block_dims):
    sub = np.ascontiguousarray(xs[:, off:off + bd])
    if ns >= 16:
      c, a, _ = _lib.train_kmeans(sub, sub[init], iters, device=dev_index)
      c, _ = _reseed_empty(sub, c, a, dev_index)
    else:
      c = sub[init]
    cb[b, :, :bd] = c
    off += bd
  return cb, block_dims


def train_ah_codebook(x, dims_per_block, iters=10, seed=0, sample=100000, device=None):
  """16-centre k-means per block on (residual) sub-vectors -> [B, 16, dpb] f32."""
  if _use_native_trainer(device):
    return train_ah_codebook_native(x, dims_per_block, iters=iters, seed=seed, sample=sample, device=device)
  dev = _dev(device)
  n, d = x.shape
  block_dims = block_layout(d, dims_per_block)
  g = np.random.default_rng(seed)
  if n > sample:
    sel = np.sort(g.choice(n, size=sample, replace=False))
    xs = torch.as_tensor(x[sel], device=dev)
  else:
    xs = torch.as_tensor(x, device=dev)
  xb = _to_blocks(xs, block_dims, dims_per_block)            # [B, n, dpb]
  nb, ns, _ = xb.shape
  init = torch.as_tensor(g.choice(ns, size=16, replace=ns < 16), device=dev)
  cb = xb[:, init, :].clone()                                 # [B, 16, dpb]
  for _ in range(iters):
    dist = torch.cdist(xb, cb) ** 2                           # [B, n, 16]
    a = dist.argmin(2)                                        # [B, n]
    onehot = torch.nn.functional.one_hot(a, 16).to(torch.float32)   # [B, n, 16]
    counts = onehot.sum(1)                                    # [B, 16]
    sums = torch.einsum("bnk,bnd->bkd", onehot, xb)
    newcb = sums / counts.clamp(min=1.0)[:, :, None]
    empty = counts == 0
    if bool(empty.any()):
      worst = dist.min(2).values.topk(16, dim=1).indices      # [B, 16]
      repl = torch.gather(xb, 1, worst[:, :, None].expand(-1, -1, xb.shape[2]))
      newcb = torch.where(empty[:, :, None], repl, newcb)
    cb = newcb
  return cb.cpu().numpy().astype(np.float32), block_dims


def encode_ah(x, codebook, block_dims, device=None, chunk=1 << 17):
  """Nearest-centre code per block -> [N, B] u8 in 0..15."""
  dev = _dev(device)
  cb = torch.as_tensor(codebook, device=dev)
  dpb = cb.shape[2]
  out = np.empty((x.shape[0], cb.shape[0]), dtype=np.uint8)
  cbn = (cb * cb).sum(2)                                      # [B, 16]
  for s in range(0, x.shape[0], chunk):
    xb = _to_blocks(torch.as_tensor(x[s:s + chunk], device=dev), block_dims, dpb)  # [B, n, dpb]
    dist = cbn[:, None, :] - 2.0 * torch.bmm(xb, cb.transpose(1, 2))              # [B, n, 16]
    out[s:s + chunk] = dist.argmin(2).T.to(torch.uint8).cpu().numpy()
  return out


def bfloat16_quantize(x):
  """`Bfloat16Quantize`: (bits + 0x8000) >> 16, saturating (`utils/bfloat16_helpers.h:30-48`)."""
  bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
  r = ((bits + 0x8000) >> 16).astype(np.uint32)
  # saturate instead of rounding a finite value up to inf
  exp_all_ones = ((r >> 7) & 0xFF) == 0xFF
  was_finite = ((bits >> 23) & 0xFF) != 0xFF
  r = np.where(exp_all_ones & was_finite, r - 1, r)
  return r.astype(np.uint16).view(np.int16)


def int8_quantize(x, chunk=1 << 18):
  """`ScalarQuantizeFloatDataset` with multiplier quantile 1.0 (`utils/scalar_quantization_helpers.cc:39-63,94-145`):
  multiplier[d] = 127 / max|x[:, d]| (1 where the column is zero), value = clamp(round_half_away(x * multiplier)).
  Returns (int8 [N, D], multipliers [D] f32)."""
  x = np.ascontiguousarray(x, dtype=np.float32)
  mx = np.zeros(x.shape[1], np.float32)
  for s in range(0, x.shape[0], chunk):
    np.maximum(mx, np.abs(x[s:s + chunk]).max(0), out=mx)
  mult = np.where(mx == 0, np.float32(1.0), np.float32(127.0) / np.where(mx == 0, np.float32(1.0), mx)).astype(np.float32)
  out = np.empty(x.shape, np.int8)
  for s in range(0, x.shape[0], chunk):
    v = x[s:s + chunk] * mult[None, :]                       # f32 product, as Int8Quantize receives it
    r = np.trunc(v)
    r += np.where(np.abs(v - r) >= np.float32(0.5), np.sign(v), np.float32(0)).astype(np.float32)   # std::round
    out[s:s + chunk] = np.clip(r, -128, 127).astype(np.int8)
  return out, mult


def quantize_centers(centers):
  """Fixed-point centres of int8 query tokenization: `KMeansTreeNode::CreateFixedPointCenters`
  (trees/kmeans_tree/kmeans_tree_node.cc:267-281) = `ScalarQuantizeFloatDataset(float_centers, 1.0, NaN)`, its
  `inverse_multiplier_by_dimension` (1.0f / multiplier, scalar_quantization_helpers.cc:133-136) and the squared norms
  of the FLOAT centres.  Returns (int8 [L, D], inverse multipliers [D] f32, squared norms [L] f32).  The library derives
  the same arrays in C++ (csrc/index.cu); this numpy restatement feeds the oracle and the tests."""
  c = np.ascontiguousarray(centers, dtype=np.float32)
  ci8, mult = int8_quantize(c)
  return ci8, (np.float32(1.0) / mult).astype(np.float32), squared_l2_norms(c)


def squared_l2_norms(x, chunk=1 << 18):
  """float(SquaredL2Norm(row)) with DenseSingleAccumulate's four strided double accumulators
  (`utils/reduction.h:357-390`): what dp_norms.npy holds (`utils/reordering_helper.cc:586-591`)."""
  x = np.ascontiguousarray(x, dtype=np.float32)
  n, d = x.shape
  out = np.empty(n, np.float32)
  d4 = d - d % 4
  for s in range(0, n, chunk):
    sq = x[s:s + chunk].astype(np.float64) ** 2
    r = [np.cumsum(sq[:, l:d4:4], axis=1)[:, -1] if d4 else np.zeros(sq.shape[0]) for l in range(4)]
    i = d4
    r[2] = r[2] + r[3]
    if i + 2 <= d:
      r[0] = r[0] + sq[:, i]
      r[1] = r[1] + sq[:, i + 1]
      i += 2
    r[1] = r[1] + r[2]
    if i < d:
      r[0] = r[0] + sq[:, i]
    out[s:s + chunk] = (r[0] + r[1]).astype(np.float32)
  return out


def build_tree_ah(db, distance="dot_product", num_leaves=100, dims_per_block=2,
                  training_sample_size=100000, tree_iters=12, ah_iters=10,
                  soar_lambda=None, overretrieve=2.0, spherical=False,
                  seed=0, device=None, keep_dataset=True, noise_shaping_threshold=float("nan"),
                  native_encode=None):
  """tree().score_ah() index: returns IndexArrays (the serialized asset set).

  Training (k-means tree, AH codebooks) runs through the C ABI on a CUDA device (`scann_b200_train_kmeans`,
  csrc/train.cu; torch Lloyd iterations only for device="cpu" fixtures and spherical k-means); the per-datapoint
  stage -- database tokenization, SOAR
  secondary assignment, residuals, AH encoding incl. noise shaping -- runs through the C ABI
  (`scann_b200_encode_database`, csrc/encode.cu) whenever a CUDA device is present (`native_encode=None`)
  and reproduces the reference's arithmetic bit for bit.  The torch encoder below it is only what CPU-only
  test fixtures are generated with: plain nearest-centre codes, primary excluded from the SOAR choice.
  """
  db = np.ascontiguousarray(db, dtype=np.float32)
  n, d = db.shape
  residual = distance == "dot_product"
  if soar_lambda is not None and distance != "dot_product":
    raise ValueError("SOAR requires dot product distance.")
  rng = np.random.default_rng(seed)
  if n > training_sample_size:
    sel = np.sort(rng.choice(n, size=training_sample_size, replace=False))
    sample = db[sel]
  else:
    sample = db
  num_leaves = min(num_leaves, n)
  centers = train_kmeans(sample, num_leaves, iters=tree_iters, seed=seed,
                         spherical=spherical, device=device)
  arr = IndexArrays(distance=distance, dataset=db if keep_dataset else None, n=n, d=d)
  arr.centers = centers.astype(np.float32)
  arr.meta["trainer"] = "scann_b200_train_kmeans" if _use_native_trainer(device, spherical) else "torch"
  arr.residual = residual
  if native_encode is None:
    native_encode = torch.cuda.is_available()
  if native_encode:
    from scann_b200 import _lib
    res_s = sample - centers[tokenize_database(sample, centers, device=device)] if residual else sample
    cb, block_dims = train_ah_codebook(res_s, dims_per_block, iters=ah_iters, seed=seed + 1,
                                       sample=training_sample_size, device=device)
    arr.codebook, arr.block_dims = cb, block_dims
    dev_index = 0 if device is None else (torch.device(device).index or 0)
    arr.tokens, arr.codes, arr.soar_codes, arr.meta["encode_stats"] = _lib.encode_database(
        db, arr.centers, cb, block_dims, residual=residual, soar_lambda=soar_lambda,
        noise_shaping_threshold=noise_shaping_threshold, device=dev_index)
    if soar_lambda is not None:
      arr.soar = True
      arr.overretrieve = float(overretrieve)
    return arr
  tok = tokenize_database(db, centers, device=device)
  if residual:
    res = db - centers[tok]
  else:
    res = db
  cb, block_dims = train_ah_codebook(res, dims_per_block, iters=ah_iters, seed=seed + 1,
                                     sample=training_sample_size, device=device)
  arr.codebook, arr.block_dims = cb, block_dims
  arr.codes = encode_ah(res, cb, block_dims, device=device)
  if soar_lambda is not None:
    sec = soar_assign(db, centers, tok, lam=soar_lambda, device=device)
    res2 = db - centers[sec]
    arr.soar_codes = encode_ah(res2, cb, block_dims, device=device)
    t2 = np.empty(2 * n, dtype=np.int32)
    # slot 2i = lower-numbered leaf, 2i+1 = the other one (`scann.cc:534-555`)
    lo = np.minimum(tok, sec)
    hi = np.maximum(tok, sec)
    t2[0::2] = lo
    t2[1::2] = hi
    # codes[] must belong to the leaf stored in slot 2i, soar_codes[] to slot 2i+1
    swap = sec < tok
    if swap.any():
      c0 = arr.codes.copy()
      arr.codes[swap] = arr.soar_codes[swap]
      arr.soar_codes[swap] = c0[swap]
    arr.tokens = t2
    arr.soar = True
    arr.overretrieve = float(overretrieve)
  else:
    arr.tokens = tok.astype(np.int32)
  return arr
