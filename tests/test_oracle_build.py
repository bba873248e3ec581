"""CPU tests: the oracle's index-build stage (database tokenization, SOAR, AH encoding)
against independent numpy / pure-Python restatements and float64 ground truth."""
import math

import numpy as np
import pytest

import oracle
from helpers import f32, fma32


def _data(n, d, L, seed, noise=0.4):
  rng = np.random.default_rng(seed)
  means = rng.standard_normal((L, d)).astype(np.float32)
  x = (means[rng.integers(0, L, n)] + noise * rng.standard_normal((n, d))).astype(np.float32)
  centers = (means + 0.05 * rng.standard_normal((L, d))).astype(np.float32)
  return x, centers


def _codebook(B, dpb, seed, scale=0.4):
  rng = np.random.default_rng(seed)
  return (scale * rng.standard_normal((B, 16, dpb))).astype(np.float32)


def np_l2_chain(x, centers):
  """many_to_many_impl.inc:236-257,522-567: ||c||^2 + ||x||^2, then fnmadd(x[d], 2 c[d], acc) sequentially."""
  cn = np.zeros(centers.shape[0], np.float32)
  for k in range(centers.shape[1]):
    cn = fma32(-centers[:, k], centers[:, k], cn)
  cn = f32(cn * np.float32(-1.0))
  qn = (x.astype(np.float64) ** 2).sum(1).astype(np.float32)
  acc = f32(cn[None, :] + qn[:, None])
  for k in range(x.shape[1]):
    acc = fma32(-x[:, k:k + 1], f32(centers[None, :, k] * np.float32(2.0)), acc)
  return acc


def test_assign_primary_matches_numpy_chain_and_f64():
  x, centers = _data(600, 24, 37, 1)
  tok, dist = oracle.assign_primary(x, centers, threads=2)
  acc = np_l2_chain(x, centers)
  np.testing.assert_array_equal(tok, acc.argmin(1).astype(np.int32))       # argmin = first minimum
  np.testing.assert_array_equal(dist.view(np.uint32), acc[np.arange(len(x)), tok].view(np.uint32))
  d64 = ((x.astype(np.float64)[:, None, :] - centers.astype(np.float64)[None]) ** 2).sum(-1)
  np.testing.assert_allclose(dist, d64.min(1), rtol=1e-4, atol=1e-4)
  assert (tok == d64.argmin(1)).mean() > 0.995


def test_assign_primary_tie_goes_to_the_smaller_centre():
  x, centers = _data(50, 8, 10, 2)
  centers[7] = centers[3]                      # two identical centres: every tie must resolve to 3
  tok, _ = oracle.assign_primary(x, centers)
  assert not (tok == 7).any()


def np_soar(x, centers, primary, lam):
  n, d = x.shape
  r = (x.astype(np.float64) - centers[primary].astype(np.float64)).astype(np.float32)
  sq = np.zeros(n, np.float64)
  for k in range(d):
    sq = sq + r[:, k].astype(np.float64) * r[:, k].astype(np.float64)
  inv = (1.0 / np.sqrt(np.where(sq < 1e-7, 1.0, sq))).astype(np.float32)
  rhat = np.where((sq < 1e-7)[:, None], np.float32(0), f32(r * inv[:, None]))
  t1 = np.zeros((n, centers.shape[0]), np.float32)
  t2 = np.zeros_like(t1)
  for k in range(d):
    diff = f32(x[:, k:k + 1] - centers[None, :, k])
    t1 = fma32(diff, diff, t1)
    t2 = fma32(diff, np.broadcast_to(rhat[:, k:k + 1], diff.shape), t2)
  cost = f32(t1 + f32(f32(np.float32(lam) * t2) * t2))
  return cost


def test_assign_soar_matches_numpy_chain_and_f64():
  x, centers = _data(400, 20, 29, 3, noise=1.5)   # overlapping clusters: the primary (cost 2.5 ||r||^2) rarely wins
  prim, _ = oracle.assign_primary(x, centers)
  sec, cost = oracle.assign_soar(x, centers, prim, 1.5, threads=2)
  ref = np_soar(x, centers, prim, 1.5)
  np.testing.assert_array_equal(sec, ref.argmin(1).astype(np.int32))
  np.testing.assert_array_equal(cost.view(np.uint32), ref[np.arange(len(x)), sec].view(np.uint32))
  # float64 statement of the SOAR cost (orthogonality_amplification_utils.h:48-68)
  r = x.astype(np.float64) - centers[prim].astype(np.float64)
  rhat = r / np.linalg.norm(r, axis=1, keepdims=True)
  diff = x.astype(np.float64)[:, None, :] - centers.astype(np.float64)[None]
  c64 = (diff ** 2).sum(-1) + 1.5 * np.einsum("nld,nd->nl", diff, rhat) ** 2
  np.testing.assert_allclose(cost, c64.min(1), rtol=1e-4)
  assert (sec == c64.argmin(1)).mean() > 0.99
  # the primary is a legal answer (cost (1 + lambda) ||r||^2); such datapoints are not spilled
  assert 0.5 < (sec != prim).mean() < 1.0


def test_soar_zero_residual_degenerates_to_nearest_centre():
  x, centers = _data(40, 8, 9, 4)
  x[:9] = centers                              # residual 0 => rhat = 0 => cost = squared L2, minimum at the primary
  prim, _ = oracle.assign_primary(x, centers)
  sec, cost = oracle.assign_soar(x, centers, prim, 1.5)
  np.testing.assert_array_equal(sec[:9], prim[:9])
  assert (cost[:9] == 0).all()


@pytest.mark.parametrize("d,dpb", [(32, 2), (33, 2), (24, 3), (32, 1)])
def test_plain_hash_is_first_nearest_centre(d, dpb):
  rng = np.random.default_rng(d * 10 + dpb)
  x = rng.standard_normal((300, d)).astype(np.float32)
  full, part = divmod(d, dpb)
  bd = np.asarray([dpb] * full + ([part] if part else []), np.int32)
  B = len(bd)
  cb = _codebook(B, dpb, 5, 1.0)
  off = np.concatenate([[0], np.cumsum(bd)])
  for b in range(B):
    cb[b, :, bd[b]:] = 0
  cb[1, 9] = cb[1, 4]                          # duplicate centre: min_element returns the first
  codes, _ = oracle.encode(x, cb, bd)
  assert not (codes[:, 1] == 9).any()
  d64 = np.stack([((x[:, off[b]:off[b + 1]].astype(np.float64)[:, None, :] -
                    cb[b, :, :bd[b]].astype(np.float64)[None]) ** 2).sum(-1) for b in range(B)], 1)   # [n, B, 16]
  chosen = np.take_along_axis(d64, codes[..., None].astype(np.int64), 2)[..., 0]
  np.testing.assert_allclose(chosen, d64.min(2), rtol=1e-5, atol=1e-6)
  if dpb == 2 and part == 0:
    # bit-exact numpy restatement for two dims: f32(f32(t0^2) + f32(t1^2)), no FMA (one_to_many_symmetric.h:691-800)
    xb = x.reshape(len(x), B, 1, 2)
    t = f32(xb - cb[None])
    dist = f32(f32(t[..., 0] * t[..., 0]) + f32(t[..., 1] * t[..., 1]))
    np.testing.assert_array_equal(codes, dist.argmin(2).astype(np.uint8))


def py_noise_shaped(res, orig, cb, bd, threshold):
  """asymmetric_hashing_impl.cc:258-503 in Python floats (IEEE double, no FMA)."""
  B = len(bd)
  off = [0]
  for v in bd:
    off.append(off[-1] + int(v))
  D = off[-1]
  cn = 0.0
  for k in range(D):
    cn += float(orig[k]) * float(orig[k])
  inv = 1.0 / math.sqrt(cn)
  norm = [[0.0] * 16 for _ in range(B)]
  par = [[0.0] * 16 for _ in range(B)]
  for b in range(B):
    for c in range(16):
      rn = p = 0.0
      for k in range(int(bd[b])):
        rc = float(res[off[b] + k]) - float(cb[b, c, k])
        rn += rc * rc
        p += rc * float(orig[off[b] + k]) * inv
      norm[b][c], par[b][c] = rn, p
  r = [0.0] * 4
  i = 0
  while i + 4 <= D:
    for l in range(4):
      r[l] += float(orig[i + l]) * float(orig[i + l])
    i += 4
  r[2] += r[3]
  if i + 2 <= D:
    r[0] += float(orig[i]) * float(orig[i])
    r[1] += float(orig[i + 1]) * float(orig[i + 1])
    i += 2
  r[1] += r[2]
  if i < D:
    r[0] += float(orig[i]) * float(orig[i])
  sqn = r[0] + r[1]
  mult = (threshold * threshold / sqn) / ((1.0 - threshold * threshold / sqn) / (D - 1.0))
  code = [min(range(16), key=lambda c: (norm[b][c], c)) for b in range(B)]
  p = 0.0
  for b in range(B):
    p += par[b][code[b]]
  order = sorted(range(B), key=lambda b: (-norm[b][code[b]], b))
  changes, rnd = True, 0
  while changes and rnd < 10:
    changes = False
    for b in order:
      cur = code[b]
      best, best_delta, best_par = cur, 0.0, p
      for c in range(16):
        if c == cur:
          continue
        new_par = p - par[b][cur] + par[b][c]
        pd = new_par * new_par - p * p
        if pd > 0.0:
          continue
        nd = norm[b][c] - norm[b][cur]
        cd = mult * pd + (nd - pd)
        if cd < best_delta:
          best, best_delta, best_par = c, cd, new_par
      if best != cur:
        p, code[b], changes = best_par, best, True
    rnd += 1
  return code


@pytest.mark.parametrize("d,dpb", [(16, 2), (17, 2), (12, 3)])
def test_noise_shaped_hash_matches_python_restatement(d, dpb):
  x, centers = _data(120, d, 7, d)
  x = f32(x / np.linalg.norm(x, axis=1, keepdims=True))          # unit norm: with T = 0.5 eta > 1 and the descent moves codes
  centers = f32(centers / np.linalg.norm(centers, axis=1, keepdims=True))
  full, part = divmod(d, dpb)
  bd = np.asarray([dpb] * full + ([part] if part else []), np.int32)
  cb = _codebook(len(bd), dpb, 6, 0.1)
  prim, _ = oracle.assign_primary(x, centers)
  codes, ties = oracle.encode(x, cb, bd, centers=centers, token=prim, threshold=0.5)
  assert ties == 0
  assert (codes != oracle.encode(x, cb, bd, centers=centers, token=prim)[0]).any()
  res = f32(x - centers[prim])
  for i in range(len(x)):
    assert codes[i].tolist() == py_noise_shaped(res[i], x[i], cb, bd, 0.5), i
  # raw (non-residual) hashing, TreeXHybridSMMD: residual == original
  codes_raw, _ = oracle.encode(x[:40], cb, bd, threshold=0.5)
  for i in range(40):
    assert codes_raw[i].tolist() == py_noise_shaped(x[i], x[i], cb, bd, 0.5), i


def test_noise_shaping_lowers_the_anisotropic_loss():
  d, dpb = 32, 2
  x, centers = _data(500, d, 11, 8)
  # unit-norm data: T = 0.2 then gives eta = (T^2 / ||x||^2) / ((1 - T^2 / ||x||^2) / (D - 1)) = 1.29 > 1
  x = f32(x / np.linalg.norm(x, axis=1, keepdims=True))
  centers = f32(centers / np.linalg.norm(centers, axis=1, keepdims=True))
  bd = np.full(d // dpb, dpb, np.int32)
  cb = _codebook(len(bd), dpb, 9, 0.08)
  prim, _ = oracle.assign_primary(x, centers)
  plain, _ = oracle.encode(x, cb, bd, centers=centers, token=prim)
  shaped, _ = oracle.encode(x, cb, bd, centers=centers, token=prim, threshold=0.2)
  assert (plain != shaped).any()
  res = (x - centers[prim]).astype(np.float64)

  def losses(codes):
    recon = np.concatenate([cb[b][codes[:, b]] for b in range(len(bd))], axis=1).astype(np.float64)
    e = res - recon
    xn = x.astype(np.float64)
    par = (e * xn).sum(1) / np.linalg.norm(xn, axis=1)
    return (e ** 2).sum(1), par ** 2

  n_p, p_p = losses(plain)
  n_s, p_s = losses(shaped)
  sq = (x.astype(np.float64) ** 2).sum(1)
  eta = (0.04 / sq) / ((1 - 0.04 / sq) / (d - 1))
  loss_p = eta * p_p + (n_p - p_p)
  loss_s = eta * p_s + (n_s - p_s)
  assert (loss_s <= loss_p + 1e-12).all()        # coordinate descent only accepts improving moves
  assert (n_s >= n_p - 1e-12).all()              # plain hashing minimises the residual norm


def test_encode_database_layout_with_soar():
  x, centers = _data(300, 16, 13, 10, noise=1.2)
  bd = np.full(8, 2, np.int32)
  cb = _codebook(8, 2, 11)
  tokens, codes, soar_codes, _ = oracle.encode_database(x, centers, cb, bd, residual=True, soar_lambda=1.5, threshold=0.2)
  prim, _ = oracle.assign_primary(x, centers)
  sec, _ = oracle.assign_soar(x, centers, prim, 1.5)
  lo, hi = tokens[0::2], tokens[1::2]
  spilled = sec != prim
  assert ((hi == -1) == ~spilled).all()
  assert (lo[spilled] < hi[spilled]).all()
  assert (np.minimum(prim, sec)[spilled] == lo[spilled]).all() and (lo[~spilled] == prim[~spilled]).all()
  assert (soar_codes[~spilled] == 0).all()
  c_lo, _ = oracle.encode(x, cb, bd, centers=centers, token=lo, threshold=0.2)
  np.testing.assert_array_equal(codes, c_lo)


# ---- committed fixtures (tests/golden/build/, generator oracle/gen_golden_build.py) ----
def build_golden_names():
  import glob
  import os
  d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "build")
  return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(d, "*.npz")))


def load_build_golden(name):
  import os
  return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "build", name + ".npz"))


@pytest.mark.parametrize("name", build_golden_names())
def test_oracle_reproduces_build_golden(name):
  z = load_build_golden(name)
  soar = None if np.isnan(z["soar_lambda"]) else float(z["soar_lambda"])
  tokens, codes, soar_codes, ties = oracle.encode_database(z["x"], z["centers"], z["codebook"], z["block_dims"],
                                                          residual=bool(z["residual"]), soar_lambda=soar,
                                                          threshold=float(z["threshold"]), threads=2)
  np.testing.assert_array_equal(tokens, z["exp_tokens"])
  np.testing.assert_array_equal(codes, z["exp_codes"])
  if soar is not None:
    np.testing.assert_array_equal(soar_codes, z["exp_soar_codes"])
  assert ties == int(z["exp_ties"])


def _kmeans_numpy(x, centers, iterations):
  """GmmUtils' loop restated with numpy: float32 first-minimum assignment through oracle.assign_primary, then double
  sums in index order inside 4 contiguous slices (1 when n < 8 k), slice sums added in slice order, times 1.0 / count."""
  import oracle
  n, d = x.shape
  k = centers.shape[0]
  c = centers.copy()
  slices = 4 if n >= 8 * k else 1
  per = (n + slices - 1) // slices
  for _ in range(iterations):
    a, _ = oracle.assign_primary(x, c, threads=2)
    for j in range(k):
      total = np.zeros(d, np.float64)
      cnt = 0
      for t in range(slices):
        lo, hi = t * per, min(n, (t + 1) * per)
        m = np.flatnonzero(a[lo:hi] == j) + lo
        part = np.zeros(d, np.float64)
        for i in m:                      # index order, one add at a time
          part += x[i].astype(np.float64)
        total += part
        cnt += len(m)
      if cnt:
        c[j] = (total * (1.0 / cnt)).astype(np.float32)
  a, _ = oracle.assign_primary(x, c, threads=2)
  return c, a


@pytest.mark.parametrize("n,d,k,iters", [(600, 6, 16, 3), (100, 5, 16, 2), (900, 33, 40, 2)])
def test_kmeans_matches_numpy_restatement(n, d, k, iters):
  import oracle
  rng = np.random.default_rng(n + k)
  means = rng.standard_normal((k, d)).astype(np.float32)
  x = (means[rng.integers(0, k, n)] + 0.5 * rng.standard_normal((n, d))).astype(np.float32)
  init = x[np.sort(rng.choice(n, k, replace=False))].copy()
  c, a, empty = oracle.kmeans(x, init, iters, threads=2)
  c_ref, a_ref = _kmeans_numpy(x, init, iters)
  assert np.array_equal(c.view(np.uint32), c_ref.view(np.uint32))
  assert np.array_equal(a, a_ref)
  # Lloyd iterations never raise the distortion
  d0 = ((x - init[oracle.assign_primary(x, init)[0]]) ** 2).sum()
  d1 = ((x - c[a]) ** 2).sum()
  assert d1 <= d0
