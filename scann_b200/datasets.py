"""Synthetic datasets of the shapes BASELINE.json names.

Everything is generated with numpy.random.default_rng(seed) so that the CPU
oracle, the CUDA path and the benchmark all see identical bytes.
"""
import numpy as np


def clustered(n, d, n_clusters, sigma=0.35, seed=0, normalize=False, chunk=1 << 18,
              centers_seed=None, threads=1):
  """Gaussian mixture: n_clusters means ~ N(0, I), point = mean + sigma * N(0, I).

  `centers_seed` pins the mixture means so database and queries (different
  `seed`) are drawn from the same mixture.  threads > 1 draws every chunk from its own stream
  (seed, chunk index) on a thread pool -- a different, equally deterministic dataset, for the large bench shapes.
  """
  crng = np.random.default_rng(seed if centers_seed is None else centers_seed)
  means = crng.standard_normal((n_clusters, d), dtype=np.float32)
  out = np.empty((n, d), dtype=np.float32)
  if threads > 1:
    from concurrent.futures import ThreadPoolExecutor

    def fill(ci):
      s = ci * chunk
      e = min(n, s + chunk)
      rng = np.random.default_rng([seed, 0x5ca77, ci])
      which = rng.integers(0, n_clusters, size=e - s)
      blk = rng.standard_normal((e - s, d), dtype=np.float32)
      blk *= sigma
      blk += means[which]
      if normalize:
        blk /= np.maximum(np.linalg.norm(blk, axis=1, keepdims=True), 1e-12)
      out[s:e] = blk

    with ThreadPoolExecutor(threads) as ex:
      list(ex.map(fill, range((n + chunk - 1) // chunk)))
    return out
  rng = np.random.default_rng([seed, 0x5ca77])
  for s in range(0, n, chunk):
    e = min(n, s + chunk)
    which = rng.integers(0, n_clusters, size=e - s)
    out[s:e] = means[which] + sigma * rng.standard_normal((e - s, d), dtype=np.float32)
  if normalize:
    out /= np.maximum(np.linalg.norm(out, axis=1, keepdims=True), 1e-12)
  return out


def config_c1(nq=10000):
  """C1: 100k x 100 f32 dot product, 100 leaves."""
  db = clustered(100_000, 100, 400, seed=1, centers_seed=101)
  q = clustered(nq, 100, 400, seed=2, centers_seed=101)
  return db, q


def config_c2(nq=10000, n=1_183_514):
  """C2: glove-100-angular shape, L2-normalised rows, dot product."""
  db = clustered(n, 100, 8000, seed=3, centers_seed=103, normalize=True)
  q = clustered(nq, 100, 8000, seed=4, centers_seed=103, normalize=True)
  return db, q


def sift_like(n, d=128, seed=7):
  rng = np.random.default_rng(seed)
  out = np.empty((n, d), dtype=np.float32)
  chunk = 1 << 18
  for s in range(0, n, chunk):
    e = min(n, s + chunk)
    out[s:e] = np.round(np.clip(np.abs(rng.standard_normal((e - s, d), dtype=np.float32)) * 40.0, 0, 218))
  return out
