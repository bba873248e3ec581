"""Generates tests/golden/build/*.npz: small fixed inputs of the index-build stage (datapoints, trained centres,
AH codebook) and the oracle's outputs (tokens, codes, SOAR codes) in the serialized layout.

Run from the repo root:  python oracle/gen_golden_build.py
The fixtures pin the oracle's build restatement against accidental changes and give the GPU encoder
(scann_b200_encode_database) a committed answer on the GPU box.  Inputs are stored, so nothing depends on a trainer.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

CASES = {
    # name: n, d, leaves, dims_per_block, noise, normalize, soar lambda, noise shaping threshold, residual
    "soar_shaped_b16": dict(n=900, d=32, L=40, dpb=2, noise=1.0, normalize=True, soar=1.5, thr=0.2, residual=True),
    "plain_varchunk_b11": dict(n=700, d=32, L=24, dpb=3, noise=1.2, normalize=False, soar=None, thr=float("nan"), residual=True),
    "raw_l2_shaped_b10": dict(n=600, d=20, L=16, dpb=2, noise=1.0, normalize=False, soar=None, thr=0.3, residual=False),
    "tensor_path_b25": dict(n=1200, d=50, L=320, dpb=2, noise=0.8, normalize=True, soar=2.0, thr=0.2, residual=True),
}


def make_inputs(c, seed):
  rng = np.random.default_rng(seed)
  means = rng.standard_normal((c["L"], c["d"])).astype(np.float32)
  x = (means[rng.integers(0, c["L"], c["n"])] + c["noise"] * rng.standard_normal((c["n"], c["d"]))).astype(np.float32)
  centers = (means + 0.05 * rng.standard_normal((c["L"], c["d"]))).astype(np.float32)
  if c["normalize"]:
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
  full, part = divmod(c["d"], c["dpb"])
  bd = np.asarray([c["dpb"]] * full + ([part] if part else []), np.int32)
  off = np.concatenate([[0], np.cumsum(bd)])
  tok, _ = oracle.assign_primary(x, centers)
  res = x - centers[tok] if c["residual"] else x
  cb = np.zeros((len(bd), 16, c["dpb"]), np.float32)
  for b in range(len(bd)):
    pick = rng.choice(len(res), 16, replace=False)
    cb[b, :, :bd[b]] = res[pick, off[b]:off[b + 1]] * np.float32(0.9)   # scaled: no exact ties with a datapoint
  return np.ascontiguousarray(x), np.ascontiguousarray(centers), cb, bd


def main():
  out_dir = os.path.join(ROOT, "tests", "golden", "build")
  os.makedirs(out_dir, exist_ok=True)
  for i, (name, c) in enumerate(CASES.items()):
    x, centers, cb, bd = make_inputs(c, 100 + i)
    tokens, codes, soar_codes, ties = oracle.encode_database(x, centers, cb, bd, residual=c["residual"],
                                                            soar_lambda=c["soar"], threshold=c["thr"], threads=4)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), x=x, centers=centers, codebook=cb, block_dims=bd,
                        residual=np.int32(c["residual"]), soar_lambda=np.float32(np.nan if c["soar"] is None else c["soar"]),
                        threshold=np.float64(c["thr"]), exp_tokens=tokens, exp_codes=codes,
                        exp_soar_codes=soar_codes if soar_codes is not None else np.zeros((0, 0), np.uint8),
                        exp_ties=np.int64(ties))
    print(name, x.shape, "spilled", None if c["soar"] is None else int((tokens[1::2] >= 0).sum()), "ties", ties)


if __name__ == "__main__":
  main()
