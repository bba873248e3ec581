"""Generates tests/golden/*.npz: small fixed indexes, queries and the oracle's outputs.

Run from the repo root:  python oracle/gen_golden.py
The fixtures pin (a) the oracle against accidental changes and (b) the CUDA path on the
GPU box, where neither /root/reference nor a rebuild of the index is available.
The index arrays are stored too, so the fixtures do not depend on torch's k-means.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from scann_b200 import datasets, index_build  # noqa: E402

CASES = {
    "dot_b16": dict(n=2000, d=32, leaves=16, dpb=2, soar=None, probe=4, pre=40, k=10, seed=21),
    "dot_soar_b25": dict(n=1500, d=50, leaves=12, dpb=2, soar=1.5, probe=5, pre=30, k=10, seed=22),
    "dot_varchunk_b11": dict(n=1200, d=32, leaves=10, dpb=3, soar=None, probe=3, pre=25, k=5, seed=23),
    "l2_b16": dict(n=2000, d=32, leaves=16, dpb=2, soar=None, probe=4, pre=40, k=10, seed=24, distance="squared_l2"),
}


def main():
  out_dir = os.path.join(ROOT, "tests", "golden")
  os.makedirs(out_dir, exist_ok=True)
  for name, c in CASES.items():
    db = datasets.clustered(c["n"], c["d"], 4 * c["leaves"], seed=c["seed"], centers_seed=100 + c["seed"])
    q = datasets.clustered(24, c["d"], 4 * c["leaves"], seed=c["seed"] + 1, centers_seed=100 + c["seed"])
    a = index_build.build_tree_ah(db, c.get("distance", "dot_product"), num_leaves=c["leaves"], dims_per_block=c["dpb"],
                                  training_sample_size=c["n"], soar_lambda=c["soar"], tree_iters=5, ah_iters=5,
                                  device="cpu")
    oi = oracle.OracleIndex(a, c["probe"], c["pre"], c["k"])
    leaf, cdist = oi.tokenize(q)
    lut, mult = oi.lut(q)
    scores0 = oi.leaf_scores(lut[0], int(leaf[0, 0]))
    cand = oi.candidates(q)
    idx, dist = oi.search_batched(q)
    np.savez_compressed(
        os.path.join(out_dir, name + ".npz"),
        dataset=db, queries=q, centers=a.centers, tokens=a.tokens, codes=a.codes,
        soar_codes=a.soar_codes if a.soar_codes is not None else np.zeros((0, 0), np.uint8),
        codebook=a.codebook, block_dims=a.block_dims, soar=np.int32(1 if a.soar else 0), distance=np.str_(a.distance),
        overretrieve=np.float32(a.overretrieve), probe=np.int32(c["probe"]), pre=np.int32(c["pre"]), k=np.int32(c["k"]),
        exp_leaf=leaf, exp_center_dist=cdist, exp_lut=lut, exp_mult=mult, exp_scores_q0_leaf0=scores0,
        exp_cand_count=cand["count"], exp_cand_leaf=cand["leaf"], exp_cand_slot=cand["slot"],
        exp_cand_dp=cand["dp"], exp_cand_score=cand["score"], exp_cand_acc=cand["acc"],
        exp_idx=idx, exp_dist=dist)
    print(name, "recall-ish first row", idx[0][:5], "bytes", os.path.getsize(os.path.join(out_dir, name + ".npz")))


if __name__ == "__main__":
  main()
