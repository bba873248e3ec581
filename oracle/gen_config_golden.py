"""Writes tests/golden/builder_configs.json: text protos produced by the REFERENCE's own
scann_builder.py (imported from /root/reference, which only exists in the build container) for a
set of builder call chains.  tests/test_api_cpu.py checks that scann_b200's builder renders the
same ScannConfig for the same calls."""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/scann/scann_ops/py/scann_builder.py"

CHAINS = {
    "tree_ah_dot": [("tree", dict(num_leaves=100, num_leaves_to_search=10, training_sample_size=5000)),
                    ("score_ah", dict(dimensions_per_block=2)), ("reorder", dict(reordering_num_neighbors=100))],
    "tree_ah_soar": [("tree", dict(num_leaves=50, num_leaves_to_search=7, soar_lambda=1.5, overretrieve_factor=1.8)),
                     ("score_ah", dict(dimensions_per_block=2, anisotropic_quantization_threshold=0.2)),
                     ("reorder", dict(reordering_num_neighbors=40))],
    "tree_ah_l2_varchunk": [("tree", dict(num_leaves=30, num_leaves_to_search=5, spherical=True, avq=None)),
                            ("score_ah", dict(dimensions_per_block=3, training_iterations=7)),
                            ("reorder", dict(reordering_num_neighbors=25))],
    "pure_ah": [("score_ah", dict(dimensions_per_block=2, hash_type="lut256"))],
    "brute_force": [("score_brute_force", dict())],
    "brute_force_bf16_tree": [("tree", dict(num_leaves=10, num_leaves_to_search=3, quantize_centroids=True,
                                            incremental_threshold=0.5)),
                              ("score_brute_force", dict(quantize="BFLOAT16"))],
}
DIST = {"tree_ah_l2_varchunk": "squared_l2"}


def run_chain(mod, name, chain):
  db = np.zeros((10, 20), np.float32)
  b = mod.ScannBuilder(db, 10, DIST.get(name, "dot_product"))
  for meth, kw in chain:
    kw = dict(kw)
    if kw.get("quantize") == "BFLOAT16":
      kw["quantize"] = mod.ReorderType.BFLOAT16
    getattr(b, meth)(**kw)
  return b.create_config()


def main():
  spec = importlib.util.spec_from_file_location("ref_scann_builder", REF)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  out = {name: run_chain(mod, name, chain) for name, chain in CHAINS.items()}
  with open(os.path.join(ROOT, "tests", "golden", "builder_configs.json"), "w") as f:
    json.dump({"chains": {k: [[m, {a: (v if not isinstance(v, float) or v == v else "nan") for a, v in kw.items()}]
                              for m, kw in c] for k, c in CHAINS.items()}, "dist": DIST, "reference_text": out}, f, indent=1)
  print("wrote", len(out), "configs")


if __name__ == "__main__":
  main()
