"""Scan-stage behaviour across probe volumes / batch densities on one index (a measurement helper, not a bench line).

  python tools/scan_regimes.py --n 20000000 --cases "10000:24:auto,10000:24:1,2000:24:auto,2000:24:1,2000:40:auto"

Each case is nq:leaves_to_search:two_phase (auto|0|1).  Builds the C5-shape index once (bench.py's generator and
builder), then prints per case: step / scan / compaction times, candidates per query, recall@10, and the scan's
algorithmic GB/s.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--workload", default="c5_deep_shape")
  ap.add_argument("--n", type=int, default=20_000_000)
  ap.add_argument("--cases", default="10000:24:auto,10000:24:1")
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--leaves-total", type=int, default=0, help="override the number of leaves of the index")
  ap.add_argument("--env", default="", help="comma-separated NAME=VALUE pairs set before each case list pass")
  args = ap.parse_args()
  import torch
  import bench
  from scann_b200 import _lib
  wl = dict(bench.WORKLOADS[args.workload])
  if args.n != wl["n"]:
    wl["leaves"] = max(16, int(round(wl["leaves"] * args.n / wl["n"])))
    wl["clusters"] = max(64, int(round(wl["clusters"] * args.n / wl["n"])))
    wl["n"] = args.n
  if args.leaves_total:
    wl["leaves"] = args.leaves_total
  for kv in [e for e in args.env.split(",") if e]:
    name, val = kv.split("=")
    os.environ[name] = val
  dev = torch.device("cuda", 0)
  db, q = bench.make_data(wl)
  arrays = bench.build_arrays(wl, db, dev)
  ix = _lib.NativeIndex(arrays, wl["probe"], wl["pre"], wl["k"])
  k = wl["k"]
  d_q = torch.from_numpy(q).to(dev)
  truth = bench.exact_topk(d_q, db, k, dev, l2=wl.get("distance") == "squared_l2")
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  base_env = {n for n in os.environ if n.startswith("SCANN_B200_")}
  for case in args.cases.split(","):
    case, *envs = case.split(";")  # nq:P:two_phase[:tokenize][;NAME=VALUE...]: per-case environment (tuning knobs)
    for name in [n for n in os.environ if n.startswith("SCANN_B200_") and n not in base_env]:
      os.environ.pop(name)
    for kv in envs:
      name, val = kv.split("=")
      os.environ[name] = val
    parts = case.split(":")
    nq_s, p_s, tp = parts[:3]
    nq, p = int(nq_s), int(p_s)
    if len(parts) > 3:
      os.environ["SCANN_B200_TOKENIZE"] = parts[3]
    else:
      os.environ.pop("SCANN_B200_TOKENIZE", None)
    if tp == "auto":
      os.environ.pop("SCANN_B200_TWO_PHASE", None)
    else:
      os.environ["SCANN_B200_TWO_PHASE"] = tp
    d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
    d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)

    def step():
      ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k, leaves=p)
      return ix.stats()
    for _ in range(3):
      step()
    rec = bench.recall_at_k(d_idx.cpu().numpy().view(np.uint32), truth[:nq])
    agg = {}
    for _ in range(args.steps):
      flush.zero_()
      torch.cuda.synchronize()
      st = step()
      for key, val in st.items():
        agg[key] = agg.get(key, 0) + val
    s = args.steps
    print(json.dumps({
        "nq": nq, "leaves_to_search": p, "two_phase": tp, "env": envs, "tokenize": os.environ.get("SCANN_B200_TOKENIZE", "auto"),
        "leaves": wl["leaves"], "recall_at_10": round(rec, 4),
        "ms_total": agg["ms_total"] / s, "ms_scan": agg["ms_scan"] / s, "ms_compact": agg["ms_compact"] / s,
        "ms_pilot": agg["ms_pilot"] / s, "ms_finalize": agg["ms_finalize"] / s, "ms_tokenize": agg["ms_tokenize"] / s,
        "qps": nq * s / (agg["ms_total"] / 1e3), "cand_per_query": agg["cand_sum"] / s / nq,
        "scan_GBps_alg": agg["scan_bytes_alg"] / (agg["ms_scan"] * 1e-3) / 1e9,
        "queries_per_probed_leaf": nq * p / wl["leaves"], "overflow_retries": agg["overflow_retries"]}), flush=True)


if __name__ == "__main__":
  main()
