"""Throughput of the GPU index-build stage (scann_b200_encode_database).

  python bench_build.py [--n N --d D --leaves L --dpb 2 --soar 1.5 --threshold 0.2 --cpu-sample 4000]

Synthetic clustered data of the bench.py shapes; centres = a k-means tree trained with the torch trainer on a
sample, codebook trained on sample residuals (both are inputs of the measured stage).  Prints one JSON line:
datapoints/s end to end through the host-buffer C-ABI call (host -> device copies of the rows and device -> host
copies of the tokens / codes inside the timed region), the per-stage CUDA-event times, and the oracle's CPU
restatement of the same stage timed on a bounded sample with all host threads (checked equal on that sample).
"""
import argparse
import json
import os
import sys
import time

import numpy as np


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--n", type=int, default=1_183_514)
  ap.add_argument("--d", type=int, default=100)
  ap.add_argument("--leaves", type=int, default=2000)
  ap.add_argument("--dpb", type=int, default=2)
  ap.add_argument("--soar", type=float, default=1.5)
  ap.add_argument("--no-soar", action="store_true")
  ap.add_argument("--threshold", type=float, default=0.2)
  ap.add_argument("--clusters", type=int, default=0)
  ap.add_argument("--train-sample", type=int, default=200000)
  ap.add_argument("--cpu-sample", type=int, default=4000)
  ap.add_argument("--repeat", type=int, default=2)
  ap.add_argument("--train-iters", type=int, default=8)
  args = ap.parse_args()
  import torch
  from scann_b200 import _lib, datasets, index_build
  import oracle

  t0 = time.time()
  clusters = args.clusters or 4 * args.leaves
  db = datasets.clustered(args.n, args.d, clusters, seed=3, centers_seed=103, normalize=True)
  log(f"data {db.shape} in {time.time() - t0:.1f}s")
  t0 = time.time()
  rng = np.random.default_rng(0)
  sel = np.sort(rng.choice(args.n, size=min(args.n, args.train_sample), replace=False))
  sample = db[sel]
  # the k-means tree through the library's trainer (scann_b200_train_kmeans), with its own CUDA-event statistics,
  # beside the same Lloyd iterations in eager torch on the same device and the oracle's CPU restatement on a bounded
  # problem (one iteration, checked equal)
  g = torch.Generator(device="cpu").manual_seed(0)
  init = sample[torch.randperm(len(sample), generator=g)[:args.leaves].numpy()]
  _lib.train_kmeans(sample[:4096], init[:16], 1)   # warm-up: context, module load
  t1 = time.time()
  centers, assign, tst = _lib.train_kmeans(sample, init, args.train_iters)
  train_wall = time.time() - t1
  t1 = time.time()
  index_build.train_kmeans(sample, args.leaves, iters=args.train_iters, seed=0, spherical=True)   # spherical = the torch path
  torch.cuda.synchronize()
  torch_wall = time.time() - t1
  mt = min(len(sample), 20000)
  kt = min(args.leaves, 1000)
  t1 = time.time()
  o_c, _, _ = oracle.kmeans(sample[:mt], init[:kt], 1, threads=os.cpu_count() or 1, want_assignment=False)
  cpu_train_s = time.time() - t1
  g_c, _, _ = _lib.train_kmeans(sample[:mt], init[:kt], 1, want_assignment=False)
  train = {"n": int(len(sample)), "k": args.leaves, "iterations": args.train_iters, "wall_s": train_wall,
           "ms_assign": tst["ms_assign"], "ms_update": tst["ms_update"], "ms_total": tst["ms_total"],
           "empty_clusters": tst["empty_clusters"], "mean_sq_distance": tst["mean_sq_distance"],
           "torch_eager_same_device_wall_s": torch_wall,
           "cpu_port": {"n": mt, "k": kt, "iterations": 1, "wall_s": cpu_train_s, "threads": os.cpu_count() or 1,
                        "equal_to_gpu": bool(np.array_equal(o_c.view(np.uint32), g_c.view(np.uint32)))}}
  log(f"k-means tree: {train}")
  centers, _ = index_build._reseed_empty(sample, centers, assign, 0)
  res = sample - centers[index_build.tokenize_database(sample, centers)]
  t1 = time.time()
  cb, bd = index_build.train_ah_codebook(res, args.dpb, iters=8, seed=1, sample=args.train_sample)
  torch.cuda.synchronize()
  train["ah_codebook_wall_s"] = time.time() - t1
  log(f"trained {args.leaves} centres + codebook {cb.shape} in {time.time() - t0:.1f}s")
  soar = None if args.no_soar else args.soar
  best = None
  for r in range(args.repeat):
    t0 = time.time()
    tokens, codes, soar_codes, st = _lib.encode_database(db, centers, cb, bd, residual=True, soar_lambda=soar,
                                                         noise_shaping_threshold=args.threshold)
    wall = time.time() - t0
    log(f"pass {r}: wall {wall:.2f}s  {st}")
    if best is None or wall < best[0]:
      best = (wall, st)
  wall, st = best
  # CPU restatement on a bounded sample (also the parity check of this run)
  m = min(args.cpu_sample, args.n)
  threads = os.cpu_count() or 1
  t0 = time.time()
  o_tokens, o_codes, o_soar, _ = oracle.encode_database(db[:m], centers, cb, bd, residual=True, soar_lambda=soar,
                                                        threshold=args.threshold, threads=threads)
  cpu_s = time.time() - t0
  npd = 2 if soar is not None else 1
  equal = bool((o_tokens == tokens[:m * npd]).all() and (o_codes == codes[:m]).all() and
               (soar is None or (o_soar == soar_codes[:m]).all()))
  nb = cb.shape[0]
  out = {
      "metric": "index-build stage: datapoints/s (database tokenization + SOAR + noise-shaped AH encoding)",
      "value": args.n / wall, "unit": "datapoints/s", "n_gpus": 1, "wall_s": wall,
      "config": {"workload": "build", "n": args.n, "d": args.d, "leaves": args.leaves, "ah_blocks": int(nb),
                 "soar_lambda": soar, "noise_shaping_threshold": args.threshold},
      "stage_ms": {k: st[k] for k in ("ms_tokenize", "ms_soar", "ms_encode", "ms_total")},
      "soar": {"spilled_frac": st["spilled"] / args.n, "cost_evaluations_per_datapoint": st["soar_evaluated"] / args.n,
               "reference_cost_evaluations_per_datapoint": args.leaves},
      "training": train,
      "tokenize_fallbacks": st["tokenize_fallbacks"], "norm_ties": st["norm_ties"],
      "h2d_bytes": int(db.nbytes), "d2h_bytes": int(tokens.nbytes + codes.nbytes + (soar_codes.nbytes if soar is not None else 0)),
      "cpu_baseline": {"value": m / cpu_s, "unit": "datapoints/s", "cores": threads, "kind": "port",
                       "sample": f"first {m} datapoints, all centres", "equal_to_gpu_on_sample": equal},
  }
  print(json.dumps(out))


if __name__ == "__main__":
  main()
