#!/usr/bin/env python
"""bench.py -- batched QPS of the B200 tree-AH query path on the glove-100-angular shape.

Contract (task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A "step" is one `search_batched` pass of the hot path over one batch of
10,000 synthetic queries.

  value        whole-job QPS with queries and outputs resident in HBM, timed with CUDA events
               recorded on the library's own stream (scann_b200_last_stats.ms_total), summed
               over the K steps, max over ranks.
  e2e          the same metric through scann_b200_search_batched (HOST buffers: pinned staging,
               H2D of the queries and D2H of ids/distances inside the timed region), wall clock.
  roofline     LUT16 scan kernel: algorithmic bytes (sum over probed (query, leaf) of
               ceil(n/32)*16*B) / its CUDA-event duration, against the measured
               HBM copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline the CPU oracle's AVX2 restatement of the reference path on this box's cores.

`--impl reference` times that CPU implementation (oracle/, AVX2 vpshufb LUT16 kernel, all host
threads, search_batched_parallel semantics) on the same workload; the reference binary itself
cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, d, leaves, probe, dims_per_block, reorder, k, nq, generator)
    # noise = AH noise_shaping_threshold (anisotropic quantization); 0.2 is what the reference documents for
    # glove-100-angular (docs/example: score_ah(2, anisotropic_quantization_threshold=0.2)).
    "c2_glove_shape": dict(n=1_183_514, d=100, leaves=2000, probe=100, dpb=2, pre=100, k=10, nq=10000,
                           clusters=8000, normalize=True, seed=3, train_sample=250000, noise=0.2),
    # C3: brute-force MIPS over a bf16 database (BASELINE.json configs[2])
    "c3_bruteforce_bf16": dict(kind="bruteforce", n=1_000_000, d=768, k=100, nq=10000, seed=5),
    # the float sibling (BruteForceSearcher<float>): f32 rows, 3-term bf16 tcgen05 pre-filter + exact fp32 chain
    "c3_bruteforce_f32": dict(kind="bruteforce", dtype="f32", n=1_000_000, d=768, k=100, nq=10000, seed=5),
    # C5 shape family (Deep1B-like: 96-d L2-normalised rows, dot product, SOAR, reorder 200) at a database size
    # one bench run can build in minutes; BASELINE.json's C5 is 100M rows / 40k leaves on 8 GPUs.  ~2,500 rows
    # per leaf as in C5; --n / --leaves rescale it.
    "c5_deep_shape": dict(n=20_000_000, d=96, leaves=8000, probe=80, dpb=2, pre=200, k=10, nq=10000,
                          clusters=80000, normalize=True, seed=9, train_sample=500000, soar=1.5, noise=0.2),
    # C4 shape (BASELINE.json configs[3]: sift 10M x 128 squared L2, k = 10): integer-valued SIFT-like rows (0..218)
    # drawn from a clustered mixture, TreeXHybridSMMD semantics (AH codes of the raw vector, no residual, no SOAR:
    # the reference's builder rejects SOAR for squared L2, scann_builder.py:200-201).  Fits one GPU (5.1 GB of rows).
    "c4_sift_shape": dict(n=10_000_000, d=128, leaves=4000, probe=64, dpb=2, pre=100, k=10, nq=10000,
                          clusters=16000, normalize=False, seed=7, train_sample=500000, distance="squared_l2", gen="sift"),
    "c1_synthetic": dict(n=100_000, d=100, leaves=100, probe=10, dpb=2, pre=100, k=10, nq=10000,
                         clusters=400, normalize=False, seed=1, train_sample=100000),
}


def log(*a):
  print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line.  Libraries loaded by the run (NCCL's version banner, for one)
# write to fd 1 directly, so fd 1 is pointed at stderr for the whole run and the result line goes to a
# saved duplicate of the original stdout.
_RESULT_FD = None


def _claim_stdout():
  global _RESULT_FD
  if _RESULT_FD is None:
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
  line = (json.dumps(obj) + "\n").encode()
  sys.stdout.flush()
  fd = 1 if _RESULT_FD is None else _RESULT_FD
  while line:
    line = line[os.write(fd, line):]


def make_data(wl):
  from scann_b200 import datasets
  db = datasets.clustered(wl["n"], wl["d"], wl["clusters"], seed=wl["seed"], centers_seed=100 + wl["seed"],
                          normalize=wl["normalize"])
  q = datasets.clustered(wl["nq"], wl["d"], wl["clusters"], seed=wl["seed"] + 1, centers_seed=100 + wl["seed"],
                         normalize=wl["normalize"])
  if wl.get("gen") == "sift":  # SIFT-like: non-negative integers up to 218 stored as f32
    for a in (db, q):
      np.abs(a, out=a)
      a *= 40.0
      np.clip(a, 0.0, 218.0, out=a)
      np.round(a, out=a)
  return db, q


def build_arrays(wl, db, device):
  from scann_b200 import index_build
  return index_build.build_tree_ah(db, wl.get("distance", "dot_product"), num_leaves=wl["leaves"], dims_per_block=wl["dpb"],
                                   training_sample_size=wl["train_sample"], tree_iters=12, ah_iters=10,
                                   soar_lambda=wl.get("soar"), seed=0, device=device,
                                   noise_shaping_threshold=wl.get("noise", float("nan")))


class ClockSampler(threading.Thread):
  """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

  FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
            "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    super().__init__(daemon=True)
    self.gpu_index = gpu_index
    self.samples = []
    self.stop_flag = threading.Event()

  def run(self):
    while not self.stop_flag.is_set():
      try:
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                              "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 7:
          self.samples.append(parts)
      except Exception:
        pass
      self.stop_flag.wait(0.2)

  def summary(self):
    if not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
    mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    try:
      return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
      pass
  return 6650.0, "fallback"


def ncu_traffic(name):
  """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
  ncu --set full capture of this workload (profiles/); None if the summary is missing."""
  try:
    t = json.load(open(os.path.join(ROOT, "profiles", name)))
    return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
  except Exception:
    return None


def recall_at_k(found, truth):
  k = truth.shape[1]
  hit = 0
  for i in range(found.shape[0]):
    hit += len(set(found[i, :k].tolist()) & set(truth[i].tolist()))
  return hit / (found.shape[0] * k)


def exact_topk(d_q, db, k, dev, rows=1 << 21, l2=False):
  """Exact f32 brute-force top-k ids on the GPU, database streamed in chunks of `rows`; dot product, or squared L2
  as argmax of <q, x> - |x|^2 / 2."""
  import torch
  nq = d_q.shape[0]
  best_v = torch.full((nq, k), -float("inf"), device=dev)
  best_i = torch.zeros((nq, k), dtype=torch.int64, device=dev)
  for r0 in range(0, db.shape[0], rows):
    d_db = torch.from_numpy(db[r0:r0 + rows]).to(dev)
    half = 0.5 * (d_db.double() ** 2).sum(1).float() if l2 else None
    for s in range(0, nq, 2000):
      sc = d_q[s:s + 2000] @ d_db.T
      if l2:
        sc -= half[None, :]
      v, i = torch.topk(sc, min(k, d_db.shape[0]), dim=1)
      cv = torch.cat([best_v[s:s + 2000], v], dim=1)
      ci = torch.cat([best_i[s:s + 2000], i + r0], dim=1)
      o = torch.topk(cv, k, dim=1).indices
      best_v[s:s + 2000] = torch.gather(cv, 1, o)
      best_i[s:s + 2000] = torch.gather(ci, 1, o)
    del d_db
  torch.cuda.empty_cache()
  return best_i.cpu().numpy()


def cpu_reference_run(oracle_index, q, sample, threads, steps, warmup):
  """Times the oracle's AVX2 path on a bounded sample; returns QPS."""
  qs = np.ascontiguousarray(q[:sample])
  for _ in range(warmup):
    oracle_index.search_batched(qs[:min(sample, 256)], impl=1, threads=threads, batch=256)
  t0 = time.perf_counter()
  for _ in range(steps):
    oracle_index.search_batched(qs, impl=1, threads=threads, batch=256)
  dt = time.perf_counter() - t0
  return sample * steps / dt, dt / steps


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default="c2_glove_shape", choices=list(WORKLOADS))
  ap.add_argument("--leaves", type=int, default=0, help="override leaves_to_search")
  ap.add_argument("--clusters", type=int, default=0, help="override the number of mixture components of the synthetic data")
  ap.add_argument("--n", type=int, default=0, help="override the database size (leaves are rescaled to keep rows per leaf)")
  ap.add_argument("--noise", type=float, default=None,
                  help="AH noise_shaping_threshold used when the index is built (the reference builder's default is 0.2)")
  ap.add_argument("--sweep-leaves", default="",
                  help="comma-separated leaves_to_search values measured after the headline run (recall / QPS trade-off)")
  ap.add_argument("--cpu-sample", type=int, default=2000)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  wl = dict(WORKLOADS[args.workload])
  if args.n > 0 and "leaves" in wl:
    wl["leaves"] = max(16, int(round(wl["leaves"] * args.n / wl["n"])))
    wl["clusters"] = max(64, int(round(wl["clusters"] * args.n / wl["n"])))
    wl["n"] = args.n
  if args.clusters > 0:
    wl["clusters"] = args.clusters
  if args.leaves > 0:
    wl["probe"] = args.leaves
  if args.noise is not None:
    wl["noise"] = args.noise

  import torch
  if wl.get("kind") == "bruteforce":
    return run_bruteforce(args, wl, rank, world, local_rank)
  if args.impl == "reference":
    if rank != 0:
      return 0
    return run_reference(args, wl)

  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  dist = None
  if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

  from scann_b200 import _lib
  t0 = time.time()
  db, q = make_data(wl)
  log(f"[rank {rank}] data {db.shape} in {time.time() - t0:.1f}s")
  t0 = time.time()
  shm = f"/dev/shm/scann_b200_bench_{os.environ.get('MASTER_PORT', '0')}_{args.workload}.npz"
  if world > 1:
    from scann_b200 import index_build
    if rank == 0:
      arrays = build_arrays(wl, db, dev)
      extra = {"soar_codes": arrays.soar_codes} if arrays.soar else {}
      np.savez(shm, centers=arrays.centers, tokens=arrays.tokens, codes=arrays.codes, codebook=arrays.codebook,
               block_dims=arrays.block_dims, **extra)
    dist.barrier()
    if rank != 0:
      z = np.load(shm)
      arrays = index_build.IndexArrays(distance=wl.get("distance", "dot_product"), dataset=db, n=db.shape[0], d=db.shape[1])
      arrays.centers, arrays.tokens, arrays.codes = z["centers"], z["tokens"], z["codes"]
      arrays.codebook, arrays.block_dims = z["codebook"], z["block_dims"]
      arrays.residual = arrays.distance == "dot_product"
      if "soar_codes" in z.files:
        arrays.soar_codes, arrays.soar, arrays.overretrieve = z["soar_codes"], True, 2.0
    dist.barrier()
    if rank == 0:
      os.unlink(shm)
  else:
    arrays = build_arrays(wl, db, dev)
  log(f"[rank {rank}] index built in {time.time() - t0:.1f}s")
  t0 = time.time()
  # Every rank holds a full replica (the C2 index is 0.5 GB) and serves its own query batches:
  # queries are independent units, so the headline N-GPU number needs no data-path collective.
  # The database-sharded mode of SURVEY.md 8e (one NCCL all-gather per batch; what C4/C5-size
  # databases need) is measured in the same run and reported under "db_sharded".
  ix = _lib.NativeIndex(arrays, wl["probe"], wl["pre"], wl["k"], device=local_rank)
  searcher = None
  if world > 1:
    from scann_b200 import distributed as sdist
    searcher = sdist.ShardedSearcher(arrays, wl["probe"], wl["pre"], wl["k"], rank, world, local_rank)
  log(f"[rank {rank}] device index in {time.time() - t0:.1f}s")

  nq, k = wl["nq"], wl["k"]
  d_q = torch.from_numpy(q).to(dev)
  d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
  d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
  torch.cuda.synchronize()

  # ground truth for recall (exact f32 brute force on the GPU)
  truth = exact_topk(d_q, db, k, dev, l2=wl.get("distance") == "squared_l2")

  def step_dev():
    ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k)
    return ix.stats()

  # end-to-end leg: the caller's query and result arrays are page-locked host memory (numpy views of
  # pinned torch tensors), as a serving front end would hold them
  q_pin_t = torch.from_numpy(q).pin_memory()
  q_pin = q_pin_t.numpy()
  oi_t = torch.empty((nq, k), dtype=torch.int32).pin_memory()
  od_t = torch.empty((nq, k), dtype=torch.float32).pin_memory()
  out_pin = (oi_t.numpy().view(np.uint32), od_t.numpy())

  def step_host():
    return ix.search_batched(q_pin, out=out_pin)

  for _ in range(args.warmup):
    step_dev()
  found = d_idx.cpu().numpy().view(np.uint32)
  rec = recall_at_k(found, truth)

  sampler = ClockSampler(local_rank)
  sampler.start()
  if dist is not None:
    dist.barrier()
  torch.cuda.synchronize()
  ms_total = 0.0
  agg = {}
  wall0 = time.perf_counter()
  for _ in range(args.steps):
    flush.zero_()  # L2 flush between timed iterations (not inside the CUDA-event interval)
    torch.cuda.synchronize()
    st = step_dev()
    ms_total += st["ms_total"]
    for key, val in st.items():
      agg[key] = agg.get(key, 0) + val
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  wall = time.perf_counter() - wall0

  # end to end through the host-buffer C ABI call
  step_host()
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  e0 = time.perf_counter()
  e2e_steps = max(3, min(args.steps, 10))
  for _ in range(e2e_steps):
    idx_h, dist_h = step_host()
  if dist is not None:
    dist.barrier()
  e2e_s = (time.perf_counter() - e0)
  sweep = []
  if args.sweep_leaves and world == 1:
    for p_ in [int(v) for v in args.sweep_leaves.split(",") if v]:
      def step_p():
        ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k, leaves=p_)
        return ix.stats()
      for _ in range(3):
        step_p()
      rec_p = recall_at_k(d_idx.cpu().numpy().view(np.uint32), truth)
      ms_p, scan_p, n_p = 0.0, 0.0, max(3, min(args.steps, 5))
      for _ in range(n_p):
        flush.zero_()
        torch.cuda.synchronize()
        st = step_p()
        ms_p += st["ms_total"]
        scan_p += st["ms_scan"]
      sweep.append({"leaves_to_search": p_, "recall_at_10": rec_p, "qps": nq * n_p / (ms_p / 1e3),
                    "ms_per_step": ms_p / n_p, "ms_scan": scan_p / n_p})
  db_sharded = None
  if searcher is not None:
    sh_dev, _ = make_sharded_steps(searcher, wl, q, d_q, d_idx, d_dist)
    for _ in range(3):
      sh_dev()
    sh_equal = bool(np.array_equal(d_idx.cpu().numpy().view(np.uint32), found))
    dist.barrier()
    sh_ms, sh_agg = 0.0, {}
    for _ in range(args.steps):
      flush.zero_()
      torch.cuda.synchronize()
      st = sh_dev()
      sh_ms += st["ms_total"]
      for key, val in st.items():
        sh_agg[key] = sh_agg.get(key, 0) + val
    dist.barrier()
    db_sharded = {"ms": sh_ms, "agg": sh_agg, "ids_equal_replica": sh_equal}
  sampler.stop_flag.set()
  sampler.join(timeout=2)

  if dist is not None:
    t = torch.tensor([ms_total, e2e_s, db_sharded["ms"] if db_sharded else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if db_sharded:
      db_sharded["ms"] = float(t[2])
  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return 0

  value = world * nq * args.steps / (ms_total / 1e3)
  e2e_value = world * nq * e2e_steps / e2e_s
  peak, peak_src = measured_peak()
  scan_launches = max(1, agg.get("scan_kernel_count", 1))
  bytes_per_launch = agg["scan_bytes_alg"] / scan_launches
  ms_per_launch = agg["ms_scan"] / scan_launches
  achieved = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9 if ms_per_launch > 0 else 0.0
  out = {
      "metric": "batched QPS at recall@10>=0.90 (tree-AH search_batched)", "value": value, "unit": "queries/s",
      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u8 LUT / int16 accumulate (scan), f32 (tokenize, reorder)", "data": "synthetic",
      "config": {"workload": args.workload, "distance": wl.get("distance", "dot_product"), "n": wl["n"], "d": wl["d"],
                 "leaves": wl["leaves"], "soar_lambda": wl.get("soar"),
                 "leaves_to_search": wl["probe"], "ah_blocks": wl["d"] // wl["dpb"], "reorder": wl["pre"],
                 "k": k, "queries_per_step": nq * world, "recall_at_10": rec,
                 "noise_shaping_threshold": wl.get("noise"),
                 "l2_flush": "256 MiB write between timed steps",
                 "parallelism": f"query-parallel x{world} (one replica and one {nq}-query batch per GPU)",
                 "wall_s_timed_region": wall},
      "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes) * world,
              "d2h_bytes_per_step": int(nq * k * 8) * world, "steps": e2e_steps},
      "gpu_launches": int(agg.get("kernel_launches", 0)),
      "clocks": sampler.summary(),
      **({"probe_sweep": sweep} if sweep else {}),
      "roofline": {"bound": "hbm", "kernel": "scan_main_kernel<W>", "achieved": achieved, "peak": peak,
                   "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak if peak else None,
                   "traffic": ncu_traffic("r01_scan_main_traffic.json") if args.workload == "c2_glove_shape" else None,
                   "alg_bytes_per_launch": bytes_per_launch, "ms_per_launch": ms_per_launch,
                   "lookups_per_s": 2 * bytes_per_launch / (ms_per_launch * 1e-3) if ms_per_launch else None},
      "stage_ms_per_step": {s: agg[s] / args.steps for s in agg if s.startswith("ms_")},
      "overflow_retries": int(agg.get("overflow_retries", 0)),
      "tokenize_fallbacks_per_step": agg.get("tokenize_fallbacks", 0) / args.steps,
      "candidates_per_query": {"mean": agg.get("cand_sum", 0) / (nq * args.steps), "max_over_steps_sum": int(agg.get("cand_max", 0))},
  }
  if db_sharded is not None:
    sa = db_sharded["agg"]
    out["db_sharded"] = {
        "value": nq * args.steps / (db_sharded["ms"] / 1e3), "unit": "queries/s", "scaling": "strong",
        "ms_per_step": db_sharded["ms"] / args.steps, "ids_equal_replica": db_sharded["ids_equal_replica"],
        "collective": "one NCCL all-gather of (id u32, tie-break key u64, exact distance f32) per batch",
        "allgather_bytes_per_rank_per_step": int(sa.get("allgather_bytes_per_rank", 0) / args.steps),
        "stage_ms_per_step": {s_: sa[s_] / args.steps for s_ in sa if s_.startswith("ms_")}}
  if not args.no_cpu_baseline and world == 1:
    import oracle
    threads = os.cpu_count() or 1
    oi = oracle.OracleIndex(arrays, wl["probe"], wl["pre"], wl["k"])
    i_cpu, _ = oi.search_batched(q[:64], impl=1)
    parity = bool(np.array_equal(i_cpu, idx_h[:64]))
    qps, _ = cpu_reference_run(oi, q, min(args.cpu_sample, nq), threads, 1, 1)
    qps1, _ = cpu_reference_run(oi, q, min(500, nq), 1, 1, 0)
    out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                           "sample": f"first {min(args.cpu_sample, nq)} queries of the step, batches of 256 "
                                     f"(search_batched_parallel semantics), AVX2 vpshufb oracle",
                           "single_thread_qps": qps1, "ids_equal_gpu_first_64": parity}
  emit(out)
  if dist is not None:
    dist.destroy_process_group()
  return 0


def make_sharded_steps(searcher, wl, q, d_q, d_idx, d_dist):
  """Sharded search: local candidates -> one NCCL all-gather -> merge."""
  import torch

  def step_dev():
    st = searcher.search_batched_device(d_q, d_idx, d_dist)
    st["ms_total"] = st["ms_total"] + st["ms_allgather"] + st.get("ms_merge", 0.0)
    return st

  hq = torch.from_numpy(q).pin_memory()

  def step_host():
    dq = hq.to(d_q.device, non_blocking=True)
    searcher.search_batched_device(dq, d_idx, d_dist)
    return d_idx.cpu().numpy().view(np.uint32), d_dist.cpu().numpy()

  return step_dev, step_host


def run_bruteforce(args, wl, rank, world, local_rank):
  """C3: bf16 brute force, 10k queries x 1M x 768, k = 100 (tcgen05 GEMM + fused top-k pre-filter)."""
  import torch
  from scann_b200 import _lib, index_build
  if world > 1 and rank != 0 and args.impl == "reference":
    return 0
  n, d, nq, k = wl["n"], wl["d"], wl["nq"], wl["k"]
  rng = np.random.default_rng(wl["seed"])
  t0 = time.time()
  f32 = wl.get("dtype") == "f32"
  bits = np.empty((n, d), np.float32 if f32 else np.int16)
  for s0 in range(0, n, 1 << 16):
    blk = rng.standard_normal((min(1 << 16, n - s0), d), dtype=np.float32)
    bits[s0:s0 + (1 << 16)] = blk if f32 else index_build.bfloat16_quantize(blk)
  q = np.random.default_rng(wl["seed"] + 1).standard_normal((nq, d), dtype=np.float32)
  log(f"[bf] data in {time.time() - t0:.1f}s")
  a = index_build.IndexArrays(distance="dot_product", dataset=bits if f32 else None, n=n, d=d)
  if not f32:
    a.bf16_dataset = bits
  if args.impl == "reference":
    import oracle
    threads = os.cpu_count() or 1
    sample = max(threads, 16)
    t0 = time.perf_counter()
    (oracle.bruteforce_f32 if f32 else oracle.bruteforce_bf16)(bits, q[:sample], k, threads=threads)
    dt = time.perf_counter() - t0
    qps = sample / dt
    emit({"impl": "reference", "metric": "batched QPS, bf16 brute-force MIPS k=100", "value": qps,
                      "unit": "queries/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 db x f32 query",
                      "data": "synthetic", "config": {"workload": args.workload, "n": n, "d": d, "k": k,
                                                      "queries_per_step": sample},
                      "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                                       "sample": f"{sample} queries, one per thread"},
                      "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return 0
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  dist = None
  if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
  ix = _lib.NativeIndex(a, 1, k, k, device=local_rank)
  d_q = torch.from_numpy(q).to(dev)
  d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
  d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  torch.cuda.synchronize()
  for _ in range(args.warmup):
    ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k)
  # recall against an f32 GEMM on the decompressed rows (first 512 queries)
  x = torch.from_numpy(bits).to(dev) if f32 else (torch.from_numpy(bits.view(np.uint16).astype(np.int32)).to(dev) << 16).view(torch.float32)
  gt = torch.topk(d_q[:512] @ x.T, k, dim=1).indices.cpu().numpy()
  del x
  torch.cuda.empty_cache()
  found = d_idx[:512].cpu().numpy().view(np.uint32)
  rec = recall_at_k(found, gt)
  sampler = ClockSampler(local_rank)
  sampler.start()
  ms_total, agg = 0.0, {}
  if dist is not None:
    dist.barrier()
  torch.cuda.synchronize()
  for _ in range(args.steps):
    flush.zero_()
    torch.cuda.synchronize()
    ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k)
    st = ix.stats()
    ms_total += st["ms_total"]
    for key, val in st.items():
      agg[key] = agg.get(key, 0) + val
  e2e_steps = max(2, min(args.steps, 5))
  ix.search_batched(q)
  if dist is not None:
    dist.barrier()
  e0 = time.perf_counter()
  for _ in range(e2e_steps):
    ix.search_batched(q)
  e2e_s = time.perf_counter() - e0
  sampler.stop_flag.set()
  sampler.join(timeout=2)
  row_sharded = None
  if dist is not None:
    # max over ranks of the device time and of the end-to-end wall time (one replica per GPU, weak scaling)
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    # the database row-sharded over the ranks: same 10k queries everywhere, local top-k, all-gather, merge
    from scann_b200 import distributed as sd
    ref_idx = d_idx.clone()
    del ix
    torch.cuda.empty_cache()
    sh = sd.ShardedBruteForce(a, k, rank, world, local_rank)
    for _ in range(args.warmup):
      sh.search_batched_device(d_q, d_idx, d_dist)
    same = bool((d_idx == ref_idx).all().item())
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sh_ms, sh_agg = 0.0, {}
    for _ in range(args.steps):
      flush.zero_()
      dist.barrier()
      torch.cuda.synchronize()
      w0 = time.perf_counter()
      st = sh.search_batched_device(d_q, d_idx, d_dist)
      torch.cuda.synchronize()
      sh_ms += (time.perf_counter() - w0) * 1e3
      for key, val in st.items():
        sh_agg[key] = sh_agg.get(key, 0) + val
    t = torch.tensor([sh_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    row_sharded = {"value": nq * args.steps / (float(t[0]) * 1e-3), "unit": "queries/s",
                   "ids_equal_replica": same, "timing": "wall clock around search + all-gather + merge, max over ranks",
                   "allgather_bytes_per_rank": sh_agg.get("allgather_bytes_per_rank", 0) // max(args.steps, 1),
                   "stage_ms_per_step": {s_: sh_agg[s_] / args.steps for s_ in sh_agg if s_.startswith("ms_")}}
    if rank != 0:
      dist.destroy_process_group()
      return 0
  flops = 2.0 * nq * n * d * (3 if f32 else 2)  # bf16 split terms per product: hi.hi + lo.hi (+ hi.lo for f32 rows)
  gemm_s = agg["ms_scan"] / args.steps * 1e-3
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  # the GEMM rounds run as ~20 ms bursts between L2 flushes, i.e. "a kernel timed alone": burst peak
  peak = float(peaks.get("bf16_tflops", 1590.0))
  peak_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
  out = {"metric": f"batched QPS, {'f32' if f32 else 'bf16'} brute-force MIPS k=100", "value": world * nq * args.steps / (ms_total * 1e-3),
         "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
         "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
         "dtype": ("(bf16 hi + lo) x (bf16 hi + lo) -> f32 (tcgen05, 3 terms), exact f32 chain re-scoring" if f32 else
                   "bf16 x (bf16 hi + bf16 lo) -> f32 (tcgen05), f32 re-scoring"), "data": "synthetic",
         "config": {"workload": args.workload, "n": n, "d": d, "k": k, "queries_per_step": nq * world,
                    "recall_at_100_first512": rec, "l2_flush": "256 MiB write between timed steps",
                    "parallelism": f"query-parallel x{world} (one replica and one {nq}-query batch per GPU)"},
         "e2e": {"value": world * nq * e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes) * world,
                 "d2h_bytes_per_step": int(nq * k * 8) * world, "steps": e2e_steps},
         "gpu_launches": int(agg["kernel_launches"]), "clocks": sampler.summary(),
         "roofline": {"bound": "tensor", "kernel": "bf::gemm_pair_kernel<1, filter>" if f32 else "bf::gemm_pair_kernel<2, filter>", "achieved": flops / gemm_s / 1e12,
                      "peak": peak, "peak_source": "measured burst" if peaks else "fallback", "unit": "TFLOP/s",
                      "frac": flops / gemm_s / 1e12 / peak, "traffic": ncu_traffic("r01_gemm_pair_traffic.json"),
                      "traffic_note": "dram bytes of the largest round's launch (497,664 rows: 764 MB compulsory)",
                      "frac_of_sustained_peak": flops / gemm_s / 1e12 / peak_sustained,
                      "note": "flops count both bf16 query terms (hi + lo); time includes the compactions between rounds",
                      "useful_tflops_f32_equivalent": flops / 2 / gemm_s / 1e12},
         "stage_ms_per_step": {s: agg[s] / args.steps for s in agg if s.startswith("ms_")}}
  if row_sharded is not None:
    out["row_sharded"] = row_sharded
    dist.destroy_process_group()
  emit(out)
  return 0


def run_reference(args, wl):
  """CPU arm: the oracle's AVX2 restatement of the reference path, all host threads."""
  import oracle
  t0 = time.time()
  db, q = make_data(wl)
  import torch
  arrays = build_arrays(wl, db, "cuda:0" if torch.cuda.is_available() else "cpu")
  log(f"[reference] data+index in {time.time() - t0:.1f}s")
  threads = os.cpu_count() or 1
  oi = oracle.OracleIndex(arrays, wl["probe"], wl["pre"], wl["k"])
  sample = min(args.cpu_sample, wl["nq"])
  steps, warmup = max(1, args.steps), max(0, args.warmup)
  # keep the whole run within a few minutes: probe the speed first
  qps_probe, _ = cpu_reference_run(oi, q, min(256, sample), threads, 1, 0)
  budget_s = 120.0
  max_steps = max(1, int(budget_s * qps_probe / sample))
  steps_run = min(steps, max_steps)
  qps, s_per_step = cpu_reference_run(oi, q, sample, threads, steps_run, min(warmup, 1))
  out = {
      "impl": "reference", "metric": "batched QPS at recall@10>=0.90 (tree-AH search_batched)", "value": qps,
      "unit": "queries/s", "n_gpus": args.gpus, "steps": steps_run, "warmup": min(warmup, 1),
      "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u8 LUT / int16 accumulate (AVX2)", "data": "synthetic",
      "config": {"workload": args.workload, "distance": wl.get("distance", "dot_product"), "n": wl["n"], "d": wl["d"],
                 "leaves": wl["leaves"], "soar_lambda": wl.get("soar"), "leaves_to_search": wl["probe"],
                 "ah_blocks": wl["d"] // wl["dpb"], "reorder": wl["pre"], "k": wl["k"],
                 "noise_shaping_threshold": wl.get("noise"),
                 "queries_per_step": sample,
                 "sample_note": f"bounded sample: the first {sample} of the workload's {wl['nq']} queries per step"},
      "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                       "sample": f"{sample} queries per step, batches of 256 over {threads} threads"},
      "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  emit(out)
  return 0


if __name__ == "__main__":
  _claim_stdout()
  sys.exit(main())
