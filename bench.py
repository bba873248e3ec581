#!/usr/bin/env python
"""bench.py -- batched QPS of the B200 tree-AH query path.

Contract (task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A "step" is one `search_batched` pass of the hot path over one batch of 10,000 synthetic queries.

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1], the glove-100-angular shape (C2), one
replica and one query batch per GPU (queries are independent units: weak scaling, no data-path collective).

  value        whole-job QPS with queries and outputs resident in HBM, timed with CUDA events recorded on the
               library's own stream (scann_b200_last_stats.ms_total), summed over the K steps, max over ranks.
  e2e          the same metric through scann_b200_search_batched (HOST buffers: H2D of the queries and D2H of
               ids/distances inside the timed region), wall clock.
  roofline     LUT16 scan kernel: algorithmic bytes (sum over probed (query, leaf) of ceil(n/32)*16*B)
               / its CUDA-event duration against the measured HBM copy bandwidth of MEASURED_PEAKS.json, plus the
               on-chip ceiling of the kernel's instruction mix (`onchip`), which is what actually binds it.
  cpu_baseline the CPU oracle's AVX2 restatement of the reference path on this box's cores.

The same run also measures, and reports inside `config` (so that they survive the driver's parsing):

  config.c5_sharded    BASELINE.json configs[4] shape (96-d, SOAR, reorder 200, k = 10; 20M rows by default, --c5-n
                       rescales up to the full 100M) with the DATABASE SHARDED BY LEAF over the N GPUs and the NCCL
                       exchange issued from the C++ library (csrc/sharded.cu): strong scaling, the same 10,000 queries
                       on every rank, ids compared with the single-GPU searcher.  At N = 1 it is the single-GPU figure
                       the curve starts from.
  config.c3_bruteforce BASELINE.json configs[2]: bf16 brute force 1M x 768, k = 100 (replicas; row-sharded at N > 1).
  config.c2_db_sharded the C2 index sharded the same way (N > 1; a 0.5 GB index gains nothing from it -- reported
                       because round 1 did).

`--impl reference` times the CPU implementation (oracle/, AVX2 vpshufb LUT16 kernel, all host threads) on the same
workload, the same 10,000 queries per step, the same warm-up, with an index built on the CPU (no product library is
loaded in that arm); the reference binary itself cannot be built in this image.
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
    "c5_deep_shape": dict(n=20_000_000, d=96, leaves=8000, probe=24, dpb=2, pre=200, k=10, nq=10000,
                          clusters=80000, normalize=True, seed=9, train_sample=500000, soar=1.5, noise=0.2, gen_threads=16),
    # C4 shape (BASELINE.json configs[3]: sift 10M x 128 squared L2, k = 10): integer-valued SIFT-like rows (0..218)
    # drawn from a clustered mixture, TreeXHybridSMMD semantics (AH codes of the raw vector, no residual, no SOAR:
    # the reference's builder rejects SOAR for squared L2, scann_builder.py:200-201).  Fits one GPU (5.1 GB of rows).
    "c4_sift_shape": dict(n=10_000_000, d=128, leaves=4000, probe=64, dpb=2, pre=100, k=10, nq=10000,
                          clusters=16000, normalize=False, seed=7, train_sample=500000, distance="squared_l2", gen="sift",
                          gen_threads=16),
    "c1_synthetic": dict(n=100_000, d=100, leaves=100, probe=10, dpb=2, pre=100, k=10, nq=10000,
                         clusters=400, normalize=False, seed=1, train_sample=100000),
}

# LUT16 scan, instruction mix per warp-wide oct lookup (8 queries x 32 slots x 1 block = 256 lookups):
# 1 LDS.64 + 2 LOP3 + 2 PRMT + 4 IMAD = 9 issue slots (scan.cu score_oct_addr).  Per SM and clock: 4 issue slots
# (9 slots per 256 lookups -> 113.8 lookups), ALU pipe 64 lanes (4 warp instructions = 8 SMSP-clocks per 256 lookups per
# SMSP -> 128), shared memory 128 B (one LDS.64 = 256 B -> 128).
ONCHIP = {"issue_slots": 4.0 / 9.0 * 256.0, "alu_pipe": 128.0, "lds_bandwidth": 128.0}
# Wide quads (sparse work lists, scan.cu score_wquad_addr): 1 LDS.64 + 2 adds = 3 issue slots per 4 queries x 32 slots
# = 128 lookups -> 170.7; one LDS.64 = 256 B of shared-memory bandwidth per 128 lookups -> 64.  Both at FULL tables:
# a leaf probed by q queries costs ceil(q / 4) quads, and `achieved` counts the useful lookups only.
ONCHIP_WIDE = {"issue_slots": 4.0 / 3.0 * 128.0, "lds_bandwidth": 64.0}


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


def make_data(wl, queries_only=False):
  from scann_b200 import datasets
  thr = int(wl.get("gen_threads", 1))
  if thr > 1:
    thr = max(2, min(thr, os.cpu_count() or 2))
  q = datasets.clustered(wl["nq"], wl["d"], wl["clusters"], seed=wl["seed"] + 1, centers_seed=100 + wl["seed"],
                         normalize=wl["normalize"])
  db = None
  if not queries_only:
    db = datasets.clustered(wl["n"], wl["d"], wl["clusters"], seed=wl["seed"], centers_seed=100 + wl["seed"],
                            normalize=wl["normalize"], threads=thr)
  if wl.get("gen") == "sift":  # SIFT-like: non-negative integers up to 218 stored as f32
    for a in (db, q):
      if a is None:
        continue
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


def build_arrays_cpu(wl, db, threads):
  """The reference arm's index: trainers on the CPU (torch), the per-datapoint stage through the CPU oracle's
  restatement of the reference builder -- libscann_b200.so is never loaded in that arm."""
  import oracle
  from scann_b200 import index_build
  distance = wl.get("distance", "dot_product")
  residual = distance == "dot_product"
  rng = np.random.default_rng(0)
  n = db.shape[0]
  sample = db[np.sort(rng.choice(n, size=wl["train_sample"], replace=False))] if n > wl["train_sample"] else db
  centers = index_build.train_kmeans(sample, min(wl["leaves"], n), iters=12, seed=0, device="cpu").astype(np.float32)
  res_s = sample - centers[index_build.tokenize_database(sample, centers, device="cpu")] if residual else sample
  cb, block_dims = index_build.train_ah_codebook(res_s, wl["dpb"], iters=10, seed=1, sample=wl["train_sample"], device="cpu")
  arr = index_build.IndexArrays(distance=distance, dataset=db, n=n, d=db.shape[1])
  arr.centers, arr.codebook, arr.block_dims, arr.residual = centers, cb, block_dims, residual
  arr.tokens, arr.codes, arr.soar_codes, _ = oracle.encode_database(
      db, centers, cb, block_dims, residual=residual, soar_lambda=wl.get("soar"),
      threshold=wl.get("noise", float("nan")), threads=threads)
  if wl.get("soar") is not None:
    arr.soar, arr.overretrieve = True, 2.0
  return arr


_ARRAY_FIELDS = ("centers", "tokens", "codes", "soar_codes", "codebook", "block_dims", "dataset")


def share_arrays(tag, arrays, rank, world, dist, distance):
  """Rank 0's index arrays -> every rank, through /dev/shm .npy files mapped read-only (one physical copy)."""
  from scann_b200 import index_build
  if world == 1:
    return arrays
  base = f"/dev/shm/scann_b200_bench_{os.environ.get('MASTER_PORT', '0')}_{tag}_"
  meta = [None]
  if rank == 0:
    for f in _ARRAY_FIELDS:
      a = getattr(arrays, f)
      if a is not None:
        np.save(base + f + ".npy", a)
    meta = [dict(n=arrays.n, d=arrays.d, soar=arrays.soar, overretrieve=arrays.overretrieve, residual=arrays.residual,
                 fields=[f for f in _ARRAY_FIELDS if getattr(arrays, f) is not None])]
  dist.broadcast_object_list(meta, src=0)
  m = meta[0]
  if rank != 0:
    arrays = index_build.IndexArrays(distance=distance, dataset=None, n=m["n"], d=m["d"])
    for f in m["fields"]:
      setattr(arrays, f, np.load(base + f + ".npy", mmap_mode="r"))
    arrays.soar, arrays.overretrieve, arrays.residual = m["soar"], m["overretrieve"], m["residual"]
  dist.barrier()  # every rank holds its mappings: the names can go (the pages live as long as the mappings)
  if rank == 0:
    for f in m["fields"]:
      os.unlink(base + f + ".npy")
  return arrays


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


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    try:
      return json.load(open(p)), "measured"
    except Exception:
      pass
  return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


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
    d_db = torch.from_numpy(np.ascontiguousarray(db[r0:r0 + rows])).to(dev)
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
  """Times the oracle's AVX2 path; returns (QPS, seconds per step)."""
  qs = np.ascontiguousarray(q[:sample])
  for _ in range(warmup):
    oracle_index.search_batched(qs, impl=1, threads=threads, batch=256)
  t0 = time.perf_counter()
  for _ in range(steps):
    oracle_index.search_batched(qs, impl=1, threads=threads, batch=256)
  dt = time.perf_counter() - t0
  return sample * steps / dt, dt / steps


def static_config(args, wl):
  """The part of `config` both arms print identically."""
  return {"workload": args.workload, "distance": wl.get("distance", "dot_product"), "n": wl["n"], "d": wl["d"],
          "leaves": wl["leaves"], "soar_lambda": wl.get("soar"), "leaves_to_search": wl["probe"],
          "ah_blocks": -(-wl["d"] // wl["dpb"]), "reorder": wl["pre"], "k": wl["k"], "queries_per_step": wl["nq"],
          "noise_shaping_threshold": wl.get("noise"),
          "l2_flush": "GPU arm: 256 MiB write between timed steps; CPU arm: none (the index exceeds its caches)"}


def scan_roofline(agg, peaks, peak_src, clocks, traffic_file=None):
  scan_launches = max(1, agg.get("scan_kernel_count", 1))
  bytes_per_launch = agg["scan_bytes_alg"] / scan_launches
  ms_per_launch = agg["ms_scan"] / scan_launches
  achieved = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9 if ms_per_launch > 0 else 0.0
  peak = float(peaks["hbm_gbs"])
  lookups_per_s = 2 * achieved * 1e9
  mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
  per_clk_sm = lookups_per_s / (148.0 * mhz * 1e6) if mhz else None
  wide = agg.get("scan_wide_launches", 0) > agg.get("scan_oct_launches", 0)
  tc = agg.get("scan_tc_launches", 0) > 0
  ceilings = ONCHIP_WIDE if wide else ONCHIP
  ceil = min(ceilings.values())
  kernel = "tc::scan_tc_kernel<W>" if tc else ("scan_main_kernel<W, NL, wide quads>" if wide else "scan_main_kernel<W, NL, octs>")
  note = ("the codes are read from DRAM about once per batch (traffic << algorithmic bytes), so the kernel is bound on "
          "chip: " + ("3 issue slots and one 256-byte shared-memory access per warp-wide wide-quad lookup (4 queries)" if wide
                      else "9 issue slots per warp-wide oct lookup (8 queries)"))
  return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "peak_source": peak_src,
          "unit": "GB/s", "frac": achieved / peak if peak else None,
          "traffic": ncu_traffic(traffic_file) if traffic_file else None,
          "alg_bytes_per_launch": bytes_per_launch, "ms_per_launch": ms_per_launch, "lookups_per_s": lookups_per_s,
          "scan_launches": {"octs": int(agg.get("scan_oct_launches", 0)), "wide_quads": int(agg.get("scan_wide_launches", 0)),
                            "tensor_cores": int(agg.get("scan_tc_launches", 0))},
          "onchip": {"unit": "(query, slot, block) lookups per clock per SM", "ceilings": ceilings,
                     "binding": min(ceilings, key=ceilings.get), "achieved": per_clk_sm, "sm_mhz_used": mhz,
                     "frac": per_clk_sm / ceil if per_clk_sm else None, "note": note}}


def e2e_host_run(ix, q_pin, nq, k, steps, callers):
  """End to end through the public host-buffer call (scann_b200_search_batched: H2D of the queries, search, D2H of ids
  and distances, all inside the call) from page-locked caller arrays.  `callers` threads each issue `steps` batches
  back to back on the same searcher, as a serving front end with that many batches in flight does: the library runs
  concurrent calls on separate lanes, so one batch's copies overlap another's kernels.  Returns (QPS, wall seconds)."""
  import torch
  outs = []
  for _ in range(callers):
    oi_t = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    od_t = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    outs.append((oi_t, od_t, (oi_t.numpy().view(np.uint32), od_t.numpy())))
  for c in range(callers):  # warm-up: creates the lanes and their workspaces
    ix.search_batched(q_pin, out=outs[c][2])
  if callers == 1:
    t0 = time.perf_counter()
    for _ in range(steps):
      ix.search_batched(q_pin, out=outs[0][2])
    dt = time.perf_counter() - t0
    return nq * steps / dt, dt, outs[0][2]
  start = threading.Barrier(callers + 1)
  done = []

  def worker(c):
    ix.search_batched(q_pin, out=outs[c][2])   # second warm-up, now concurrently
    start.wait()
    for _ in range(steps):
      ix.search_batched(q_pin, out=outs[c][2])
    done.append(time.perf_counter())
  ths = [threading.Thread(target=worker, args=(c,)) for c in range(callers)]
  for th in ths:
    th.start()
  start.wait()
  t0 = time.perf_counter()
  for th in ths:
    th.join()
  dt = max(done) - t0
  return nq * steps * callers / dt, dt, outs[0][2]


def timed_steps(step, steps, flush, torch, dist):
  ms, agg = 0.0, {}
  for _ in range(steps):
    flush.zero_()  # L2 flush between timed iterations (not inside the CUDA-event interval)
    torch.cuda.synchronize()
    if dist is not None:
      dist.barrier()
    st = step()
    ms += st["ms_total"]
    for key, val in st.items():
      agg[key] = agg.get(key, 0) + val
  return ms, agg


def sharded_section(tag, wl, arrays, q, truth, full_index, rank, world, local_rank, dev, dist, steps, warmup, flush):
  """Database-sharded search of `arrays` over the ranks (leaf sharding, NCCL from the C++ library), strong scaling:
  the same queries on every rank.  Returns the report dict on rank 0."""
  import torch
  from scann_b200 import distributed as sdist
  nq, k = wl["nq"], wl["k"]
  d_q = torch.from_numpy(q).to(dev)
  d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
  d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
  t0 = time.time()
  sh = sdist.ShardedIndex(arrays, wl["probe"], wl["pre"], wl["k"], rank, world, local_rank)
  log(f"[rank {rank}] {tag}: shard index in {time.time() - t0:.1f}s")
  torch.cuda.synchronize()
  report = {}
  for light in (False, True):
    step = lambda: sh.search_batched_device(d_q, d_idx, d_dist, light=light)
    for _ in range(warmup):
      torch.cuda.synchronize()
      dist.barrier()
      step()
    found = d_idx.cpu().numpy().view(np.uint32)
    ms, agg = timed_steps(step, steps, flush, torch, dist)
    # end to end from pinned host queries to host results (wall clock, barrier on both sides)
    hq = torch.from_numpy(q).pin_memory()
    hi = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    hd = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    e_steps = max(3, min(steps, 10))
    torch.cuda.synchronize()
    dist.barrier()
    e0 = time.perf_counter()
    for _ in range(e_steps):
      d_q.copy_(hq, non_blocking=True)
      torch.cuda.synchronize()
      step()
      hi.copy_(d_idx, non_blocking=True)
      hd.copy_(d_dist, non_blocking=True)
      torch.cuda.synchronize()
    dist.barrier()
    e2e_s = time.perf_counter() - e0
    t = torch.tensor([ms, e2e_s] + [agg.get(s_, 0.0) for s_ in STAGES], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = torch.tensor([float(agg.get("scan_bytes_alg", 0)), float(agg.get("cand_sum", 0))], dtype=torch.float64, device=dev)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
      ms_max = float(tmax[0])
      r = {"qps": nq * steps / (ms_max / 1e3), "ms_per_step": ms_max / steps,
           "e2e_qps": nq * e_steps / float(tmax[1]),
           "stage_ms_per_step_max_over_ranks": {s_: float(tmax[2 + i]) / steps for i, s_ in enumerate(STAGES)},
           "exchange_bytes_per_rank_per_step": int(agg.get("exchange_bytes", 0) / steps),
           "scan_GBps_alg_all_ranks": float(tsum[0]) / steps / (float(tmax[2 + STAGES.index("ms_scan")]) / steps * 1e-3) / 1e9
                                      if float(tmax[2 + STAGES.index("ms_scan")]) > 0 else None,
           "candidates_per_query_all_ranks": float(tsum[1]) / steps / nq,
           "overflow_retries": int(agg.get("overflow_retries", 0))}
      if truth is not None:
        r["recall_at_10"] = recall_at_k(found, truth)
      if full_index is not None:
        r["ids_equal_single_gpu"] = bool(np.array_equal(found, full_index))
      report["light" if light else "parity"] = r
  del sh
  torch.cuda.empty_cache()
  return report


STAGES = ["ms_tokenize", "ms_lut", "ms_pilot", "ms_worklist", "ms_scan", "ms_compact", "ms_finalize", "ms_exchange", "ms_merge"]


def c5_section(args, rank, world, local_rank, dev, dist, flush, peaks, peak_src):
  """BASELINE.json configs[4] shape, database sharded over the ranks; at N = 1 the single-GPU figure."""
  import torch
  from scann_b200 import _lib
  wl = dict(WORKLOADS["c5_deep_shape"])
  if args.c5_n and args.c5_n != wl["n"]:
    wl["leaves"] = max(16, int(round(wl["leaves"] * args.c5_n / wl["n"])))
    wl["clusters"] = max(64, int(round(wl["clusters"] * args.c5_n / wl["n"])))
    wl["n"] = args.c5_n
  if args.c5_leaves:
    wl["probe"] = args.c5_leaves
  if args.c5_nq:
    wl["nq"] = args.c5_nq
  nq, k = wl["nq"], wl["k"]
  t0 = time.time()
  arrays = None
  db, q = make_data(wl, queries_only=rank != 0)
  if rank == 0:
    log(f"[c5] data {db.shape} in {time.time() - t0:.1f}s")
    t0 = time.time()
    arrays = build_arrays(wl, db, dev)
    log(f"[c5] index built in {time.time() - t0:.1f}s")
  arrays = share_arrays("c5", arrays, rank, world, dist, "dot_product")
  steps = max(3, min(args.steps, 10))
  out = {"workload": "c5_deep_shape", "n": wl["n"], "d": wl["d"], "leaves": wl["leaves"], "leaves_to_search": wl["probe"],
         "soar_lambda": wl["soar"], "reorder": wl["pre"], "k": k, "queries_per_step": nq, "steps": steps,
         "scaling": "strong (database sharded by leaf over the GPUs; every rank sees all queries)",
         "collectives": "NCCL from C++: all-gather (leaves), all-reduce min (tau), all-to-all (16 B records), all-gather (results)"}
  found_full, truth = None, None
  d_q = torch.from_numpy(q).to(dev)
  if rank == 0:
    # single-GPU searcher on the whole database: the reference result of the sharded runs (and the N = 1 figure)
    t0 = time.time()
    ix = _lib.NativeIndex(arrays, wl["probe"], wl["pre"], wl["k"], device=local_rank)
    log(f"[c5] single-GPU index in {time.time() - t0:.1f}s")
    truth = exact_topk(d_q, arrays.dataset, k, dev)
    d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
    d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)

    def step1():
      ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k)
      return ix.stats()
    for _ in range(3):
      step1()
    found_full = d_idx.cpu().numpy().view(np.uint32)
    if world == 1 or args.c5_single_gpu_timing:
      ms, agg = timed_steps(step1, steps, flush, torch, None)
      scan_ms = agg["ms_scan"] / steps
      single = {"qps": nq * steps / (ms / 1e3), "ms_per_step": ms / steps,
                "recall_at_10": recall_at_k(found_full, truth),
                "stage_ms_per_step": {s_: agg[s_] / steps for s_ in STAGES if s_ in agg},
                "scan_GBps_alg": agg["scan_bytes_alg"] / steps / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None,
                "scan_frac_of_hbm_peak": agg["scan_bytes_alg"] / steps / (scan_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]) if scan_ms > 0 else None,
                "candidates_per_query": agg["cand_sum"] / steps / nq,
                "queries_per_probed_leaf": nq * wl["probe"] / wl["leaves"],
                "scan_roofline": scan_roofline(agg, peaks, peak_src, None,
                                               "r02_scan_c5sparse_traffic.json" if nq * wl["probe"] / wl["leaves"] <= 8 else None)}
      # host buffers end to end
      hq = torch.from_numpy(q).pin_memory().numpy()
      oi = torch.empty((nq, k), dtype=torch.int32).pin_memory()
      od = torch.empty((nq, k), dtype=torch.float32).pin_memory()
      outp = (oi.numpy().view(np.uint32), od.numpy())
      ix.search_batched(hq, out=outp)
      e0 = time.perf_counter()
      for _ in range(steps):
        ix.search_batched(hq, out=outp)
      single["e2e_qps"] = nq * steps / (time.perf_counter() - e0)
      out["single_gpu"] = single
    ix.close()
    del ix
    torch.cuda.empty_cache()
  if world > 1:
    rep = sharded_section("c5", wl, arrays, q, truth, found_full, rank, world, local_rank, dev, dist, steps, 3, flush)
    if rank == 0:
      out["n_gpus"] = world
      out["sharded"] = rep
      out["qps"] = rep["parity"]["qps"]
      scan = rep["parity"].get("scan_GBps_alg_all_ranks")
      out["scan_frac_of_hbm_peak_per_gpu"] = scan / world / float(peaks["hbm_gbs"]) if scan else None
  elif rank == 0:
    out["n_gpus"] = 1
    out["qps"] = out["single_gpu"]["qps"]
  return out if rank == 0 else None


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default=os.environ.get("SCANN_B200_BENCH_WORKLOAD", "c2_glove_shape"), choices=list(WORKLOADS))
  ap.add_argument("--leaves", type=int, default=0, help="override leaves_to_search")
  ap.add_argument("--clusters", type=int, default=0, help="override the number of mixture components of the synthetic data")
  ap.add_argument("--n", type=int, default=0, help="override the database size (leaves are rescaled to keep rows per leaf)")
  ap.add_argument("--nq", type=int, default=0, help="override the queries per step")
  ap.add_argument("--noise", type=float, default=None,
                  help="AH noise_shaping_threshold used when the index is built (the reference builder's default is 0.2)")
  ap.add_argument("--sweep-leaves", default="",
                  help="comma-separated leaves_to_search values measured after the headline run (recall / QPS trade-off)")
  ap.add_argument("--cpu-sample", type=int, default=2000)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-c5", action="store_true", help="skip the C5-shape (sharded) section")
  ap.add_argument("--no-c3", action="store_true", help="skip the C3 brute-force section")
  ap.add_argument("--c5-n", type=int, default=0, help="database size of the C5-shape section (default 20M; BASELINE's C5 is 100M)")
  ap.add_argument("--c5-leaves", type=int, default=0, help="leaves_to_search of the C5-shape section (default 24)")
  ap.add_argument("--c5-nq", type=int, default=0, help="queries per step of the C5-shape section (default 10000)")
  ap.add_argument("--c5-single-gpu-timing", action="store_true", help="also time the single-GPU searcher when N > 1")
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
  if args.nq > 0:
    wl["nq"] = args.nq
  if args.noise is not None:
    wl["noise"] = args.noise

  import torch
  if wl.get("kind") == "bruteforce":
    if args.impl == "reference":
      return 0 if rank != 0 else run_bruteforce_reference(args, wl)
    out = run_bruteforce(args, wl, rank, world, local_rank)
    if rank == 0:
      emit(out)
    return 0
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
  db, q = make_data(wl, queries_only=rank != 0)
  arrays = None
  if rank == 0:
    log(f"[rank {rank}] data {db.shape} in {time.time() - t0:.1f}s")
    t0 = time.time()
    arrays = build_arrays(wl, db, dev)
    log(f"[rank {rank}] index built in {time.time() - t0:.1f}s")
  arrays = share_arrays("main", arrays, rank, world, dist, wl.get("distance", "dot_product"))
  db = arrays.dataset
  t0 = time.time()
  # Every rank holds a full replica (the C2 index is 0.5 GB) and serves its own query batches:
  # queries are independent units, so the headline N-GPU number needs no data-path collective.
  ix = _lib.NativeIndex(arrays, wl["probe"], wl["pre"], wl["k"], device=local_rank)
  log(f"[rank {rank}] device index in {time.time() - t0:.1f}s")

  nq, k = wl["nq"], wl["k"]
  d_q = torch.from_numpy(q).to(dev)
  d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
  d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
  torch.cuda.synchronize()

  # ground truth for recall (exact f32 brute force on the GPU)
  truth = exact_topk(d_q, db, k, dev, l2=wl.get("distance") == "squared_l2") if rank == 0 else None

  def step_dev():
    ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k)
    return ix.stats()

  # end-to-end leg: the caller's query and result arrays are page-locked host memory (numpy views of
  # pinned torch tensors), as a serving front end would hold them
  q_pin_t = torch.from_numpy(q).pin_memory()
  q_pin = q_pin_t.numpy()

  for _ in range(args.warmup):
    step_dev()
  found = d_idx.cpu().numpy().view(np.uint32)
  rec = recall_at_k(found, truth) if rank == 0 else None

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

  # end to end through the host-buffer C ABI call: one caller, and two callers with a batch in flight each
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  e2e_steps = max(3, min(args.steps, 10))
  e2e1_qps, e2e1_s, (idx_h, dist_h) = e2e_host_run(ix, q_pin, nq, k, e2e_steps, 1)
  if dist is not None:
    dist.barrier()
  e2e2_qps, e2e_s, _ = e2e_host_run(ix, q_pin, nq, k, e2e_steps, 2)
  if dist is not None:
    dist.barrier()
  sampler.stop_flag.set()
  sampler.join(timeout=2)
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
  peaks, peak_src = measured_peaks()
  c2_sharded = None
  if world > 1:
    rep = sharded_section("c2", wl, arrays, q, truth, found if rank == 0 else None, rank, world, local_rank, dev, dist,
                          max(3, min(args.steps, 10)), 3, flush)
    if rank == 0:
      c2_sharded = rep

  if dist is not None:
    t = torch.tensor([ms_total, e2e_s, e2e1_s], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, e2e1_s = float(t[0]), float(t[1]), float(t[2])

  # the headline index is not needed any more: make room for the other sections
  cpu_base = None
  if rank == 0 and not args.no_cpu_baseline:
    import oracle
    threads = os.cpu_count() or 1
    oi = oracle.OracleIndex(arrays, wl["probe"], wl["pre"], wl["k"])
    i_cpu, _ = oi.search_batched(q[:64], impl=1)
    parity = bool(np.array_equal(i_cpu, idx_h[:64]))
    qps, _ = cpu_reference_run(oi, q, min(args.cpu_sample, nq), threads, 1, 1)
    qps1, _ = cpu_reference_run(oi, q, min(500, nq), 1, 1, 0)
    cpu_base = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                "sample": f"first {min(args.cpu_sample, nq)} queries of the step, batches of 256 "
                          f"(search_batched_parallel semantics), AVX2 vpshufb oracle; exact top-N candidate "
                          f"contract (DESIGN.md section 2), not the reference's order-dependent int16 pre-filter",
                "single_thread_qps": qps1, "ids_equal_gpu_first_64": parity}
    del oi
  ix.close()
  del ix
  torch.cuda.empty_cache()

  c5 = None
  if not args.no_c5 and args.workload == "c2_glove_shape":
    if dist is not None:
      dist.barrier()
    c5 = c5_section(args, rank, world, local_rank, dev, dist, flush, peaks, peak_src)
  c3 = None
  if not args.no_c3 and args.workload == "c2_glove_shape":
    if dist is not None:
      dist.barrier()
    wl3 = dict(WORKLOADS["c3_bruteforce_bf16"])
    full = run_bruteforce(args, wl3, rank, world, local_rank, dist=dist, flush=flush)
    if rank == 0:
      c3 = {"qps": full["value"], "e2e_qps": full["e2e"]["value"], "ms_per_step": full["ms_per_step"],
            "n": wl3["n"], "d": wl3["d"], "k": wl3["k"], "queries_per_step_per_gpu": wl3["nq"],
            "recall_at_100_first512": full["config"]["recall_at_100_first512"],
            "tensor_tflops": full["roofline"]["achieved"], "tensor_frac_of_burst_peak": full["roofline"]["frac"],
            "tensor_frac_of_sustained_peak": full["roofline"]["frac_of_sustained_peak"],
            "useful_tflops_f32_equivalent": full["roofline"]["useful_tflops_f32_equivalent"],
            "bf_widenings": full.get("bf_widenings"), "row_sharded": full.get("row_sharded"),
            "parallelism": full["config"]["parallelism"]}

  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return 0

  value = world * nq * args.steps / (ms_total / 1e3)
  e2e_value = world * nq * e2e_steps * 2 / e2e_s          # two callers
  e2e_single = world * nq * e2e_steps / e2e1_s            # one caller
  clocks = sampler.summary()
  cfg = static_config(args, wl)
  cfg.update({"recall_at_10": rec, "queries_per_step_all_gpus": nq * world,
              "parallelism": f"query-parallel x{world} (one replica and one {nq}-query batch per GPU)",
              "wall_s_timed_region": wall})
  if c2_sharded is not None:
    cfg["c2_db_sharded"] = c2_sharded
  if c5 is not None:
    cfg["c5_sharded"] = c5
  if c3 is not None:
    cfg["c3_bruteforce"] = c3
  out = {
      "metric": "batched QPS at recall@10>=0.90 (tree-AH search_batched)", "value": value, "unit": "queries/s",
      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u8 LUT / int16 accumulate (scan), f32 (tokenize, reorder)", "data": "synthetic",
      "config": cfg,
      "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes) * world,
              "d2h_bytes_per_step": int(nq * k * 8) * world, "steps": e2e_steps * 2, "callers_per_gpu": 2,
              "single_caller_value": e2e_single,
              "note": "scann_b200_search_batched from page-locked host arrays; every step copies its queries in and "
                      "its results out inside the call; two caller threads per searcher keep one batch in flight "
                      "each (the library overlaps them on separate lanes), single_caller_value = one caller"},
      "gpu_launches": int(agg.get("kernel_launches", 0)),
      "clocks": clocks,
      **({"probe_sweep": sweep} if sweep else {}),
      "roofline": scan_roofline(agg, peaks, peak_src, clocks,
                                "r02_scan_main_c2_traffic.json" if args.workload == "c2_glove_shape" else None),
      "stage_ms_per_step": {s: agg[s] / args.steps for s in agg if s.startswith("ms_")},
      "overflow_retries": int(agg.get("overflow_retries", 0)),
      "tokenize_fallbacks_per_step": agg.get("tokenize_fallbacks", 0) / args.steps,
      "candidates_per_query": {"mean": agg.get("cand_sum", 0) / (nq * args.steps), "max_over_steps_sum": int(agg.get("cand_max", 0))},
  }
  if cpu_base is not None:
    out["cpu_baseline"] = cpu_base
  emit(out)
  if dist is not None:
    dist.destroy_process_group()
  return 0


def gen_normal(n, d, seed, threads):
  """[n, d] standard normals, every 64k-row chunk from its own stream (seed, chunk) on a thread pool."""
  from concurrent.futures import ThreadPoolExecutor
  out = np.empty((n, d), np.float32)
  chunk = 1 << 16

  def fill(ci):
    s = ci * chunk
    e = min(n, s + chunk)
    out[s:e] = np.random.default_rng([seed, ci]).standard_normal((e - s, d), dtype=np.float32)
  with ThreadPoolExecutor(max(1, threads)) as ex:
    list(ex.map(fill, range((n + chunk - 1) // chunk)))
  return out


def bruteforce_data(wl, threads):
  from scann_b200 import index_build
  n, d, nq = wl["n"], wl["d"], wl["nq"]
  f32 = wl.get("dtype") == "f32"
  x = gen_normal(n, d, wl["seed"], threads)
  bits = x if f32 else index_build.bfloat16_quantize(x)
  q = np.random.default_rng(wl["seed"] + 1).standard_normal((nq, d), dtype=np.float32)
  a = index_build.IndexArrays(distance="dot_product", dataset=bits if f32 else None, n=n, d=d)
  if not f32:
    a.bf16_dataset = bits
  return a, bits, q


def run_bruteforce_reference(args, wl):
  """CPU arm of C3: the oracle's exact f32-query x bf16-row brute force, all host threads, >= 256 queries."""
  import oracle
  threads = os.cpu_count() or 1
  n, d, nq, k = wl["n"], wl["d"], wl["nq"], wl["k"]
  f32 = wl.get("dtype") == "f32"
  a, bits, q = bruteforce_data(wl, threads)
  sample = min(nq, max(256, 16 * threads))
  fn = oracle.bruteforce_f32 if f32 else oracle.bruteforce_bf16
  fn(bits, q[:threads], k, threads=threads)  # warm-up
  t0 = time.perf_counter()
  fn(bits, q[:sample], k, threads=threads)
  dt = time.perf_counter() - t0
  qps = sample / dt
  emit({"impl": "reference", "metric": f"batched QPS, {'f32' if f32 else 'bf16'} brute-force MIPS k=100", "value": qps,
        "unit": "queries/s", "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 db x f32 query",
        "data": "synthetic", "config": {"workload": args.workload, "n": n, "d": d, "k": k, "queries_per_step": nq},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the step's {nq} queries (the CPU needs ~{nq / qps:.0f} s for all of them)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
  return 0


def run_bruteforce(args, wl, rank, world, local_rank, dist=None, flush=None):
  """C3: bf16 brute force, 10k queries x 1M x 768, k = 100 (tcgen05 GEMM + fused top-k pre-filter).  Returns the
  report dict (rank 0 emits it when C3 is the selected workload; the default run nests it under config)."""
  import torch
  from scann_b200 import _lib
  n, d, nq, k = wl["n"], wl["d"], wl["nq"], wl["k"]
  f32 = wl.get("dtype") == "f32"
  t0 = time.time()
  a, bits, q = bruteforce_data(wl, max(1, (os.cpu_count() or 1) // max(world, 1)))
  log(f"[bf rank {rank}] data in {time.time() - t0:.1f}s")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  own_group = False
  if world > 1 and dist is None:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    own_group = True
  ix = _lib.NativeIndex(a, 1, k, k, device=local_rank)
  d_q = torch.from_numpy(q).to(dev)
  d_idx = torch.zeros((nq, k), dtype=torch.int32, device=dev)
  d_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
  if flush is None:
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
  steps = args.steps
  for _ in range(steps):
    flush.zero_()
    torch.cuda.synchronize()
    ix.search_batched_device(d_q.data_ptr(), nq, d_idx.data_ptr(), d_dist.data_ptr(), k)
    st = ix.stats()
    ms_total += st["ms_total"]
    for key, val in st.items():
      agg[key] = agg.get(key, 0) + val
  e2e_steps = max(2, min(steps, 5))
  q_pin = torch.from_numpy(q).pin_memory().numpy()
  oi_t = torch.empty((nq, k), dtype=torch.int32).pin_memory()
  od_t = torch.empty((nq, k), dtype=torch.float32).pin_memory()
  out_pin = (oi_t.numpy().view(np.uint32), od_t.numpy())
  ix.search_batched(q_pin, out=out_pin)
  if dist is not None:
    dist.barrier()
  e0 = time.perf_counter()
  for _ in range(e2e_steps):
    ix.search_batched(q_pin, out=out_pin)
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
    ix.close()
    del ix
    torch.cuda.empty_cache()
    sh = sd.ShardedBruteForce(a, k, rank, world, local_rank)
    for _ in range(args.warmup):
      sh.search_batched_device(d_q, d_idx, d_dist)
    same = bool((d_idx == ref_idx).all().item())
    dist.barrier()
    torch.cuda.synchronize()
    sh_ms, sh_agg = 0.0, {}
    for _ in range(steps):
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
    row_sharded = {"value": nq * steps / (float(t[0]) * 1e-3), "unit": "queries/s",
                   "ids_equal_replica": same, "timing": "wall clock around search + all-gather + merge, max over ranks",
                   "allgather_bytes_per_rank": sh_agg.get("allgather_bytes_per_rank", 0) // max(steps, 1),
                   "stage_ms_per_step": {s_: sh_agg[s_] / steps for s_ in sh_agg if s_.startswith("ms_")}}
    del sh
    torch.cuda.empty_cache()
    if own_group and rank != 0:
      dist.destroy_process_group()
    if rank != 0:
      return None
  else:
    ix.close()
    del ix
    torch.cuda.empty_cache()
  flops = 2.0 * nq * n * d * (3 if f32 else 2)  # bf16 split terms per product: hi.hi + lo.hi (+ hi.lo for f32 rows)
  gemm_s = agg["ms_scan"] / steps * 1e-3
  peaks, peak_src = measured_peaks()
  # the GEMM rounds run as ~20 ms bursts between L2 flushes, i.e. "a kernel timed alone": burst peak
  peak = float(peaks.get("bf16_tflops", 1590.0))
  peak_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
  out = {"metric": f"batched QPS, {'f32' if f32 else 'bf16'} brute-force MIPS k=100", "value": world * nq * steps / (ms_total * 1e-3),
         "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
         "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
         "dtype": ("(bf16 hi + lo) x (bf16 hi + lo) -> f32 (tcgen05, 3 terms), exact f32 chain re-scoring" if f32 else
                   "bf16 x (bf16 hi + bf16 lo) -> f32 (tcgen05), f32 re-scoring"), "data": "synthetic",
         "config": {"workload": "c3_bruteforce_f32" if f32 else "c3_bruteforce_bf16", "n": n, "d": d, "k": k, "queries_per_step": nq,
                    "recall_at_100_first512": rec, "l2_flush": "256 MiB write between timed steps",
                    "parallelism": f"query-parallel x{world} (one replica and one {nq}-query batch per GPU)"},
         "e2e": {"value": world * nq * e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes) * world,
                 "d2h_bytes_per_step": int(nq * k * 8) * world, "steps": e2e_steps},
         "gpu_launches": int(agg["kernel_launches"]), "clocks": sampler.summary(),
         "bf_widenings": int(agg.get("bf_widenings", 0)), "bf_exact_fallbacks": int(agg.get("bf_exact_fallbacks", 0)),
         "roofline": {"bound": "tensor", "kernel": "bf::gemm_pair_kernel<1, filter>" if f32 else "bf::gemm_pair_kernel<2, filter>", "achieved": flops / gemm_s / 1e12,
                      "peak": peak, "peak_source": "measured burst" if peak_src == "measured" else "fallback", "unit": "TFLOP/s",
                      "frac": flops / gemm_s / 1e12 / peak, "traffic": ncu_traffic("r01_gemm_pair_traffic.json"),
                      "traffic_note": "dram bytes of the largest round's launch (497,664 rows: 764 MB compulsory)",
                      "frac_of_sustained_peak": flops / gemm_s / 1e12 / peak_sustained,
                      "note": "flops count both bf16 query terms (hi + lo); time includes the compactions between rounds",
                      "useful_tflops_f32_equivalent": flops / 2 / gemm_s / 1e12},
         "stage_ms_per_step": {s: agg[s] / steps for s in agg if s.startswith("ms_")}}
  if row_sharded is not None:
    out["row_sharded"] = row_sharded
    if own_group:
      dist.destroy_process_group()
  return out


def run_reference(args, wl):
  """CPU arm: the oracle's AVX2 restatement of the reference path, all host threads, the same queries per step and
  the same warm-up as the GPU arm.  The index is built on the CPU; libscann_b200.so is not loaded."""
  import oracle
  threads = os.cpu_count() or 1
  t0 = time.time()
  db, q = make_data(wl)
  arrays = build_arrays_cpu(wl, db, threads)
  log(f"[reference] data + CPU-built index in {time.time() - t0:.1f}s")
  assert "scann_b200._lib" not in sys.modules, "the reference arm must not load the product library"
  oi = oracle.OracleIndex(arrays, wl["probe"], wl["pre"], wl["k"])
  nq = wl["nq"]
  steps, warmup = max(1, args.steps), max(0, args.warmup)
  # keep the whole run within a few minutes: probe the speed first
  qps_probe, _ = cpu_reference_run(oi, q, min(512, nq), threads, 1, 0)
  budget_s = 150.0
  max_steps = max(1, int(budget_s * qps_probe / nq) - warmup)
  steps_run = min(steps, max_steps)
  qps, s_per_step = cpu_reference_run(oi, q, nq, threads, steps_run, warmup)
  cfg = static_config(args, wl)
  out = {
      "impl": "reference", "metric": "batched QPS at recall@10>=0.90 (tree-AH search_batched)", "value": qps,
      "unit": "queries/s", "n_gpus": args.gpus, "steps": steps_run, "warmup": warmup,
      "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u8 LUT / int16 accumulate (AVX2)", "data": "synthetic",
      "config": cfg,
      "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                       "sample": f"all {nq} queries of the step, batches of 256 over {threads} threads; index built on "
                                 f"the CPU (torch trainers + the oracle's restatement of the reference builder); exact "
                                 f"top-N candidate contract, not the reference's order-dependent int16 pre-filter"},
      "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  emit(out)
  return 0


if __name__ == "__main__":
  _claim_stdout()
  sys.exit(main())
