"""Shared helpers: golden fixture loading and an independent numpy restatement of the path."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
  return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
  from scann_b200 import index_build
  z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
  dist = str(z["distance"]) if "distance" in z.files else "dot_product"
  a = index_build.IndexArrays(distance=dist, dataset=z["dataset"], n=z["dataset"].shape[0],
                              d=z["dataset"].shape[1])
  a.centers, a.tokens, a.codes = z["centers"], z["tokens"], z["codes"]
  a.codebook, a.block_dims = z["codebook"], z["block_dims"]
  a.soar = bool(int(z["soar"]))
  a.soar_codes = z["soar_codes"] if a.soar else None
  a.overretrieve = float(z["overretrieve"])
  a.residual = dist == "dot_product"
  return a, z


def f32(x):
  return np.asarray(x, dtype=np.float32)


def fma32(a, b, c):
  """float32 fused multiply-add emulated in float64 (product of two f32 is exact in f64)."""
  return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def np_center_distances(q, centers):
  """many_to_many_impl.inc:551-563: acc = fnmadd(q[d], c[d], acc), sequential in d."""
  acc = np.zeros((q.shape[0], centers.shape[0]), np.float32)
  for d in range(q.shape[1]):
    acc = fma32(-q[:, d:d + 1], centers[None, :, d], acc)
  return acc


def np_center_distances_i8(q, centers, distance):
  """KMeansTreeNode::GetAllDistancesInt8 (trees/kmeans_tree/kmeans_tree_node.h:222-256), an independent numpy
  restatement: fixed-point centres as CreateFixedPointCenters derives them, the three-at-a-time AVX2 order of
  OneToManyAsymmetricTemplate<.., int8_t> for centres [0, 3 (L / 3)) and the one-to-one order of
  DenseDotProductInt8FloatAvxImpl (dot_product_impl.inc:3-54) for the last L mod 3."""
  from scann_b200 import index_build
  ci8, inv, sqn = index_build.quantize_centers(centers)
  l2 = distance != "dot_product"
  scale = f32(inv * np.float32(2.0)) if l2 else inv
  qp = f32(q * scale[None, :])                                # [nq, D]
  c = ci8.astype(np.float32)                                  # [L, D]
  nq, D = q.shape
  L = c.shape[0]
  L3 = L // 3 * 3
  out = np.empty((nq, L), np.float32)
  # main kernel: eight fnmadd lanes
  a = np.zeros((8, nq, L), np.float32)
  j = 0
  while j + 8 <= D:
    for l in range(8):
      a[l] = fma32(-qp[:, j + l, None], c[None, :, j + l], a[l])
    j += 8
  if j + 4 <= D:
    for l in range(4):
      a[l] = fma32(-qp[:, j + l, None], c[None, :, j + l], a[l])
    j += 4
  r = f32(f32(f32(a[0] + a[4]) + f32(a[2] + a[6])) + f32(f32(a[1] + a[5]) + f32(a[3] + a[7])))
  while j < D:
    r = fma32(-qp[:, j, None], c[None, :, j], r)
    j += 1
  out[:] = r
  # tail rows: one-to-one kernel, two accumulators over 16 dims, rounded product added in the 4-wide step
  if L3 < L:
    ct = c[L3:]
    a0 = np.zeros((8, nq, L - L3), np.float32)
    a1 = np.zeros((8, nq, L - L3), np.float32)
    j = 0
    while j + 16 <= D:
      for l in range(8):
        a0[l] = fma32(ct[None, :, j + l], qp[:, j + l, None], a0[l])
        a1[l] = fma32(ct[None, :, j + 8 + l], qp[:, j + 8 + l, None], a1[l])
      j += 16
    if j + 8 <= D:
      for l in range(8):
        a0[l] = fma32(ct[None, :, j + l], qp[:, j + l, None], a0[l])
      j += 8
    if j + 4 <= D:
      for l in range(4):
        a0[l] = f32(a0[l] + f32(ct[None, :, j + l] * qp[:, j + l, None]))
      j += 4
    v = f32(a0 + a1)
    s = f32(f32(f32(v[0] + v[4]) + f32(v[2] + v[6])) + f32(f32(v[1] + v[5]) + f32(v[3] + v[7])))
    while j < D:
      s = fma32(ct[None, :, j], qp[:, j, None], s)
      j += 1
    out[:, L3:] = -s
  if l2:
    qn = index_build.squared_l2_norms(q)
    out = f32(out + f32(qn[:, None] + sqn[None, :]))
  return out


def i8_tok_arrays(L, D, distance, seed=0):
  """A minimal tree-AH asset set whose only interesting part is the centres (the tokenizer reads nothing else)."""
  from scann_b200 import index_build
  rng = np.random.default_rng(seed)
  n = max(4 * L, 64)
  db = rng.standard_normal((n, D)).astype(np.float32)
  a = index_build.IndexArrays(distance=distance, dataset=db, n=n, d=D)
  a.centers = (rng.standard_normal((L, D)) * rng.uniform(0.2, 3.0, D)[None, :]).astype(np.float32)
  if D > 2:
    a.centers[:, 1] = 0.0                                    # an all-zero column: multiplier 1
  a.tokens = (np.arange(n) % L).astype(np.int32)
  a.codes = rng.integers(0, 16, (n, D), dtype=np.uint8)       # one dim per block
  a.codebook = rng.standard_normal((D, 16, 1)).astype(np.float32)
  a.block_dims = np.ones(D, np.int32)
  a.soar, a.soar_codes, a.overretrieve, a.residual = False, None, 2.0, distance == "dot_product"
  a.int8_tokenization = True
  return a, rng.standard_normal((9, D)).astype(np.float32)


def np_lut_dpb2(q, codebook):
  """asymmetric_hashing_impl.cc:505-645 for dims_per_block == 2 (two products, one add, no FMA)."""
  nq, B = q.shape[0], codebook.shape[0]
  qb = q.reshape(nq, B, 1, 2)
  p0 = f32(qb[..., 0] * codebook[None, :, :, 0])
  p1 = f32(qb[..., 1] * codebook[None, :, :, 1])
  raw = f32(f32(-p0) - p1)                                   # [nq, B, 16]
  mx = np.abs(raw).reshape(nq, -1).max(1)
  denom = np.maximum(mx, np.sqrt(np.float32(np.finfo(np.float32).eps))).astype(np.float32)
  mult = (np.float32(127.0) / denom).astype(np.float32)
  v = f32(raw * mult[:, None, None])
  r = np.where(v >= 0, np.floor(v + np.float32(0.5)), np.ceil(v - np.float32(0.5)))  # round half away
  # floor(v+0.5) in f32 can differ from round() only when v+0.5 rounds up to the next integer;
  # redo those in float64, which represents v + 0.5 exactly
  v64 = v.astype(np.float64)
  r = np.where(v64 >= 0, np.floor(v64 + 0.5), np.ceil(v64 - 0.5))
  return (r + 128).astype(np.uint8), mult


def np_slots(tokens, n_leaves, soar):
  """datapoints_by_token in file order (scann.cc:88-98)."""
  mult = 2 if soar else 1
  out = [[] for _ in range(n_leaves)]
  for j, t in enumerate(tokens.tolist()):
    if t >= 0:
      out[t].append(j // mult)
  return [np.asarray(x, dtype=np.int64) for x in out]


def np_search(arrays, q, probe, pre, k):
  """End-to-end numpy restatement (dims_per_block == 2, dot product)."""
  a = arrays
  L, B = a.centers.shape[0], a.codes.shape[1]
  cd = np_center_distances(q, a.centers)
  lut, mult = np_lut_dpb2(q, a.codebook)
  slots = np_slots(a.tokens, L, a.soar)
  disjoint = not a.soar
  nover = pre if disjoint else int(float(pre) * float(np.float32(a.overretrieve)))
  ids = np.zeros((q.shape[0], k), np.uint32)
  dists = np.full((q.shape[0], k), np.nan, np.float32)
  cands = []
  for i in range(q.shape[0]):
    order = np.lexsort((np.arange(L), cd[i]))[:probe]
    inv = np.float32(1.0 / np.float64(mult[i]))
    rows = []
    for leaf in order.tolist():
      dps = slots[leaf]
      if len(dps) == 0:
        continue
      codes = a.codes[dps]
      if a.soar:
        sec = a.tokens[2 * dps + 1] == leaf
        codes = np.where(sec[:, None], a.soar_codes[dps], codes)
      acc = lut[i][np.arange(B)[None, :], codes].astype(np.int32).sum(1) - 128 * B
      score = f32(f32(acc.astype(np.float32) * inv) + cd[i, leaf])
      for s in range(len(dps)):
        rows.append((float(score[s]), leaf, s, int(dps[s]), np.float32(score[s])))
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    rows = rows[:nover]
    cands.append(rows)
    best = {}
    if disjoint:
      best = {r[3]: r[4] for r in rows}
    else:
      for r in rows:
        if r[3] in best:
          best[r[3]] = np.float32(np.float32(0.5) * best[r[3]] + np.float32(0.5) * r[4])
        else:
          best[r[3]] = r[4]
      keep = sorted(best.items(), key=lambda kv: (float(kv[1]), kv[0]))[:pre]
      best = dict(keep)
    ex = []
    for dp in best:
      ex.append((-float(np.dot(q[i].astype(np.float64), a.dataset[dp].astype(np.float64))), dp))
    ex.sort()
    for j, (dist, dp) in enumerate(ex[:k]):
      ids[i, j] = dp
      dists[i, j] = -dist
  return ids, dists, cd, lut, mult, cands


def reference_pipeline_search(c, q):
  """SearchBatched of the tree-AH index of `c` (a conftest.Case with >= 8, even dims per block, no SOAR) in which every
  NUMBER comes from the reference's own compiled code (oracle/_ref): centre distances (many-to-many accumulation), the
  lookup table (one-to-many AVX2 kernel, SSE4 one-to-one kernel for centre 15, fixed-point conversion), per-leaf scores
  (LUT16 AVX2 kernel on the reference's packing), exact reordering distances (one-to-many AVX2 kernel).  Only the
  selections are written here: top P leaves by (distance, leaf), top N' candidates by (score, leaf, slot) -- (score, id)
  for squared L2, TreeXHybridSMMD -- and top k by (distance, id).  Returns (ids [nq, k] u32, distances [nq, k] f32) as
  the API reports them (dot product: similarities, result_multiplier_ of scann_ops/cc/scann.cc:365-369)."""
  from oracle import ref
  a = c.arrays
  assert not a.soar
  l2 = a.distance == "squared_l2"
  L = a.centers.shape[0]
  B, _, dpb = a.codebook.shape
  assert dpb >= 8 and dpb % 2 == 0 and a.block_dims.min() == dpb
  canon = lambda x: np.float32(x) + np.float32(0.0)           # DistanceComparator: -0.0 == +0.0
  cd_all = ref.many_to_many_f32(q, a.centers, squared_l2=l2)
  ids = np.zeros((len(q), c.k), np.uint32)
  dists = np.full((len(q), c.k), np.nan, np.float32)
  for i in range(len(q)):
    leaves = np.lexsort((np.arange(L), canon(cd_all[i])))[:c.probe]
    raw = np.empty((B, 16), np.float32)
    for b in range(B):
      qb = q[i, b * dpb:(b + 1) * dpb]
      raw[b, :15] = ref.one_to_many_f32(qb, a.codebook[b, :15], squared_l2=l2)
      one = ref.one_to_one_sse4(qb, a.codebook[b, 15], squared_l2=l2)
      raw[b, 15] = np.float32(one if l2 else -one)
    lut, mult = ref.lut_to_fixed_point(raw)
    rows = []
    for leaf in leaves.tolist():
      dps = c.oracle.leaf_datapoints(leaf)
      if len(dps) == 0:
        continue
      packed = ref.pack_dataset(a.codes[dps])
      bias = np.float32(0.0) if l2 else cd_all[i, leaf]       # squared L2 (SMMD): no leaf bias
      (idx, dist), = ref.lut16_top_float(packed, len(dps), B, [lut], [bias], [mult])
      assert len(idx) == len(dps)
      for s, d in zip(idx.tolist(), dist.tolist()):
        rows.append((float(canon(d)), (int(dps[s]),) if l2 else (leaf, s), int(dps[s])))
    rows.sort(key=lambda t: (t[0], t[1]))
    cand = np.asarray([r[2] for r in rows[:c.pre]], np.int64)
    pad = (-len(cand)) % 3                                     # the kernel takes three rows at a time
    padded = np.concatenate([cand, np.repeat(cand[-1:], pad)])
    exact = ref.one_to_many_f32(q[i], a.dataset[padded], squared_l2=l2)[:len(cand)]
    order = np.lexsort((cand, canon(exact)))[:c.k]
    ids[i, :len(order)] = cand[order]
    dists[i, :len(order)] = exact[order] if l2 else -exact[order]
  return ids, dists
