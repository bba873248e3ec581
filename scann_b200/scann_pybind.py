"""`scann_pybind.ScannNumpy` -- the class `scann_ops_pybind.py` talks to.

Mirror of the pybind class bound in scann/scann_ops/cc/python/scann_pybind.cc:28-53 and
implemented in scann_ops/cc/scann_npy.cc: same constructor overloads, method names, argument
meaning, return shapes/dtypes and exception types/prefixes (scann_npy.cc:41-47,62-76,214-256).
Everything on the query path goes through the C ABI (include/scann_b200.h); this file only
marshals numpy buffers, like scann_npy.cc does.  Index training (k-means, AH codebooks) is done by
`index_build` (out of scope of the hot path).
"""
import ctypes as C
import math
import os

import numpy as np

from . import _lib, config as cfgmod, index_build


def _runtime(prefix, exc):
  msg = exc.message if isinstance(exc, _lib.ScannB200Error) else str(exc)
  return RuntimeError(prefix + msg)


class _Plan:
  """What the ScannConfig asks for, reduced to the fields the query path needs."""

  def __init__(self, text):
    try:
      c = cfgmod.parse(text)
    except cfgmod.TextProtoError as e:
      raise RuntimeError(f"Error initializing searcher: Failed to parse config: {e}")
    self.msg = c
    name = c.path("distance_measure", "distance_measure", default="SquaredL2Distance")
    dist = {"DotProductDistance": "dot_product", "SquaredL2Distance": "squared_l2"}.get(name)
    if dist is None:
      raise RuntimeError(f"Error initializing searcher: UNIMPLEMENTED: distance measure {name}")
    self.distance = dist
    self.num_neighbors = cfgmod.as_int(c.get("num_neighbors"), 1)
    self.partitioning = c.get("partitioning")
    self.ah = c.path("hash", "asymmetric_hash")
    self.brute_force = c.get("brute_force")
    self.reordering = c.get("exact_reordering")
    self.autopilot = c.get("autopilot")

  def check_supported(self):
    def unimpl(what):
      raise RuntimeError(f"Error initializing searcher: UNIMPLEMENTED: {what} is outside the scann_b200 hot path")
    if self.autopilot is not None:
      unimpl("autopilot")
    if self.is_brute_force():
      # score_brute_force(): BruteForceSearcher<float> / Bfloat16BruteForceSearcher (brute_force/*.cc)
      if self.distance != "dot_product" and self.bf16_brute_force():
        unimpl("bfloat16 brute force with a distance other than dot product")  # bfloat16_brute_force.cc:60-75: MIPS only
      if cfgmod.as_bool(self.brute_force.path("fixed_point", "enabled"), False):
        unimpl("int8 brute force")
      if self.reordering is not None:
        unimpl("reordering after brute force")
      return
    if self.partitioning is None or self.ah is None:
      unimpl("a searcher without tree() + score_ah() or score_brute_force()")
    p, ah = self.partitioning, self.ah
    if p.has("projection") or p.has("bottom_up_top_level_partitioner"):
      unimpl("PCA/truncate projections and upper_tree")
    # tree(quantize_centroids=True): FIXED_POINT_INT8 query tokenization (csrc/prep.cu tokenize_i8_kernel)
    if p.get("query_tokenization_type", "FLOAT") not in ("FLOAT", "FIXED_POINT_INT8"):
      unimpl("query_tokenization_type " + str(p.get("query_tokenization_type")))
    if p.get("database_tokenization_type", "FLOAT") != "FLOAT":
      unimpl("database_tokenization_type other than FLOAT")
    # training options the GPU trainer does not honour are refused, not dropped (a user who asks for anisotropic
    # centroids must not get plain k-means silently)
    avq = cfgmod.as_float(p.get("avq"), math.nan)
    if avq is not None and not math.isnan(avq):
      unimpl("anisotropic (AVQ) partitioning")
    if p.get("single_machine_center_initialization", "RANDOM_INITIALIZATION") != "RANDOM_INITIALIZATION":
      unimpl("k-means centre initialisation other than RANDOM_INITIALIZATION")
    if p.path("query_spilling", "spilling_type") not in ("FIXED_NUMBER_OF_CENTERS",):
      unimpl("query spilling other than FIXED_NUMBER_OF_CENTERS")
    if ah.get("lookup_type") != "INT8_LUT16" or cfgmod.as_int(ah.get("num_clusters_per_block"), 256) != 16:
      unimpl("lut256 / non-LUT16 asymmetric hashing")
    if ah.path("projection", "projection_type") not in ("CHUNK", "VARIABLE_CHUNK"):
      unimpl("AH projections other than CHUNK / VARIABLE_CHUNK")
    residual = cfgmod.as_bool(ah.get("use_residual_quantization"), False)
    if self.distance == "dot_product" and not residual:
      unimpl("non-residual dot-product tree-AH")
    if self.distance == "squared_l2" and residual:
      unimpl("residual squared-L2 tree-AH")
    r = self.reordering
    if r is not None:
      if self.int8_reorder():
        thr = cfgmod.as_float(r.path("fixed_point", "noise_shaping_threshold"), math.nan)
        if thr is not None and not math.isnan(thr):
          unimpl("noise-shaped int8 quantization of the reordering dataset")
        quantile = cfgmod.as_float(r.path("fixed_point", "fixed_point_multiplier_quantile"), 1.0)
        if quantile is not None and abs(quantile - 1.0) >= 0.001:
          unimpl("fixed_point_multiplier_quantile != 1")
      if self.bf16_reorder():
        thr = cfgmod.as_float(r.path("bfloat16", "noise_shaping_threshold"), math.nan)
        if thr is not None and not math.isnan(thr):
          unimpl("noise-shaped bfloat16 quantization of the reordering dataset")

  def is_brute_force(self):
    return self.brute_force is not None and self.partitioning is None and self.ah is None

  def bf16_brute_force(self):
    return self.is_brute_force() and cfgmod.as_bool(self.brute_force.path("bfloat16", "enabled"), False)

  def bf16_reorder(self):
    """exact_reordering { bfloat16 { enabled: true } } (Bfloat16ReorderingHelper, utils/reordering_helper.cc:720-757)."""
    r = self.reordering
    return r is not None and cfgmod.as_bool(r.path("bfloat16", "enabled"), False)

  def int8_reorder(self):
    """exact_reordering { fixed_point { enabled: true } } (FixedPointFloatDense*ReorderingHelper,
    utils/reordering_helper.cc:384-441,581-618; base/reordering_helper_factory.cc:106-175)."""
    r = self.reordering
    return r is not None and cfgmod.as_bool(r.path("fixed_point", "enabled"), False)

  def int8_tokenization(self):
    """partitioning { query_tokenization_type: FIXED_POINT_INT8 } (scann_builder.py:231, partitioner_factory.cc:95-98)."""
    p = self.partitioning
    return p is not None and p.get("query_tokenization_type", "FLOAT") == "FIXED_POINT_INT8"

  def dims_per_block(self):
    proj = self.ah.get("projection")
    if proj.get("projection_type") == "CHUNK":
      return cfgmod.as_int(proj.get("num_dims_per_block"))
    blocks = proj.all("variable_blocks")
    return cfgmod.as_int(blocks[0].get("num_dims_per_block"))


class ScannNumpy:
  """ScannNumpy(artifacts_dir: str, assets_pbtxt: str) | ScannNumpy(db, config_text, training_threads)."""

  def __init__(self, *args):
    self._index = None
    self._arrays = None
    self._assets = None
    if len(args) == 2 and isinstance(args[0], str):
      self._load(args[0], args[1])
    elif len(args) == 3:
      self._build(args[0], args[1], args[2])
    else:
      raise TypeError("ScannNumpy(artifacts_dir, assets_pbtxt) or ScannNumpy(dataset, config, training_threads)")

  # ---- construction ----
  def _build(self, db, config_text, training_threads):
    del training_threads  # training runs on the GPU (or torch CPU threads)
    db = np.asarray(db)
    if db.ndim != 2:
      raise ValueError("Dataset input must be two-dimensional")  # scann_npy.cc:70-71
    plan = _Plan(config_text)
    plan.check_supported()
    self._config_text = config_text
    db = np.ascontiguousarray(db, dtype=np.float32)
    if plan.is_brute_force():
      try:
        arrays = index_build.IndexArrays(distance=plan.distance, dataset=None, n=db.shape[0], d=db.shape[1])
        if plan.bf16_brute_force():
          arrays.bf16_dataset = index_build.bfloat16_quantize(db)   # bfloat16_brute_force.cc:60-75
        else:
          arrays.dataset = db
        self._finish(arrays, plan)
      except _lib.ScannB200Error as e:
        raise _runtime("Error initializing searcher: ", e)
      return
    p, ah = plan.partitioning, plan.ah
    soar = p.path("database_spilling", "spilling_type") in ("TWO_CENTER_ORTHOGONALITY_AMPLIFIED", "SOAR")
    lam = cfgmod.as_float(p.path("database_spilling", "orthogonality_amplification_lambda"), 1.5) if soar else None
    thr = cfgmod.as_float(ah.get("noise_shaping_threshold"), math.nan)
    if thr is None:
      thr = math.nan
    # the threshold selects Indexer::HashWithNoiseShaping in scann_b200_encode_database (csrc/encode.cu)
    try:
      arrays = index_build.build_tree_ah(
          db, plan.distance, num_leaves=cfgmod.as_int(p.get("num_children")),
          dims_per_block=plan.dims_per_block(),
          training_sample_size=cfgmod.as_int(p.get("expected_sample_size"), 100000),
          tree_iters=cfgmod.as_int(p.get("max_clustering_iterations"), 12),
          ah_iters=cfgmod.as_int(ah.get("max_clustering_iterations"), 10),
          soar_lambda=lam, overretrieve=cfgmod.as_float(p.path("database_spilling", "overretrieve_factor"), 2.0),
          spherical=p.get("partitioning_type", "GENERIC") == "SPHERICAL",
          keep_dataset=plan.reordering is not None, noise_shaping_threshold=thr)
      if plan.bf16_reorder():
        # reordering_helper.cc:729-730: the reordering dataset is Bfloat16QuantizeFloatDataset(original)
        arrays.bf16_dataset = index_build.bfloat16_quantize(db)
        arrays.dataset = None
      elif plan.int8_reorder():
        # reordering_helper.cc:384-397,581-595: ScalarQuantizeFloatDataset(original, quantile 1.0) (+ row norms for L2)
        arrays.int8_dataset, arrays.int8_multipliers = index_build.int8_quantize(db)
        if plan.distance == "squared_l2":
          arrays.dp_norms = index_build.squared_l2_norms(db)
        arrays.dataset = None
      self._finish(arrays, plan)
    except _lib.ScannB200Error as e:
      raise _runtime("Error initializing searcher: ", e)

  def _finish(self, arrays, plan):
    final_nn = plan.num_neighbors
    if plan.is_brute_force():
      self._arrays, self._plan = arrays, plan
      self._n, self._d = arrays.n, arrays.d
      self._index = _lib.NativeIndex(arrays, 1, final_nn, final_nn)
      return
    arrays.int8_tokenization = plan.int8_tokenization()
    leaves = cfgmod.as_int(plan.partitioning.path("query_spilling", "max_spill_centers"), arrays.centers.shape[0])
    pre = cfgmod.as_int(plan.reordering.get("approx_num_neighbors"), final_nn) if plan.reordering is not None else final_nn
    self._arrays = arrays
    self._plan = plan
    self._n, self._d = arrays.n, arrays.d
    self._index = _lib.NativeIndex(arrays, leaves, pre, final_nn)

  def _load(self, artifacts_dir, assets_pbtxt):
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.scann_b200_assets_load(artifacts_dir.encode(), assets_pbtxt.encode(), C.byref(h))
    if rc:
      raise RuntimeError("Error loading artifacts: " + L.scann_b200_last_error().decode("utf-8", "replace"))
    try:
      desc = _lib.IndexDesc()
      rc = L.scann_b200_assets_describe(h, C.byref(desc))
      if rc:
        raise RuntimeError("Error loading artifacts: " + L.scann_b200_last_error().decode("utf-8", "replace"))
      self._config_text = L.scann_b200_assets_config(h).decode()
      plan = _Plan(self._config_text)
      plan.check_supported()
      arrays = _arrays_from_desc(desc, plan)
      self._finish(arrays, plan)
    except _lib.ScannB200Error as e:
      raise _runtime("Error loading artifacts: ", e)
    finally:
      L.scann_b200_assets_free(h)

  # ---- queries (scann_npy.cc:209-270) ----
  def search(self, q, final_nn, pre_reorder_nn, leaves):
    q = np.asarray(q)
    if q.ndim != 1:
      raise ValueError("Query must be one-dimensional")
    idx, dist = self.search_batched(q[None, :], final_nn, pre_reorder_nn, leaves, False, 0)
    valid = ~np.isnan(dist[0])
    return idx[0][valid], dist[0][valid]

  def search_batched(self, queries, final_nn, pre_reorder_nn, leaves, parallel=False, batch_size=256):
    del parallel, batch_size  # one GPU path; the CPU library's thread fan-out has no analogue here
    queries = np.asarray(queries)
    if queries.ndim != 2:
      raise ValueError("Queries must be in two-dimensional array")
    if queries.shape[1] != self._d:
      raise RuntimeError("Error during search: Query doesn't match dataset dimsensionality")  # scann.cc:467-468
    try:
      return self._index.search_batched(queries, final_nn, pre_reorder_nn, leaves)
    except _lib.ScannB200Error as e:
      raise _runtime("Error during search: ", e)

  # ---- serialization (scann_npy.cc:272-282) ----
  def serialize(self, path, relative_path=False):
    a = self._arrays
    L = _lib.lib()
    keep = []

    def own(x, dt):
      if x is None:
        return None
      y = np.ascontiguousarray(x, dtype=dt)
      keep.append(y)
      return _lib.ptr(y)

    d = _lib.IndexDesc()
    d.distance = 0 if a.distance == "dot_product" else 1
    d.n, d.d = a.n, a.d
    if a.centers is not None:
      d.n_leaves, d.n_blocks, d.dims_per_block = a.centers.shape[0], a.codes.shape[1], a.codebook.shape[2]
    d.block_dims = own(a.block_dims, np.int32)
    d.centers = own(a.centers, np.float32)
    d.tokens = own(a.tokens, np.int32)
    d.soar = 1 if a.soar else 0
    d.codes = own(a.codes, np.uint8)
    d.soar_codes = own(a.soar_codes, np.uint8)
    d.codebook = own(a.codebook, np.float32)
    d.dataset = own(a.dataset, np.float32)
    d.bf16_dataset = own(a.bf16_dataset, np.int16)
    d.int8_dataset = own(a.int8_dataset, np.int8)
    d.int8_multipliers = own(a.int8_multipliers, np.float32)
    d.dp_norms = own(a.dp_norms, np.float32)
    d.overretrieve = a.overretrieve
    buf = C.create_string_buffer(1 << 16)
    rc = L.scann_b200_assets_save(path.encode(), C.byref(d), self._config_text.encode(), 1 if relative_path else 0,
                                  buf, len(buf))
    if rc:
      raise RuntimeError("Failed to extract SingleMachineFactoryOptions: " + L.scann_b200_last_error().decode())
    with open(os.path.join(path, "scann_assets.pbtxt"), "w") as f:
      f.write(buf.value.decode())

  # ---- misc surface ----
  def config(self):
    return self._config_text

  def size(self):
    return self._n

  def set_num_threads(self, num_threads):
    del num_threads

  def stats(self):
    return self._index.stats()

  def _unsupported(self, what):
    raise RuntimeError(f"{what} is not supported by the scann_b200 query path (out of scope, SURVEY.md section 8)")

  def upsert(self, *a, **k):
    self._unsupported("upsert")

  def delete(self, *a, **k):
    self._unsupported("delete")

  def rebalance(self, *a, **k):
    self._unsupported("rebalance")

  def reserve(self, *a, **k):
    self._unsupported("reserve")

  def get_health_stats(self):
    self._unsupported("get_health_stats")

  def initialize_health_stats(self):
    self._unsupported("initialize_health_stats")

  @staticmethod
  def suggest_autopilot(config, n, dim):
    raise RuntimeError("suggest_autopilot is not supported by the scann_b200 query path")


def _arrays_from_desc(desc, plan):
  """Copies the arrays a loaded-assets descriptor points at into an IndexArrays."""

  def arr(ptr, shape, dt):
    if not ptr:
      return None
    n = int(np.prod(shape))
    buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt).reshape(shape).copy()

  n, d, L, B, S = desc.n, desc.d, desc.n_leaves, desc.n_blocks, desc.dims_per_block
  a = index_build.IndexArrays(distance="dot_product" if desc.distance == 0 else "squared_l2", dataset=None, n=n, d=d)
  a.dataset = arr(desc.dataset, (n, d), np.float32)
  a.bf16_dataset = arr(desc.bf16_dataset, (n, d), np.int16)
  a.int8_dataset = arr(desc.int8_dataset, (n, d), np.int8)
  a.int8_multipliers = arr(desc.int8_multipliers, (d,), np.float32)
  a.dp_norms = arr(desc.dp_norms, (n,), np.float32)
  a.centers = arr(desc.centers, (L, d), np.float32)
  a.soar = bool(desc.soar)
  a.tokens = arr(desc.tokens, (n * (2 if a.soar else 1),), np.int32)
  a.codes = arr(desc.codes, (n, B), np.uint8)
  a.soar_codes = arr(desc.soar_codes, (n, B), np.uint8)
  a.codebook = arr(desc.codebook, (B, 16, S), np.float32)
  a.block_dims = arr(desc.block_dims, (B,), np.int32)
  a.overretrieve = float(desc.overretrieve)
  a.residual = True
  return a
