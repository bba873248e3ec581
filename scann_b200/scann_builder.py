"""ScannBuilder: the `builder(db, k, dist).tree(...).score_ah(...).reorder(...).build()` surface.

API-compatible with the reference's config generator (scann/scann_ops/py/scann_builder.py:57-469):
same method names, argument names, defaults and error messages, and the emitted text proto uses
the same ScannConfig fields, so either library's `create_config()` output can be fed to the other.
The implementation is independent: stages are recorded as plain dicts and rendered through
`scann_b200.config.emit` instead of f-string templates.
"""
import enum
import math

from .config import Enum, emit


class ReorderType(enum.Enum):
  FLOAT32 = 1
  INT8 = 2
  BFLOAT16 = 3


class IncrementalMode(enum.Enum):
  NONE = 1
  ONLINE = 2
  ONLINE_INCREMENTAL = 3


_DISTANCES = {"dot_product": "DotProductDistance", "squared_l2": "SquaredL2Distance"}


def _norm_quantize(q):
  if q is True:
    return ReorderType.INT8
  if q is False:
    return ReorderType.FLOAT32
  return q


class ScannBuilder(object):
  """Collects the stages of a searcher and renders the ScannConfig text proto."""

  def __init__(self, db, num_neighbors, distance_measure):
    self.params = {}
    self.training_threads = 0
    self.builder_lambda = None
    self.db = db
    self.num_neighbors = num_neighbors
    self.distance_measure = distance_measure

  # ---- plumbing ----
  def _stage(self, key, **kwargs):
    if key in self.params:
      raise Exception(f"{key} has already been configured")
    self.params[key] = kwargs
    return self

  def set_n_training_threads(self, threads):
    self.training_threads = threads
    return self

  def set_builder_lambda(self, builder_lambda):
    """builder_lambda(db, config_text, training_threads, **kwargs) -> searcher."""
    self.builder_lambda = builder_lambda
    return self

  # ---- stages (signatures follow scann_builder.py:107-383) ----
  def pca(self, reduction_dim=None, pca_significance_threshold=0.80, pca_truncation_threshold=0.6):
    if (reduction_dim is None) == (pca_significance_threshold is None):
      raise ValueError("pca must be called with either reduction_dim or pca_significance_threshold")
    return self._stage("pca", reduction_dim=reduction_dim, pca_significance_threshold=pca_significance_threshold,
                       pca_truncation_threshold=pca_truncation_threshold)

  def truncate(self, reduction_dim):
    if reduction_dim >= self.db.shape[1]:
      raise ValueError(f"reduction_dim must be less than {self.db.shape[1]}")
    return self._stage("truncate", reduction_dim=reduction_dim)

  def upper_tree(self, num_leaves, num_leaves_to_search, avq=float("nan"), soar_lambda=None,
                 overretrieve_factor=None, scoring_mode=ReorderType.INT8,
                 anisotropic_quantization_threshold=float("nan")):
    return self._stage("upper_tree", num_leaves=num_leaves, num_leaves_to_search=num_leaves_to_search, avq=avq,
                       soar_lambda=soar_lambda, overretrieve_factor=overretrieve_factor, scoring_mode=scoring_mode,
                       anisotropic_quantization_threshold=anisotropic_quantization_threshold)

  def tree(self, num_leaves, num_leaves_to_search, training_sample_size=100000, min_partition_size=50,
           training_iterations=12, spherical=False, quantize_centroids=False, random_init=True,
           incremental_threshold=None, avq=None, soar_lambda=None, overretrieve_factor=None):
    if avq is not None and self.distance_measure != "dot_product":
      raise ValueError("AVQ only applies to dot product distance.")
    if soar_lambda is not None and self.distance_measure != "dot_product":
      raise ValueError("SOAR requires dot product distance.")
    return self._stage("tree", num_leaves=num_leaves, num_leaves_to_search=num_leaves_to_search,
                       training_sample_size=training_sample_size, min_partition_size=min_partition_size,
                       training_iterations=training_iterations, spherical=spherical,
                       quantize_centroids=quantize_centroids, random_init=random_init,
                       incremental_threshold=incremental_threshold, avq=avq, soar_lambda=soar_lambda,
                       overretrieve_factor=overretrieve_factor)

  def score_ah(self, dimensions_per_block, anisotropic_quantization_threshold=float("nan"),
               training_sample_size=100000, min_cluster_size=100, hash_type="lut16", training_iterations=10,
               residual_quantization=None):
    del min_cluster_size  # deprecated in the reference as well
    if hash_type not in ("lut16", "lut256"):
      raise ValueError("hash_type must be one of ['lut16', 'lut256']")
    kw = dict(dimensions_per_block=dimensions_per_block,
              anisotropic_quantization_threshold=anisotropic_quantization_threshold,
              training_sample_size=training_sample_size, hash_type=hash_type,
              training_iterations=training_iterations)
    if residual_quantization is not None:
      kw["residual_quantization"] = residual_quantization
    return self._stage("score_ah", **kw)

  def score_brute_force(self, quantize=ReorderType.FLOAT32):
    return self._stage("score_bf", quantize=_norm_quantize(quantize))

  def reorder(self, reordering_num_neighbors, quantize=ReorderType.FLOAT32,
              anisotropic_quantization_threshold=float("nan")):
    return self._stage("reorder", reordering_num_neighbors=reordering_num_neighbors,
                       quantize=_norm_quantize(quantize),
                       anisotropic_quantization_threshold=anisotropic_quantization_threshold)

  def autopilot(self, mode=IncrementalMode.NONE, quantize=ReorderType.FLOAT32):
    return self._stage("autopilot", mode=mode, quantize=quantize)

  # ---- rendering ----
  def _projection(self):
    dim = self.db.shape[1]
    pca, trunc = self.params.get("pca"), self.params.get("truncate")
    if pca is not None and trunc is not None:
      raise ValueError("Exactly 1 of pca or truncate must be set")
    if pca is not None:
      body = [("projection_type", Enum("PCA")), ("input_dim", dim)]
      if pca["reduction_dim"] is not None:
        body.append(("num_dims_per_block", pca["reduction_dim"]))
      else:
        body += [("pca_significance_threshold", float(pca["pca_significance_threshold"])),
                 ("pca_truncation_threshold", float(pca["pca_truncation_threshold"]))]
      return body
    if trunc is not None:
      return [("projection_type", Enum("TRUNCATE")), ("num_dims_per_block", trunc["reduction_dim"]),
              ("input_dim", dim)]
    return None

  def _quantized_stanza(self, quantize, threshold=None):
    name = "bfloat16" if quantize == ReorderType.BFLOAT16 else "fixed_point"
    body = [("enabled", quantize != ReorderType.FLOAT32)]
    if threshold is not None:
      body.append(("noise_shaping_threshold", float(threshold)))
    return (name, body)

  def _render_tree(self, t, distance_cfg, projection):
    part = [
        ("num_children", t["num_leaves"]),
        ("min_cluster_size", t["min_partition_size"]),
        ("max_clustering_iterations", t["training_iterations"]),
        ("single_machine_center_initialization",
         Enum("RANDOM_INITIALIZATION" if t["random_init"] else "DEFAULT_KMEANS_PLUS_PLUS")),
        ("partitioning_distance", [("distance_measure", "SquaredL2Distance")]),
        ("query_spilling", [("spilling_type", Enum("FIXED_NUMBER_OF_CENTERS")),
                            ("max_spill_centers", t["num_leaves_to_search"])]),
        ("expected_sample_size", t["training_sample_size"]),
        ("query_tokenization_distance_override", distance_cfg),
        ("partitioning_type", Enum("SPHERICAL" if t["spherical"] else "GENERIC")),
        ("query_tokenization_type", Enum("FIXED_POINT_INT8" if t["quantize_centroids"] else "FLOAT")),
    ]
    inc = t["incremental_threshold"]
    if isinstance(inc, bool):
      inc = None
    if isinstance(inc, int):
      part.append(("incremental_training_config", [("number_of_datapoints", inc)]))
    elif isinstance(inc, float):
      part.append(("incremental_training_config", [("fraction", inc)]))
    if t["avq"] is not None:
      part.append(("avq", float(t["avq"])))
    if t["soar_lambda"] is not None:
      soar = [("spilling_type", Enum("TWO_CENTER_ORTHOGONALITY_AMPLIFIED")),
              ("orthogonality_amplification_lambda", float(t["soar_lambda"]))]
      if t["overretrieve_factor"] is not None:
        soar.append(("overretrieve_factor", float(t["overretrieve_factor"])))
      part.append(("database_spilling", soar))
    if projection is not None:
      part.append(("projection", projection))
    up = self.params.get("upper_tree")
    if up is not None:
      mode = {ReorderType.INT8: "FIXED8", ReorderType.BFLOAT16: "BFLOAT16", ReorderType.FLOAT32: "FLOAT32"}[
          up["scoring_mode"]]
      part.append(("bottom_up_top_level_partitioner", [
          ("enabled", True), ("num_centroids", up["num_leaves"]),
          ("num_centroids_to_search", up["num_leaves_to_search"]), ("avq", float(up["avq"])),
          ("soar", [("enabled", up["soar_lambda"] is not None), ("lambda", float(up["soar_lambda"] or 1.5)),
                    ("overretrieve_factor", float(up["overretrieve_factor"] or 2.0))]),
          ("quantization", Enum(mode)),
          ("noise_shaping_threshold", float(up["anisotropic_quantization_threshold"]))]))
    return ("partitioning", part)

  def _render_ah(self, ah, projection):
    n_dims = self.db.shape[1]
    dpb = ah["dimensions_per_block"]
    lut16 = ah["hash_type"] == "lut16"
    full_blocks, partial = divmod(n_dims, dpb)
    if projection is not None:
      proj = [("projection_type", Enum("CHUNK")), ("num_dims_per_block", dpb)]
    elif partial == 0:
      proj = [("input_dim", n_dims), ("projection_type", Enum("CHUNK")), ("num_blocks", full_blocks),
              ("num_dims_per_block", dpb)]
    else:
      proj = [("input_dim", n_dims), ("projection_type", Enum("VARIABLE_CHUNK")),
              ("variable_blocks", [("num_blocks", full_blocks), ("num_dims_per_block", dpb)]),
              ("variable_blocks", [("num_blocks", 1), ("num_dims_per_block", partial)])]
    residual = bool(ah["residual_quantization"])
    global_topn = bool(lut16 and (full_blocks + (partial > 0)) <= 256 and residual)
    body = [
        ("lookup_type", Enum("INT8_LUT16" if lut16 else "INT8")),
        ("use_residual_quantization", residual),
        ("use_global_topn", global_topn),
        ("quantization_distance", [("distance_measure", "SquaredL2Distance")]),
        ("num_clusters_per_block", 16 if lut16 else 256),
        ("projection", proj),
        ("fixed_point_lut_conversion_options", [("float_to_int_conversion_method", Enum("ROUND"))]),
        ("noise_shaping_threshold", float(ah["anisotropic_quantization_threshold"])),
        ("expected_sample_size", ah["training_sample_size"]),
        ("max_clustering_iterations", ah["training_iterations"]),
    ]
    return ("hash", [("asymmetric_hash", body)])

  def create_config(self):
    """Returns the text-format ScannConfig for the configured stages."""
    dist_name = _DISTANCES.get(self.distance_measure)
    if dist_name is None:
      raise ValueError(f"distance_measure must be one of {list(_DISTANCES.keys())}")
    distance_cfg = [("distance_measure", dist_name)]
    cfg = [("num_neighbors", self.num_neighbors), ("distance_measure", distance_cfg)]
    ap = self.params.get("autopilot")
    if ap is not None:
      cfg.append(("autopilot", [("tree_ah", [("incremental_mode", Enum(ap["mode"].name)),
                                             ("reordering_dtype", Enum(ap["quantize"].name))])]))
      return emit(cfg) + "\n"
    projection = self._projection()
    t = self.params.get("tree")
    if t is not None:
      cfg.append(self._render_tree(t, distance_cfg, projection))
    ah, bf = self.params.get("score_ah"), self.params.get("score_bf")
    if (ah is None) == (bf is None):
      raise ValueError("Exactly 1 of score_ah or score_brute_force must be set")
    if ah is not None:
      ah = dict(ah)
      if "residual_quantization" not in ah:
        ah["residual_quantization"] = t is not None and self.distance_measure == "dot_product"
      cfg.append(self._render_ah(ah, projection))
    else:
      cfg.append(("brute_force", [self._quantized_stanza(bf["quantize"])]))
    r = self.params.get("reorder")
    if r is not None:
      cfg.append(("exact_reordering", [("approx_num_neighbors", r["reordering_num_neighbors"]),
                                       self._quantized_stanza(r["quantize"],
                                                              r["anisotropic_quantization_threshold"])]))
    return emit(cfg) + "\n"

  def build(self, docids=None, **kwargs):
    """Calls builder_lambda(db, config, training_threads, docids=..., **kwargs)."""
    if self.builder_lambda is None:
      raise Exception("build() called but no builder lambda was set.")
    config = self.create_config()
    return self.builder_lambda(self.db, config, self.training_threads, docids=docids, **kwargs)
