"""CPU tests of the reference-facing surface: builder config, text/binary ScannConfig, asset files."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, load_golden
from scann_b200 import _lib, config as cfgmod, scann_builder, scann_ops_pybind


def normalize(m):
  """Msg -> comparable nested structure (bools/numbers canonicalised, field order ignored)."""
  out = []
  for name, v in m.fields:
    if isinstance(v, cfgmod.Msg):
      out.append((name, normalize(v)))
    else:
      s = str(v)
      if s in ("True", "true"):
        s = "true"
      elif s in ("False", "false"):
        s = "false"
      else:
        try:
          f = float(s)
          s = "nan" if math.isnan(f) else repr(f)
        except ValueError:
          pass
      out.append((name, s))
  return sorted(out, key=lambda kv: (kv[0], str(kv[1])))


def golden_configs():
  with open(os.path.join(GOLDEN_DIR, "builder_configs.json")) as f:
    return json.load(f)


@pytest.mark.parametrize("name", sorted(golden_configs()["chains"].keys()))
def test_builder_renders_the_reference_config(name):
  g = golden_configs()
  db = np.zeros((10, 20), np.float32)
  b = scann_builder.ScannBuilder(db, 10, g["dist"].get(name, "dot_product"))
  for meth, kw in g["chains"][name]:
    kw = {k: (float("nan") if v == "nan" else v) for k, v in kw.items()}
    if kw.get("quantize") == "BFLOAT16":
      kw["quantize"] = scann_builder.ReorderType.BFLOAT16
    getattr(b, meth)(**kw)
  ours = normalize(cfgmod.parse(b.create_config()))
  ref = normalize(cfgmod.parse(g["reference_text"][name]))
  assert ours == ref


def test_builder_errors_match_reference_behaviour():
  db = np.zeros((10, 20), np.float32)
  b = scann_builder.ScannBuilder(db, 10, "dot_product").tree(10, 2)
  with pytest.raises(Exception, match="tree has already been configured"):
    b.tree(10, 2)
  with pytest.raises(ValueError, match="Exactly 1 of score_ah or score_brute_force"):
    scann_builder.ScannBuilder(db, 10, "dot_product").tree(10, 2).create_config()
  with pytest.raises(ValueError, match="SOAR requires dot product"):
    scann_builder.ScannBuilder(db, 10, "squared_l2").tree(10, 2, soar_lambda=1.5)
  with pytest.raises(ValueError, match="distance_measure must be one of"):
    scann_builder.ScannBuilder(db, 10, "cosine").score_ah(2).create_config()
  with pytest.raises(Exception, match="no builder lambda"):
    scann_builder.ScannBuilder(db, 10, "dot_product").score_ah(2).build()


@pytest.mark.parametrize("name", sorted(golden_configs()["reference_text"].keys()))
def test_config_text_binary_round_trip_through_c_abi(name):
  text = golden_configs()["reference_text"][name]
  L = _lib.lib()
  n = C.c_size_t()
  buf = C.create_string_buffer(1 << 16)
  assert L.scann_b200_config_text_to_binary(text.encode(), buf, len(buf), C.byref(n)) == 0, L.scann_b200_last_error()
  out = C.create_string_buffer(1 << 16)
  assert L.scann_b200_config_binary_to_text(buf, n.value, out, len(out)) == 0, L.scann_b200_last_error()
  assert normalize(cfgmod.parse(out.value.decode())) == normalize(cfgmod.parse(text))


def test_binary_config_is_valid_protobuf_wire_format():
  """Decode the bytes with an independent reader (google.protobuf's wire decoder)."""
  from google.protobuf.internal import decoder
  text = golden_configs()["reference_text"]["tree_ah_soar"]
  L = _lib.lib()
  n = C.c_size_t()
  buf = C.create_string_buffer(1 << 16)
  assert L.scann_b200_config_text_to_binary(text.encode(), buf, len(buf), C.byref(n)) == 0
  data = buf.raw[:n.value]

  def walk(b):
    pos, fields = 0, {}
    while pos < len(b):
      tag, pos = decoder._DecodeVarint(b, pos)
      num, wt = tag >> 3, tag & 7
      if wt == 0:
        v, pos = decoder._DecodeVarint(b, pos)
      elif wt == 2:
        ln, pos = decoder._DecodeVarint(b, pos)
        v, pos = b[pos:pos + ln], pos + ln
      elif wt == 5:
        v, pos = np.frombuffer(b[pos:pos + 4], np.float32)[0], pos + 4
      elif wt == 1:
        v, pos = np.frombuffer(b[pos:pos + 8], np.float64)[0], pos + 8
      else:
        raise AssertionError(f"bad wire type {wt}")
      fields.setdefault(num, []).append(v)
    return fields

  top = walk(data)
  assert top[3] == [10]                                      # num_neighbors
  assert walk(top[5][0])[1] == [b"DotProductDistance"]       # distance_measure.distance_measure
  part = walk(top[8][0])
  assert part[3] == [50]                                     # num_children
  spill = walk(part[20][0])
  assert spill[1] == [4] and abs(spill[4][0] - 1.5) < 1e-6 and abs(spill[5][0] - 1.8) < 1e-6
  ah = walk(walk(top[13][0])[5][0])
  assert ah[20] == [3] and ah[22] == [1] and abs(ah[28][0] - 0.2) < 1e-12
  assert walk(top[17][0])[1] == [40]                         # exact_reordering.approx_num_neighbors


def test_text_proto_parser_accepts_builder_spellings():
  m = cfgmod.parse('a: True b { c: nan d {k: "v"} } e: < f: 1 > # comment\n g: -inf, h: \'x\\"y\'')
  assert cfgmod.as_bool(m.get("a")) and math.isnan(cfgmod.as_float(m.path("b", "c")))
  assert m.path("b", "d", "k") == "v" and m.path("e", "f") == "1"
  assert cfgmod.as_float(m.get("g")) == -math.inf and m.get("h") == 'x"y'
  with pytest.raises(cfgmod.TextProtoError):
    cfgmod.parse("a { b: 1")


@pytest.mark.parametrize("name,relative", [("dot_b16", True), ("dot_soar_b25", False), ("dot_varchunk_b11", True)])
def test_asset_files_round_trip(name, relative, tmp_path):
  a, z = load_golden(name)
  L = _lib.lib()
  keep = []

  def own(x, dt):
    y = np.ascontiguousarray(x, dtype=dt)
    keep.append(y)
    return _lib.ptr(y)

  d = _lib.IndexDesc()
  d.distance, d.n, d.d = 0, a.n, a.d
  d.n_leaves, d.n_blocks, d.dims_per_block = a.centers.shape[0], a.codes.shape[1], a.codebook.shape[2]
  d.block_dims, d.centers, d.tokens = own(a.block_dims, np.int32), own(a.centers, np.float32), own(a.tokens, np.int32)
  d.soar = 1 if a.soar else 0
  d.codes = own(a.codes, np.uint8)
  if a.soar:
    d.soar_codes = own(a.soar_codes, np.uint8)
  d.codebook, d.dataset = own(a.codebook, np.float32), own(a.dataset, np.float32)
  b = scann_builder.ScannBuilder(a.dataset, 10, "dot_product").tree(
      a.centers.shape[0], int(z["probe"]), soar_lambda=1.5 if a.soar else None).score_ah(
          int(a.codebook.shape[2])).reorder(int(z["pre"]))
  cfg = b.create_config()
  buf = C.create_string_buffer(1 << 16)
  rc = L.scann_b200_assets_save(str(tmp_path).encode(), C.byref(d), cfg.encode(), 1 if relative else 0, buf, len(buf))
  assert rc == 0, L.scann_b200_last_error()
  manifest = buf.value.decode()
  # files are the reference's: names, npy dtypes/shapes (scann.cc:504-601, io_npy.h:39-73)
  assert np.array_equal(np.load(tmp_path / "hashed_dataset.npy"), a.codes)
  assert np.load(tmp_path / "hashed_dataset.npy").shape[1] == a.codes.shape[1]       # one byte per AH block
  assert np.array_equal(np.load(tmp_path / "datapoint_to_token.npy"), a.tokens)
  assert np.load(tmp_path / "datapoint_to_token.npy").dtype == np.int32
  assert np.array_equal(np.load(tmp_path / "dataset.npy"), a.dataset)
  assert (tmp_path / "hashed_dataset_soar.npy").exists() == a.soar
  for f in ("scann_config.pb", "ah_codebook.pb", "serialized_partitioner.pb"):
    assert (tmp_path / f).stat().st_size > 0
  with open(tmp_path / "dataset.npy", "rb") as f:
    head = f.read(10)
    assert head[:8] == b"\x93NUMPY\x01\x00" and (10 + int.from_bytes(head[8:10], "little")) % 64 == 0
  assert ("asset_path: \"dataset.npy\"" in manifest) == relative
  # load back
  h = C.c_void_p()
  assert L.scann_b200_assets_load(str(tmp_path).encode(), manifest.encode(), C.byref(h)) == 0, L.scann_b200_last_error()
  d2 = _lib.IndexDesc()
  assert L.scann_b200_assets_describe(h, C.byref(d2)) == 0
  from scann_b200.scann_pybind import _Plan, _arrays_from_desc
  plan = _Plan(L.scann_b200_assets_config(h).decode())
  b2 = _arrays_from_desc(d2, plan)
  L.scann_b200_assets_free(h)
  assert (d2.n, d2.d, d2.n_leaves, d2.n_blocks, d2.soar) == (a.n, a.d, d.n_leaves, d.n_blocks, d.soar)
  assert d2.default_leaves == int(z["probe"]) and d2.default_pre_nn == int(z["pre"]) and d2.default_final_nn == 10
  np.testing.assert_array_equal(b2.centers, a.centers)       # f32 -> double -> f32 is exact
  np.testing.assert_array_equal(b2.codebook, a.codebook)
  np.testing.assert_array_equal(b2.block_dims, a.block_dims)
  np.testing.assert_array_equal(b2.tokens, a.tokens)
  np.testing.assert_array_equal(b2.codes, a.codes)
  np.testing.assert_array_equal(b2.dataset, a.dataset)
  if a.soar:
    np.testing.assert_array_equal(b2.soar_codes, a.soar_codes)
  assert normalize(plan.msg) == normalize(cfgmod.parse(cfg))


def test_load_errors_are_reported_not_fatal(tmp_path):
  L = _lib.lib()
  h = C.c_void_p()
  rc = L.scann_b200_assets_load(str(tmp_path).encode(), b"", C.byref(h))
  assert rc != 0 and b"scann_config.pb" in L.scann_b200_last_error()
  with pytest.raises(ValueError, match="is not a directory"):
    scann_ops_pybind.load_searcher(str(tmp_path / "missing"))
  with pytest.raises(ValueError, match="No scann_assets.pbtxt"):
    scann_ops_pybind.load_searcher(str(tmp_path))


def test_int8_reordering_assets_round_trip(tmp_path):
  """int8_dataset.npy / int8_multipliers.npy / dp_norms.npy (scann.cc:568-593): names, dtypes, manifest, load."""
  from scann_b200 import index_build
  a, z = load_golden("l2_b16")
  q8, mult = index_build.int8_quantize(a.dataset)
  norms = index_build.squared_l2_norms(a.dataset)
  L = _lib.lib()
  keep = []

  def own(x, dt):
    y = np.ascontiguousarray(x, dtype=dt)
    keep.append(y)
    return _lib.ptr(y)

  d = _lib.IndexDesc()
  d.distance, d.n, d.d = 1, a.n, a.d
  d.n_leaves, d.n_blocks, d.dims_per_block = a.centers.shape[0], a.codes.shape[1], a.codebook.shape[2]
  d.block_dims, d.centers, d.tokens = own(a.block_dims, np.int32), own(a.centers, np.float32), own(a.tokens, np.int32)
  d.codes, d.codebook = own(a.codes, np.uint8), own(a.codebook, np.float32)
  d.int8_dataset, d.int8_multipliers, d.dp_norms = own(q8, np.int8), own(mult, np.float32), own(norms, np.float32)
  cfg = scann_builder.ScannBuilder(a.dataset, 10, "squared_l2").tree(a.centers.shape[0], 5).score_ah(2).reorder(
      50, quantize=scann_builder.ReorderType.INT8).create_config()
  assert "fixed_point" in cfg
  buf = C.create_string_buffer(1 << 16)
  assert L.scann_b200_assets_save(str(tmp_path).encode(), C.byref(d), cfg.encode(), 1, buf, len(buf)) == 0, L.scann_b200_last_error()
  manifest = buf.value.decode()
  for t in ("INT8_DATASET_NPY", "INT8_MULTIPLIERS_NPY", "INT8_NORMS_NPY"):
    assert t in manifest
  assert not (tmp_path / "dataset.npy").exists()
  f8 = np.load(tmp_path / "int8_dataset.npy")
  assert f8.dtype == np.int8 and np.array_equal(f8, q8)
  assert np.array_equal(np.load(tmp_path / "int8_multipliers.npy"), mult)
  assert np.array_equal(np.load(tmp_path / "dp_norms.npy"), norms)
  h = C.c_void_p()
  assert L.scann_b200_assets_load(str(tmp_path).encode(), manifest.encode(), C.byref(h)) == 0, L.scann_b200_last_error()
  d2 = _lib.IndexDesc()
  assert L.scann_b200_assets_describe(h, C.byref(d2)) == 0
  from scann_b200.scann_pybind import _Plan, _arrays_from_desc
  plan = _Plan(L.scann_b200_assets_config(h).decode())
  assert plan.int8_reorder()
  b2 = _arrays_from_desc(d2, plan)
  L.scann_b200_assets_free(h)
  assert (d2.n, d2.d) == (a.n, a.d) and b2.dataset is None
  np.testing.assert_array_equal(b2.int8_dataset, q8)
  np.testing.assert_array_equal(b2.int8_multipliers, mult)
  np.testing.assert_array_equal(b2.dp_norms, norms)


def test_plan_reads_reordering_and_noise_shaping_options():
  """Host logic of ScannNumpy(db, config, threads): which searcher / reordering helper / encoder the config selects
  (base/reordering_helper_factory.cc:106-200, tree_ah_hybrid_residual.cc:414-428)."""
  from scann_b200.scann_pybind import _Plan
  import math
  db = np.zeros((10, 8), np.float32)
  R = scann_builder.ReorderType
  mk = lambda **kw: scann_builder.ScannBuilder(db, 5, kw.pop("dist", "dot_product")).tree(4, 2).score_ah(2, **kw)
  p = _Plan(mk(anisotropic_quantization_threshold=0.2).reorder(30).create_config())
  p.check_supported()
  assert not p.int8_reorder() and not p.bf16_reorder() and not p.is_brute_force()
  assert abs(cfgmod.as_float(p.ah.get("noise_shaping_threshold"), math.nan) - 0.2) < 1e-12
  p = _Plan(mk().reorder(30, quantize=R.INT8).create_config())
  p.check_supported()
  assert p.int8_reorder() and not p.bf16_reorder()
  assert math.isnan(cfgmod.as_float(p.ah.get("noise_shaping_threshold"), math.nan))
  p = _Plan(mk(dist="squared_l2").reorder(30, quantize=R.BFLOAT16).create_config())
  p.check_supported()
  assert p.bf16_reorder() and not p.int8_reorder() and p.distance == "squared_l2"
  # noise-shaped quantization of the REORDERING rows is not built: rejected loudly, not ignored
  with pytest.raises(Exception, match="noise-shaped"):
    _Plan(mk().reorder(30, quantize=R.INT8, anisotropic_quantization_threshold=0.2).create_config()).check_supported()
  bf = _Plan(scann_builder.ScannBuilder(db, 5, "dot_product").score_brute_force(R.BFLOAT16).create_config())
  assert bf.is_brute_force() and bf.bf16_brute_force()


def test_int8_quantizer_semantics():
  """ScalarQuantizeFloatDataset at quantile 1.0 (utils/scalar_quantization_helpers.cc:39-63, .h:40-50)."""
  from scann_b200 import index_build
  x = np.array([[0.5, -2.0, 0.0, 1.0], [-0.25, 1.0, 0.0, 0.996], [0.0019685, 0.0, 0.0, -1.0]], np.float32)
  q, m = index_build.int8_quantize(x)
  np.testing.assert_array_equal(m, np.array([254.0, 63.5, 1.0, 127.0], np.float32))   # 127 / max|column|, 1 for a zero column
  # 0.5 * 254 = 127; -0.25 * 254 = -63.5 -> -64 (round half away from zero); 0.0019685 * 254 = 0.49999 -> 0
  np.testing.assert_array_equal(q[:, 0], np.array([127, -64, 0], np.int8))
  np.testing.assert_array_equal(q[:, 1], np.array([-127, 64, 0], np.int8))             # 1.0 * 63.5 = 63.5 -> 64
  np.testing.assert_array_equal(q[:, 2], np.zeros(3, np.int8))
  np.testing.assert_array_equal(q[:, 3], np.array([127, 126, -127], np.int8))          # 0.996 * 127 = 126.49 -> 126
  n = index_build.squared_l2_norms(x)
  np.testing.assert_allclose(n, (x.astype(np.float64) ** 2).sum(1), rtol=1e-7)


def test_quantized_centroids_config_reaches_the_index_descriptor(tmp_path):
  """tree(quantize_centroids=True) -> `query_tokenization_type: FIXED_POINT_INT8` (scann_builder.py:231) survives
  serialization (binary scann_config.pb) and comes back in scann_b200_index_desc.query_tokenization_type; nothing extra
  is written -- the fixed-point centres are derived at load time, as KMeansTreeNode::CreateFixedPointCenters does."""
  from scann_b200.scann_pybind import _Plan
  a, z = load_golden("dot_b16")
  L = _lib.lib()
  keep = []

  def own(x, dt):
    y = np.ascontiguousarray(x, dtype=dt)
    keep.append(y)
    return _lib.ptr(y)

  for quantized in (False, True):
    d = _lib.IndexDesc()
    d.distance, d.n, d.d = 0, a.n, a.d
    d.n_leaves, d.n_blocks, d.dims_per_block = a.centers.shape[0], a.codes.shape[1], a.codebook.shape[2]
    d.block_dims, d.centers, d.tokens = own(a.block_dims, np.int32), own(a.centers, np.float32), own(a.tokens, np.int32)
    d.codes, d.codebook, d.dataset = own(a.codes, np.uint8), own(a.codebook, np.float32), own(a.dataset, np.float32)
    cfg = scann_builder.ScannBuilder(a.dataset, 10, "dot_product").tree(
        a.centers.shape[0], 4, quantize_centroids=quantized).score_ah(2).reorder(40).create_config()
    assert ("FIXED_POINT_INT8" in cfg) == quantized
    out = tmp_path / ("q" if quantized else "f")
    out.mkdir()
    buf = C.create_string_buffer(1 << 16)
    assert L.scann_b200_assets_save(str(out).encode(), C.byref(d), cfg.encode(), 0, buf, len(buf)) == 0, L.scann_b200_last_error()
    assert sorted(p.name for p in out.iterdir()) == sorted(
        ["scann_config.pb", "serialized_partitioner.pb", "ah_codebook.pb", "datapoint_to_token.npy",
         "hashed_dataset.npy", "dataset.npy"])  # (the manifest text is returned to the caller, who writes it)
    h = C.c_void_p()
    assert L.scann_b200_assets_load(str(out).encode(), buf.value, C.byref(h)) == 0, L.scann_b200_last_error()
    d2 = _lib.IndexDesc()
    assert L.scann_b200_assets_describe(h, C.byref(d2)) == 0
    plan = _Plan(L.scann_b200_assets_config(h).decode())
    L.scann_b200_assets_free(h)
    assert d2.query_tokenization_type == (1 if quantized else 0)
    plan.check_supported()
    assert plan.int8_tokenization() == quantized


def test_plan_refuses_what_is_not_built_around_the_tree():
  """Two-level trees (upper_tree) and a quantized DATABASE tokenization are refused, not ignored."""
  from scann_b200.scann_pybind import _Plan
  db = np.zeros((10, 8), np.float32)
  ok = scann_builder.ScannBuilder(db, 5, "dot_product").tree(4, 2, quantize_centroids=True).score_ah(2).create_config()
  _Plan(ok).check_supported()
  with pytest.raises(Exception, match="upper_tree"):
    _Plan(ok.replace("query_tokenization_type: FIXED_POINT_INT8",
                     "query_tokenization_type: FIXED_POINT_INT8\n bottom_up_top_level_partitioner { enabled: true num_centroids: 2 }")).check_supported()
  with pytest.raises(Exception, match="database_tokenization_type"):
    _Plan(ok.replace("query_tokenization_type: FIXED_POINT_INT8",
                     "query_tokenization_type: FIXED_POINT_INT8\n database_tokenization_type: FIXED_POINT_INT8")).check_supported()
