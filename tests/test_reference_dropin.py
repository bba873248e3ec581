"""Drop-in check against the reference's own, unmodified Python layer.

`/root/reference/scann/scann_ops/py/scann_ops_pybind.py:27` does `import scann_pybind` and drives a class
`ScannNumpy` (bound in scann_ops/cc/python/scann_pybind.cc:28-53).  Here that file is imported as it is, with
`scann_b200.scann_pybind` registered under the name it asks for, and

  * every attribute of the extension module the reference wrapper touches exists on our module / class,
  * `builder(...).tree().score_ah().reorder().build()`, `create_searcher`, `load_searcher` and the `ScannSearcher`
    methods reach our `ScannNumpy` with the arguments the pybind class would get (constructor overloads, the
    parallel / batch-size flags of `search_batched`, `serialize`),
  * the config the reference's builder hands over is, field by field, what our own builder mirror renders.

The reference tree only exists in the build container: the module skips itself elsewhere (the GPU box runs the same
API through tests/test_gpu_api.py).  No GPU work happens here -- the searcher behind ScannNumpy is a recorder.
"""
import ast
import importlib
import os
import sys
import types

import numpy as np
import pytest

REF_ROOT = "/root/reference"
REF_WRAPPER = os.path.join(REF_ROOT, "scann/scann_ops/py/scann_ops_pybind.py")

pytestmark = pytest.mark.skipif(not os.path.isfile(REF_WRAPPER), reason="reference tree not present")


@pytest.fixture()
def ref(monkeypatch):
  """The reference's scann_ops_pybind module bound to scann_b200.scann_pybind."""
  from scann_b200 import scann_pybind as ours
  monkeypatch.syspath_prepend(REF_ROOT)
  # generated protobuf module of the back-compat shim (scann_ops_pybind_backcompat.py:20); only touched when a
  # directory has no scann_assets.pbtxt, which these tests do not exercise
  pb2 = types.ModuleType("scann.scann_ops.scann_assets_pb2")
  monkeypatch.setitem(sys.modules, "scann.scann_ops.scann_assets_pb2", pb2)
  monkeypatch.setitem(sys.modules, "scann_pybind", ours)
  for name in [n for n in sys.modules if n == "scann" or n.startswith("scann.")]:
    if name != "scann.scann_ops.scann_assets_pb2":
      monkeypatch.delitem(sys.modules, name, raising=False)
  mod = importlib.import_module("scann.scann_ops.py.scann_ops_pybind")
  assert os.path.realpath(mod.__file__) == os.path.realpath(REF_WRAPPER)
  assert mod.scann_pybind is ours
  yield mod
  for name in [n for n in sys.modules if n == "scann" or n.startswith("scann.")]:
    sys.modules.pop(name, None)


class Recorder:
  """Stands in for the searcher behind ScannNumpy: records what the reference wrapper passes down."""
  calls = []

  def __init__(self, *args):
    Recorder.calls.append(("__init__", args))
    self.n = args[0].shape[0] if isinstance(args[0], np.ndarray) else 0

  def search(self, q, final_nn, pre_nn, leaves):
    Recorder.calls.append(("search", (q.shape, final_nn, pre_nn, leaves)))
    return np.arange(3, dtype=np.uint32), np.zeros(3, np.float32)

  def search_batched(self, q, final_nn, pre_nn, leaves, parallel, batch_size):
    Recorder.calls.append(("search_batched", (q.shape, final_nn, pre_nn, leaves, parallel, batch_size)))
    return np.zeros((q.shape[0], 3), np.uint32), np.zeros((q.shape[0], 3), np.float32)

  def serialize(self, artifacts_dir, relative_path):
    Recorder.calls.append(("serialize", (artifacts_dir, relative_path)))

  def size(self):
    return self.n

  def set_num_threads(self, n):
    Recorder.calls.append(("set_num_threads", (n,)))

  def config(self):
    return "cfg"


def _extension_attributes_used(path):
  """Names the reference wrapper reads off `scann_pybind` / `scann_pybind.ScannNumpy` / `self.searcher`."""
  tree = ast.parse(open(path).read())
  module_attrs, class_attrs, searcher_attrs = set(), set(), set()
  for node in ast.walk(tree):
    if not isinstance(node, ast.Attribute):
      continue
    v = node.value
    if isinstance(v, ast.Name) and v.id == "scann_pybind":
      module_attrs.add(node.attr)
    elif isinstance(v, ast.Attribute) and isinstance(v.value, ast.Name) and v.value.id == "scann_pybind":
      class_attrs.add(node.attr)
    elif isinstance(v, ast.Attribute) and v.attr == "searcher" and isinstance(v.value, ast.Name) and v.value.id == "self":
      searcher_attrs.add(node.attr)
  return module_attrs, class_attrs, searcher_attrs


def test_every_extension_attribute_the_reference_wrapper_uses_exists():
  from scann_b200 import scann_pybind as ours
  module_attrs, class_attrs, searcher_attrs = _extension_attributes_used(REF_WRAPPER)
  assert module_attrs == {"ScannNumpy"}
  assert searcher_attrs >= {"search", "search_batched", "serialize", "size", "config", "set_num_threads"}
  for name in class_attrs | searcher_attrs:
    assert callable(getattr(ours.ScannNumpy, name, None)), f"scann_pybind.ScannNumpy.{name} is missing"


def test_reference_builder_and_wrapper_drive_our_scann_numpy(ref, monkeypatch, tmp_path):
  from scann_b200 import scann_ops_pybind as ours_api
  monkeypatch.setattr(ref.scann_pybind, "ScannNumpy", Recorder)
  Recorder.calls = []
  db = np.random.default_rng(0).standard_normal((500, 32)).astype(np.float32)

  def chain(b):
    return (b.tree(num_leaves=20, num_leaves_to_search=5, training_sample_size=500, soar_lambda=1.5)
            .score_ah(2, anisotropic_quantization_threshold=0.2).reorder(40))

  s = chain(ref.builder(db, 10, "dot_product")).build()
  kind, args = Recorder.calls[0]
  assert kind == "__init__" and args[0] is db and args[2] == 0       # ScannNumpy(db, config, training_threads)
  # the reference's builder and our mirror render the same ScannConfig for the same calls (field by field: the two
  # differ in indentation only)
  from scann_b200 import config as cfgmod
  from test_api_cpu import normalize
  assert normalize(cfgmod.parse(args[1])) == normalize(cfgmod.parse(chain(ours_api.builder(db, 10, "dot_product")).create_config()))

  q = db[:7]
  s.search(q[0], final_num_neighbors=5)
  s.search_batched(q)
  s.search_batched(q, final_num_neighbors=4, pre_reorder_num_neighbors=30, leaves_to_search=9)
  s.search_batched_parallel(q, batch_size=128)
  s.set_num_threads(3)
  assert s.size() == 500 and s.config() == "cfg"
  s.serialize(str(tmp_path))
  assert Recorder.calls[1:] == [
      ("search", ((32,), 5, -1, -1)),
      ("search_batched", ((7, 32), -1, -1, -1, False, 0)),
      ("search_batched", ((7, 32), 4, 30, 9, False, 0)),
      ("search_batched", ((7, 32), -1, -1, -1, True, 128)),
      ("set_num_threads", (3,)),
      ("serialize", (str(tmp_path), False)),
  ]

  # load_searcher: ScannNumpy(artifacts_dir, <text of scann_assets.pbtxt>)
  (tmp_path / "scann_assets.pbtxt").write_text('assets { asset_type: AH_CENTERS asset_path: "x" }\n')
  Recorder.calls = []
  ref.load_searcher(str(tmp_path))
  assert Recorder.calls == [("__init__", (str(tmp_path), 'assets { asset_type: AH_CENTERS asset_path: "x" }\n'))]


def test_reference_wrapper_surfaces_our_error_without_a_device(ref):
  """No recorder here: the reference wrapper constructs the real ScannNumpy.  Without a CUDA device the product
  refuses loudly (there is no CPU fallback) with the RuntimeError prefix scann_npy.cc:62-76 uses."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("needs a host without a CUDA device")
  db = np.random.default_rng(1).standard_normal((300, 16)).astype(np.float32)
  with pytest.raises(RuntimeError, match="Error initializing searcher"):
    ref.builder(db, 5, "dot_product").tree(num_leaves=10, num_leaves_to_search=3, training_sample_size=300) \
        .score_ah(2).reorder(20).build()
