"""Python surface of the searcher: `builder`, `create_searcher`, `load_searcher`, `ScannSearcher`.

Same names, arguments and behaviour as scann/scann_ops/py/scann_ops_pybind.py:38-273, so user
code switches with `from scann_b200 import scann_ops_pybind` instead of
`from scann.scann_ops.py import scann_ops_pybind`.
"""
import os
import pickle

import numpy as np

from . import scann_builder, scann_pybind


class ScannSearcher(object):
  """Thin convenience wrapper around a `ScannNumpy` (docid mapping, defaults as None)."""

  def __init__(self, searcher, docids=None):
    self.searcher = searcher
    self.docids = docids
    if docids is not None:
      self.docid_to_id = {docid: i for i, docid in enumerate(docids)}
      if len(docids) != len(self.docid_to_id):
        raise ValueError("Duplicates found in docids.")

  @staticmethod
  def _dflt(v):
    return -1 if v is None else v

  def search(self, q, final_num_neighbors=-1, pre_reorder_num_neighbors=-1, leaves_to_search=-1):
    """Single query; -1 keeps the searcher's default for that parameter."""
    idx, dist = self.searcher.search(q, final_num_neighbors, pre_reorder_num_neighbors, leaves_to_search)
    if self.docids is not None:
      idx = [self.docids[j] for j in idx]
    return idx, dist

  def _batched(self, queries, final_num_neighbors, pre_reorder_num_neighbors, leaves_to_search, parallel, batch_size):
    idx, dist = self.searcher.search_batched(queries, self._dflt(final_num_neighbors),
                                             self._dflt(pre_reorder_num_neighbors), self._dflt(leaves_to_search),
                                             parallel, batch_size)
    if self.docids is not None:
      idx = [[self.docids[j] for j in row] for row in idx]
    return idx, dist

  def search_batched(self, queries, final_num_neighbors=None, pre_reorder_num_neighbors=None,
                     leaves_to_search=None):
    return self._batched(queries, final_num_neighbors, pre_reorder_num_neighbors, leaves_to_search, False, 0)

  def search_batched_parallel(self, queries, final_num_neighbors=None, pre_reorder_num_neighbors=None,
                              leaves_to_search=None, batch_size=256):
    return self._batched(queries, final_num_neighbors, pre_reorder_num_neighbors, leaves_to_search, True,
                         batch_size)

  def serialize(self, artifacts_dir, relative_path=False):
    self.searcher.serialize(artifacts_dir, relative_path)
    if self.docids is not None:
      with open(os.path.join(artifacts_dir, "scann_docids.pkl"), "wb") as f:
        pickle.dump(self.docids, f)

  def get_health_stats(self):
    return self.searcher.get_health_stats()

  def initialize_health_stats(self):
    return self.searcher.initialize_health_stats()

  def upsert(self, docids, database, batch_size=1):
    return self.searcher.upsert(docids, database, batch_size)

  def delete(self, docids):
    return self.searcher.delete(docids)

  def rebalance(self, config=None):
    return self.searcher.rebalance("" if config is None else config)

  def reserve(self, num_datapoints):
    return self.searcher.reserve(num_datapoints)

  def size(self):
    return self.searcher.size()

  def set_num_threads(self, num_threads):
    self.searcher.set_num_threads(num_threads)

  def config(self):
    return self.searcher.config()


def builder(db, num_neighbors, distance_measure):
  """`builder(db, k, "dot_product").tree(...).score_ah(...).reorder(...).build()`."""

  def builder_lambda(db, config, training_threads, **kwargs):
    return create_searcher(db, config, training_threads, **kwargs)

  return scann_builder.ScannBuilder(db, num_neighbors, distance_measure).set_builder_lambda(builder_lambda)


def create_searcher(db, scann_config, training_threads=0, docids=None, **unused_kwargs):
  if docids is not None and len(docids) != db.shape[0]:
    raise ValueError(f"docid and database size mismatch: {len(docids)} != {db.shape[0]}.")
  return ScannSearcher(scann_pybind.ScannNumpy(db, scann_config, training_threads), docids=docids)


def load_searcher(artifacts_dir, assets_backcompat_shim=True):
  """Loads the assets written by `ScannSearcher.serialize` (either library's)."""
  del assets_backcompat_shim  # pre-1.2 asset directories without a manifest are not supported
  if not os.path.isdir(artifacts_dir):
    raise ValueError(f"{artifacts_dir} is not a directory.")
  assets_pbtxt = os.path.join(artifacts_dir, "scann_assets.pbtxt")
  if not os.path.isfile(assets_pbtxt):
    raise ValueError("No scann_assets.pbtxt found.")
  docids_path = os.path.join(artifacts_dir, "scann_docids.pkl")
  docids = None
  if os.path.isfile(docids_path):
    with open(docids_path, "rb") as f:
      docids = pickle.load(f)
  with open(assets_pbtxt, "r") as f:
    return ScannSearcher(scann_pybind.ScannNumpy(artifacts_dir, f.read()), docids)
