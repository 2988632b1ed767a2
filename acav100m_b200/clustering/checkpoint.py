"""Centroid checkpoints ``cache_epoch_{e}_{name}`` in both layouts the reference knows (run_clustering.py:55-116).

* **ver2** (``--clustering.save_scheme_ver2=True``, our default): a tree ``{model: {layer: KMeans.get_attrs()}}`` of
  plain dicts with numpy arrays -- unpickles anywhere.
* **ver1** (what the reference writes with its own defaults: its config has no ``save_scheme_ver2`` key, so
  ``torch.save`` pickles the ``KMeans`` OBJECTS, class path ``sgd_clustering.KMeans``).  Reading it needs that class;
  here a restricted unpickler maps ``sgd_clustering.KMeans`` to a stand-in that only keeps the instance ``__dict__``, so a
  reference checkpoint loads without the reference on ``sys.path``.  Writing it (``save_scheme_ver2=False``) pickles
  stand-in objects under the same class path with exactly the reference's attributes (CPU tensors), which the
  reference's ``torch.load`` + ``.to(device)`` accepts (run_clustering.py:93, sgd_clustering.py:59-61).
"""
import pickle
import sys
import types

import numpy as np
import torch

REFERENCE_MODULE = "sgd_clustering"
REFERENCE_ATTRS = ("args", "centers", "counts", "count", "lr", "initial_rounds", "reinit", "fallback", "sequential")


class _ReferenceKMeansState:
    """Stand-in for the reference's ``sgd_clustering.KMeans`` instances inside a ver1 checkpoint."""

    def get_attrs(self):
        d = {k: getattr(self, k) for k in REFERENCE_ATTRS if hasattr(self, k)}
        for k in ("centers", "counts"):
            if torch.is_tensor(d.get(k)):
                d[k] = d[k].detach().cpu().numpy()
        return d


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == REFERENCE_MODULE and name == "KMeans":
            return _ReferenceKMeansState
        if module.split(".")[0] in ("munch", "argparse") or name in ("Munch", "DefaultMunch", "Namespace"):
            return types.SimpleNamespace if name == "Namespace" else dict          # the pickled `args` tree is not used
        return super().find_class(module, name)


class _PickleModule:
    """What ``torch.load(pickle_module=...)`` needs: ``load`` and ``Unpickler`` (legacy and zip checkpoints)."""
    __name__ = "acav_checkpoint_pickle"
    Unpickler = _Unpickler

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()


def load_tree(path):
    """-> ``{model: {layer: attrs dict}}`` from a ver1 or ver2 checkpoint file."""
    tree = torch.load(str(path), pickle_module=_PickleModule, weights_only=False, map_location="cpu")
    out = {}
    for m, per in tree.items():
        out[m] = {}
        for layer, v in per.items():
            out[m][layer] = dict(v) if isinstance(v, dict) else v.get_attrs()
    return out


def save_tree_ver1(tree_attrs, path):
    """Write ``{model: {layer: attrs}}`` as the reference's default (object) layout."""
    mod = sys.modules.get(REFERENCE_MODULE)
    installed = mod is None
    if installed:
        mod = types.ModuleType(REFERENCE_MODULE)
        sys.modules[REFERENCE_MODULE] = mod
    had = getattr(mod, "KMeans", None)
    stand_in = type("KMeans", (), {"__module__": REFERENCE_MODULE})
    mod.KMeans = stand_in
    try:
        objs = {}
        for m, per in tree_attrs.items():
            objs[m] = {}
            for layer, attrs in per.items():
                o = stand_in()
                for k in REFERENCE_ATTRS:
                    if k in attrs:
                        v = attrs[k]
                        if k in ("centers", "counts"):
                            v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
                        setattr(o, k, v)
                objs[m][layer] = o
        torch.save(objs, str(path))
    finally:
        if installed:
            del sys.modules[REFERENCE_MODULE]
        elif had is not None:
            mod.KMeans = had
        else:
            del mod.KMeans
