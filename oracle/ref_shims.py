"""Import the UNMODIFIED reference operators from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by ``oracle/gen_golden.py`` and by the
``-m "not gpu"`` tests that re-validate the restatements when ``/root/reference`` is mounted.  The GPU
box has no ``/root/reference``; nothing on a GPU code path may call this module.

The reference's k-means file needs three third-party modules that are not installed here
(SURVEY.md section 8c).  Each shim below is the smallest stand-in that lets
``clustering/code/sgd_clustering.py`` run on CPU; none of them changes the arithmetic:

1. ``torch_scatter.scatter_add``  (torch-scatter==2.0.5, pinned at reference README.md:48, not
   vendored): upstream ``scatter_sum`` broadcasts ``index`` to ``src`` and calls
   ``out.scatter_add_(dim, index, src)``.  The shim does exactly that.
2. ``mps.distributed``  (reference file imports the absent ``diffdist``): identity
   ``all_reduce`` / ``all_gather`` -- single-process semantics.
3. the hard-coded ``.cuda()`` at ``clustering/code/sgd_clustering.py:113`` is neutralised by making
   ``Tensor.cuda`` the identity while the reference runs (restored afterwards).
"""
import contextlib
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("ACAV_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "clustering/code/sgd_clustering.py"))


def _broadcast_index(src, index, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand_as(src)


def _scatter_add(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast_index(src, index, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def _install_kmeans_shims():
    if "torch_scatter" not in sys.modules:
        mod = types.ModuleType("torch_scatter")
        mod.scatter_add = _scatter_add
        sys.modules["torch_scatter"] = mod
    if "mps" not in sys.modules:
        pkg = types.ModuleType("mps")
        pkg.__path__ = []
        dist = types.ModuleType("mps.distributed")
        dist.all_reduce = lambda tensors, average=True: tensors
        dist.all_gather = lambda tensors: tensors
        pkg.distributed = dist
        sys.modules["mps"] = pkg
        sys.modules["mps.distributed"] = dist


def _import_from(path, name):
    """Import module `name` from directory `path` without leaving `path` on sys.path."""
    sys.path.insert(0, path)
    try:
        for stale in [m for m in sys.modules if m == name or m.startswith(name + ".")]:
            del sys.modules[stale]
        return importlib.import_module(name)
    finally:
        sys.path.remove(path)


def load_reference_kmeans():
    """Return the reference ``KMeans`` class (clustering/code/sgd_clustering.py:10)."""
    _install_kmeans_shims()
    mod = _import_from(os.path.join(REFERENCE_ROOT, "clustering/code"), "sgd_clustering")
    return mod.KMeans


def load_reference_measures():
    """Return (get_measure, get_cluster_pairing) of subset_selection/code (imports unmodified)."""
    code = os.path.join(REFERENCE_ROOT, "subset_selection/code")
    measures = _import_from(code, "measures")
    pairing = _import_from(code, "pairing")
    return measures.get_measure, pairing.get_cluster_pairing


@contextlib.contextmanager
def cuda_is_identity():
    """Shim 3: ``x.cuda()`` returns ``x`` while the reference's ``KMeans.add`` runs on CPU."""
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = saved


def reference_kmeans_args(device="cpu", num_gpus=1):
    comp = types.SimpleNamespace(device=device, num_gpus=num_gpus)
    return types.SimpleNamespace(computation=comp)
