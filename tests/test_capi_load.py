"""The C-ABI library builds for sm_100a, loads without a GPU and exports exactly what
include/acav_b200.h declares (no compute calls here)."""
import os
import re
import subprocess

import pytest

from acav100m_b200 import _lib, build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "acav_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acav_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    return build_mod.build()


def test_header_declares_functions():
    names = declared_functions()
    assert "acav_kmeans_assign" in names and "acav_mi_run" in names and len(names) >= 20


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing
    stray = [n for n in exported if n.startswith("acav_") and n not in declared_functions()]
    assert not stray, stray


def test_ctypes_table_matches_header(lib_path):
    assert sorted(_lib.SIGNATURES) == declared_functions()
    lib = _lib.load()
    assert lib.acav_abi_version() == 1
    assert _lib.status_string(0) == "ok"
    assert "unsupported" in _lib.status_string(-2)
    assert _lib.status_string(2)            # cudaErrorMemoryAllocation text comes from cudart


def test_library_is_self_contained(lib_path):
    """cudart is static and libcuda is not linked: loadable on a CPU-only box."""
    out = subprocess.check_output(["ldd", lib_path], text=True)
    assert "libcuda.so" not in out and "libcudart" not in out


def test_built_for_sm100a_only(lib_path):
    out = subprocess.check_output(["cuobjdump", "-lelf", lib_path], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_operators_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from acav100m_b200.clustering import KMeans
    from acav100m_b200.subset_selection import get_measure
    km = KMeans(None, 8, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        km.add(torch.zeros(16, 8))
    import numpy as np
    for name in ("mem_mi", "batch_mi", "mi", "ami"):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            get_measure(name)(np.zeros((4, 2), dtype=np.int64), ncentroids=2, batch_size=2, selection_size=1)
    rc = _lib.load().acav_device_info(None, None, None)
    assert rc != 0                            # a cudaError, not a silent success
