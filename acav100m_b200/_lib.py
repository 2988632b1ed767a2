"""ctypes binding of libacav_b200.so (C ABI in include/acav_b200.h).

There is no fallback: if the library is missing, was not built for this machine, or no CUDA device
is present, every operator raises.  Build with ``python -m acav100m_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacav_b200.so")

c_i32, c_i64, c_f32, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/acav_b200.h one to one (tests/test_capi_load.py
# parses the header and checks that this table and the .so export exactly the declared symbols).
SIGNATURES = {
    "acav_abi_version": (ctypes.c_int, []),
    "acav_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "acav_device_info": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_kmeans_create": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i64]),
    "acav_kmeans_destroy": (ctypes.c_int, [c_vp]),
    "acav_kmeans_workspace_bytes": (c_i64, [c_vp]),
    "acav_kmeans_set_tile_variant": (ctypes.c_int, [c_vp, c_i32]),
    "acav_kmeans_assign": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_f32, c_f32,
                                          c_vp, c_vp, c_vp, c_vp, c_i32, c_vp]),
    "acav_kmeans_prepare_centers": (ctypes.c_int, [c_vp, c_vp, c_vp, c_f32, c_f32, c_vp]),
    "acav_kmeans_prepare_batch": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp]),
    "acav_kmeans_assign_prepared": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_f32, c_f32,
                                                   c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acav_kmeans_comm_create": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32]),
    "acav_kmeans_comm_destroy": (ctypes.c_int, [c_vp]),
    "acav_kmeans_comm_handle_bytes": (ctypes.c_int, []),
    "acav_kmeans_comm_export": (ctypes.c_int, [c_vp, c_vp]),
    "acav_kmeans_comm_connect": (ctypes.c_int, [c_vp, c_vp]),
    "acav_kmeans_comm_arena": (c_vp, [c_vp]),
    "acav_kmeans_comm_connect_ptrs": (ctypes.c_int, [c_vp, c_vp]),
    "acav_kmeans_comm_status": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_kmeans_update_p2p": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_f64, c_vp, c_vp, c_vp, c_vp]),
    "acav_kmeans_underused_flags": (ctypes.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "acav_kmeans_assign_noise": (ctypes.c_int, [c_vp, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "acav_kmeans_histogram": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "acav_kmeans_update_fused": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_f64, c_vp, c_vp, c_vp, c_vp]),
    "acav_kmeans_update_sequential": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_f64, c_vp, c_vp, c_vp]),
    "acav_kmeans_update_local": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_f64, c_vp, c_vp, c_vp,
                                                c_vp, c_vp]),
    "acav_kmeans_apply_deltas": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "acav_mi_create": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_i64, c_i64]),
    "acav_mi_destroy": (ctypes.c_int, [c_vp]),
    "acav_mi_load_candidates": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_mi_set_tables": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "acav_mi_add_sample": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp]),
    "acav_mi_local_best": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_mi_apply": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "acav_mi_run": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_i32, c_vp]),
    "acav_mi_read_state": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acav_mi_pairs_create": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_i64, c_i64]),
    "acav_mi_pairs_destroy": (ctypes.c_int, [c_vp]),
    "acav_mi_pairs_load_candidates": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_mi_pairs_set_tables": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "acav_mi_pairs_add_sample": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_mi_pairs_record_words": (ctypes.c_int, [c_vp]),
    "acav_mi_pairs_local_best": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "acav_mi_pairs_apply": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "acav_mi_pairs_run": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "acav_mi_pairs_read_state": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "acav_mi_dense_create": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp]),
    "acav_mi_dense_destroy": (ctypes.c_int, [c_vp]),
    "acav_mi_dense_add": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "acav_mi_dense_score": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "acav_mi_dense_score_exact": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "acav_mi_dense_score_ami": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "acav_mi_debug_timers": (ctypes.c_int, [c_vp, c_vp]),
    "acav_mi_comm_handle_bytes": (ctypes.c_int, []),
    "acav_mi_comm_export": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp]),
    "acav_mi_comm_connect": (ctypes.c_int, [c_vp, c_vp]),
    "acav_mi_loop_supported": (ctypes.c_int, [c_i32, c_i32, c_i32]),
    "acav_mi_prepare": (ctypes.c_int, [c_vp, c_i32, c_vp]),
    "acav_mi_set_stream_variant": (ctypes.c_int, [c_vp, c_i32, c_i32]),
    "acav_mi_status": (ctypes.c_int, [c_vp, c_vp, c_vp]),
}

ASSIGN_EXACT, ASSIGN_TENSOR = 0, 1
TILE_AUTO, TILE_SINGLE, TILE_PAIR_256, TILE_PAIR_512 = 0, 1, 2, 3
MI_LOOP_KERNELS, MI_LOOP_PERSISTENT, MI_LOOP_CELLS, MI_LOOP_BYTES = 0, 1, 2, 3
E_INVALID, E_UNSUPPORTED, E_STATE, E_NO_DEVICE = -1, -2, -3, -4         # ACAV_E_* of include/acav_b200.h
PERSISTENT_READY = True         # persistent greedy-MI kernel validated against the C oracle on a B200
TENSOR_PATH_READY = True        # tcgen05 assignment validated against the exact kernel on a B200

_lib = None


class AcavError(RuntimeError):
    def __init__(self, fn, status, text):
        super().__init__(f"{fn} failed with status {status}: {text}")
        self.status = status


def load():
    """Load the shared library (CPU-safe: does not touch the CUDA driver)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing -- build it with `python -m acav100m_b200.build`. "
                "acav100m_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def status_string(status):
    return load().acav_status_string(int(status)).decode()


def call(name, *args):
    """Call a status-returning entry point; raise AcavError on a non-zero status."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise AcavError(name, rc, status_string(rc))


def require_cuda(device=None):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("acav100m_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"acav100m_b200 operators run on CUDA devices only, got {dev}")
    return dev


def stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, row_strided=False):
    """Device/host pointer of a contiguous torch tensor (or None).  `row_strided` admits a 2-D
    tensor whose rows are dense but separated by a larger stride (the ABI takes `ldx`)."""
    if t is None:
        return None
    if row_strided:
        assert t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]
    else:
        assert t.is_contiguous(), "non-contiguous tensor passed to the C ABI"
    return ctypes.c_void_p(t.data_ptr())
