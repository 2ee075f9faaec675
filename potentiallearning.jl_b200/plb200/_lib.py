"""ctypes binding of libplb200.so — the same C ABI a Julia `ccall` binds (include/plb200.h).

The library is the product: if it is missing, or no sm_100 GPU is present, every call raises. There is no CPU path.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLB200_LIB", os.path.join(_HERE, "libplb200.so"))

PL_KERNEL_LINEAR, PL_KERNEL_DOT, PL_KERNEL_RBF, PL_KERNEL_IMQ = 0, 1, 2, 3
PL_CINV_IDENTITY, PL_CINV_DIAGONAL, PL_CINV_FULL = 0, 1, 2
PL_FILL_RM_UPPER, PL_FILL_RM_LOWER, PL_FILL_FULL = 1, 2, 3
PL_FILL_JULIA_U = PL_FILL_RM_LOWER

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/plb200.h one to one (tests/test_abi.py checks the header against this)
PROTOTYPES = {
    "pl_init": (_int, [_int]),
    "pl_init_all": (_int, [_int, ctypes.POINTER(ctypes.c_int)]),
    "pl_device_count": (_int, []),
    "pl_comm_unique_id": (_int, [_vp]),
    "pl_comm_init_rank": (_int, [_int, _int, _vp]),
    "pl_comm_info": (_int, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "pl_comm_destroy": (_int, []),
    "pl_comm_peer_enabled": (_int, []),
    "pl_last_coll_ms": (_int, [_dp]),
    "pl_normal_eq_sharded": (_int, [_dp, _i64, _i64, _i64, _dp, _i64, _dbl, _dbl, _dp, _int, _dbl, _dp, _dp]),
    "pl_normal_eq_sharded_dev": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _dbl, _dbl, _vp, _int, _dbl, _vp, _vp, _vp]),
    "pl_cov_sharded": (_int, [_dp, _i64, _i64, _i64, _i64, _dp, _int, _int, _dp]),
    "pl_cov_sharded_dev": (_int, [_vp, _i64, _i64, _i64, _i64, _vp, _int, _int, _vp, _vp]),
    "pl_allreduce_dev": (_int, [_vp, _i64, _vp]),
    "pl_gram_resident": (_int, [_int, _dp, _i64, _i64, _i64, _int, _int, _dp, _dbl, _dbl, ctypes.POINTER(_vp)]),
    "pl_gram_resident_info": (_int, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "pl_gram_resident_shard": (_int, [_vp, _int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_i64), ctypes.POINTER(_vp)]),
    "pl_gram_fetch_block": (_int, [_vp, _i64, _i64, _i64, _i64, _dp, _i64]),
    "pl_gram_free": (_int, [_vp]),
    "pl_finalize": (None, []),
    "pl_last_error": (ctypes.c_char_p, []),
    "pl_last_timing": (_int, [_dp, _dp, _dp]),
    "pl_last_host_ms": (_int, [_dp]),
    "pl_launch_count": (_i64, []),
    "pl_workspace_bytes": (_i64, []),
    "pl_workspace_trim": (_int, [_int]),
    "pl_syrk_schedule": (_int, [_i64, _i64, _int, _int, _int, ctypes.c_double, _i64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "pl_set_profiling": (_int, [_int]),
    "pl_last_main_kernel_ms": (_int, [_dp]),
    "pl_host_alloc": (_vp, [_i64]),
    "pl_host_free": (_int, [_vp]),
    "pl_gram_part": (_int, [_int, _dp, _i64, _i64, _i64, _int, _int, _dp, _dbl, _dbl, _int, _int, _dp]),
    "pl_gram_dot": (_int, [_dp, _i64, _i64, _i64, _int, _dp, _i64, _int]),
    "pl_gram_rbf": (_int, [_dp, _i64, _i64, _i64, _int, _dp, _dbl, _dbl, _dp, _i64, _int]),
    "pl_gram_imq": (_int, [_dp, _i64, _i64, _i64, _int, _dp, _dbl, _dbl, _dp, _i64, _int]),
    "pl_gram_cross": (_int, [_int, _dp, _i64, _i64, _dp, _i64, _i64, _i64, _int, _int, _dp, _dbl, _dbl, _dp, _i64]),
    "pl_normal_eq": (_int, [_dp, _i64, _i64, _i64, _dp, _i64, _dbl, _dbl, _dp, _int, _dbl, _dp, _dp]),
    "pl_normal_eq_blocks": (_int, [ctypes.POINTER(_dp), ctypes.POINTER(_i64), _i64, _i64, _dp, _i64, _dbl, _dbl, _dp, _int, _dbl, _i64, _dp, _dp]),
    "pl_neq_begin": (_int, [_i64]),
    "pl_neq_add": (_int, [_dp, _i64, _i64, _dp, _i64, _dbl, _dbl, _dp]),
    "pl_neq_finish": (_int, [_int, _dbl, _dp, _dp]),
    "pl_global_feature": (_int, [_dp, _i64, ctypes.POINTER(_i64), _i64, _i64, _int, _dp, _i64]),
    "pl_ksd_rbf": (_int, [_dp, _dp, _i64, _i64, _i64, _i64, _int, _dp, _dbl, _dbl, _dp]),
    "pl_correlation_feature": (_int, [_dp, _i64, ctypes.POINTER(_i64), _i64, _i64, _dp]),
    "pl_global_feature_dev": (_int, [_vp, _i64, _vp, _i64, _i64, _int, _vp, _i64, _i64, _vp]),
    "pl_cov": (_int, [_dp, _i64, _i64, _i64, _dp, _int, _int, _dp]),
    "pl_eigh": (_int, [_dp, _i64, _i64, _int, _dp, _dp, _i64]),
    "pl_eigh_dev": (_int, [_vp, _i64, _i64, _int, _int, _vp, _vp]),
    "pl_ell_ensemble": (_int, [_int, _dp, _i64, _i64, _i64, _int, _int, _dp, _dbl, _dbl, _dp, _dp, _i64]),
    "pl_ell_inclusion_prob": (_int, [_dbl, _i64, _dp]),
    "pl_ell_eigvecs": (_int, [ctypes.POINTER(_i64), _i64, _i64, _dp, _i64]),
    "pl_ell_sample": (_int, [ctypes.POINTER(_i64), _i64, _i64, _dp, ctypes.POINTER(_i64)]),
    "pl_ell_greedy": (_int, [_i64, _i64, ctypes.POINTER(_i64)]),
    "pl_esp_select": (_int, [_dp, _i64, _i64, _dp, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "pl_rmsd_periodic": (_int, [_dp, _i64, _i64, _dp, _dp, _i64]),
    "pl_rmsd_kabsch": (_int, [_dp, _i64, _i64, _dp, _i64]),
    "pl_metrics": (_int, [_dp, _dp, _i64, _dp]),
    "pl_metrics_dev": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "pl_predict": (_int, [_dp, _i64, _i64, _i64, _dp, _dbl, _i64, _dp]),
    "pl_predict_dev": (_int, [_vp, _i64, _i64, _i64, _vp, _dbl, _i64, _vp, _vp]),
    "pl_gram_dev": (_int, [_int, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _int, _int, _vp, _dbl, _dbl, _vp, _i64, _int, _int, _int, _int, _vp]),
    "pl_gram_tile_count": (_i64, [_i64, _i64, _int, _int]),
    "pl_gram_tile_coords": (_int, [_i64, _i64, _int, _int, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "pl_syrk_dev": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp]),
    "pl_cov_shifted_dev": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "pl_cov_recentre_dev": (_int, [_vp, _i64, _vp, _vp, _dbl, _vp]),
    "pl_normal_eq_finish_dev": (_int, [_vp, _vp, _i64, _int, _dbl, _vp, _vp, _vp]),
    "pl_colsum_dev": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "pl_scale_div_dev": (_int, [_vp, _dbl, _i64, _vp, _vp]),
    "pl_sym_scale_div_dev": (_int, [_vp, _i64, _dbl, _vp, _vp]),
}


class PLB200Error(RuntimeError):
    pass


_lib = None
_inited = False
_device = None  # CUDA device index the library's first context is bound to (None: whatever was current at pl_init(-1))


def load():
    """dlopen libplb200.so and attach prototypes. Raises if the library has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PLB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). plb200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().pl_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise PLB200Error(f"{what} failed (code {rc}): {last_error()}")


def init(device: int = -1):
    """pl_init: bind to one B200. Raises PLB200Error when there is no sm_100 device."""
    global _inited, _device
    lib = load()
    if not _inited:
        check(lib.pl_init(int(device)), "pl_init")
        _inited = True
        _device = int(device) if device is not None and int(device) >= 0 else None
    elif device is not None and int(device) >= 0 and _device is not None and int(device) != _device:
        raise PLB200Error(f"libplb200 is bound to CUDA device {_device}; cannot re-bind it to device {device} (one process drives one GPU, "
                          "or use init_all)")
    return lib


def init_all(ndev: int = 0, device_ids=None):
    """pl_init_all: ONE process drives `ndev` GPUs (0: all visible); the host entry points then shard large inputs over them."""
    global _inited, _device
    lib = load()
    if not _inited:
        ids = None if device_ids is None else (ctypes.c_int * len(device_ids))(*device_ids)
        check(lib.pl_init_all(int(ndev if device_ids is None else len(device_ids)), ids), "pl_init_all")
        _inited = True
        _device = 0 if device_ids is None else int(device_ids[0])
    return lib


def device_count() -> int:
    return int(load().pl_device_count())


def comm_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    check(load().pl_comm_unique_id(buf), "pl_comm_unique_id")
    return buf.raw


def comm_init_rank(nranks: int, rank: int, uid: bytes):
    """collective: every rank of a one-process-per-GPU job calls it with rank 0's pl_comm_unique_id bytes"""
    assert len(uid) == 128
    check(lib().pl_comm_init_rank(int(nranks), int(rank), ctypes.create_string_buffer(uid, 128)), "pl_comm_init_rank")


def comm_init_from_torch(group=None):
    """Create the library's NCCL communicator over the ranks of an initialised torch.distributed job: rank 0's unique id travels
    through torch's object broadcast (control plane only); the data-path collectives are then issued by libplb200 itself."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    comm_init_rank(world, rank, box[0])
    return world, rank


def comm_info():
    n, r, v = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    load().pl_comm_info(ctypes.byref(n), ctypes.byref(r), ctypes.byref(v))
    return {"nranks": n.value, "rank": r.value, "nccl_version": v.value, "peer_exchange": bool(load().pl_comm_peer_enabled())}


def comm_destroy():
    check(load().pl_comm_destroy(), "pl_comm_destroy")


def last_coll_ms() -> float:
    ms = _dbl()
    check(load().pl_last_coll_ms(ctypes.byref(ms)), "pl_last_coll_ms")
    return ms.value


def lib():
    return init()


def finalize():
    global _inited
    global _device
    if _lib is not None and _inited:
        _lib.pl_finalize()
        _inited = False
        _device = None


def last_timing():
    k, h, d = _dbl(), _dbl(), _dbl()
    load().pl_last_timing(ctypes.byref(k), ctypes.byref(h), ctypes.byref(d))
    w = _dbl()
    load().pl_last_host_ms(ctypes.byref(w))
    return {"kernel_ms": k.value, "h2d_ms": h.value, "d2h_ms": d.value, "host_wall_ms": w.value}


def set_profiling(on: bool):
    check(load().pl_set_profiling(1 if on else 0), "pl_set_profiling")


def last_main_kernel_ms() -> float:
    ms = _dbl()
    check(load().pl_last_main_kernel_ms(ctypes.byref(ms)), "pl_last_main_kernel_ms")
    return ms.value


def launch_count() -> int:
    return int(load().pl_launch_count())


def workspace_bytes() -> int:
    """device scratch currently held by the library"""
    return int(load().pl_workspace_bytes())


def dptr(a: np.ndarray | None):
    """host pointer of a C-contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def pinned_zeros(shape):
    """zeros(shape) in pinned host memory owned by the library (pl_host_alloc): results land there at the PCIe rate. The array keeps
    its allocation alive; it is returned to the driver when the array (and every view of it) is garbage collected."""
    import weakref

    n = int(np.prod(shape))
    lib_ = lib()
    ptr = lib_.pl_host_alloc(max(n, 1) * 8)
    if not ptr:
        raise PLB200Error(f"pl_host_alloc failed: {last_error()}")
    buf = (ctypes.c_double * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
    weakref.finalize(buf, lib_.pl_host_free, ptr)
    return arr

