"""ctypes binding of the C-ABI CUDA library (``libhacidx.so``, built in-tree by
``__graft_entry__.build()`` / ``haconvdr_b200/csrc/Makefile``).

There is no CPU fallback: if the library is missing or cannot be loaded the
import of any engine object fails loudly with the reason.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhacidx.so")
CSRC = os.path.join(_HERE, "csrc")

HAC_PATH_AUTO, HAC_PATH_GEMV, HAC_PATH_MMA, HAC_PATH_I8 = 0, 1, 2, 3
HAC_MAX_K = 1024

HAC_ABI_VERSION = 3
HAC_EXCHANGE_WORDS_PER_QUERY = 16

c_i64 = ctypes.c_int64
c_f32p = ctypes.POINTER(ctypes.c_float)
c_i64p = ctypes.POINTER(ctypes.c_int64)


class HacStats(ctypes.Structure):
    _fields_ = [
        ("path", ctypes.c_int32), ("retries", ctypes.c_int32), ("n_chunks", ctypes.c_int32),
        ("kernel_launches", ctypes.c_int32), ("candidates_emitted", ctypes.c_int64),
        ("candidates_rescored", ctypes.c_int64), ("margin_max", ctypes.c_float),
        ("screen_err_max", ctypes.c_float), ("scan_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
        ("ntotal", ctypes.c_int64), ("bytes_fp32", ctypes.c_int64), ("bytes_shadow", ctypes.c_int64),
        ("bytes_i8", ctypes.c_int64), ("n_sync_chunks", ctypes.c_int32), ("pipelined", ctypes.c_int32),
        ("tail_ms", ctypes.c_float), ("warm_rows", ctypes.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/hac_index.h declares: name -> (restype, argtypes)
_VP = ctypes.c_void_p
SIGNATURES = {
    "hac_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_VP)]),
    "hac_destroy": (ctypes.c_int, [_VP]),
    "hac_reserve": (ctypes.c_int, [_VP, c_i64]),
    "hac_add": (ctypes.c_int, [_VP, c_i64, _VP]),
    "hac_add_device": (ctypes.c_int, [_VP, c_i64, _VP, _VP]),
    "hac_reset": (ctypes.c_int, [_VP]),
    "hac_add_synthetic": (ctypes.c_int, [_VP, c_i64, ctypes.c_uint64, c_i64, ctypes.c_int]),
    "hac_synth_fill_device": (ctypes.c_int, [ctypes.c_int, _VP, c_i64, ctypes.c_int, ctypes.c_uint64, c_i64,
                                             ctypes.c_int, _VP]),
    "hac_set_id_base": (ctypes.c_int, [_VP, c_i64]),
    "hac_set_id_table": (ctypes.c_int, [_VP, _VP, c_i64]),
    "hac_search": (ctypes.c_int, [_VP, c_i64, _VP, ctypes.c_int, _VP, _VP]),
    "hac_search_device": (ctypes.c_int, [_VP, c_i64, _VP, ctypes.c_int, _VP, _VP, _VP]),
    "hac_search_ex": (ctypes.c_int, [_VP, c_i64, _VP, ctypes.c_int, _VP, _VP, ctypes.c_int]),
    "hac_search_device_ex": (ctypes.c_int, [_VP, c_i64, _VP, ctypes.c_int, _VP, _VP, _VP, ctypes.c_int]),
    "hac_merge_topk_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_i64, ctypes.c_int, _VP, _VP,
                                             ctypes.c_int, _VP, _VP, _VP]),
    "hac_merge_topk_peers_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_i64, ctypes.c_int,
                                                   ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.c_int, _VP, _VP, _VP]),
    "hac_enable_peer_access": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "hac_shards_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_int, ctypes.POINTER(_VP)]),
    "hac_shards_destroy": (ctypes.c_int, [_VP]),
    "hac_shards_search": (ctypes.c_int, [_VP, c_i64, _VP, ctypes.c_int, _VP, _VP]),
    "hac_shards_search_device": (ctypes.c_int, [_VP, c_i64, _VP, ctypes.c_int, _VP, _VP, _VP]),
    "hac_shards_set_exchange": (ctypes.c_int, [_VP, ctypes.c_int]),
    "hac_shards_peer_access": (ctypes.c_int, [_VP]),
    "hac_shards_last_phases": (ctypes.c_int, [_VP, c_f32p, ctypes.c_int]),
    "hac_gather_ids_device": (ctypes.c_int, [ctypes.c_int, _VP, c_i64, _VP, c_i64, _VP, _VP]),
    "hac_set_threshold_exchange": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(_VP), ctypes.c_int, c_i64]),
    "hac_reciprocal_rank_device": (ctypes.c_int, [ctypes.c_int, _VP, c_i64, ctypes.c_int, _VP, _VP, _VP, _VP, _VP]),
    "hac_save_shard": (ctypes.c_int, [_VP, ctypes.c_char_p]),
    "hac_load_shard": (ctypes.c_int, [_VP, ctypes.c_char_p]),
    "hac_pinned_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_VP)]),
    "hac_pinned_free": (ctypes.c_int, [_VP]),
    "hac_set_option": (ctypes.c_int, [_VP, ctypes.c_char_p, c_i64]),
    "hac_ntotal": (c_i64, [_VP]),
    "hac_dim": (ctypes.c_int, [_VP]),
    "hac_device": (ctypes.c_int, [_VP]),
    "hac_get_stats": (ctypes.c_int, [_VP, ctypes.POINTER(HacStats)]),
    "hac_abi_version": (ctypes.c_int, []),
    "hac_last_error": (ctypes.c_char_p, []),
}

_LIB = None


def build(verbose: bool = False) -> str:
    """Compile the library for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libhacidx.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib():
    """Load the CUDA library; raises (no fallback) if it is not there."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "haconvdr_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C haconvdr_b200/csrc`). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the ABI lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error() -> str:
    return (lib().hac_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = "%s failed (%d): %s" % (what or "hac call", rc, last_error())
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise MemoryError(msg)
    raise RuntimeError(msg)
