"""faiss-restatement CPU IndexFlatIP: BLAS sgemm blocks + C heaps (TEST INFRASTRUCTURE ONLY).

This is the timed CPU baseline ("port") of bench.py: it follows faiss 1.7.x
``IndexFlatIP::search`` for nq >= 20 - fp32 ``sgemm`` over 4096-query x 1024-row
blocks, then per-query min-heap updates with strict ``>`` - which is what
`/root/reference/src/test_HAConvDR_topiocqa.py:102` executes on CPU when
``use_gpu`` is false.  The sgemm is torch's MKL ``mm`` (numpy's BLAS when torch is absent);
the heap half is ``flat_ip_c.c`` (OpenMP over queries).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_flat_ip.so")
    src = os.path.join(_HERE, "flat_ip_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_flat_ip.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i64, f32p, i64p = ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        L.oracle_heap_init.argtypes = [i64, ctypes.c_int, f32p, i64p]
        L.oracle_heap_addn.argtypes = [i64, ctypes.c_int, f32p, i64p, f32p, i64, i64, i64]
        L.oracle_heap_reorder.argtypes = [i64, ctypes.c_int, f32p, i64p]
        L.oracle_search_naive.argtypes = [i64, i64, ctypes.c_int, f32p, f32p, ctypes.c_int, f32p, i64p]
        for f in (L.oracle_heap_init, L.oracle_heap_addn, L.oracle_heap_reorder, L.oracle_search_naive):
            f.restype = None
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def search_blas(q: np.ndarray, x: np.ndarray, k: int, query_block: int = 4096, corpus_block: int = 1024):
    """faiss-style blocked search: returns (D float32 [nq,k], I int64 [nq,k])."""
    L = lib()
    q = np.ascontiguousarray(q, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    nq, n = q.shape[0], x.shape[0]
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    L.oracle_heap_init(nq, k, _fp(D), _ip(I))
    buf = np.empty((min(query_block, nq), corpus_block), np.float32)
    try:                                   # MKL sgemm("N","T") through torch: no transposed-view penalty
        import torch
        tq, tx, tbuf = torch.from_numpy(q), torch.from_numpy(x), torch.from_numpy(buf)
    except ImportError:
        torch = None
    for q0 in range(0, nq, query_block):
        q1 = min(nq, q0 + query_block)
        for j0 in range(0, n, corpus_block):
            j1 = min(n, j0 + corpus_block)
            ip = buf[: q1 - q0, : j1 - j0]
            if torch is not None:
                torch.mm(tq[q0:q1], tx[j0:j1].T, out=tbuf[: q1 - q0, : j1 - j0])
            else:
                np.matmul(q[q0:q1], x[j0:j1].T, out=ip)
            L.oracle_heap_addn(q1 - q0, k, _fp(D[q0:q1]), _ip(I[q0:q1]), _fp(ip), j1 - j0, buf.shape[1], j0)
    L.oracle_heap_reorder(nq, k, _fp(D), _ip(I))
    return D, I


def search_naive(q: np.ndarray, x: np.ndarray, k: int):
    L = lib()
    q = np.ascontiguousarray(q, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    D = np.empty((q.shape[0], k), np.float32)
    I = np.empty((q.shape[0], k), np.int64)
    L.oracle_search_naive(q.shape[0], x.shape[0], q.shape[1], _fp(q), _fp(x), k, _fp(D), _ip(I))
    return D, I


class FlatIPPort:
    """faiss surface over ``search_blas`` (used as the CPU index in the baseline)."""

    def __init__(self, d):
        self.d, self.ntotal, self._x = int(d), 0, []

    def add(self, x):
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape[1] == self.d
        self._x.append(x.copy())
        self.ntotal += len(x)

    def reset(self):
        self._x, self.ntotal = [], 0

    def search(self, q, k):
        x = self._x[0] if len(self._x) == 1 else np.concatenate(self._x, 0)
        return search_blas(q, x, k)
