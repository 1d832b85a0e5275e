"""``FlatIPIndex`` - the reference-facing shim over the C-ABI CUDA library.

Mirrors the part of ``faiss.IndexFlatIP`` / ``GpuIndexFlatIP`` the reference relies on
(`/root/reference/src/test_HAConvDR_topiocqa.py:52` ctor, `:98` ``add``, `:102`
``search(q, k) -> (D, I)``, `:122` ``reset``; attributes ``d`` and ``ntotal``):

* ``add`` copies the rows (caller may free its array afterwards, `:123-124`), ids are
  insertion order;
* ``search`` returns ``D`` float32 ``[nq, k]`` sorted descending and ``I`` int64 ``[nq, k]``;
  unfilled slots are ``-3.4028235e38`` / ``-1``; ties are ordered by ascending id;
* dimension mismatches raise ``AssertionError`` like the faiss Python wrapper
  (``assert d == self.d``), a k outside ``[1, HAC_MAX_K]`` raises ``ValueError``.

NumPy in -> NumPy out (host buffers, copies inside the call); CUDA torch tensors in ->
CUDA torch tensors out (``data_ptr()`` handoff on the current stream, no host round trip).
PyTorch is optional and only used for that tensor handoff.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import HAC_MAX_K, HAC_PATH_AUTO, HacStats, check


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class FlatIPIndex:
    def __init__(self, d: int = 768, device: int = 0, reserve: int = 0):
        self._h = ctypes.c_void_p()
        self._lib = _lib.lib()
        check(self._lib.hac_create(int(d), int(device), ctypes.byref(self._h)), "hac_create")
        self.d = int(d)
        self.device = int(device)
        self.is_trained = True
        if reserve:
            self.reserve(reserve)

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hac_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ntotal(self) -> int:
        return int(self._lib.hac_ntotal(self._h))

    def reserve(self, n_rows: int):
        check(self._lib.hac_reserve(self._h, int(n_rows)), "hac_reserve")

    # -- add / reset --------------------------------------------------------------------------
    def add(self, x):
        if _is_torch_tensor(x) and x.is_cuda:
            import torch
            assert x.dim() == 2 and x.shape[1] == self.d, "add: expected [n, %d]" % self.d
            if x.device.index != self.device:
                raise ValueError("add: tensor lives on cuda:%s, index on cuda:%d" % (x.device.index, self.device))
            x = x.contiguous().float()
            stream = torch.cuda.current_stream(x.device).cuda_stream
            check(self._lib.hac_add_device(self._h, x.shape[0], x.data_ptr(), stream), "hac_add_device")
            return
        if _is_torch_tensor(x):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, "add: expected [n, %d] float32" % self.d
        check(self._lib.hac_add(self._h, x.shape[0], x.ctypes.data), "hac_add")

    def add_synthetic(self, n: int, seed: int = 42, row0: int = 0, dist: int = 0):
        """Append ``n`` device-generated rows; row ``row0 + i`` depends only on (seed, row0 + i)."""
        check(self._lib.hac_add_synthetic(self._h, int(n), int(seed), int(row0), int(dist)), "hac_add_synthetic")

    def reset(self):
        check(self._lib.hac_reset(self._h), "hac_reset")

    # -- shard files (engine-native corpus format: fp32 rows + int8 image + tile constants + centre) -----------
    def save_shard(self, path: str):
        check(self._lib.hac_save_shard(self._h, str(path).encode()), "hac_save_shard")

    def load_shard(self, path: str):
        """Fill this (empty) index from a shard file by plain DMA - no conversion, no statistics pass."""
        check(self._lib.hac_load_shard(self._h, str(path).encode()), "hac_load_shard")

    # -- id translation (replaces passage_embedding2id[I], reference :110) ---------------------------
    def set_id_base(self, base: int):
        check(self._lib.hac_set_id_base(self._h, int(base)), "hac_set_id_base")

    def set_id_table(self, ids):
        if ids is None:
            check(self._lib.hac_set_id_table(self._h, None, 0), "hac_set_id_table")
            return
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        check(self._lib.hac_set_id_table(self._h, ids.ctypes.data, ids.shape[0]), "hac_set_id_table")

    # -- search -------------------------------------------------------------------------------
    def search(self, q, k: int, path: int = HAC_PATH_AUTO, out=None, D=None, I=None):
        """``out=(D, I)``: optional preallocated CUDA result tensors (float32 / int64 ``[nq, k]``,
        contiguous) for the tensor path, e.g. views of a symmetric-memory buffer peers read from.
        ``D=``, ``I=``: preallocated NumPy result arrays for the host path, as ``faiss.Index.search`` accepts
        (page-locked arrays receive the results by direct DMA)."""
        k = int(k)
        if k <= 0 or k > HAC_MAX_K:
            raise ValueError("search: k=%d outside [1, %d]" % (k, HAC_MAX_K))
        if _is_torch_tensor(q) and q.is_cuda:
            import torch
            assert q.dim() == 2 and q.shape[1] == self.d, "search: expected [nq, %d]" % self.d
            if q.device.index != self.device:
                raise ValueError("search: tensor lives on cuda:%s, index on cuda:%d" % (q.device.index, self.device))
            q = q.contiguous().float()
            if out is not None:
                D, I = out
                assert D.is_cuda and I.is_cuda and D.is_contiguous() and I.is_contiguous()
                assert D.dtype == torch.float32 and I.dtype == torch.int64
                assert tuple(D.shape) == (q.shape[0], k) and tuple(I.shape) == (q.shape[0], k)
            else:
                D = torch.empty((q.shape[0], k), dtype=torch.float32, device=q.device)
                I = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
            stream = torch.cuda.current_stream(q.device).cuda_stream
            check(self._lib.hac_search_device_ex(self._h, q.shape[0], q.data_ptr(), k, D.data_ptr(), I.data_ptr(),
                                                 stream, int(path)), "hac_search_device")
            return D, I
        if _is_torch_tensor(q):
            q = q.detach().cpu().numpy()
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.d, "search: expected [nq, %d] float32" % self.d
        if D is None:
            D = np.empty((q.shape[0], k), dtype=np.float32)
        if I is None:
            I = np.empty((q.shape[0], k), dtype=np.int64)
        assert D.shape == (q.shape[0], k) and D.dtype == np.float32 and D.flags.c_contiguous
        assert I.shape == (q.shape[0], k) and I.dtype == np.int64 and I.flags.c_contiguous
        check(self._lib.hac_search_ex(self._h, q.shape[0], q.ctypes.data, k, D.ctypes.data, I.ctypes.data,
                                      int(path)), "hac_search")
        return D, I

    def set_threshold_exchange(self, mine_ptr: int, peer_ptrs, capacity: int):
        """Cross-shard threshold exchange buffers (raw device pointers; see ``hac_set_threshold_exchange``)."""
        n = len(peer_ptrs)
        arr = (ctypes.c_void_p * max(n, 1))(*peer_ptrs)
        check(self._lib.hac_set_threshold_exchange(self._h, mine_ptr, arr, n, int(capacity)),
              "hac_set_threshold_exchange")

    def set_option(self, name: str, value: int):
        check(self._lib.hac_set_option(self._h, name.encode(), int(value)), "hac_set_option")

    def stats(self) -> dict:
        st = HacStats()
        check(self._lib.hac_get_stats(self._h, ctypes.byref(st)), "hac_get_stats")
        return st.as_dict()


def merge_topk_device(D_lists, I_lists, k_out: int):
    """k-way merge on the device: ``D_lists`` float32 / ``I_lists`` int64 CUDA tensors of shape
    ``[n_lists, nq, k]`` -> ``(D [nq, k_out], I [nq, k_out])`` by (score desc, id asc)."""
    import torch
    assert D_lists.is_cuda and I_lists.is_cuda and D_lists.shape == I_lists.shape and D_lists.dim() == 3
    D_lists = D_lists.contiguous().float()
    I_lists = I_lists.contiguous().long()
    n_lists, nq, k = D_lists.shape
    D = torch.empty((nq, k_out), dtype=torch.float32, device=D_lists.device)
    I = torch.empty((nq, k_out), dtype=torch.int64, device=D_lists.device)
    stream = torch.cuda.current_stream(D_lists.device).cuda_stream
    check(_lib.lib().hac_merge_topk_device(D_lists.device.index, n_lists, nq, k, D_lists.data_ptr(),
                                           I_lists.data_ptr(), int(k_out), D.data_ptr(), I.data_ptr(), stream),
          "hac_merge_topk_device")
    return D, I


def merge_topk_peers_device(D_ptrs, I_ptrs, nq: int, k: int, k_out: int, device):
    """Merge lists that live in different buffers (raw device pointers, e.g. the peers' symmetric-memory
    result buffers, read in-kernel over NVLink) -> ``(D [nq, k_out], I [nq, k_out])`` on ``device``."""
    import torch
    n = len(D_ptrs)
    assert n == len(I_ptrs) and n >= 1
    dev = torch.device(device)
    D = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
    I = torch.empty((nq, k_out), dtype=torch.int64, device=dev)
    arr_t = ctypes.c_void_p * n
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().hac_merge_topk_peers_device(dev.index, n, nq, k, arr_t(*D_ptrs), arr_t(*I_ptrs), int(k_out),
                                                 D.data_ptr(), I.data_ptr(), stream),
          "hac_merge_topk_peers_device")
    return D, I


def gather_ids_device(table, ids):
    """offset -> pid on the device (`/root/reference/src/test_HAConvDR_topiocqa.py:250`): ``table`` int64 CUDA
    tensor (offset2pid), ``ids`` int64 CUDA tensor of offsets (any shape); unfilled slots (-1) map to -1."""
    import torch
    assert table.is_cuda and ids.is_cuda and table.dtype == torch.int64 and ids.dtype == torch.int64
    ids_c = ids.contiguous()
    out = torch.empty_like(ids_c)
    stream = torch.cuda.current_stream(ids.device).cuda_stream
    check(_lib.lib().hac_gather_ids_device(ids.device.index, table.contiguous().data_ptr(), table.numel(),
                                           ids_c.data_ptr(), ids_c.numel(), out.data_ptr(), stream),
          "hac_gather_ids_device")
    return out


def synth_rows_device(n: int, d: int, seed: int, row0: int = 0, dist: int = 0, device: int = 0):
    """Rows of the device generator as a CUDA tensor (queries, and read-back for parity tests)."""
    import torch
    out = torch.empty((n, d), dtype=torch.float32, device="cuda:%d" % device)
    stream = torch.cuda.current_stream(out.device).cuda_stream
    check(_lib.lib().hac_synth_fill_device(device, out.data_ptr(), n, d, int(seed), int(row0), int(dist), stream),
          "hac_synth_fill_device")
    return out
