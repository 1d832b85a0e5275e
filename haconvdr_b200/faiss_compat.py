"""Stand-ins for the faiss names the reference driver touches, so that
``build_faiss_index`` (`/root/reference/src/test_HAConvDR_topiocqa.py:39-71`) runs
unmodified with ``import haconvdr_b200.faiss_compat as faiss``:

    StandardGpuResources().setTempMemory(n)        :46-50
    IndexFlatIP(768)                               :52
    GpuMultipleClonerOptions() .shard .usePrecomputed   :55-57
    GpuResourcesVector().push_back / Int32Vector().push_back   :59-63
    index_cpu_to_gpu_multiple(vres, vdev, cpu_index, co)       :64-66

There is no CPU index: ``IndexFlatIP`` binds lazily to a GPU shard on first use (device 0
unless re-targeted by ``index_cpu_to_gpu_multiple``), and a multi-device clone shards rows
contiguously across the devices of this process, as faiss ``IndexShards`` does with
``co.shard = True``.
"""
from __future__ import annotations

import numpy as np

from .index import FlatIPIndex


class StandardGpuResources:
    def __init__(self):
        self.temp_memory = None

    def setTempMemory(self, nbytes):   # accepted for compatibility; the engine sizes its own workspace
        self.temp_memory = int(nbytes)

    def noTempMemory(self):
        self.temp_memory = 0


class GpuMultipleClonerOptions:
    def __init__(self):
        self.shard = False
        self.usePrecomputed = False
        self.useFloat16 = False
        self.indicesOptions = 0
        self.verbose = False


GpuClonerOptions = GpuMultipleClonerOptions


class _Vector(list):
    def push_back(self, v):
        self.append(v)

    def size(self):
        return len(self)

    def at(self, i):
        return self[i]


class GpuResourcesVector(_Vector):
    pass


class Int32Vector(_Vector):
    pass


def get_num_gpus() -> int:
    import torch
    return torch.cuda.device_count()


class IndexFlatIP:
    """faiss.IndexFlatIP surface; storage is created on the target GPU at first use."""

    def __init__(self, d: int, device: int = 0):
        self.d = int(d)
        self._device = int(device)
        self._impl = None
        self.is_trained = True

    def _get(self) -> FlatIPIndex:
        if self._impl is None:
            self._impl = FlatIPIndex(self.d, self._device)
        return self._impl

    @property
    def ntotal(self):
        return 0 if self._impl is None else self._impl.ntotal

    def add(self, x):
        self._get().add(x)

    def search(self, q, k, D=None, I=None):
        return self._get().search(q, k, D=D, I=I)

    def reset(self):
        if self._impl is not None:
            self._impl.reset()

    def stats(self):
        return self._get().stats()


class ShardedInProcessIndex:
    """faiss IndexShards equivalent inside one process (`index_cpu_to_gpu_multiple` with ``co.shard = True``,
    `/root/reference/src/test_HAConvDR_topiocqa.py:55-66`): rows of every ``add`` are split into contiguous ranges,
    one per device.  ``search`` is ONE call into the library (``hac_shards_search``, `csrc/hac_shards.cu`), the same
    protocol as one process per GPU (``haconvdr_b200.sharded``):

    * one persistent native host thread per shard (as faiss' own IndexShards keeps) copies the queries to its device
      and runs the single-shard search - no Python, no GIL on the search path;
    * the shards publish their best scores to each other WHILE they scan (``hac_set_threshold_exchange`` over
      peer-mapped device pointers owned by the group), so every shard rescores only its share of the global top-k;
    * results stay in per-device buffers; ONE merge kernel on the first device reads all of them in place over NVLink
      - no gather copies, no host round trip before the final D2H into the caller's arrays."""

    def __init__(self, d: int, devices):
        import ctypes
        from concurrent.futures import ThreadPoolExecutor
        from . import _lib
        self.d = int(d)
        self.devices = [int(v) for v in devices]
        self.shards = [FlatIPIndex(self.d, dev) for dev in self.devices]
        self._rows = [[] for _ in self.devices]          # per shard: (global_start, n) of every add
        self.ntotal = 0
        self.is_trained = True
        self.threshold_exchange = True
        self._lib = _lib.lib()
        self._pool = ThreadPoolExecutor(len(self.devices), thread_name_prefix="hac-shard-add")
        handles = (ctypes.c_void_p * len(self.shards))(*[sh._h.value for sh in self.shards])
        self._grp = ctypes.c_void_p()
        _lib.check(self._lib.hac_shards_create(handles, len(self.shards), ctypes.byref(self._grp)), "hac_shards_create")
        self.peer_access = bool(self._lib.hac_shards_peer_access(self._grp))
        self.last_phase_ms = None

    def _set_ids(self, g):
        rows = self._rows[g]
        if len(rows) == 1:
            self.shards[g].set_id_base(rows[0][0])
        else:                                            # several adds: ids are no longer base + local row
            self.shards[g].set_id_table(np.concatenate([np.arange(s, s + n, dtype=np.int64) for s, n in rows]))

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, "add: expected [n, %d] float32" % self.d
        n, G = x.shape[0], len(self.shards)
        bounds = [(g * n) // G for g in range(G + 1)]

        def one(g):
            lo, hi = bounds[g], bounds[g + 1]
            if hi > lo:
                self.shards[g].add(x[lo:hi])
                self._rows[g].append((self.ntotal + lo, hi - lo))
                self._set_ids(g)
        list(self._pool.map(one, range(G)))              # the shards copy and convert their rows concurrently
        self.ntotal += n

    def reset(self):
        for g, sh in enumerate(self.shards):
            sh.reset()
            self._rows[g] = []
        self.ntotal = 0

    def search(self, q, k, D=None, I=None):
        """NumPy in -> NumPy out (``D=`` / ``I=``: preallocated result arrays, as ``faiss.Index.search`` accepts);
        CUDA tensors on the first shard's device in -> CUDA tensors out."""
        import ctypes
        from ._lib import HAC_MAX_K, check
        from .index import _is_torch_tensor
        k = int(k)
        if k <= 0 or k > HAC_MAX_K:
            raise ValueError("search: k=%d outside [1, %d]" % (k, HAC_MAX_K))
        check(self._lib.hac_shards_set_exchange(self._grp, 1 if self.threshold_exchange else 0), "hac_shards_set_exchange")
        if _is_torch_tensor(q) and q.is_cuda:
            import torch
            assert q.dim() == 2 and q.shape[1] == self.d, "search: expected [nq, %d]" % self.d
            if q.device.index != self.devices[0]:
                raise ValueError("search: tensor lives on cuda:%s, first shard on cuda:%d" % (q.device.index, self.devices[0]))
            q = q.contiguous().float()
            D = torch.empty((q.shape[0], k), dtype=torch.float32, device=q.device)
            I = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
            stream = torch.cuda.current_stream(q.device).cuda_stream
            check(self._lib.hac_shards_search_device(self._grp, q.shape[0], q.data_ptr(), k, D.data_ptr(), I.data_ptr(),
                                                     stream), "hac_shards_search_device")
        else:
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
            check(self._lib.hac_shards_search(self._grp, q.shape[0], q.ctypes.data, k, D.ctypes.data, I.ctypes.data),
                  "hac_shards_search")
        ph = (ctypes.c_float * 4)()
        self._lib.hac_shards_last_phases(self._grp, ph, 4)
        # host-side milestones of the search, ms since the call began
        self.last_phase_ms = {"fastest_shard_done": ph[3], "slowest_shard_done": ph[0], "merge_enqueued": ph[1],
                              "results_on_host": ph[2]}
        return D, I

    def stats(self):
        return [sh.stats() for sh in self.shards]

    def close(self):
        if getattr(self, "_grp", None) is not None and self._grp.value:
            self._lib.hac_shards_destroy(self._grp)      # before the shards it borrows
            self._grp = None
            self._pool.shutdown(wait=True)
        for sh in self.shards:
            sh.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def index_cpu_to_gpu_multiple(vres, vdev, cpu_index, co=None):
    devices = [int(v) for v in vdev]
    if not devices:
        raise ValueError("index_cpu_to_gpu_multiple: empty device list")
    if getattr(cpu_index, "ntotal", 0):
        raise NotImplementedError("cloning a populated index is not needed by the reference path")
    if len(devices) == 1 or (co is not None and not co.shard):
        return IndexFlatIP(cpu_index.d, devices[0])     # replica mode degenerates to one device here
    return ShardedInProcessIndex(cpu_index.d, devices)


def index_cpu_to_gpu(res, device, cpu_index, co=None):
    return IndexFlatIP(cpu_index.d, int(device))


def index_cpu_to_all_gpus(cpu_index, co=None, ngpu=-1):
    n = get_num_gpus() if ngpu < 0 else ngpu
    c = co or GpuMultipleClonerOptions()
    return index_cpu_to_gpu_multiple([None] * n, list(range(n)), cpu_index, c)
