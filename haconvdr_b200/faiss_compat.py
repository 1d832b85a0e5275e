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

from .index import FlatIPIndex, merge_topk_device


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

    def search(self, q, k):
        return self._get().search(q, k)

    def reset(self):
        if self._impl is not None:
            self._impl.reset()

    def stats(self):
        return self._get().stats()


class ShardedInProcessIndex:
    """faiss IndexShards equivalent inside one process: rows of every ``add`` are split into
    contiguous ranges, one per device; ``search`` runs on every device, gathers the G x Q x k
    candidates on the first device (peer copies over NVLink) and merges them there."""

    def __init__(self, d: int, devices):
        self.d = int(d)
        self.devices = [int(v) for v in devices]
        self.shards = [FlatIPIndex(self.d, dev) for dev in self.devices]
        self._ids = [np.zeros(0, np.int64) for _ in self.devices]
        self.ntotal = 0
        self.is_trained = True

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, "add: expected [n, %d] float32" % self.d
        n, G = x.shape[0], len(self.shards)
        bounds = [(g * n) // G for g in range(G + 1)]
        for g, sh in enumerate(self.shards):
            lo, hi = bounds[g], bounds[g + 1]
            if hi > lo:
                sh.add(x[lo:hi])
                self._ids[g] = np.concatenate([self._ids[g], np.arange(self.ntotal + lo, self.ntotal + hi)])
                sh.set_id_table(self._ids[g])
        self.ntotal += n

    def reset(self):
        for g, sh in enumerate(self.shards):
            sh.reset()
            self._ids[g] = np.zeros(0, np.int64)
        self.ntotal = 0

    def search(self, q, k):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.d, "search: expected [nq, %d] float32" % self.d
        dev0 = torch.device("cuda", self.devices[0])
        q_host = torch.from_numpy(q).pin_memory()

        def one(dev, sh):                      # one host thread per shard, as faiss IndexShards does;
            with torch.cuda.device(dev):       # the C-ABI call releases the GIL, so the devices run concurrently
                qd = q_host.to(torch.device("cuda", dev), non_blocking=True)
                return sh.search(qd, k)

        with ThreadPoolExecutor(len(self.shards)) as pool:
            parts = list(pool.map(one, self.devices, self.shards))
        D = torch.stack([p[0].to(dev0) for p in parts])
        I = torch.stack([p[1].to(dev0) for p in parts])
        with torch.cuda.device(dev0):
            Dm, Im = merge_topk_device(D, I, k)
        return Dm.cpu().numpy(), Im.cpu().numpy()


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
