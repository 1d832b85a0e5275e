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
    """faiss IndexShards equivalent inside one process (`index_cpu_to_gpu_multiple` with ``co.shard = True``,
    `/root/reference/src/test_HAConvDR_topiocqa.py:55-66`): rows of every ``add`` are split into contiguous ranges,
    one per device.  ``search`` is the same protocol as one process per GPU (``haconvdr_b200.sharded``):

    * one persistent host thread and one CUDA stream per shard (the C-ABI calls release the GIL, so the devices run
      concurrently); queries go up once per device from a reused page-locked buffer;
    * the shards publish their best scores to each other WHILE they scan (``hac_set_threshold_exchange``: peer-mapped
      device pointers, ``hac_enable_peer_access``), so every shard rescores only its share of the global top-k;
    * results stay in per-device buffers; ONE merge kernel on the first device reads all of them in place over NVLink
      (``hac_merge_topk_peers_device``) - no gather copies, no host round trip before the final D2H."""

    def __init__(self, d: int, devices):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        from . import _lib
        self.d = int(d)
        self.devices = [int(v) for v in devices]
        self.shards = [FlatIPIndex(self.d, dev) for dev in self.devices]
        self._rows = [[] for _ in self.devices]          # per shard: (global_start, n) of every add
        self.ntotal = 0
        self.is_trained = True
        self.threshold_exchange = True
        self._torch = torch
        self._pool = ThreadPoolExecutor(len(self.devices), thread_name_prefix="hac-shard")
        self._streams = [torch.cuda.Stream(device=dev) for dev in self.devices]
        self._events = [torch.cuda.Event() for _ in self.devices]
        self._peer_ok = True
        for a in sorted(set(self.devices)):              # kernels on a read (merge) and write (exchange) memory of b
            for b in sorted(set(self.devices)):
                if a != b and _lib.lib().hac_enable_peer_access(a, b) != 0:
                    self._peer_ok = False                # no peer access on this system: gather copies, no exchange
        self._epoch = 0
        self._words = None        # per shard: int64 CUDA tensor of exchange words, capacity in queries
        self._words_cap = 0
        self._res = None          # per shard: (D [cap] float32, I [cap] int64) result buffers
        self._res_cap = 0
        self._q_pinned = None
        self._out_pinned = None
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

    def _ensure_buffers(self, nq, k):
        torch = self._torch
        from ._lib import HAC_EXCHANGE_WORDS_PER_QUERY as WPQ
        if nq * k > self._res_cap:
            cap = 1 << 16
            while cap < nq * k:
                cap *= 2
            self._res = [(torch.empty(cap, dtype=torch.float32, device="cuda:%d" % dev),
                          torch.empty(cap, dtype=torch.int64, device="cuda:%d" % dev)) for dev in self.devices]
            self._res_cap = cap
        if self.threshold_exchange and self._peer_ok and nq > self._words_cap:
            cap = 4096
            while cap < nq:
                cap *= 2
            self._words = [torch.zeros(cap * WPQ, dtype=torch.int64, device="cuda:%d" % dev) for dev in self.devices]
            for dev in set(self.devices):
                torch.cuda.synchronize(dev)
            for g, sh in enumerate(self.shards):
                sh.set_threshold_exchange(self._words[g].data_ptr(),
                                          [w.data_ptr() for h, w in enumerate(self._words) if h != g], cap * WPQ)
            self._words_cap = cap

    def search(self, q, k):
        import time
        torch = self._torch
        from .index import merge_topk_peers_device
        t_start = time.perf_counter()
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.d, "search: expected [nq, %d] float32" % self.d
        nq, k, G = q.shape[0], int(k), len(self.shards)
        if self._q_pinned is None or self._q_pinned.shape[0] < nq:
            self._q_pinned = torch.empty((max(nq, 256), self.d), dtype=torch.float32).pin_memory()
        qp = self._q_pinned[:nq]
        qp.copy_(torch.from_numpy(q))
        self._ensure_buffers(nq, k)
        self._epoch = self._epoch % 0x0FFFFFFF + 1
        armed = self.threshold_exchange and self._peer_ok and self._words is not None and G > 1
        outs = [(D[:nq * k].view(nq, k), I[:nq * k].view(nq, k)) for D, I in self._res]

        t_shard = [None] * G

        def one(g):
            t0 = time.perf_counter()
            dev, sh = self.devices[g], self.shards[g]
            with torch.cuda.device(dev), torch.cuda.stream(self._streams[g]):
                qd = qp.to(torch.device("cuda", dev), non_blocking=True)
                sh.set_option("exchange_epoch", self._epoch if armed else 0)
                t1 = time.perf_counter()
                try:
                    sh.search(qd, k, out=outs[g])
                finally:
                    sh.set_option("exchange_epoch", 0)
                self._events[g].record(self._streams[g])
            t_shard[g] = (t0 - t_start, t1 - t_start, time.perf_counter() - t_start)
        t_prep = time.perf_counter()
        list(self._pool.map(one, range(G)))              # re-raises the first shard failure after all have finished
        t_threads = time.perf_counter()
        dev0 = self.devices[0]
        with torch.cuda.device(dev0), torch.cuda.stream(self._streams[0]):
            for ev in self._events[1:]:
                self._streams[0].wait_event(ev)
            if self._peer_ok:
                Dm, Im = merge_topk_peers_device([o[0].data_ptr() for o in outs], [o[1].data_ptr() for o in outs],
                                                 nq, k, k, torch.device("cuda", dev0))
            else:
                Dm, Im = merge_topk_device(torch.stack([o[0].to("cuda:%d" % dev0) for o in outs]),
                                           torch.stack([o[1].to("cuda:%d" % dev0) for o in outs]), k)
            if self._out_pinned is None or self._out_pinned[0].numel() < nq * k:
                n = max(nq * k, 1 << 16)
                self._out_pinned = (torch.empty(n, dtype=torch.float32).pin_memory(),
                                    torch.empty(n, dtype=torch.int64).pin_memory())
            Dh, Ih = self._out_pinned[0][:nq * k].view(nq, k), self._out_pinned[1][:nq * k].view(nq, k)
            Dh.copy_(Dm, non_blocking=True)
            Ih.copy_(Im, non_blocking=True)
            self._streams[0].synchronize()
        t_merge = time.perf_counter()
        out = Dh.numpy().copy(), Ih.numpy().copy()
        # host-side phases of the last search, ms since entry: query staging, per-shard (thread start, C call start, C call
        # end), merge + D2H, result copies
        self.last_phase_ms = {"prep": 1e3 * (t_prep - t_start), "shards": [[round(1e3 * v, 3) for v in t] for t in t_shard],
                              "threads_done": 1e3 * (t_threads - t_start), "merge_d2h_done": 1e3 * (t_merge - t_start),
                              "total": 1e3 * (time.perf_counter() - t_start)}
        return out

    def stats(self):
        return [sh.stats() for sh in self.shards]

    def close(self):
        self._pool.shutdown(wait=True)
        for sh in self.shards:
            sh.close()


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
