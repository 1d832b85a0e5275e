"""Corpus sharding across GPUs, one process per GPU (``torch.distributed``).

The reference shards its index with faiss ``IndexShards`` (``co.shard = True``,
`/root/reference/src/test_HAConvDR_topiocqa.py:55-66`): contiguous row ranges per device, a
host-side merge of the per-shard top-k lists.  Here rank ``r`` of ``G`` owns global rows
``[r*N/G, (r+1)*N/G)`` (ids = global offsets through ``set_id_base``), queries are replicated,
and one search is: local exact top-k on every rank -> ONE all-gather of the ``Q x k``
(score, id) candidates -> k-way merge kernel on the device.  The corpus rows never cross a link.

Two exchange implementations:
  * ``"p2p"`` (default on GPUs when symmetric memory is available): every rank writes its local
    result straight into a symmetric-memory buffer, one cross-GPU barrier, then ONE merge kernel reads
    all G peers' lists in place over NVLink (``hac_merge_topk_peers_device``) - no NCCL launch, no
    staging copy of the candidates.  Buffers alternate between two slots per search, so a single
    barrier per search orders both the reads and the next overwrite.
  * ``"nccl"``: ``all_gather_into_tensor`` of scores and ids, then the merge kernel (also the path
    the gloo CPU tests drive with an injected merge).

The local index and the merge are injected so the partitioning / gather / ordering logic can be
exercised under ``gloo`` on CPU with the oracle standing in (tests only); the defaults are the
CUDA engine and the device merge kernel - there is no CPU path in the product.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_rows: int, world_size: int):
    """Contiguous partition of ``n_rows`` global rows: bounds[r] .. bounds[r+1] belongs to rank r."""
    return [(r * n_rows) // world_size for r in range(world_size + 1)]


class ShardedFlatIPIndex:
    def __init__(self, d: int, local_index=None, device=None, group=None, merge=None, exchange="auto"):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.d = int(d)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        if local_index is None:
            from .index import FlatIPIndex
            local_index = FlatIPIndex(self.d, device if device is not None else torch.cuda.current_device())
        self.local = local_index
        if merge is None:
            from .index import merge_topk_device
            merge = merge_topk_device
        self._merge = merge
        self.ntotal = 0            # global rows
        self._lo = 0               # global offset of the next add's first row
        self._bases = []           # (local_row_start, global_row_start, n) per add
        self._pinned_q = None      # page-locked staging for host queries
        self._pinned_out = None    # page-locked staging for host results
        self.exchange = exchange   # "auto" | "p2p" | "nccl"
        self._symm = None          # (handle, buffer, nq, k) of the symmetric result buffer
        self._thr = None           # (handle, buffer, capacity) of the symmetric threshold-exchange buffer
        self._epoch = 0            # search counter, identical on all ranks (tags the exchanged thresholds)
        self.threshold_exchange = True   # p2p path: shards share their k-th-best bounds while they scan
        self._symm_failed = False
        self._step = 0
        self.profile = False       # True: record per-phase CUDA-event times of each search (diagnostics)
        self.last_phase_ms = None

    # -- building the shard ---------------------------------------------------------------------
    def _after_add(self, lo_global: int, n_local: int):
        start_local = sum(b[2] for b in self._bases)
        self._bases.append((start_local, lo_global, n_local))
        if len(self._bases) == 1:
            self.local.set_id_base(lo_global)
        else:                                   # several adds: ids are no longer base + row
            table = np.concatenate([np.arange(g, g + n, dtype=np.int64) for _, g, n in self._bases])
            self.local.set_id_table(table)

    def add(self, x):
        """Every rank passes the same global block (as the reference's single process does);
        each keeps its contiguous slice."""
        n = int(x.shape[0])
        b = shard_bounds(n, self.world_size)
        lo, hi = b[self.rank], b[self.rank + 1]
        if hi > lo:
            self.local.add(x[lo:hi])
            self._after_add(self._lo + lo, hi - lo)
        self._lo += n
        self.ntotal += n

    def add_synthetic(self, n_global: int, seed: int = 42, dist_kind: int = 0):
        b = shard_bounds(n_global, self.world_size)
        lo, hi = b[self.rank], b[self.rank + 1]
        if hasattr(self.local, "reserve"):
            self.local.reserve(hi - lo)
        self.local.add_synthetic(hi - lo, seed=seed, row0=self._lo + lo, dist=dist_kind)
        self._after_add(self._lo + lo, hi - lo)
        self._lo += n_global
        self.ntotal += n_global

    def reset(self):
        self.local.reset()
        self.ntotal, self._lo, self._bases = 0, 0, []

    # -- search ---------------------------------------------------------------------------------
    def search(self, q, k: int, D=None, I=None):
        """``q``: the same queries on every rank (numpy -> numpy results, CUDA tensor -> CUDA tensors).
        Returns the merged global (D [Q,k], I [Q,k]) on every rank.  ``D=``, ``I=``: optional preallocated
        NumPy result arrays for the host path (page-locked ones receive the results by direct DMA).

        Host queries are copied straight from the caller's array when it is page-locked, otherwise
        through a reused pinned staging buffer; host results come back through pinned buffers too."""
        torch, dist = self._torch, self._dist
        out_D, out_I = D, I                                        # caller's result arrays (host path), may be None
        as_numpy = not (hasattr(q, "is_cuda") and q.is_cuda)
        if as_numpy and self._on_gpu():
            if self.world_size == 1:
                return self.local.search(q, k, D=out_D, I=out_I)   # plain host-buffer C-ABI call
            qh = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
            if not qh.is_pinned():
                if self._pinned_q is None or self._pinned_q.shape != qh.shape:
                    self._pinned_q = torch.empty(qh.shape, dtype=torch.float32).pin_memory()
                self._pinned_q.copy_(qh)
                qh = self._pinned_q
            q = qh.to(self._device(), non_blocking=True)
        prof = self.profile and self._on_gpu() and self.world_size > 1
        if prof:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        use_p2p = (self.world_size > 1 and self._on_gpu() and self.exchange in ("auto", "p2p")
                   and self._ensure_symm(int(q.shape[0]), int(k)))
        if use_p2p:
            from .index import merge_topk_peers_device
            hdl, buf, nq, _ = self._symm
            slot = self._step & 1
            self._step += 1
            d_bytes, i_bytes = nq * k * 4, nq * k * 8
            slot_bytes = (d_bytes + i_bytes + 255) // 256 * 256
            off = slot * slot_bytes
            D_loc = buf[off:off + d_bytes].view(torch.float32).view(nq, k)
            I_loc = buf[off + d_bytes:off + d_bytes + i_bytes].view(torch.int64).view(nq, k)
            # the shards publish their ceil(k/G)-th best scores to each other while they scan; the minimum bounds the
            # GLOBAL k-th best, so every shard emits / rescores ~1/G of the candidates it would alone; its list then
            # holds only its share of the global top-k (fillers behind), which is all the merge below needs
            self._epoch += 1
            armed = self.threshold_exchange and self._ensure_thr(nq)
            self.local.set_option("exchange_epoch", self._epoch if armed else 0)
            try:
                self.local.search(q, k, out=(D_loc, I_loc))      # results land in the symmetric buffer
            finally:
                if armed:
                    self.local.set_option("exchange_epoch", 0)
            if prof:
                ev[1].record()
            hdl.barrier(channel=0)                               # every rank's lists are complete
            if prof:
                ev[2].record()
            D, I = merge_topk_peers_device([p + off for p in hdl.buffer_ptrs],
                                           [p + off + d_bytes for p in hdl.buffer_ptrs], nq, k, k, self._device())
            if prof:
                ev[3].record()
                torch.cuda.synchronize()
                self.last_phase_ms = {"local_search": ev[0].elapsed_time(ev[1]), "barrier": ev[1].elapsed_time(ev[2]),
                                      "p2p_merge": ev[2].elapsed_time(ev[3])}
        else:
            D, I = self.local.search(q, k)
            if prof:
                ev[1].record()
        if self.world_size > 1 and not use_p2p:
            if not torch.is_tensor(D):
                D, I = torch.from_numpy(np.ascontiguousarray(D)), torch.from_numpy(np.ascontiguousarray(I))
            nq = D.shape[0]
            Dg = torch.empty((self.world_size * nq, k), dtype=D.dtype, device=D.device)
            Ig = torch.empty((self.world_size * nq, k), dtype=I.dtype, device=I.device)
            dist.all_gather_into_tensor(Dg, D.contiguous(), group=self.group)   # rank-major: [G][Q][k]
            dist.all_gather_into_tensor(Ig, I.contiguous(), group=self.group)
            if prof:
                ev[2].record()
            D, I = self._merge(Dg.view(self.world_size, nq, k), Ig.view(self.world_size, nq, k), k)
            if prof:
                ev[3].record()
                torch.cuda.synchronize()
                self.last_phase_ms = {"local_search": ev[0].elapsed_time(ev[1]), "all_gather": ev[1].elapsed_time(ev[2]),
                                      "merge": ev[2].elapsed_time(ev[3])}
        if as_numpy and torch.is_tensor(D):
            if D.is_cuda:
                if out_D is not None and out_I is not None:
                    tD, tI = torch.from_numpy(out_D), torch.from_numpy(out_I)
                    if tD.is_pinned() and tI.is_pinned():          # caller's page-locked arrays: direct DMA
                        tD.copy_(D, non_blocking=True)
                        tI.copy_(I, non_blocking=True)
                        torch.cuda.current_stream(D.device).synchronize()
                        return out_D, out_I
                if self._pinned_out is None or self._pinned_out[0].shape != D.shape:
                    self._pinned_out = (torch.empty(D.shape, dtype=D.dtype).pin_memory(),
                                        torch.empty(I.shape, dtype=I.dtype).pin_memory())
                self._pinned_out[0].copy_(D, non_blocking=True)
                self._pinned_out[1].copy_(I, non_blocking=True)
                torch.cuda.current_stream(D.device).synchronize()
                if out_D is not None and out_I is not None:
                    out_D[...] = self._pinned_out[0].numpy()
                    out_I[...] = self._pinned_out[1].numpy()
                    return out_D, out_I
                return self._pinned_out[0].numpy().copy(), self._pinned_out[1].numpy().copy()
            return D.numpy(), I.numpy()
        return D, I

    def _ensure_symm(self, nq: int, k: int) -> bool:
        """Symmetric-memory result buffer (2 slots) for (nq, k); False if unavailable (-> NCCL path)."""
        if self._symm_failed:
            return False
        if self._symm is not None and self._symm[2] == nq and self._symm[3] == k:
            return True
        torch, dist = self._torch, self._dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            slot_bytes = (nq * k * 12 + 255) // 256 * 256
            with torch.cuda.device(self._device()):
                buf = symm_mem.empty((2 * slot_bytes,), dtype=torch.uint8, device=self._device())
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            self._symm = (hdl, buf, nq, k)
            self._step = 0
            return True
        except Exception as e:          # no symmetric memory on this system: use the NCCL exchange
            if self.exchange == "p2p":
                raise
            self._symm_failed = True
            import warnings
            warnings.warn("symmetric memory unavailable (%s); using the NCCL all-gather exchange" % (e,))
            return False

    def _ensure_thr(self, nq: int) -> bool:
        """Symmetric-memory buffer of per-query threshold words, registered with the local engine."""
        if self._thr is not None and self._thr[2] >= nq:
            return True
        torch, dist = self._torch, self._dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            cap = max(int(nq), 4096)
            with torch.cuda.device(self._device()):
                buf = symm_mem.empty((cap,), dtype=torch.int64, device=self._device())
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            torch.cuda.synchronize(self._device())
            hdl.barrier(channel=1)                               # every rank's buffer is zeroed before anyone reads it
            ptrs = list(hdl.buffer_ptrs)
            mine = ptrs[self.rank]
            peers = [p for r, p in enumerate(ptrs) if r != self.rank]
            self.local.set_threshold_exchange(mine, peers, cap)
            self._thr = (hdl, buf, cap)
            return True
        except Exception as e:          # not available: every shard keeps its own thresholds (still exact)
            import warnings
            warnings.warn("threshold exchange unavailable (%s)" % (e,))
            self.threshold_exchange = False
            return False

    def _on_gpu(self) -> bool:
        return hasattr(self.local, "device") and hasattr(self.local, "_h")

    def _device(self):
        return self._torch.device("cuda", self.local.device)
