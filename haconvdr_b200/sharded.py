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
        self._symm = None          # (handle, buffer, slot_bytes) of the symmetric result buffer (2 slots)
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
        nq_, k_ = int(q.shape[0]), int(k)
        use_p2p = (self.world_size > 1 and self._on_gpu() and self.exchange in ("auto", "p2p")
                   and self._ensure_symm(nq_ * k_ * 12 + 16))
        err = None                                                 # a failure of THIS rank's local search
        if use_p2p:
            from .index import merge_topk_peers_device
            hdl, buf, slot_bytes = self._symm
            nq = nq_
            slot = self._step & 1
            self._step += 1
            d_bytes, i_bytes = (nq * k * 4 + 15) // 16 * 16, nq * k * 8      # ids start 16-byte aligned
            off = slot * slot_bytes
            D_loc = buf[off:off + nq * k * 4].view(torch.float32).view(nq, k)
            I_loc = buf[off + d_bytes:off + d_bytes + i_bytes].view(torch.int64).view(nq, k)
            # the shards publish their best scores at a few ranks to each other while they scan; together they bound the
            # GLOBAL k-th best, so every shard emits / rescores ~1/G of the candidates it would alone; its list then
            # holds only its share of the global top-k (fillers behind), which is all the merge below needs
            self._epoch = self._epoch % 0x0FFFFFFF + 1             # 1 .. 2^28 - 1, identical on all ranks
            armed = self.threshold_exchange and self._ensure_thr(nq)
            try:
                self.local.set_option("exchange_epoch", self._epoch if armed else 0)
                self.local.search(q, k, out=(D_loc, I_loc))      # results land in the symmetric buffer
            except Exception as e:   # noqa: BLE001 - the peers are about to wait for this rank: keep the protocol going
                err = e
                D_loc.fill_(-3.4028234663852886e38)
                I_loc.fill_(-1)
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
            try:
                D, I = self.local.search(q, k)
            except Exception as e:   # noqa: BLE001
                if self.world_size == 1:
                    raise
                err = e
                D = torch.full((nq_, k_), -3.4028234663852886e38, dtype=torch.float32, device=self._gather_device())
                I = torch.full((nq_, k_), -1, dtype=torch.int64, device=self._gather_device())
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
        if self.world_size > 1:
            # every rank went through the same collectives whatever happened locally; now agree on the outcome, so
            # that a failure on one rank (overflow, out of memory) raises on ALL ranks instead of hanging the others
            if not self._agree(err is None):
                if err is not None:
                    raise err
                raise RuntimeError("sharded search failed on another rank")
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

    def _agree(self, ok: bool) -> bool:
        """True iff `ok` on every rank (one tiny all-reduce; every rank must call it at the same point)."""
        torch, dist = self._torch, self._dist
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self._gather_device())
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def _gather_device(self):
        return self._device() if self._on_gpu() else self._torch.device("cpu")

    def _ensure_symm(self, need_bytes: int) -> bool:
        """Symmetric-memory result buffer (2 slots), sized by capacity: it only ever grows, so a change of batch size
        or k (online turns, the k sweep) reuses it.  False if unavailable on ANY rank (-> NCCL path on all ranks)."""
        if self._symm_failed:
            return False
        if self._symm is not None and self._symm[2] >= need_bytes:
            return True
        torch, dist = self._torch, self._dist
        slot_bytes = 1 << 20
        while slot_bytes < need_bytes:
            slot_bytes *= 2
        new, why = None, ""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            if self._symm is not None:
                # peers may still be reading the old buffer (merge kernel of the previous search)
                torch.cuda.synchronize(self._device())
                dist.barrier(group=self.group)
            with torch.cuda.device(self._device()):
                buf = symm_mem.empty((2 * slot_bytes,), dtype=torch.uint8, device=self._device())
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            new = (hdl, buf, slot_bytes)
        except Exception as e:          # noqa: BLE001 - decided together with the other ranks below
            why = str(e)
        if self._agree(new is not None):
            self._symm = new            # `_step` keeps running: slot parity stays aligned across ranks
            return True
        if self.exchange == "p2p":
            raise RuntimeError("symmetric memory unavailable on at least one rank (%s)" % (why or "peer failure"))
        self._symm_failed = True
        self._symm = None
        import warnings
        warnings.warn("symmetric memory unavailable (%s); using the NCCL all-gather exchange" % (why or "peer failure"))
        return False

    def _ensure_thr(self, nq: int) -> bool:
        """Symmetric-memory buffer of per-query threshold words, registered with the local engine
        (HAC_EXCHANGE_WORDS_PER_QUERY words per query).  Armed on all ranks or on none."""
        from ._lib import HAC_EXCHANGE_WORDS_PER_QUERY as WPQ
        if not self.threshold_exchange:
            return False
        if self._thr is not None and self._thr[2] >= nq:
            return True
        torch, dist = self._torch, self._dist
        new, why = None, ""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            cap = 4096
            while cap < nq:
                cap *= 2
            if self._thr is not None:
                torch.cuda.synchronize(self._device())
                dist.barrier(group=self.group)
            with torch.cuda.device(self._device()):
                buf = symm_mem.empty((cap * WPQ,), dtype=torch.int64, device=self._device())
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            torch.cuda.synchronize(self._device())
            new = (hdl, buf, cap)
        except Exception as e:          # noqa: BLE001 - decided together with the other ranks below
            why = str(e)
        if self._agree(new is not None):
            hdl, buf, cap = new
            hdl.barrier(channel=1)                               # every rank's buffer is zeroed before anyone reads it
            ptrs = list(hdl.buffer_ptrs)
            self.local.set_threshold_exchange(ptrs[self.rank], [p for r, p in enumerate(ptrs) if r != self.rank], cap * WPQ)
            self._thr = new
            return True
        import warnings
        warnings.warn("threshold exchange unavailable (%s): every shard keeps its own thresholds (still exact)" % (
            why or "peer failure"))
        self.threshold_exchange = False        # not available: decided identically on every rank
        return False

    def _on_gpu(self) -> bool:
        return hasattr(self.local, "device") and hasattr(self.local, "_h")

    def _device(self):
        return self._torch.device("cuda", self.local.device)
