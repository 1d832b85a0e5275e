"""CPU, world_size 2 over gloo: the N>1 host logic of haconvdr_b200.sharded (row partition, id bases,
all-gather layout, merge order) with the oracle index and a NumPy merge injected in place of the
CUDA engine.  The CUDA pieces themselves are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard:
    """oracle FlatIP + the id-translation surface of FlatIPIndex (test stand-in for one GPU shard)."""

    def __init__(self, d):
        from oracle.flat_ip import FlatIP
        self._ix, self.d, self._base, self._table = FlatIP(d), d, 0, None

    @property
    def ntotal(self):
        return self._ix.ntotal

    def add(self, x):
        self._ix.add(x)

    def reset(self):
        self._ix.reset()
        self._base, self._table = 0, None

    def set_id_base(self, b):
        self._base = int(b)

    def set_id_table(self, t):
        self._table = np.asarray(t, np.int64)

    def search(self, q, k):
        D, I = self._ix.search(np.asarray(q, np.float32), k)
        valid = I >= 0
        out = I.copy()
        out[valid] = self._table[I[valid]] if self._table is not None else I[valid] + self._base
        return D, out


def numpy_merge(Dg, Ig, k):
    """[G,Q,k] -> [Q,k] by (score desc, id asc), fillers (id -1) last."""
    D = Dg.numpy().transpose(1, 0, 2).reshape(Dg.shape[1], -1)
    I = Ig.numpy().transpose(1, 0, 2).reshape(Ig.shape[1], -1)
    key_i = np.where(I < 0, np.iinfo(np.int64).max, I)
    order = np.lexsort((key_i, -D.astype(np.float64)), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(D, order, 1)), torch.from_numpy(np.take_along_axis(I, order, 1))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from haconvdr_b200.sharded import ShardedFlatIPIndex, shard_bounds
        from oracle.flat_ip import brute_force_fp64
        rng = np.random.default_rng(123)                      # same data on every rank
        blocks = [rng.standard_normal((n, 64)).astype(np.float32) for n in (501, 333)]
        q = rng.standard_normal((17, 64)).astype(np.float32)
        x = np.concatenate(blocks, 0)
        idx = ShardedFlatIPIndex(64, local_index=OracleShard(64), merge=numpy_merge)
        for b in blocks:
            idx.add(b)
        assert idx.ntotal == 834
        lo0, hi0 = shard_bounds(501, world)[rank], shard_bounds(501, world)[rank + 1]
        lo1, hi1 = shard_bounds(333, world)[rank], shard_bounds(333, world)[rank + 1]
        assert idx.local.ntotal == (hi0 - lo0) + (hi1 - lo1)
        D, I = idx.search(q, 20)
        D64, I64 = brute_force_fp64(q, x, 20)
        assert np.array_equal(np.asarray(I), I64), (rank, np.asarray(I)[0], I64[0])
        np.testing.assert_allclose(np.asarray(D), D64, rtol=1e-5, atol=1e-5)
        # k larger than a shard and than the corpus: fillers must stay last after the merge
        idx.reset()
        idx.add(blocks[0][:5])
        D, I = idx.search(q, 8)
        D64, I64 = brute_force_fp64(q, blocks[0][:5], 8)
        assert np.array_equal(np.asarray(I), I64)
        np.save(os.path.join(tmp, "ok_%d.npy" % rank), np.asarray([1]))
    finally:
        dist.destroy_process_group()


def test_sharded_index_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(2))


class FlakyShard(OracleShard):
    """Raises once, on the chosen rank only (stands in for an overflow / out-of-memory of one GPU's local search)."""

    def __init__(self, d, fail_now):
        super().__init__(d)
        self.fail_now = fail_now

    def search(self, q, k):
        if self.fail_now():
            raise RuntimeError("local search failed on this rank")
        return super().search(q, k)


def _worker_failure(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from haconvdr_b200.sharded import ShardedFlatIPIndex
        from oracle.flat_ip import brute_force_fp64
        rng = np.random.default_rng(7)
        x = rng.standard_normal((400, 64)).astype(np.float32)
        q = rng.standard_normal((9, 64)).astype(np.float32)
        calls = {"n": 0}

        def fail_now():
            calls["n"] += 1
            return rank == 1 and calls["n"] == 2             # the second search fails, on rank 1 only
        idx = ShardedFlatIPIndex(64, local_index=FlakyShard(64, fail_now), merge=numpy_merge)
        idx.add(x)
        D64, I64 = brute_force_fp64(q, x, 10)
        D, I = idx.search(q, 10)
        assert np.array_equal(np.asarray(I), I64)
        # one rank fails: EVERY rank must raise (the failing one its own error), none may hang in a collective
        try:
            idx.search(q, 10)
            raised = None
        except RuntimeError as e:
            raised = str(e)
        assert raised is not None, "rank %d returned a result although rank 1 failed" % rank
        assert ("this rank" in raised) == (rank == 1), (rank, raised)
        # and the protocol is still aligned afterwards: the next search works on all ranks
        D, I = idx.search(q, 10)
        assert np.array_equal(np.asarray(I), I64)
        np.save(os.path.join(tmp, "fail_ok_%d.npy" % rank), np.asarray([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_a_failure_on_one_rank_raises_on_every_rank_gloo(tmp_path):
    """ADVICE (round 1): a local failure on one rank must not leave the others waiting in a collective."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_failure, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "fail_ok_%d.npy" % r)) for r in range(2))


def test_shard_bounds_cover_rows_exactly():
    from haconvdr_b200.sharded import shard_bounds
    for n in (0, 1, 7, 25_700_592, 54_573_064):
        for g in (1, 2, 4, 8):
            b = shard_bounds(n, g)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(g))
            assert max(b[i + 1] - b[i] for i in range(g)) - min(b[i + 1] - b[i] for i in range(g)) <= 1


def test_share_rule_bounds_the_global_kth_best():
    """The bound the shards exchange while they scan (hac_set_threshold_exchange): if every one of G shards holds at
    least ceil(k/G) rows scoring >= L_r, then min_r L_r <= the global k-th best score - for any split of the rows,
    and with equality-ish tightness for evenly mixed shards."""
    rng = np.random.default_rng(11)
    for G, k, n in ((2, 100, 5000), (8, 100, 4000), (8, 10, 999), (3, 7, 50), (16, 100, 3000)):
        s = rng.standard_normal(n).astype(np.float32)
        ks = -(-k // G)
        for split in ("even", "skewed"):
            if split == "even":
                shards = np.array_split(rng.permutation(s), G)
            else:                                   # all the best rows in one shard: the bound must stay valid (if weak)
                order = np.sort(s)[::-1]
                cuts = np.sort(rng.choice(np.arange(ks * 2, n - ks * 2), G - 1, replace=False))
                shards = np.split(order, cuts)
            if any(len(x) < ks for x in shards):
                continue
            bound = min(np.sort(x)[::-1][ks - 1] for x in shards)
            kth = np.sort(s)[::-1][k - 1]
            assert bound <= kth
            if split == "even" and n >= 50 * k:
                assert np.sum(s >= bound) <= 4 * k     # close to the k-th best for mixed shards


def _exchange_ranks(k, G):
    """The ranks every shard publishes (mirror of search_batch_i8 in csrc/hac_api.cu)."""
    ks = -(-k // G)
    r = [min(k, max(1, int(m * ks + 0.5))) for m in (0.3, 0.5, 0.7, 0.85, 1.0, 1.2, 1.5, 2.0, 3.0, 5.0)]
    r += [min(k, max(1, k // 2)), k]
    return sorted(set(r))[-12:]


def _claims_bound(shards, ranks, k):
    """refresh_kernel's rule: the largest published score T at which sum_r max{c : L_r(c) >= T} >= k (None if none)."""
    claims = []
    for x in shards:
        top = np.sort(x)[::-1]
        claims.append([(c, top[c - 1]) for c in ranks if len(top) >= c])
    best = None
    for T in sorted({v for cl in claims for _, v in cl}, reverse=True):
        total = sum(max([c for c, v in cl if v >= T], default=0) for cl in claims)
        if total >= k:
            best = T
            break
    return best


def test_multi_rank_claims_bound_the_global_kth_best_and_beat_the_share_rule():
    """Cross-shard exchange of round 2: every shard publishes its best scores at a few ranks; the largest T at which
    the claims add up to k rows never exceeds the global k-th best (valid for ANY split of the rows), is at least as
    tight as the single-rank share rule, and for evenly mixed shards admits far fewer rows above it."""
    rng = np.random.default_rng(12)
    gain = []
    for G, k, n in ((2, 100, 5000), (8, 100, 40000), (8, 10, 999), (3, 7, 50), (16, 100, 30000), (4, 1, 100), (8, 128, 20000)):
        ranks = _exchange_ranks(k, G)
        assert ranks[-1] == k and len(ranks) <= 12 and all(1 <= c <= k for c in ranks)
        ks = -(-k // G)
        assert ks in ranks or G == 1
        for trial in range(6):
            s = rng.standard_normal(n).astype(np.float32)
            if trial % 2 == 0:
                shards = np.array_split(rng.permutation(s), G)
            else:
                order = np.sort(s)[::-1]
                cuts = np.sort(rng.choice(np.arange(1, n - 1), G - 1, replace=False))
                shards = np.split(order, cuts)
            kth = np.sort(s)[::-1][k - 1]
            T = _claims_bound(shards, ranks, k)
            assert T is None or T <= kth                           # no claim set reaching k = no bound (still valid)
            if T is not None and all(len(x) >= ks for x in shards):
                share = min(np.sort(x)[::-1][ks - 1] for x in shards)
                assert T >= share
                if trial % 2 == 0 and n >= 200 * k:
                    gain.append((np.sum(s >= share), np.sum(s >= T), k))
    # evenly mixed shards: rows above the bound (what a shard still has to rescore) shrink towards k
    assert np.mean([b / k for _, b, k in gain]) < np.mean([a / k for a, _, k in gain])
    assert np.mean([b / k for _, b, k in gain]) < 1.6
