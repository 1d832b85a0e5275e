"""GPU (-m gpu): the multi-shard paths.  The first tests run on ONE device (two shards = two handles on cuda:0, the
"peer" pointers are same-device buffers), so the exchange / peer-merge logic is exercised on a single-GPU box too;
the rest needs >= 2 devices (skipped otherwise): the real multi-GPU paths.
  * torchrun, one process per GPU, NCCL all-gather of Q x k candidates + device merge
    (haconvdr_b200.sharded.ShardedFlatIPIndex) against the fp64 arbiter;
  * the in-process faiss IndexShards stand-in (faiss_compat.index_cpu_to_gpu_multiple, n_gpu = 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["HAC_ROOT"])
from haconvdr_b200.sharded import ShardedFlatIPIndex
from haconvdr_b200 import FlatIPIndex
from oracle.flat_ip import brute_force_fp64
from oracle.compare import assert_parity
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(77)
blocks = [rng.standard_normal((n, 768), dtype=np.float32) for n in (30001, 17777)]
q = rng.standard_normal((150, 768), dtype=np.float32)
x = np.concatenate(blocks, 0)
idx = ShardedFlatIPIndex(768, FlatIPIndex(768, local))
for b in blocks:
    idx.add(b)
D, I = idx.search(q, 100)
D64, I64 = brute_force_fp64(q, x, 100)
x64 = x.astype(np.float64)
assert_parity(D64, I64, D, I, rtol=1e-5, ref_scores_of=lambda qi, ids: x64[ids] @ q[qi].astype(np.float64))
Dd, Id = idx.search(torch.from_numpy(q).cuda(), 100)
assert np.array_equal(Id.cpu().numpy(), I) and np.array_equal(Dd.cpu().numpy(), D)
# both exchange implementations (symmetric-memory P2P merge, NCCL all-gather + merge) agree bitwise
assert idx._symm is not None, "P2P exchange was not used"
idx.exchange = "nccl"
Dn, In = idx.search(q, 100)
assert np.array_equal(In, I) and np.array_equal(Dn, D)
idx.exchange = "p2p"
for _ in range(3):                      # alternating slots
    Dp, Ip = idx.search(q, 100)
    assert np.array_equal(Ip, I) and np.array_equal(Dp, D)
idx.exchange = "auto"
# threshold exchange over NVLink during the scan: same merged result, never more rescoring than without it
st_on = idx.local.stats()
idx.threshold_exchange = False
Dx, Ix = idx.search(q, 100)
st_off = idx.local.stats()
assert np.array_equal(Ix, I) and np.array_equal(Dx, D)
assert st_on["candidates_rescored"] <= st_off["candidates_rescored"], (st_on, st_off)
idx.threshold_exchange = True
Dx, Ix = idx.search(q, 100)
assert np.array_equal(Ix, I) and np.array_equal(Dx, D)
# faiss-style preallocated outputs, page-locked and ordinary
Dp = torch.empty((150, 100), dtype=torch.float32).pin_memory().numpy(); Ip = torch.empty((150, 100), dtype=torch.int64).pin_memory().numpy()
r = idx.search(q, 100, D=Dp, I=Ip)
assert r[0] is Dp and np.array_equal(Ip, I) and np.array_equal(Dp, D)
Do = np.empty((150, 100), np.float32); Io = np.empty((150, 100), np.int64)
idx.search(q, 100, D=Do, I=Io)
assert np.array_equal(Io, I) and np.array_equal(Do, D)

# synthetic shards reproduce the same global corpus for any world size
idx.reset()
idx.add_synthetic(100000, seed=42)
from haconvdr_b200.index import synth_rows_device
xs = synth_rows_device(100000, 768, seed=42, device=local).cpu().numpy()
D, I = idx.search(q, 10)
D64, I64 = brute_force_fp64(q, xs, 10)
xs64 = xs.astype(np.float64)
assert_parity(D64, I64, D, I, rtol=1e-5, ref_scores_of=lambda qi, ids: xs64[ids] @ q[qi].astype(np.float64))
dist.barrier()
if rank == 0:
    print("MULTI_OK")
dist.destroy_process_group()
'''


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_torchrun_nccl_sharded_search(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HAC_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "MULTI_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_in_process_index_shards_two_devices():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from haconvdr_b200 import faiss_compat as faiss
    from oracle.compare import assert_parity
    from oracle.flat_ip import brute_force_fp64
    co = faiss.GpuMultipleClonerOptions()
    co.shard = True
    vres, vdev = faiss.GpuResourcesVector(), faiss.Int32Vector()
    for i in range(2):
        vdev.push_back(i)
        vres.push_back(faiss.StandardGpuResources())
    index = faiss.index_cpu_to_gpu_multiple(vres, vdev, faiss.IndexFlatIP(768), co)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((25001, 768), dtype=np.float32)
    q = rng.standard_normal((70, 768), dtype=np.float32)
    index.add(x[:9000])
    index.add(x[9000:])
    D, I = index.search(q, 100)
    D64, I64 = brute_force_fp64(q, x, 100)
    x64 = x.astype(np.float64)
    assert_parity(D64, I64, D, I, rtol=1e-5, ref_scores_of=lambda qi, ids: x64[ids] @ q[qi].astype(np.float64))
    index.reset()
    assert index.ntotal == 0


# ---- one device, two shards: threshold exchange + peer-pointer merge through the C ABI -----------------------------
def _two_shards(x, split):
    import haconvdr_b200 as hb
    a, b = hb.FlatIPIndex(768, 0), hb.FlatIPIndex(768, 0)
    a.add(x[:split])
    b.add(x[split:])
    b.set_id_base(split)
    return a, b


@pytest.mark.parametrize("nq,n,k,concurrent,warm", [(150, 60000, 100, True, 0), (150, 60000, 100, False, 0),
                                                    (3, 40000, 10, True, 0), (300, 90001, 1, True, 0),
                                                    (40, 30000, 128, False, 0), (300, 90001, 100, True, 8192),
                                                    (150, 60000, 100, False, 4096), (130, 60000, 1000, True, 0),
                                                    (40, 30000, 300, False, 4096), (140, 70000, 1000, True, 16384)])
def test_threshold_exchange_and_peer_merge_two_shards_one_device(nq, n, k, concurrent, warm):
    """hac_set_threshold_exchange + hac_merge_topk_peers_device with same-device "peer" pointers: two shards search
    (concurrently from two host threads on two streams, or one after the other), publish their ceil(k/2)-th best
    scores to each other, return only their share of the global top-k, and the peer-pointer merge of the two lists
    equals the brute-force result of the whole corpus.  Same protocol as one process per GPU over NVLink."""
    import threading
    import torch
    from haconvdr_b200.index import merge_topk_peers_device
    from oracle.compare import assert_parity
    from oracle.flat_ip import brute_force_fp64
    rng = np.random.default_rng(nq + n + k)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    split = n // 2 + 17
    shards = _two_shards(x, split)
    dev = torch.device("cuda", 0)
    from haconvdr_b200._lib import HAC_EXCHANGE_WORDS_PER_QUERY
    cap = 512 * HAC_EXCHANGE_WORDS_PER_QUERY          # u64 words: 16 per query
    words = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(2)]
    for r, sh in enumerate(shards):
        sh.set_threshold_exchange(words[r].data_ptr(), [words[1 - r].data_ptr()], cap)
        sh.set_option("i8_warm_rows", warm)      # > 0: the shard's first rows go through the f16 screen first
        sh.set_option("i8_large_k_rows_per_k", 0)   # k > 128 on the int8 screen too (shards this small default to f16)
    qd = torch.from_numpy(q).to(dev)
    D64, I64 = brute_force_fp64(q, x, k)
    x64 = x.astype(np.float64)
    scores_of = lambda qi, ids: x64[ids] @ q[qi].astype(np.float64)   # noqa: E731

    def run(epoch):
        outs = [(torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
                for _ in range(2)]
        errs = []

        def one(r):
            try:
                torch.cuda.set_device(0)
                with torch.cuda.stream(torch.cuda.Stream(dev)):
                    shards[r].set_option("exchange_epoch", epoch)
                    shards[r].search(qd, k, out=outs[r])
                    shards[r].set_option("exchange_epoch", 0)
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        if concurrent:
            ts = [threading.Thread(target=one, args=(r,)) for r in range(2)]
            [t.start() for t in ts]
            [t.join() for t in ts]
        else:
            one(0)
            one(1)
        assert not errs, errs
        torch.cuda.synchronize()
        stats = [sh.stats() for sh in shards]
        D, I = merge_topk_peers_device([o[0].data_ptr() for o in outs], [o[1].data_ptr() for o in outs], nq, k, k, dev)
        return D.cpu().numpy(), I.cpu().numpy(), stats, outs

    D0, I0, st_off, _ = run(0)                                   # exchange off: every shard returns its own full top-k
    assert_parity(D64, I64, D0, I0, rtol=1e-5, ref_scores_of=scores_of)
    for epoch in (1, 2, 7):
        D1, I1, st_on, outs = run(epoch)
        assert np.array_equal(I1, I0) and np.array_equal(D1, D0), epoch
        for st in st_on:
            assert st["retries"] == 0 and st["screen_err_max"] <= st["margin_max"], st
            assert st["warm_rows"] == warm, st
    if not concurrent:
        # shard 1 searched after shard 0 had published its final bounds: it keeps only rows that can reach the global
        # top-k, so it returns fillers where its own k-th best would have been
        assert int((outs[1][1] < 0).sum().item()) > 0
        assert st_on[1]["candidates_rescored"] <= st_off[1]["candidates_rescored"]
    # a stale epoch is ignored (tags differ): results unchanged
    D2, I2, _, _ = run(3)
    assert np.array_equal(I2, I0) and np.array_equal(D2, D0)
    for sh in shards:
        sh.set_threshold_exchange(0, [], 0)
        sh.close()


def test_in_process_index_shards_on_one_device():
    """faiss_compat.ShardedInProcessIndex (what `index_cpu_to_gpu_multiple` returns for n_gpu > 1) with both shards
    placed on cuda:0: add / search / reset against the brute-force oracle."""
    from haconvdr_b200 import faiss_compat as faiss
    from oracle.compare import assert_parity
    from oracle.flat_ip import brute_force_fp64
    index = faiss.ShardedInProcessIndex(768, [0, 0])
    rng = np.random.default_rng(6)
    x = rng.standard_normal((33001, 768), dtype=np.float32)
    q = rng.standard_normal((90, 768), dtype=np.float32)
    index.add(x[:12000])
    index.add(x[12000:])
    assert index.ntotal == 33001
    x64 = x.astype(np.float64)
    for k in (100, 1):
        D, I = index.search(q, k)
        D64, I64 = brute_force_fp64(q, x, k)
        assert_parity(D64, I64, D, I, rtol=1e-5, ref_scores_of=lambda qi, ids: x64[ids] @ q[qi].astype(np.float64))
    D2, I2 = index.search(q, 1)
    assert np.array_equal(I2, I) and np.array_equal(D2, D)
    ph = index.last_phase_ms
    assert 0.0 < ph["fastest_shard_done"] <= ph["slowest_shard_done"] <= ph["results_on_host"]
    # faiss-style preallocated outputs, the CUDA-tensor form, and the independent-shards mode (no exchange): all bitwise equal
    import torch
    D100, I100 = index.search(q, 100)
    Dp, Ip = np.empty((90, 100), np.float32), np.empty((90, 100), np.int64)
    assert index.search(q, 100, D=Dp, I=Ip)[0] is Dp
    assert np.array_equal(Ip, I100) and np.array_equal(Dp, D100)
    Dt, It = index.search(torch.from_numpy(q).cuda(), 100)
    assert Dt.is_cuda and np.array_equal(It.cpu().numpy(), I100) and np.array_equal(Dt.cpu().numpy(), D100)
    index.threshold_exchange = False
    Dn, In = index.search(q, 100)
    assert np.array_equal(In, I100) and np.array_equal(Dn, D100)
    index.threshold_exchange = True
    for st in index.stats():
        assert st["retries"] == 0 and st["screen_err_max"] <= st["margin_max"], st
    # a batch larger than the exchange buffers' first size (4096 queries) regrows them; k outside the range is refused
    qb = rng.standard_normal((4500, 768), dtype=np.float32)
    Db, Ib = index.search(qb, 10)
    D64b, I64b = brute_force_fp64(qb[:64], x, 10)
    assert_parity(D64b, I64b, Db[:64], Ib[:64], rtol=1e-5, ref_scores_of=lambda qi, ids: x64[ids] @ qb[qi].astype(np.float64))
    with pytest.raises(ValueError):
        index.search(q, 2000)
    index.reset()
    assert index.ntotal == 0
    D0, I0 = index.search(q[:3], 5)                      # empty shards: faiss' unfilled slots
    assert (I0 == -1).all() and (D0 == np.float32(-3.4028235e38)).all()
    index.add(x[:5000])
    D5, I5 = index.search(q, 10)
    D64s, I64s = brute_force_fp64(q, x[:5000], 10)
    assert_parity(D64s, I64s, D5, I5, rtol=1e-5, ref_scores_of=lambda qi, ids: x64[ids] @ q[qi].astype(np.float64))
    index.close()


@pytest.mark.parametrize("n_shards,n,nq,k", [(3, 20011, 130, 100), (5, 9000, 7, 10), (4, 3, 5, 4)])
def test_in_process_index_with_more_shards_than_devices(n_shards, n, nq, k):
    """The native shard group with 3-5 shards on cuda:0 (uneven splits, shards that stay empty, several peers per
    shard in the threshold exchange): same results as one index over all rows."""
    import haconvdr_b200 as hb
    from haconvdr_b200 import faiss_compat as faiss
    from oracle.compare import assert_parity
    from oracle.flat_ip import brute_force_fp64
    rng = np.random.default_rng(n_shards * 1000 + n)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    index = faiss.ShardedInProcessIndex(768, [0] * n_shards)
    index.add(x[: n // 2])
    index.add(x[n // 2:])
    assert index.ntotal == n and sum(sh.ntotal for sh in index.shards) == n
    D, I = index.search(q, k)
    one = hb.FlatIPIndex(768, 0)
    one.add(x)
    D1, I1 = one.search(q, k)
    assert np.array_equal(I, I1) and np.array_equal(D, D1)          # bitwise what a single shard returns
    if n >= 4 * k:
        x64 = x.astype(np.float64)
        D64, I64 = brute_force_fp64(q, x, k)
        assert_parity(D64, I64, D, I, rtol=1e-5, ref_scores_of=lambda qi, ids: x64[ids] @ q[qi].astype(np.float64))
    else:                                                           # k > rows: faiss' unfilled slots behind the rows
        assert (I[:, n:] == -1).all() and (np.sort(I[:, :n], axis=1) == np.arange(n)).all()
    one.close()
    index.close()


def test_shard_group_through_the_raw_c_abi():
    """hac_shards_* called as a C host would: borrowed handles, mismatching dimensions refused, per-shard id bases and
    options respected, argument errors reported through hac_last_error()."""
    import ctypes
    import haconvdr_b200 as hb
    from haconvdr_b200 import _lib
    L = _lib.lib()
    a, b = hb.FlatIPIndex(768, 0), hb.FlatIPIndex(256, 0)
    handles = (ctypes.c_void_p * 2)(a._h.value, b._h.value)
    grp = ctypes.c_void_p()
    assert L.hac_shards_create(handles, 2, ctypes.byref(grp)) == -1 and "dimensions differ" in _lib.last_error()
    c = hb.FlatIPIndex(768, 0)
    handles = (ctypes.c_void_p * 2)(a._h.value, c._h.value)
    assert L.hac_shards_create(handles, 2, ctypes.byref(grp)) == 0
    x = np.random.default_rng(2).standard_normal((3000, 768), dtype=np.float32)
    a.add(x[:1500])
    c.add(x[1500:])
    c.set_id_base(1500)
    c.set_option("default_path", hb.HAC_PATH_GEMV)       # one shard on the fp32 GEMV path (batches of 4), the other on int8
    q = x[:9].copy()
    D, I = np.empty((9, 5), np.float32), np.empty((9, 5), np.int64)
    assert L.hac_shards_search(grp, 9, q.ctypes.data, 5, D.ctypes.data, I.ctypes.data) == 0, _lib.last_error()
    assert np.array_equal(I[:, 0], np.arange(9))         # every row is its own best match
    assert L.hac_shards_search(grp, 9, q.ctypes.data, 0, D.ctypes.data, I.ctypes.data) == -1
    assert "k outside" in _lib.last_error()
    assert L.hac_shards_destroy(grp) == 0
    for h in (a, b, c):
        h.close()
