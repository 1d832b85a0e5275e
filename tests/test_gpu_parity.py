"""GPU (-m gpu): parity of the CUDA path against the oracle, through the shim / C ABI.

Bars (BASELINE.json north_star): ids and order identical to the reference outside tolerance
groups (reference scores within 1e-5 relative of each other), scores within 1e-5 relative,
recall@k overlap 1.0; integer-valued cases are bit-exact."""
import tempfile
import types

import numpy as np
import pytest

from oracle.compare import assert_parity
from oracle.flat_ip import FlatIP, NEG_FLT_MAX, brute_force_fp64
from helpers import GOLDEN_MERGE_CASES, load_golden, write_blocks

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _engine():
    import haconvdr_b200 as hb
    return hb


def _scores_of(q, x):
    x64 = x.astype(np.float64)
    return lambda qi, ids: x64[ids] @ q[qi].astype(np.float64)


def _check(q, x, k, D, I, also_fp32_oracle=True):
    D64, I64 = brute_force_fp64(q, x, k)
    # when (nearly) the whole corpus is returned the tail scores pass through zero, where a relative
    # bound is meaningless: allow 1e-5 of the score scale there (never the case at BASELINE sizes)
    atol = RTOL * float(np.abs(D64[I64 >= 0]).max()) if 4 * k >= x.shape[0] else 0.0
    rep = assert_parity(D64, I64, D, I, rtol=RTOL, ref_scores_of=_scores_of(q, x), atol=atol)
    if also_fp32_oracle:
        ref = FlatIP(x.shape[1])
        ref.add(x)
        Dr, Ir = ref.search(q, k)
        assert_parity(Dr, Ir, D, I, rtol=RTOL, ref_scores_of=_scores_of(q, x), atol=atol)
    assert np.all(np.diff(D, axis=1) <= 0)
    return rep


@pytest.mark.parametrize("name", GOLDEN_MERGE_CASES)
def test_golden_fixtures_through_reference_loop(name):
    """The reference's block loop (our mirror of it) over the CUDA index reproduces the golden
    outputs made by the reference's own function."""
    hb = _engine()
    from haconvdr_b200 import faiss_compat as faiss
    from haconvdr_b200.retrieval import search_one_by_one_with_faiss
    g = load_golden(name)
    d = g["q"].shape[1]
    index = faiss.index_cpu_to_gpu_multiple([None], [0], faiss.IndexFlatIP(d), faiss.GpuMultipleClonerOptions())
    with tempfile.TemporaryDirectory() as tmp:
        write_blocks(tmp, g["blocks"], g["id_start"])
        args = types.SimpleNamespace(passage_block_num=int(g.get("block_num", len(g["blocks"]) + 3)))
        D, I = search_one_by_one_with_faiss(args, tmp, index, g["q"], g["k"])
    assert D.dtype == np.float64 and I.dtype == np.int64 and D.shape == g["D"].shape
    integer_case = name.startswith("kat_int") or "ties" in name
    if integer_case:
        assert np.array_equal(I, g["I"])
        assert np.array_equal(D, g["D"])
    else:
        k = g["k"]
        x = np.concatenate(g["blocks"][: int(g.get("block_num", len(g["blocks"])))], 0)
        sc = _scores_of(g["q"], x)
        valid_cols = min(k, x.shape[0]) if len(g["blocks"]) == 1 else k
        if "short" in name:
            # unfilled slots and the emb2id[-1] wrap must match exactly where the reference is deterministic
            assert np.array_equal(I, g["I"])
            # (k > rows and d = 64: scores pass through zero, hence the absolute term 1e-5 * max|score|)
            np.testing.assert_allclose(D, g["D"], rtol=RTOL, atol=RTOL * float(np.abs(g["D"][g["D"] > -1e30]).max()))
        else:
            off = g["id_start"]
            assert_parity(g["D"][:, :valid_cols], g["I"][:, :valid_cols] - off, D[:, :valid_cols],
                          I[:, :valid_cols] - off, rtol=RTOL, ref_scores_of=sc)


@pytest.mark.parametrize("nq", [1, 2, 3, 4])
def test_small_batch_gemv_path(nq):
    hb = _engine()
    rng = np.random.default_rng(10 + nq)
    x = rng.standard_normal((30011, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 100, path=hb.HAC_PATH_GEMV)
    assert idx.stats()["path"] == hb.HAC_PATH_GEMV
    _check(q, x, 100, D, I)
    # the tensor-core path ends in the same exact rescore: bitwise identical results
    Dm, Im = idx.search(q, 100, path=hb.HAC_PATH_MMA)
    assert np.array_equal(Im, I) and np.array_equal(Dm, D)


@pytest.mark.parametrize("nq,n,k", [(5, 257, 10), (130, 50000, 100), (300, 120001, 100), (129, 4096, 1),
                                    (64, 70000, 1000), (257, 9000, 7), (33, 30000, 1024)])
def test_large_batch_mma_path(nq, n, k):
    hb = _engine()
    rng = np.random.default_rng(nq * 7 + k)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, k, path=hb.HAC_PATH_MMA)
    st = idx.stats()
    assert st["path"] == hb.HAC_PATH_MMA and st["retries"] == 0
    assert st["screen_err_max"] <= st["margin_max"], st      # the rigorous margin really bounds the screen error
    _check(q, x, k, D, I)


def test_anisotropic_corpus_common_mean_component():
    """ANCE-like geometry: a large shared mean plus small noise, so scores crowd together."""
    hb = _engine()
    rng = np.random.default_rng(5)
    mu = rng.standard_normal(768).astype(np.float32)
    x = (mu + 0.3 * rng.standard_normal((60000, 768))).astype(np.float32)
    q = (mu + 0.3 * rng.standard_normal((140, 768))).astype(np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 100)
    st = idx.stats()
    assert st["screen_err_max"] <= st["margin_max"], st
    _check(q, x, 100, D, I)


def test_ties_integer_scores_are_bit_exact():
    hb = _engine()
    rng = np.random.default_rng(3)
    base = rng.integers(-3, 4, size=(500, 768)).astype(np.float32)
    x = np.concatenate([base, base[::-1], base[:250]], 0)          # every row duplicated 2-3 times
    q = rng.integers(-3, 4, size=(150, 768)).astype(np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    exact = q.astype(np.int64) @ x.astype(np.int64).T
    order = np.argsort(-exact, axis=1, kind="stable")[:, :100]
    for path in (hb.HAC_PATH_MMA,):
        D, I = idx.search(q, 100, path=path)
        assert np.array_equal(I, order)
        assert np.array_equal(D, np.take_along_axis(exact, order, 1).astype(np.float32))
    Dg, Ig = idx.search(q[:4], 100, path=hb.HAC_PATH_GEMV)
    assert np.array_equal(Ig, order[:4])


def test_segments_reset_and_id_translation():
    hb = _engine()
    rng = np.random.default_rng(8)
    x = rng.standard_normal((9000, 768), dtype=np.float32)
    q = rng.standard_normal((40, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    for lo, hi in ((0, 1), (1, 300), (300, 4097), (4097, 9000)):     # ragged appends -> several segments
        idx.add(x[lo:hi])
    assert idx.ntotal == 9000
    D, I = idx.search(q, 50)
    _check(q, x, 50, D, I)
    idx.set_id_base(1000)
    D2, I2 = idx.search(q, 50)
    assert np.array_equal(I2, I + 1000) and np.array_equal(D2, D)
    table = np.arange(9000, dtype=np.int64)[::-1].copy() * 3
    idx.set_id_table(table)
    D3, I3 = idx.search(q, 50)
    assert np.array_equal(I3, table[I])
    idx.reset()
    assert idx.ntotal == 0
    De, Ie = idx.search(q, 5)
    assert np.all(Ie == -1) and np.all(De == NEG_FLT_MAX)
    idx.add(x[:777])                                              # capacity is reused after reset
    D4, I4 = idx.search(q, 50)
    _check(q, x[:777], 50, D4, I4)


def test_k_larger_than_ntotal_fills_like_faiss():
    hb = _engine()
    rng = np.random.default_rng(9)
    x = rng.standard_normal((37, 768), dtype=np.float32)
    q = rng.standard_normal((6, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    for path in (hb.HAC_PATH_AUTO, hb.HAC_PATH_MMA):
        D, I = idx.search(q, 100, path=path)
        D64, I64 = brute_force_fp64(q, x, 100)
        assert np.array_equal(I, I64)
        assert np.all(D[:, 37:] == NEG_FLT_MAX) and np.all(I[:, 37:] == -1)


def test_error_behaviour_mirrors_faiss_wrapper():
    hb = _engine()
    idx = hb.FlatIPIndex(768)
    with pytest.raises(AssertionError):
        idx.add(np.zeros((4, 767), np.float32))
    idx.add(np.zeros((4, 768), np.float32))
    with pytest.raises(AssertionError):
        idx.search(np.zeros((1, 64), np.float32), 10)
    for bad_k in (0, -1, hb.HAC_MAX_K + 1):
        with pytest.raises(ValueError):
            idx.search(np.zeros((1, 768), np.float32), bad_k)
    with pytest.raises(ValueError):
        hb.FlatIPIndex(100)                       # d must be a multiple of 64


def test_torch_tensor_handoff_and_device_merge():
    import torch
    hb = _engine()
    from haconvdr_b200.index import merge_topk_device
    rng = np.random.default_rng(12)
    x = rng.standard_normal((20000, 768), dtype=np.float32)
    q = rng.standard_normal((33, 768), dtype=np.float32)
    shards = []
    for g, (lo, hi) in enumerate(((0, 6000), (6000, 13000), (13000, 20000))):
        s = hb.FlatIPIndex(768)
        s.add(torch.from_numpy(x[lo:hi]).cuda())
        s.set_id_base(lo)
        shards.append(s)
    qd = torch.from_numpy(q).cuda()
    parts = [s.search(qd, 100) for s in shards]
    assert all(p[0].is_cuda and p[1].dtype == torch.int64 for p in parts)
    D, I = merge_topk_device(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 100)
    _check(q, x, 100, D.cpu().numpy(), I.cpu().numpy())


def test_synthetic_rows_do_not_depend_on_shard_split():
    import torch
    hb = _engine()
    from haconvdr_b200.index import synth_rows_device
    a = synth_rows_device(1000, 768, seed=42, row0=0)
    b = torch.cat([synth_rows_device(300, 768, seed=42, row0=0), synth_rows_device(700, 768, seed=42, row0=300)])
    assert torch.equal(a, b)
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1.0) < 0.01
    x = a.cpu().numpy()
    q = synth_rows_device(20, 768, seed=4242).cpu().numpy()
    idx = hb.FlatIPIndex(768)
    idx.add_synthetic(1000, seed=42, row0=0)
    D, I = idx.search(q, 10)
    _check(q, x, 10, D, I)


def test_mass_duplicates_recovered_in_careful_mode():
    """30 000 identical rows tie for the top: the shortlist overflows in fast mode, the careful mode
    (roll back, split, exact compaction) must still return the exact (score desc, id asc) answer."""
    hb = _engine()
    rng = np.random.default_rng(21)
    v = rng.standard_normal(768).astype(np.float32)
    x = np.concatenate([rng.standard_normal((5000, 768), dtype=np.float32),
                        np.tile(v, (30000, 1)),
                        rng.standard_normal((5000, 768), dtype=np.float32)], 0)
    q = rng.standard_normal((130, 768), dtype=np.float32)
    q[:40] = (v * rng.uniform(0.5, 2.0, size=(40, 1)) + 0.05 * rng.standard_normal((40, 768))).astype(np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 100)
    st = idx.stats()
    assert st["retries"] >= 1, st                       # the fast mode did overflow
    D64, I64 = brute_force_fp64(q, x, 100)
    assert np.array_equal(I[:40], np.tile(np.arange(5000, 5100), (40, 1)))     # first 100 duplicates, id order
    assert_parity(D64, I64, D, I, rtol=RTOL, ref_scores_of=_scores_of(q, x))
    # k = 1000 with the larger shortlist as well
    D, I = idx.search(q[:8], 1000)
    D64, I64 = brute_force_fp64(q[:8], x, 1000)
    assert_parity(D64, I64, D, I, rtol=RTOL, ref_scores_of=_scores_of(q[:8], x))


def test_resident_path_to_trec_run_matches_reference_bytes():
    """Blocks -> HBM-resident shard -> one search -> offset2pid + dedup -> TREC run: byte-identical to the
    run file the reference's output_test_res wrote for the same inputs (golden)."""
    import os
    hb = _engine()
    from haconvdr_b200 import retrieval
    g = load_golden("trec_run_dedup_d64")
    with tempfile.TemporaryDirectory() as tmp:
        write_blocks(tmp, g["blocks"], 0)
        idx = hb.FlatIPIndex(64)
        n_blocks, n_rows = retrieval.load_resident(idx, tmp, 26)
        assert (n_blocks, n_rows) == (2, 320) and idx.ntotal == 320
        D, I = retrieval.search_resident(idx, g["q"], g["k"])
        assert np.array_equal(I, g["I"][:, : g["k"]])
        ranked = retrieval.rank_pids(D, I, g["offset2pid"].tolist(), g["k"])
        run = retrieval.write_trec_run(os.path.join(tmp, "run.trec"), g["qids"].tolist(), ranked, g["k"])
        got, want = open(run).read().split("\n"), str(g["run_text"]).split("\n")
        assert len(got) == len(want)
        for a, b in zip(got, want):
            fa, fb = a.split(" "), b.split(" ")
            assert fa[:5] == fb[:5] and fa[6:] == fb[6:]                 # qid Q0 pid rank 200-rank ... tag
            if len(fa) > 5:
                assert abs(float(fa[5]) - float(fb[5])) <= RTOL * abs(float(fb[5])) + 1e-5


def test_query_batches_beyond_the_internal_batch_size():
    """More queries than one internal batch (16384): results of every sub-batch land at the right offset."""
    hb = _engine()
    rng = np.random.default_rng(31)
    x = rng.standard_normal((3000, 768), dtype=np.float32)
    q = rng.standard_normal((16384 + 700, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 5)
    sel = np.r_[0:50, 16300:16450, 17000:17084]
    _check(q[sel], x, 5, D[sel], I[sel], also_fp32_oracle=False)


def test_operand_scaling_extreme_magnitudes():
    """Power-of-two operand scaling keeps the f16 screen usable for tiny and huge embeddings."""
    hb = _engine()
    rng = np.random.default_rng(33)
    base_x = rng.standard_normal((20000, 768), dtype=np.float32)
    base_q = rng.standard_normal((70, 768), dtype=np.float32)
    for sx, sq in ((1e-6, 1.0), (3e4, 1e-3), (1.0, 5e3)):
        x, q = (base_x * np.float32(sx)), (base_q * np.float32(sq))
        idx = hb.FlatIPIndex(768)
        idx.add(x)
        for path in (hb.HAC_PATH_MMA, hb.HAC_PATH_I8):
            D, I = idx.search(q, 50, path=path)
            st = idx.stats()
            assert st["path"] == path and st["retries"] == 0 and st["screen_err_max"] <= st["margin_max"], (sx, sq, st)
            _check(q, x, 50, D, I, also_fp32_oracle=False)


def test_all_zero_corpus_is_one_big_tie():
    hb = _engine()
    x = np.zeros((9000, 768), np.float32)
    q = np.random.default_rng(2).standard_normal((10, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 100)
    assert np.array_equal(I, np.tile(np.arange(100), (10, 1))) and np.all(D == 0)


def test_device_api_on_a_side_stream():
    import torch
    hb = _engine()
    rng = np.random.default_rng(41)
    x = rng.standard_normal((40000, 768), dtype=np.float32)
    q = rng.standard_normal((200, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        xd = torch.from_numpy(x).cuda(non_blocking=True)
        idx.add(xd)
        qd = torch.from_numpy(q).cuda(non_blocking=True) * 1.0        # produced on the side stream
        D, I = idx.search(qd, 100)
        Dh, Ih = D.cpu().numpy(), I.cpu().numpy()
    _check(q, x, 100, Dh, Ih, also_fp32_oracle=False)


@pytest.mark.parametrize("n_lists,k,k_out", [(2, 100, 100), (8, 100, 100), (3, 7, 20), (8, 1000, 1000), (5, 1, 1)])
def test_device_merge_against_numpy(n_lists, k, k_out):
    """hac_merge_topk_device on sorted lists with exact ties and filler slots."""
    import torch
    from haconvdr_b200.index import merge_topk_device
    rng = np.random.default_rng(n_lists * 1000 + k)
    nq = 37
    D = np.round(rng.standard_normal((n_lists, nq, k)) * 4).astype(np.float32)        # coarse values -> many ties
    I = rng.permutation(n_lists * nq * k).reshape(n_lists, nq, k).astype(np.int64)   # unique ids
    fill = rng.integers(0, k + 1, size=(n_lists, nq))                                 # valid entries per list
    for l in range(n_lists):
        for qi in range(nq):
            D[l, qi, fill[l, qi]:] = NEG_FLT_MAX
            I[l, qi, fill[l, qi]:] = -1
            order = np.lexsort((np.where(I[l, qi] < 0, np.iinfo(np.int64).max, I[l, qi]), -D[l, qi].astype(np.float64)))
            D[l, qi], I[l, qi] = D[l, qi][order], I[l, qi][order]
    Dm, Im = merge_topk_device(torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda(), k_out)
    catD = D.transpose(1, 0, 2).reshape(nq, -1)
    catI = I.transpose(1, 0, 2).reshape(nq, -1)
    order = np.lexsort((np.where(catI < 0, np.iinfo(np.int64).max, catI), -catD.astype(np.float64)), axis=1)
    wantD, wantI = np.take_along_axis(catD, order, 1), np.take_along_axis(catI, order, 1)
    if k_out > catD.shape[1]:
        pad = k_out - catD.shape[1]
        wantD = np.concatenate([wantD, np.full((nq, pad), NEG_FLT_MAX, np.float32)], 1)
        wantI = np.concatenate([wantI, np.full((nq, pad), -1, np.int64)], 1)
    assert np.array_equal(Im.cpu().numpy(), wantI[:, :k_out])
    assert np.array_equal(Dm.cpu().numpy(), wantD[:, :k_out])


@pytest.mark.parametrize("nq,n,k", [(1, 30011, 100), (5, 257, 10), (130, 50000, 100), (300, 120001, 100),
                                    (129, 4096, 1), (64, 70000, 1000), (257, 9000, 7)])
def test_int8_screen_path(nq, n, k):
    """The int8 tensor-core screen ends in the same exact fp32 scores: bitwise equal to the f16 path."""
    hb = _engine()
    rng = np.random.default_rng(nq * 11 + k)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.set_option("build_i8", 1)
    idx.add(x[: n // 3])
    idx.add(x[n // 3:])                                  # second append re-quantises the straddled tile
    D8, I8 = idx.search(q, k, path=hb.HAC_PATH_I8)
    st = idx.stats()
    assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, st
    assert st["screen_err_max"] <= st["margin_max"], st
    Dm, Im = idx.search(q, k, path=hb.HAC_PATH_MMA)
    assert np.array_equal(I8, Im) and np.array_equal(D8, Dm)
    _check(q, x, k, D8, I8, also_fp32_oracle=False)


def test_int8_screen_anisotropic_and_scaled_data():
    hb = _engine()
    rng = np.random.default_rng(55)
    mu = rng.standard_normal(768).astype(np.float32)
    x = (mu + 0.3 * rng.standard_normal((60000, 768))).astype(np.float32) * np.float32(37.0)
    q = (mu + 0.3 * rng.standard_normal((140, 768))).astype(np.float32) * np.float32(1e-3)
    idx = hb.FlatIPIndex(768)
    idx.set_option("build_i8", 1)
    idx.add(x)
    D8, I8 = idx.search(q, 100, path=hb.HAC_PATH_I8)
    st = idx.stats()
    assert st["screen_err_max"] <= st["margin_max"], st
    Dm, Im = idx.search(q, 100, path=hb.HAC_PATH_MMA)
    assert np.array_equal(I8, Im) and np.array_equal(D8, Dm)
    _check(q, x, 100, D8, I8, also_fp32_oracle=False)


def test_int8_screen_mass_duplicates_fall_back():
    hb = _engine()
    rng = np.random.default_rng(21)
    v = rng.standard_normal(768).astype(np.float32)
    x = np.concatenate([rng.standard_normal((5000, 768), dtype=np.float32), np.tile(v, (30000, 1))], 0)
    q = (v * rng.uniform(0.5, 2.0, size=(40, 1)) + 0.05 * rng.standard_normal((40, 768))).astype(np.float32)
    idx = hb.FlatIPIndex(768)
    idx.set_option("build_i8", 1)
    idx.add(x)
    D, I = idx.search(q, 100, path=hb.HAC_PATH_I8)
    assert idx.stats()["retries"] >= 1
    assert np.array_equal(I, np.tile(np.arange(5000, 5100), (40, 1)))


@pytest.mark.parametrize("d", [64, 128, 192, 256, 512, 1024])
def test_other_dimensions_all_paths(d):
    """d is fixed to 768 in the reference; the engine takes any multiple of 64 up to 1024."""
    hb = _engine()
    rng = np.random.default_rng(d)
    x = rng.standard_normal((21000, d), dtype=np.float32)
    q = rng.standard_normal((70, d), dtype=np.float32)
    idx = hb.FlatIPIndex(d)
    idx.set_option("build_i8", 1)
    idx.add(x)
    D, I = idx.search(q, 20, path=hb.HAC_PATH_MMA)
    _check(q, x, 20, D, I, also_fp32_oracle=False)
    Dg, Ig = idx.search(q[:3], 20, path=hb.HAC_PATH_GEMV)
    assert np.array_equal(Ig, I[:3]) and np.array_equal(Dg, D[:3])
    D8, I8 = idx.search(q, 20, path=hb.HAC_PATH_I8)          # falls back to the f16 screen when d % 128 != 0
    assert np.array_equal(I8, I) and np.array_equal(D8, D)


def test_preallocated_host_outputs_like_faiss():
    hb = _engine()
    rng = np.random.default_rng(6)
    x = rng.standard_normal((5000, 768), dtype=np.float32)
    q = rng.standard_normal((9, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D0, I0 = idx.search(q, 10)
    D = np.empty((9, 10), np.float32)
    I = np.empty((9, 10), np.int64)
    r = idx.search(q, 10, D=D, I=I)
    assert r[0] is D and r[1] is I and np.array_equal(D, D0) and np.array_equal(I, I0)


def test_offset2pid_gather_on_device():
    import torch
    from haconvdr_b200.index import gather_ids_device
    rng = np.random.default_rng(4)
    table = rng.integers(0, 10**9, size=5000).astype(np.int64)
    ids = rng.integers(-1, 5000, size=(37, 100)).astype(np.int64)
    out = gather_ids_device(torch.from_numpy(table).cuda(), torch.from_numpy(ids).cuda()).cpu().numpy()
    want = np.where(ids >= 0, table[np.maximum(ids, 0)], -1)
    assert np.array_equal(out, want)


def test_centred_screen_shrinks_the_margin_and_changes_nothing_else():
    """ANCE-like rows share a large mean: the f16 image holds x - c, the scan adds q.c back.  Results are
    bitwise those of the uncentred screen (both end in the exact fp32 rescore), the margin is several times
    smaller, and far fewer rows need rescoring."""
    hb = _engine()
    rng = np.random.default_rng(77)
    mu = (3.0 * rng.standard_normal(768)).astype(np.float32)
    x = (mu + 0.3 * rng.standard_normal((80000, 768))).astype(np.float32)
    q = (mu + 0.3 * rng.standard_normal((200, 768))).astype(np.float32)
    out = {}
    for centred in (1, 0):
        idx = hb.FlatIPIndex(768)
        idx.set_option("center_screen", centred)
        idx.add(x[:50000])
        idx.add(x[50000:])                               # later adds use the centre of the first one
        D, I = idx.search(q, 100, path=hb.HAC_PATH_MMA)
        st = idx.stats()
        assert st["screen_err_max"] <= st["margin_max"], st
        out[centred] = (D, I, st)
        idx.close()
    assert np.array_equal(out[1][1], out[0][1]) and np.array_equal(out[1][0], out[0][0])
    assert out[1][2]["margin_max"] < 0.5 * out[0][2]["margin_max"], (out[1][2], out[0][2])
    assert out[1][2]["candidates_rescored"] <= out[0][2]["candidates_rescored"]
    _check(q, x, 100, out[1][0], out[1][1])


def test_centre_of_an_unrepresentative_first_add_stays_exact():
    """Any centre keeps the bound rigorous: here it comes from a single outlier row, later blocks look nothing
    like it, and a reset picks a new one."""
    hb = _engine()
    rng = np.random.default_rng(78)
    x = rng.standard_normal((40000, 768), dtype=np.float32)
    x[0] = 25.0
    q = rng.standard_normal((70, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x[:1])
    idx.add(x[1:])
    D, I = idx.search(q, 100)
    st = idx.stats()
    assert st["screen_err_max"] <= st["margin_max"], st
    _check(q, x, 100, D, I)
    idx.reset()
    idx.add(x[1:])
    D2, I2 = idx.search(q, 100)
    assert idx.stats()["margin_max"] < st["margin_max"]
    _check(q, x[1:], 100, D2, I2)


def test_careful_mode_is_sticky_until_reset():
    hb = _engine()
    rng = np.random.default_rng(79)
    v = rng.standard_normal(768).astype(np.float32)
    x = np.concatenate([rng.standard_normal((5000, 768), dtype=np.float32), np.tile(v, (30000, 1))], 0)
    q = (v * rng.uniform(0.5, 2.0, size=(40, 1)) + 0.05 * rng.standard_normal((40, 768))).astype(np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 100, path=hb.HAC_PATH_MMA)
    assert idx.stats()["retries"] == 1
    D2, I2 = idx.search(q, 100, path=hb.HAC_PATH_MMA)
    assert idx.stats()["retries"] == 0                   # started in careful mode: no wasted fast pass
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    assert np.array_equal(I, np.tile(np.arange(5000, 5100), (40, 1)))
    idx.reset()
    idx.add(x[:5000])
    idx.search(q, 100, path=hb.HAC_PATH_MMA)
    assert idx.stats()["retries"] == 0


@pytest.mark.parametrize("nq,k", [(1, 100), (4, 100), (32, 10), (129, 128), (300, 100), (64, 129), (40, 1000), (200, 300)])
def test_auto_path_policy(nq, k):
    """The int8 image is built by default; AUTO takes its screen at every batch size for k <= 128, and for larger k on
    shards of >= 12288 rows per k ("i8_large_k_rows_per_k"; smaller ones run the f16 screen).  Same results either way."""
    hb = _engine()
    rng = np.random.default_rng(80 + nq)
    x = rng.standard_normal((90000, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    idx.set_option("i8_large_k_rows_per_k", 0)                 # no shard-size limit: the int8 screen at every k
    D, I = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, st
    assert st["bytes_i8"] >= 90000 * 768                       # the int8 image is built on add ...
    assert st["bytes_shadow"] == 0                             # ... the f16 image only when a search needs it
    idx.set_option("i8_large_k_rows_per_k", 12288)             # the default: 90 000 rows are too few for k > 128
    Dp, Ip = idx.search(q, k)
    assert idx.stats()["path"] == (hb.HAC_PATH_I8 if k <= 128 else hb.HAC_PATH_MMA)
    assert np.array_equal(Ip, I) and np.array_equal(Dp, D)
    idx.set_option("i8_large_k_rows_per_k", 90000 // k)        # exactly enough rows per k
    idx.search(q, k)
    assert idx.stats()["path"] == (hb.HAC_PATH_I8 if k <= 128 or nq >= 128 or 256 * nq <= 90000 // k else hb.HAC_PATH_MMA)
    Dm, Im = idx.search(q, k, path=hb.HAC_PATH_MMA)
    assert idx.stats()["bytes_shadow"] >= 90000 * 768 * 2
    assert np.array_equal(I, Im) and np.array_equal(D, Dm)
    _check(q, x, k, D, I, also_fp32_oracle=False)
    idx.set_option("i8_auto_max_k", 0)
    idx.search(q, k)
    assert idx.stats()["path"] == hb.HAC_PATH_MMA
    # without the image: the f16 screen, and an explicit HAC_PATH_I8 request falls back to it
    idx2 = hb.FlatIPIndex(768)
    idx2.set_option("build_i8", 0)
    idx2.add(x[:20000])
    D2, I2 = idx2.search(q, min(k, 128), path=hb.HAC_PATH_I8)
    assert idx2.stats()["path"] == hb.HAC_PATH_MMA and idx2.stats()["bytes_i8"] == 0
    assert 20000 * 768 * 2 <= idx2.stats()["bytes_shadow"] < 20224 * 768 * 3
    _check(q, x[:20000], min(k, 128), D2, I2, also_fp32_oracle=False)


def test_int8_overflow_is_remembered_until_reset():
    hb = _engine()
    rng = np.random.default_rng(21)
    v = rng.standard_normal(768).astype(np.float32)
    x = np.concatenate([rng.standard_normal((5000, 768), dtype=np.float32), np.tile(v, (30000, 1))], 0)
    q = (v * rng.uniform(0.5, 2.0, size=(40, 1)) + 0.05 * rng.standard_normal((40, 768))).astype(np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    D, I = idx.search(q, 100)                            # int8 screen overflows, f16 fast mode overflows, careful mode
    assert idx.stats()["path"] == hb.HAC_PATH_MMA and idx.stats()["retries"] >= 2
    D2, I2 = idx.search(q, 100)                          # straight to the f16 screen in careful mode
    assert idx.stats()["path"] == hb.HAC_PATH_MMA and idx.stats()["retries"] == 0
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    assert np.array_equal(I, np.tile(np.arange(5000, 5100), (40, 1)))
    idx.reset()
    idx.add(x[:5000])
    idx.search(q, 100)
    assert idx.stats()["path"] == hb.HAC_PATH_I8 and idx.stats()["retries"] == 0


@pytest.mark.parametrize("direct", [False, True])
def test_native_block_files_stream_into_the_index(direct):
    """A native .hacb block streamed through pinned staging - with O_DIRECT where the file
    system allows it - gives the same index as the pickle of the same rows, whole or as a row slice."""
    hb = _engine()
    from haconvdr_b200 import loader
    rng = np.random.default_rng(31)
    x = rng.standard_normal((45003, 768), dtype=np.float32)      # not a multiple of 4 rows: padded tail page
    q = rng.standard_normal((20, 768), dtype=np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        write_blocks(tmp, [x], 0)
        nat = loader.convert_block_to_native(tmp, 0)
        ref = hb.FlatIPIndex(768)
        loader.stream_block_into(ref, loader.block_paths(tmp, 0)[0])
        Dr, Ir = ref.search(q, 50)
        idx = hb.FlatIPIndex(768)
        st = {}
        n = loader.stream_block_into(idx, nat, chunk_bytes=8 << 20, direct=direct, stats=st)
        assert n == 45003 and idx.ntotal == 45003 and (not st["direct"] or direct)
        D, I = idx.search(q, 50)
        assert np.array_equal(I, Ir) and np.array_equal(D, Dr)
        _check(q, x, 50, D, I, also_fp32_oracle=False)
        idx.reset()
        loader.stream_block_into(idx, nat, chunk_bytes=8 << 20, direct=direct, row_range=(10000, 30001))
        D2, I2 = idx.search(q, 50)
        _check(q, x[10000:30001], 50, D2, I2, also_fp32_oracle=False)


def test_int8_screen_on_anisotropic_rows_needs_the_centre():
    """x = mu + 0.3*eps: uncentred, the int8 margin (~14 score units) dwarfs the score spread (~9) and nearly every
    row is emitted; centred, the int8 screen prunes as on isotropic data and stays on its own path."""
    hb = _engine()
    rng = np.random.default_rng(91)
    mu = rng.standard_normal(768).astype(np.float32)
    x = (mu + 0.3 * rng.standard_normal((200000, 768))).astype(np.float32)
    q = (mu + 0.3 * rng.standard_normal((8, 768))).astype(np.float32)
    res = {}
    for centred in (1, 0):
        idx = hb.FlatIPIndex(768)
        idx.set_option("build_i8", 1)
        idx.set_option("center_screen", centred)
        idx.add(x)
        D, I = idx.search(q, 100, path=hb.HAC_PATH_I8)
        res[centred] = (D, I, idx.stats())
        idx.close()
    st = res[1][2]
    assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, st
    assert st["screen_err_max"] <= st["margin_max"], st
    assert st["candidates_rescored"] < 8 * 40000, st                 # a fraction of the 200 000 rows per query
    assert res[0][2]["candidates_rescored"] > 2 * st["candidates_rescored"] or res[0][2]["retries"] >= 1, res[0][2]
    assert np.array_equal(res[1][1], res[0][1]) and np.array_equal(res[1][0], res[0][0])
    _check(q, x, 100, res[1][0], res[1][1], also_fp32_oracle=False)


@pytest.mark.parametrize("nq,n,k", [(130, 50000, 100), (300, 120001, 100), (256, 4096, 1), (64, 70000, 1000),
                                    (700, 90000, 10)])
def test_int8_screen_cta_pairs_and_unit_schedules(nq, n, k):
    """cta_group::2 int8 scan (CTA pairs share each MMA) and all unit schedules: bitwise the same results."""
    hb = _engine()
    rng = np.random.default_rng(nq + n + k)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.set_option("build_i8", 1)
    idx.add(x)
    Dm, Im = idx.search(q, k, path=hb.HAC_PATH_MMA)
    for cg in (1, 2):
        for tile_major in (0, 1, 2):                                 # 2 = query-stationary pairs (cg = 2), else as 1
            idx.set_option("i8_cta_group", cg)
            idx.set_option("scan_tile_major", tile_major)
            D8, I8 = idx.search(q, k, path=hb.HAC_PATH_I8)
            st = idx.stats()
            assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, (cg, tile_major, st)
            assert np.array_equal(I8, Im) and np.array_equal(D8, Dm), (cg, tile_major)
            Df, If = idx.search(q, k, path=hb.HAC_PATH_MMA)          # the f16 scan under the same schedule
            assert np.array_equal(If, Im) and np.array_equal(Df, Dm), (cg, tile_major)
    _check(q, x, k, Dm, Im, also_fp32_oracle=False)


@pytest.mark.parametrize("nq,n,k", [(1, 200003, 100), (4, 150000, 10), (130, 120001, 100), (300, 260000, 100),
                                    (520, 90000, 1), (64, 70000, 1000)])
def test_int8_pipelined_search_equals_the_synchronous_one(nq, n, k):
    """The pipelined int8 search (scans back to back, rescore + refresh of chunk i on a side stream beside the scan of
    chunk i+1, stale thresholds) returns bitwise what the chunk-synchronous schedule and the f16 screen return; with
    `i8_pipe_min_rows` lowered the overlap engages on a corpus the oracle can still brute-force."""
    hb = _engine()
    rng = np.random.default_rng(nq * 7 + k)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x[: n // 2 + 11])
    idx.add(x[n // 2 + 11:])
    idx.set_option("i8_pipeline", 0)
    D0, I0 = idx.search(q, k, path=hb.HAC_PATH_I8)
    st0 = idx.stats()
    assert st0["path"] == hb.HAC_PATH_I8 and st0["retries"] == 0 and st0["pipelined"] == 0, st0
    _check(q, x, k, D0, I0, also_fp32_oracle=False)
    idx.set_option("i8_pipeline", 1)
    for min_rows, growth, dist in ((4096, 125, 2), (8192, 500, 2), (4096, 125, 1), (0, 125, 2)):
        idx.set_option("i8_pipe_min_rows", min_rows)
        idx.set_option("i8_pipe_growth_x1000", growth)
        idx.set_option("i8_pipe_dist", dist)
        for _ in range(2):
            D1, I1 = idx.search(q, k, path=hb.HAC_PATH_I8)
            st = idx.stats()
            assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, (min_rows, growth, dist, st)
            assert st["screen_err_max"] <= st["margin_max"], st
            assert np.array_equal(I1, I0) and np.array_equal(D1, D0), (min_rows, growth, dist)
        if min_rows and dist == 2 and n >= 100000:
            assert st["pipelined"] == 1 and st["n_sync_chunks"] < st["n_chunks"], st
    Dm, Im = idx.search(q, k, path=hb.HAC_PATH_MMA)
    assert np.array_equal(Im, I0) and np.array_equal(Dm, D0)
    # the device-tensor API on a caller stream goes through the same two internal streams
    import torch
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        qd = torch.from_numpy(q).cuda()
        Dd, Id = idx.search(qd, k, path=hb.HAC_PATH_I8)
        Dd, Id = Dd.cpu().numpy(), Id.cpu().numpy()
    assert np.array_equal(Id, I0) and np.array_equal(Dd, D0)


@pytest.mark.parametrize("nq,n,k,warm", [(300, 150000, 100, 16384), (130, 90001, 10, 4096), (520, 70000, 1, 65536),
                                         (64, 60000, 100, 8192), (3, 50000, 100, 8192)])
def test_int8_warm_start_changes_nothing_but_the_work(nq, n, k, warm):
    """Warm start of the int8 search: the first rows go through the f16 screen (tight margin), the int8 scan continues
    behind them from an exact threshold.  Results are bitwise those of the plain int8 search and of the f16 screen,
    with fewer rescored pairs; the slab follows resets, appends and a first segment smaller than the request."""
    hb = _engine()
    rng = np.random.default_rng(nq + n + k + warm)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.reserve(n)
    idx.add(x[: n // 2 + 7])
    idx.add(x[n // 2 + 7:])
    idx.set_option("i8_warm_rows", 0)
    D0, I0 = idx.search(q, k, path=hb.HAC_PATH_I8)
    st0 = idx.stats()
    assert st0["path"] == hb.HAC_PATH_I8 and st0["retries"] == 0 and st0["warm_rows"] == 0, st0
    _check(q, x, k, D0, I0, also_fp32_oracle=False)
    for pipeline in (0, 1):
        idx.set_option("i8_pipeline", pipeline)
        idx.set_option("i8_pipe_min_rows", 4096 if pipeline else 0)
        idx.set_option("i8_warm_rows", warm)
        for _ in range(2):
            D1, I1 = idx.search(q, k, path=hb.HAC_PATH_I8)
            st = idx.stats()
            assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, st
            assert st["warm_rows"] == min(warm, n // 256 * 256), st
            assert st["screen_err_max"] <= st["margin_max"], st
            assert np.array_equal(I1, I0) and np.array_equal(D1, D0), (pipeline, st)
        if nq >= 64 and pipeline == 0:
            assert st["candidates_rescored"] < st0["candidates_rescored"], (st, st0)
    idx.set_option("i8_pipeline", 0)
    Dm, Im = idx.search(q, k, path=hb.HAC_PATH_MMA)
    assert np.array_equal(Im, I0) and np.array_equal(Dm, D0)
    # appended rows: the slab (first rows) stays valid, the new rows are scanned by the int8 screen
    extra = rng.standard_normal((5000, 768), dtype=np.float32)
    idx.add(extra)
    x2 = np.concatenate([x, extra])
    D2, I2 = idx.search(q, k, path=hb.HAC_PATH_I8)
    assert idx.stats()["warm_rows"] > 0
    _check(q, x2, k, D2, I2, also_fp32_oracle=False)
    # reset + another corpus in ragged segments (no reserve): the slab is rebuilt and limited to the first segment
    idx.reset()
    y = rng.standard_normal((40000, 768), dtype=np.float32)
    idx.add(y[:9000])
    idx.add(y[9000:])
    D3, I3 = idx.search(q, k, path=hb.HAC_PATH_I8)
    st3 = idx.stats()
    assert st3["path"] == hb.HAC_PATH_I8 and st3["retries"] == 0 and 0 < st3["warm_rows"] <= 40000, st3
    _check(q, y, k, D3, I3, also_fp32_oracle=False)
    idx.set_option("i8_warm_rows", 0)
    D4, I4 = idx.search(q, k, path=hb.HAC_PATH_I8)
    assert np.array_equal(I4, I3) and np.array_equal(D4, D3)


def test_f16_image_is_built_lazily_and_kept_up_to_date():
    """With the int8 image present the f16 image (2 bytes per element of HBM) does not exist until a search needs the
    f16 screen; once built, later adds keep it complete; `lazy_f16 = 0` builds it on add as before."""
    hb = _engine()
    rng = np.random.default_rng(404)
    x = rng.standard_normal((70000, 768), dtype=np.float32)
    q = rng.standard_normal((140, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x[:30000])
    D8, I8 = idx.search(q, 100)
    assert idx.stats()["path"] == hb.HAC_PATH_I8 and idx.stats()["bytes_shadow"] == 0
    Dm, Im = idx.search(q, 100, path=hb.HAC_PATH_MMA)          # builds the image
    assert idx.stats()["bytes_shadow"] >= 30000 * 768 * 2
    assert np.array_equal(Im, I8) and np.array_equal(Dm, D8)
    idx.add(x[30000:30001])                                    # appended into the existing segment / a new one
    idx.add(x[30001:])
    D8, I8 = idx.search(q, 100)
    Dm, Im = idx.search(q, 100, path=hb.HAC_PATH_MMA)
    assert np.array_equal(Im, I8) and np.array_equal(Dm, D8)
    _check(q, x, 100, Dm, Im, also_fp32_oracle=False)
    Dk, Ik = idx.search(q[:40], 300)                           # AUTO with k > 128 on a small shard: the f16 screen
    assert idx.stats()["path"] == hb.HAC_PATH_MMA
    _check(q[:40], x, 300, Dk, Ik, also_fp32_oracle=False)
    idx.reset()
    idx.add(x[:5000])
    Dm, Im = idx.search(q, 10, path=hb.HAC_PATH_MMA)
    _check(q, x[:5000], 10, Dm, Im, also_fp32_oracle=False)
    eager = hb.FlatIPIndex(768)
    eager.set_option("lazy_f16", 0)
    eager.add(x[:9000])
    assert eager.stats()["bytes_shadow"] >= 9000 * 768 * 2
    De, Ie = eager.search(q, 10, path=hb.HAC_PATH_MMA)
    _check(q, x[:9000], 10, De, Ie, also_fp32_oracle=False)


@pytest.mark.parametrize("nq,n,k", [(130, 140000, 300), (200, 150000, 1000)])
def test_large_k_runs_the_int8_screen_behind_a_warm_slab_sized_by_k(nq, n, k):
    """k > 128: AUTO = int8 screen, the automatic f16 warm slab is min(6144 k, rows / 4); bitwise the f16 screen's result."""
    hb = _engine()
    rng = np.random.default_rng(n + k)
    x = rng.standard_normal((n, 768), dtype=np.float32)
    q = rng.standard_normal((nq, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    idx.set_option("i8_large_k_rows_per_k", 0)                 # (the default keeps shards this small on the f16 screen)
    D, I = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0 and st["bytes_shadow"] == 0, st
    assert st["warm_rows"] == (n // 4) // 256 * 256 and st["screen_err_max"] <= st["margin_max"], st
    Dm, Im = idx.search(q, k, path=hb.HAC_PATH_MMA)
    assert np.array_equal(I, Im) and np.array_equal(D, Dm)
    idx.set_option("i8_warm_rows", 0)
    D0, I0 = idx.search(q, k)
    assert idx.stats()["warm_rows"] == 0 and np.array_equal(I0, I) and np.array_equal(D0, D)
    _check(q[:48], x, k, D[:48], I[:48], also_fp32_oracle=False)


def test_default_path_gemv_batches_by_four():
    """`default_path = GEMV` must split an AUTO search into batches of 4 (it used to fail for nq > 4)."""
    hb = _engine()
    rng = np.random.default_rng(9)
    x = rng.standard_normal((6000, 768), dtype=np.float32)
    q = rng.standard_normal((11, 768), dtype=np.float32)
    idx = hb.FlatIPIndex(768)
    idx.add(x)
    idx.set_option("default_path", hb.HAC_PATH_GEMV)
    idx.set_option("i8_auto_max_k", 0)
    D, I = idx.search(q, 10)
    assert idx.stats()["path"] == hb.HAC_PATH_GEMV
    _check(q, x, 10, D, I, also_fp32_oracle=False)


def test_shard_file_round_trip_is_a_plain_dma_reload(tmp_path):
    """hac_save_shard / hac_load_shard: one file per shard with the fp32 rows, the int8 image, its tile
    constants and the centre.  A reloaded index answers bitwise like the one that was saved - on every scan path -
    without running a conversion kernel; multi-segment shards are saved without the image and rebuilt on load."""
    hb = _engine()
    rng = np.random.default_rng(77)
    mu = rng.standard_normal(768).astype(np.float32)
    x = (0.5 * mu + rng.standard_normal((41003, 768))).astype(np.float32)       # a centre that is not zero
    q = (0.5 * mu + rng.standard_normal((140, 768))).astype(np.float32)
    src = hb.FlatIPIndex(768, reserve=41003)
    src.add(x[:20000])
    src.add(x[20000:])
    D0, I0 = src.search(q, 100)
    st0 = src.stats()
    assert st0["path"] == hb.HAC_PATH_I8
    path = str(tmp_path / "shard0.hacs")
    src.save_shard(path)
    import os
    assert os.path.getsize(path) % 4096 == 0 and os.path.getsize(path) >= 41003 * 768 * 5
    dst = hb.FlatIPIndex(768)
    dst.load_shard(path)
    assert dst.ntotal == 41003 and dst.stats()["bytes_shadow"] == 0 and dst.stats()["bytes_i8"] > 0
    D1, I1 = dst.search(q, 100)
    st1 = dst.stats()
    assert st1["path"] == hb.HAC_PATH_I8 and st1["retries"] == 0
    assert np.array_equal(I1, I0) and np.array_equal(D1, D0)
    assert st1["candidates_emitted"] == st0["candidates_emitted"]              # the very same image and thresholds
    for path_id in (hb.HAC_PATH_MMA, hb.HAC_PATH_GEMV):
        Dp, Ip = dst.search(q[:4], 100, path=path_id)
        assert np.array_equal(Ip, I0[:4]) and np.array_equal(Dp, D0[:4])
    _check(q, x, 100, D1, I1, also_fp32_oracle=False)
    dst.add(x[:500])                                                           # a loaded shard keeps growing
    assert dst.ntotal == 41503
    D2, I2 = dst.search(q, 10)
    _check(q, np.concatenate([x, x[:500]], 0), 10, D2, I2, also_fp32_oracle=False)
    with pytest.raises(RuntimeError):
        dst.load_shard(path)                                                   # not empty
    other_d = hb.FlatIPIndex(256)
    with pytest.raises(ValueError):
        other_d.load_shard(path)
    # several segments: rows only, image rebuilt on load
    ragged = hb.FlatIPIndex(768)
    ragged.add(x[:9000])
    ragged.add(x[9000:30001])
    path2 = str(tmp_path / "shard_ragged.hacs")
    ragged.save_shard(path2)
    back = hb.FlatIPIndex(768)
    back.load_shard(path2)
    Dr, Ir = ragged.search(q, 100)
    Db, Ib = back.search(q, 100)
    assert np.array_equal(Ib, Ir) and np.array_equal(Db, Dr)
    # no int8 image wanted on the loading side: the sections are skipped
    plain = hb.FlatIPIndex(768)
    plain.set_option("build_i8", 0)
    plain.load_shard(path)
    Dq, Iq = plain.search(q, 100)
    assert plain.stats()["path"] == hb.HAC_PATH_MMA and plain.stats()["bytes_i8"] == 0
    assert np.array_equal(Iq, I0) and np.array_equal(Dq, D0)
