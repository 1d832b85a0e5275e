"""GPU (-m gpu): BASELINE.json's full size (25.7M x 768, one B200) through size-independent properties.

The CPU oracle cannot hold 79 GB, so parity at this size is pinned by: planted rows with an
analytically known rank and score, an independent exact re-scoring of a few queries with torch
(fp32, TF32 off) over the regenerated corpus, agreement of the two scan paths, prefix consistency
across k, and idempotence."""
import numpy as np
import pytest

from oracle.compare import assert_parity
from fullscale_util import DIM, N_QUERIES, N_ROWS, arbiter_topk, queries_over_all_tiles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_index():
    import torch
    import haconvdr_b200 as hb
    from haconvdr_b200.index import synth_rows_device
    free, _ = torch.cuda.mem_get_info()
    if free < 130e9:
        pytest.skip("needs ~125 GB of free HBM")
    n_planted = 16
    q = synth_rows_device(40, DIM, seed=4242)
    idx = hb.FlatIPIndex(DIM, 0, reserve=N_ROWS)
    idx.add_synthetic(N_ROWS - n_planted, seed=42, row0=0)
    planted = 3.0 * q[:n_planted]                    # row N-16+i = 3*q_i  ->  score 3*|q_i|^2 >> any random row
    idx.add(planted)
    assert idx.ntotal == N_ROWS
    yield idx, q, n_planted
    idx.close()


def test_planted_rows_rank_first_with_exact_score(big_index):
    import torch
    idx, q, n_planted = big_index
    D, I = idx.search(q, 100)
    st = idx.stats()
    assert st["retries"] == 0 and st["screen_err_max"] <= st["margin_max"]
    D, I = D.cpu().numpy(), I.cpu().numpy()
    qn = q.cpu().numpy().astype(np.float64)
    for i in range(n_planted):
        assert I[i, 0] == N_ROWS - n_planted + i
        np.testing.assert_allclose(D[i, 0], 3.0 * (qn[i] @ qn[i]), rtol=1e-5)
    assert np.all(np.diff(D, axis=1) <= 0)
    # rank-100 score of an N(0,1) corpus sits near 4.5 sigma = 4.5 * sqrt(768) ~ 124
    assert 110 < float(np.median(D[n_planted:, 99])) < 140


def test_headline_batch_all_query_tiles_against_fp64_arbiter(big_index):
    """The shape bench.py times: 2514 queries (20 query tiles -> the CTA-pair int8 scan, tile-major, pipelined chunks)
    over 25.7M rows.  260 queries spread over ALL 20 tiles are compared - ids, ORDER and scores, tolerance groups per
    SURVEY 8d - with an fp64 re-scoring of the regenerated corpus; the rest through idempotence and planted rows."""
    import torch
    import haconvdr_b200 as hb
    from haconvdr_b200.index import synth_rows_device
    idx, q, n_planted = big_index
    q_all = synth_rows_device(N_QUERIES, DIM, seed=4242)
    assert torch.equal(q_all[:40], q)
    D, I = idx.search(q_all, 100)
    st = idx.stats()
    assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, st
    assert st["screen_err_max"] <= st["margin_max"], st
    assert st["bytes_shadow"] == 0, "the f16 image must not exist before a search needs it"
    sel = queries_over_all_tiles()
    assert len(sel) >= 256 and len(set((sel // 128).tolist())) == 20
    sel_t = torch.from_numpy(sel).cuda()
    ref_D, ref_I, scores_of = arbiter_topk(q_all[sel_t], N_ROWS - n_planted, 100, planted=3.0 * q[:n_planted])
    got_D, got_I = D[sel_t].cpu().numpy(), I[sel_t].cpu().numpy()
    rep = assert_parity(ref_D, ref_I, got_D, got_I, rtol=1e-5, ref_scores_of=scores_of)
    assert rep.recall == 1.0 and rep.n_queries == len(sel)
    Dn, In = D.cpu().numpy(), I.cpu().numpy()
    assert np.all(np.diff(Dn, axis=1) <= 0) and In.min() >= 0 and In.max() < N_ROWS
    for i in range(n_planted):
        assert In[i, 0] == N_ROWS - n_planted + i
    # the pipelined schedule, the search without the f16 warm start and the single-CTA scan return bitwise the same
    # 2514 x 100 result as the default (chunk-synchronous schedule behind a warm start, CTA pairs)
    assert st["pipelined"] == 0 and st["warm_rows"] > 0, st
    idx.set_option("i8_pipeline", 1)
    D2, I2 = idx.search(q_all, 100)
    assert idx.stats()["pipelined"] == 1
    idx.set_option("i8_pipeline", 0)
    assert torch.equal(I2, I) and torch.equal(D2, D)
    for opt, val, back in (("i8_warm_rows", 0, -1), ("i8_warm_rows", 65536, -1), ("i8_cta_group", 1, 2)):
        idx.set_option(opt, val)
        D3, I3 = idx.search(q_all, 100)
        idx.set_option(opt, back)
        assert torch.equal(I3, I) and torch.equal(D3, D), (opt, val)
    # small batches of the same queries (single-tile scan, HBM-bound) agree with the rows of the big batch
    for nq in (1, 4, 32):
        Ds, Is = idx.search(q_all[:nq], 100)
        assert torch.equal(Is, I[:nq]) and torch.equal(Ds, D[:nq]), nq


def test_full_corpus_rescoring_with_torch_agrees(big_index):
    """Independent exact path: regenerate the corpus in slabs, fp64 matmul + topk (fullscale_util.arbiter_topk)."""
    idx, q, n_planted = big_index
    sel = q[n_planted:n_planted + 8]
    ref_D, ref_I, scores_of = arbiter_topk(sel, N_ROWS - n_planted, 100, planted=3.0 * q[:n_planted])
    D, I = idx.search(sel, 100)
    rep = assert_parity(ref_D, ref_I, D.cpu().numpy(), I.cpu().numpy(), rtol=1e-5, ref_scores_of=scores_of)
    assert rep.recall == 1.0


def test_scan_paths_agree_bitwise_and_results_are_stable(big_index):
    import torch
    import haconvdr_b200 as hb
    idx, q, n_planted = big_index
    q4 = q[20:24]
    Dg, Ig = idx.search(q4, 100, path=hb.HAC_PATH_GEMV)
    Dm, Im = idx.search(q4, 100, path=hb.HAC_PATH_MMA)
    assert torch.equal(Ig, Im) and torch.equal(Dg, Dm)
    D1, I1 = idx.search(q, 100)
    D2, I2 = idx.search(q, 100)
    assert torch.equal(I1, I2) and torch.equal(D1, D2)             # idempotent
    D10, I10 = idx.search(q, 10)
    assert torch.equal(I10, I1[:, :10]) and torch.equal(D10, D1[:, :10])   # prefix consistency across k
    Dk, Ik = idx.search(q[:8], 1000)
    assert torch.equal(Ik[:, :100], I1[:8]) and torch.equal(Dk[:, :100], D1[:8])
