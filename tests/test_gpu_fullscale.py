"""GPU (-m gpu): BASELINE.json's full size (25.7M x 768, one B200) through size-independent properties.

The CPU oracle cannot hold 79 GB, so parity at this size is pinned by: planted rows with an
analytically known rank and score, an independent exact re-scoring of a few queries with torch
(fp32, TF32 off) over the regenerated corpus, agreement of the two scan paths, prefix consistency
across k, and idempotence."""
import numpy as np
import pytest

from oracle.compare import assert_parity

pytestmark = pytest.mark.gpu

N_ROWS = 25_700_592
DIM = 768


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


def test_full_corpus_rescoring_with_torch_agrees(big_index):
    """Independent exact path: regenerate the corpus in slabs, fp32 matmul (TF32 off) + topk."""
    import torch
    from haconvdr_b200.index import synth_rows_device
    idx, q, n_planted = big_index
    torch.backends.cuda.matmul.allow_tf32 = False
    sel = q[n_planted:n_planted + 8]
    k = 100
    best_s = torch.full((sel.shape[0], k), -float("inf"), device="cuda")
    best_i = torch.full((sel.shape[0], k), -1, dtype=torch.int64, device="cuda")
    slab = 2_000_000
    n_syn = N_ROWS - n_planted
    sel64 = sel.double()
    slabs = [(r0, min(slab, n_syn - r0)) for r0 in range(0, n_syn, slab)] + [(n_syn, n_planted)]
    for r0, n in slabs:
        # the last slab is the planted rows (3 * q_i), which also score high against other queries
        x = synth_rows_device(n, DIM, seed=42, row0=r0) if r0 < n_syn else 3.0 * q[:n_planted]
        s = (sel64 @ x.double().T)                       # fp64 arbiter on the device
        s, i = torch.topk(s, min(k, n), dim=1)
        cat_s = torch.cat([best_s.double(), s], 1)
        cat_i = torch.cat([best_i, i + r0], 1)
        top = torch.topk(cat_s, k, dim=1)
        best_s, best_i = top.values, torch.gather(cat_i, 1, top.indices)
        del x, s
    D, I = idx.search(sel, k)
    ref_D, ref_I = best_s.cpu().numpy(), best_i.cpu().numpy()
    order = np.lexsort((ref_I, -ref_D), axis=1)          # (score desc, id asc)
    ref_D, ref_I = np.take_along_axis(ref_D, order, 1), np.take_along_axis(ref_I, order, 1)
    rep = assert_parity(ref_D, ref_I, D.cpu().numpy(), I.cpu().numpy(), rtol=1e-5)
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
