"""GPU (-m gpu): the full-size corpus (25.7M x 768) appended in ragged pieces, so that the shard is THREE
segments of different sizes and the scans / rescores cross segment boundaries at the size bench.py runs.
Parity against the fp64 arbiter over the regenerated corpus (tests/fullscale_util.py)."""
import numpy as np
import pytest

from oracle.compare import assert_parity
from fullscale_util import DIM, N_QUERIES, N_ROWS, arbiter_topk, queries_over_all_tiles

pytestmark = pytest.mark.gpu


def test_three_ragged_segments_at_full_size():
    import torch
    import haconvdr_b200 as hb
    from haconvdr_b200.index import synth_rows_device
    free, _ = torch.cuda.mem_get_info()
    if free < 130e9:
        pytest.skip("needs ~125 GB of free HBM")
    idx = hb.FlatIPIndex(DIM, 0)                       # no reserve: every add grows the shard
    pieces = [5_000_000, 7_000_077, N_ROWS - 12_000_077]
    row0 = 0
    for n in pieces:
        idx.add_synthetic(n, seed=42, row0=row0)
        row0 += n
    assert idx.ntotal == N_ROWS
    st0 = idx.stats()
    assert st0["bytes_fp32"] >= N_ROWS * DIM * 4 and st0["bytes_shadow"] == 0
    q_all = synth_rows_device(N_QUERIES, DIM, seed=4242)
    D, I = idx.search(q_all, 100)
    st = idx.stats()
    assert st["path"] == hb.HAC_PATH_I8 and st["retries"] == 0, st
    assert st["screen_err_max"] <= st["margin_max"], st
    sel = queries_over_all_tiles()
    sel_t = torch.from_numpy(sel).cuda()
    ref_D, ref_I, scores_of = arbiter_topk(q_all[sel_t], N_ROWS, 100)
    rep = assert_parity(ref_D, ref_I, D[sel_t].cpu().numpy(), I[sel_t].cpu().numpy(), rtol=1e-5, ref_scores_of=scores_of)
    assert rep.recall == 1.0 and rep.n_queries >= 256
    # rows of all three segments appear among the results
    In = I.cpu().numpy()
    assert (In < pieces[0]).any() and ((In >= pieces[0]) & (In < pieces[0] + pieces[1])).any() \
        and (In >= pieces[0] + pieces[1]).any()
    # single-tile / HBM-bound batches and the exact GEMV scan over the same three segments
    for nq in (1, 4):
        Ds, Is = idx.search(q_all[:nq], 100)
        assert torch.equal(Is, I[:nq]) and torch.equal(Ds, D[:nq])
        Dg, Ig = idx.search(q_all[:nq], 100, path=hb.HAC_PATH_GEMV)
        assert torch.equal(Ig, I[:nq]) and torch.equal(Dg, D[:nq])
    idx.close()
