"""CPU: pin the oracle against the golden fixtures (made by the reference's own
functions, tests/golden/make_golden.py), analytic known answers and the fp64 arbiter."""
import tempfile

import numpy as np
import pytest

from oracle import cpu_port
from oracle.compare import assert_parity, compare_topk
from oracle.flat_ip import (FlatIP, NEG_FLT_MAX, brute_force_fp64, offsets_to_ranked_pids,
                            search_one_by_one, trec_lines)
from helpers import GOLDEN_MERGE_CASES, load_golden, write_blocks


@pytest.mark.parametrize("name", GOLDEN_MERGE_CASES)
def test_restated_loop_matches_reference_golden(name):
    g = load_golden(name)
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, g["blocks"], g["id_start"])
        nb = int(g.get("block_num", len(g["blocks"]) + 3))
        D, I = search_one_by_one(nb, d, FlatIP(g["q"].shape[1]), g["q"], g["k"])
    assert D.dtype == np.float64 and I.dtype == np.int64
    assert D.shape == g["D"].shape and I.shape == g["I"].shape
    assert np.array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", GOLDEN_MERGE_CASES)
def test_cost_faithful_python_loop_matches_reference_golden(name):
    """oracle/ref_loop.py (tuples + deepcopy + two-pointer merges: what bench.py --impl reference times when the
    reference module is not mounted) returns exactly what the reference's own function returned."""
    from oracle.ref_loop import search_blocks_python
    g = load_golden(name)
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, g["blocks"], g["id_start"])
        nb = int(g.get("block_num", len(g["blocks"]) + 3))
        D, I = search_blocks_python(nb, d, FlatIP(g["q"].shape[1]), g["q"], g["k"])
    assert D.dtype == np.float64 and I.dtype == np.int64
    assert D.shape == g["D"].shape and np.array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], rtol=1e-6, atol=0)


def test_integer_known_answers_are_exact():
    g = load_golden("kat_int_d768_1block")
    x, q = g["blocks"][0], g["q"]
    exact = (q.astype(np.int64) @ x.astype(np.int64).T)          # exact integer scores
    order = np.argsort(-exact, axis=1, kind="stable")[:, : g["k"]]
    assert np.array_equal(g["I"], order)
    assert np.array_equal(g["D"], np.take_along_axis(exact, order, 1).astype(np.float64))


def test_trec_lines_match_reference_run_file():
    g = load_golden("trec_run_dedup_d64")
    ranked = offsets_to_ranked_pids(g["D"], g["I"], g["offset2pid"].tolist(), g["k"])
    assert "".join(trec_lines(g["qids"].tolist(), ranked, g["k"])) == str(g["run_text"])


def test_prj_golden_uses_the_same_loop():
    g = load_golden("prj_run_d64")
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, g["blocks"], 0)
        D, I = search_one_by_one(10, d, FlatIP(64), g["q"], g["k"])
    assert np.array_equal(I, g["I"])


def test_flat_ip_against_fp64_arbiter_and_c_port():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((5000, 768), dtype=np.float32)
    q = rng.standard_normal((23, 768), dtype=np.float32)
    idx = FlatIP(768)
    idx.add(x[:1234])
    idx.add(x[1234:])
    assert idx.ntotal == 5000
    D, I = idx.search(q, 100)
    assert D.dtype == np.float32 and I.dtype == np.int64
    assert np.all(np.diff(D, axis=1) <= 0)
    Da, Ia = brute_force_fp64(q, x, 100)
    xs = x.astype(np.float64)
    rep = assert_parity(Da, Ia, D, I, ref_scores_of=lambda qi, ids: xs[ids] @ q[qi].astype(np.float64))
    assert rep.recall == 1.0
    Db, Ib = cpu_port.search_blas(q, x, 100)
    assert np.array_equal(Ib, I) and np.array_equal(Db, D)
    Dn, In = cpu_port.search_naive(q, x, 100)
    assert_parity(Da, Ia, Dn, In)
    idx.reset()
    assert idx.ntotal == 0


def test_short_corpus_fill_and_ties():
    x = np.zeros((7, 64), np.float32)
    x[:, 0] = [3, 1, 3, 2, 3, 0, 1]
    q = np.zeros((2, 64), np.float32)
    q[0, 0], q[1, 0] = 1, -1
    for search in (lambda: FlatIP(64), None):
        if search is None:
            D, I = cpu_port.search_blas(q, x, 10)
        else:
            idx = search()
            idx.add(x)
            D, I = idx.search(q, 10)
        assert I[0].tolist() == [0, 2, 4, 3, 1, 6, 5, -1, -1, -1]
        assert I[1].tolist() == [5, 1, 6, 3, 0, 2, 4, -1, -1, -1]
        assert np.all(D[:, 7:] == NEG_FLT_MAX)


def test_comparator_rules():
    ref_D = np.array([[10.0, 9.0, 9.0 * (1 + 5e-6), 5.0]])
    ref_D = -np.sort(-ref_D, axis=1)
    ref_I = np.array([[4, 7, 8, 1]])
    # swap inside the near-tie group: accepted
    assert compare_topk(ref_D, ref_I, ref_D, np.array([[4, 8, 7, 1]])).ok
    # swap across groups: rejected
    assert not compare_topk(ref_D, ref_I, ref_D, np.array([[7, 4, 8, 1]])).ok
    # boundary substitution accepted only with a within-tolerance true score
    got_I = np.array([[4, 7, 8, 2]])
    assert not compare_topk(ref_D, ref_I, ref_D, got_I).ok
    assert compare_topk(ref_D, ref_I, ref_D, got_I, ref_scores_of=lambda qi, ids: np.full(len(ids), 5.0)).ok
    assert not compare_topk(ref_D, ref_I, ref_D, got_I, ref_scores_of=lambda qi, ids: np.full(len(ids), 4.9)).ok
    # score drift beyond tolerance: rejected
    assert not compare_topk(ref_D, ref_I, ref_D * (1 + 3e-5), ref_I).ok
