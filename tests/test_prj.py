"""PRJ drivers' use of the search result: `improve_judge` mirror against golden outputs of the
reference's own functions (CPU), fused reciprocal rank against the oracle restatement (GPU)."""
import json
import os

import numpy as np
import pytest

from oracle import prj_eval

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _judge_golden():
    with open(os.path.join(GOLDEN, "prj_improve_judge.json")) as f:
        return json.load(f)


def test_improve_judge_matches_both_reference_variants(tmp_path):
    from haconvdr_b200.prj import improve_judge
    g = _judge_golden()
    out_t = improve_judge(g["ids"], g["scores"])
    assert [[k, v] for k, v in out_t.items()] == g["topiocqa"]
    qrel = tmp_path / "qrel.json"
    qrel.write_text("".join(json.dumps({"sample_id": s}) + "\n" for s in g["qrel_ids"]))
    qfile = tmp_path / "q.json"
    qfile.write_text("".join(json.dumps({"id": s}) + "\n" for s in g["ids"]))
    out_q = improve_judge(str(qfile), g["scores"], str(qrel))
    assert [[k, v] for k, v in out_q.items()] == g["qrecc"]
    assert g["topiocqa"] != g["qrecc"]            # the fixture does exercise the variants' difference


def test_oracle_recip_rank_follows_the_run_file_quirks():
    # plain case
    ranked = [(7, 1.0), (3, 0.9), (9, 0.8), (0, 0), (0, 0)]
    assert prj_eval.recip_rank(ranked, 5, [9]) == (1.0 / 3, 3)
    assert prj_eval.recip_rank(ranked, 5, [4]) == (0.0, 0)
    # the padding's pid 0 is a document of the run: ranked once, behind the real ones
    assert prj_eval.recip_rank(ranked, 5, [0]) == (1.0 / 4, 4)
    # a real pid 0 ranked first is overwritten by the padding lines and moves to the end
    ranked = [(0, 1.0), (3, 0.9), (0, 0), (0, 0)]
    assert prj_eval.recip_rank(ranked, 4, [0]) == (1.0 / 2, 2)
    assert prj_eval.recip_rank(ranked, 4, [3]) == (1.0, 1)
    # no padding: pid 0 keeps its place
    ranked = [(0, 1.0), (3, 0.9)]
    assert prj_eval.recip_rank(ranked, 2, [0]) == (1.0, 1)


@pytest.mark.gpu
def test_fused_reciprocal_rank_matches_the_run_file_route():
    import torch
    from haconvdr_b200.prj import reciprocal_ranks
    from haconvdr_b200.retrieval import rank_pids
    rng = np.random.default_rng(99)
    nq, k, n_off = 300, 100, 5000
    offset2pid = rng.integers(0, 900, size=n_off).astype(np.int64)        # many offsets share a pid: dedup + padding
    offset2pid[:50] = 0                                                   # and pid 0 is a real passage
    I = np.stack([rng.choice(n_off, size=k, replace=False) for _ in range(nq)]).astype(np.int64)
    D = -np.sort(-rng.standard_normal((nq, k)), axis=1)
    relevant = [rng.choice(900, size=int(rng.integers(0, 4)), replace=False).tolist() for _ in range(nq)]
    for qi in range(0, nq, 7):
        relevant[qi] = relevant[qi] + [0]
    ranked = rank_pids(D, I, offset2pid, k)
    want = [prj_eval.recip_rank(ranked[qi], k, relevant[qi]) for qi in range(nq)]
    rr, rank = reciprocal_ranks(torch.from_numpy(I).cuda(), torch.from_numpy(offset2pid).cuda(), relevant)
    assert rank.cpu().tolist() == [w[1] for w in want]
    np.testing.assert_array_equal(rr.cpu().numpy(), np.asarray([w[0] for w in want], dtype=np.float32))
    assert sum(1 for w in want if w[1] > 0) > 50
    # injective mapping, no padding: plain first-hit rank
    I2 = torch.arange(40, dtype=torch.int64, device="cuda").repeat(3, 1)
    rr2, rank2 = reciprocal_ranks(I2, None, [[5], [39, 2], []])
    assert rank2.cpu().tolist() == [6, 3, 0]
