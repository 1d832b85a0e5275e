"""CPU restatement of the PRJ drivers' per-query evaluation (TEST INFRASTRUCTURE ONLY - never imported by
the product path).

`/root/reference/src/test_PRJ_topiocqa.py:232-255` builds, per query, ``topN`` ``(pid, score)`` tuples
(duplicate pids skipped, unfilled trailing slots ``(0, 0)``), `:290-299` writes them as run lines whose 5th column
is ``200 - rank``, `:318-338` reads the lines back into ``runs[qid][pid] = int(col5)`` (a repeated pid keeps its
LAST line) and lets pytrec_eval compute ``recip_rank`` against the binarised qrels.

pytrec_eval is not installed in this image and is not under /root/reference (parity of this function is
**unpinned**: restated from trec_eval's documented behaviour - documents ranked by score descending, ties by
document id descending; ``recip_rank`` = 1 / rank of the first document judged relevant, 0 if none).
"""
from __future__ import annotations


def run_scores(ranked, top_n):
    """``runs[qid]`` as the reference's reader builds it from its own run lines (`:318-327`)."""
    run = {}
    for i in range(top_n):
        pid, _ = ranked[i]
        run[str(pid)] = -i - 1 + 200
    return run


def recip_rank(ranked, top_n, relevant_pids):
    run = run_scores(ranked, top_n)
    rel = {str(p) for p in relevant_pids}
    # trec_eval: score descending, then docno descending
    order = sorted(run.items(), key=lambda kv: kv[0], reverse=True)
    order.sort(key=lambda kv: kv[1], reverse=True)
    for r, (pid, _) in enumerate(order, 1):
        if pid in rel:
            return 1.0 / r, r
    return 0.0, 0
