"""Tolerance-aware top-k comparator (TEST INFRASTRUCTURE ONLY).

Implements the parity bar of BASELINE.json ``north_star`` / SURVEY.md section 8d:
returned ids and their order must be identical to the reference's, except where
reference scores tie or fall within ``rtol`` relative of each other; scores must
agree within ``rtol`` relative; recall@k overlap with the reference must be 1.0.

The rule, per query:
  * reference ranks are grouped: consecutive ranks whose scores differ by at most
    ``rtol * |score|`` chain into one tolerance group (so a group is a maximal run
    of near-ties);
  * inside a group the candidate may permute ids freely; across groups the order
    must be identical;
  * the last group is open-ended: any id whose reference score (taken from
    ``ref_scores_of``, e.g. the fp64 arbiter over the full corpus) is within
    tolerance of the k-th reference score may stand in for a member of that group;
  * every candidate score must be within ``rtol`` relative of the reference score
    of the same id.
"""
from __future__ import annotations

import numpy as np


class ParityReport:
    def __init__(self):
        self.n_queries = 0
        self.n_exact_rows = 0          # rows with identical id order
        self.n_tolerated_rows = 0      # rows equal up to tolerance groups
        self.failures: list[str] = []
        self.max_rel_score_err = 0.0
        self.recall = 1.0

    @property
    def ok(self) -> bool:
        return not self.failures

    def __repr__(self):
        return ("ParityReport(queries=%d exact=%d tolerated=%d failures=%d "
                "max_rel_score_err=%.3g recall=%.6f)" % (
                    self.n_queries, self.n_exact_rows, self.n_tolerated_rows,
                    len(self.failures), self.max_rel_score_err, self.recall))


def _groups(scores: np.ndarray, rtol: float, atol: float = 0.0):
    """Yield (start, stop) index ranges of chained near-tie groups of a descending row."""
    k = len(scores)
    start = 0
    for r in range(1, k + 1):
        if r == k or abs(scores[r - 1] - scores[r]) > rtol * max(abs(scores[r - 1]), abs(scores[r])) + atol:
            yield start, r
            start = r


def compare_topk(ref_D, ref_I, got_D, got_I, rtol: float = 1e-5, ref_scores_of=None,
                 max_report: int = 8, atol: float = 0.0) -> ParityReport:
    """Compare candidate (got_D, got_I) with reference (ref_D, ref_I).

    ``ref_scores_of(qi, ids) -> scores`` (optional) returns reference-quality scores
    for arbitrary ids of query ``qi``; it lets the comparator accept a boundary
    substitution whose true score is within tolerance of the k-th reference score.
    Filler entries (id -1) must match exactly in position.

    ``atol`` (default 0) adds an absolute term to every tolerance.  It exists only for
    synthetic cases whose returned scores pass through zero (k >= corpus size, tiny d): a
    relative bound is meaningless for a score that is itself a cancellation residue of
    size 1e-7 * sum|q_i x_i|.  The workloads of BASELINE.json never return such scores.
    """
    ref_D = np.asarray(ref_D, dtype=np.float64)
    got_D = np.asarray(got_D, dtype=np.float64)
    ref_I = np.asarray(ref_I)
    got_I = np.asarray(got_I)
    rep = ParityReport()
    if ref_I.shape != got_I.shape or ref_D.shape != got_D.shape:
        rep.failures.append("shape mismatch: ref %s/%s got %s/%s" % (
            ref_D.shape, ref_I.shape, got_D.shape, got_I.shape))
        return rep
    nq, k = ref_I.shape
    rep.n_queries = nq
    hit = tot = 0
    for qi in range(nq):
        rI, gI, rD, gD = ref_I[qi], got_I[qi], ref_D[qi], got_D[qi]
        valid = rI >= 0
        nv = int(valid.sum())
        if nv > 1 and np.any(np.diff(gD[:nv]) > 0):
            rep.failures.append("q%d: returned scores are not non-increasing" % qi)
            continue
        if not np.array_equal(gI[nv:], rI[nv:]):
            rep.failures.append("q%d: filler slots differ" % qi)
            continue
        tot += nv
        if np.array_equal(rI, gI):
            rep.n_exact_rows += 1
            hit += nv
            err = np.maximum(np.abs(gD[:nv] - rD[:nv]) - atol, 0.0) / np.maximum(np.abs(rD[:nv]), 1e-30)
            if nv:
                rep.max_rel_score_err = max(rep.max_rel_score_err, float(err.max()))
                if err.max() > rtol:
                    rep.failures.append("q%d: score rel err %.3g > %g" % (qi, err.max(), rtol))
            continue
        ok = True
        ref_score_by_id = {int(i): float(s) for i, s in zip(rI[:nv], rD[:nv])}
        groups = list(_groups(rD[:nv], rtol, atol))
        for gi, (a, b) in enumerate(groups):
            rset, gset = set(rI[a:b].tolist()), set(gI[a:b].tolist())
            if rset == gset:
                hit += b - a
                continue
            last = gi == len(groups) - 1
            extra = sorted(gset - rset)
            if last and ref_scores_of is not None and len(gset) == b - a:
                s_extra = np.asarray(ref_scores_of(qi, np.asarray(extra, dtype=np.int64)), dtype=np.float64)
                kth = rD[nv - 1]
                tol = rtol * max(abs(kth), 1e-30) + atol
                if np.all(np.abs(s_extra - kth) <= tol) and np.all(s_extra <= rD[a] + tol):
                    for i, s in zip(extra, s_extra):
                        ref_score_by_id[int(i)] = float(s)
                    hit += b - a
                    continue
            ok = False
            hit += len(rset & gset)
            if len(rep.failures) < max_report:
                rep.failures.append("q%d ranks[%d:%d): ref ids %s != got ids %s" % (
                    qi, a, b, sorted(rset - gset)[:6], extra[:6]))
            else:
                rep.failures.append("q%d" % qi)
        if ok:
            rep.n_tolerated_rows += 1
            for i, s in zip(gI[:nv].tolist(), gD[:nv].tolist()):
                rs = ref_score_by_id.get(int(i))
                if rs is None:
                    continue
                e = max(abs(s - rs) - atol, 0.0) / max(abs(rs), 1e-30)
                rep.max_rel_score_err = max(rep.max_rel_score_err, e)
                if e > rtol:
                    rep.failures.append("q%d id %d: score rel err %.3g > %g" % (qi, i, e, rtol))
                    break
    rep.recall = (hit / tot) if tot else 1.0
    return rep


def assert_parity(ref_D, ref_I, got_D, got_I, rtol: float = 1e-5, ref_scores_of=None, atol: float = 0.0):
    rep = compare_topk(ref_D, ref_I, got_D, got_I, rtol=rtol, ref_scores_of=ref_scores_of, atol=atol)
    assert rep.ok and rep.recall == 1.0, "%r\n%s" % (rep, "\n".join(rep.failures[:12]))
    return rep
