"""NumPy restatement of the reference retrieval path (TEST INFRASTRUCTURE ONLY).

What is restated, and from where:

* ``FlatIP`` - ``faiss.IndexFlatIP`` as the reference uses it
  (`/root/reference/src/test_HAConvDR_topiocqa.py:52` ctor, `:98` ``add``,
  `:102` ``search``, `:122` ``reset``).  faiss itself is absent (third-party,
  "faiss-gpu 1.7.2", `/root/reference/README.md:12`), so the published algorithm
  of faiss 1.7.x ``IndexFlatIP::search`` is restated: score = fp32 inner product,
  no normalisation; queries are processed in blocks of 4096 against corpus blocks
  of 1024 rows through an fp32 ``sgemm``; every query keeps a size-k min-heap that
  replaces its root only when ``score > heap_min`` (strict), rows visited in
  increasing index order; the heap is finally reordered to descending score;
  unfilled slots carry id -1 and score -FLT_MAX; outputs are float32 [Q,k] and
  int64 [Q,k].  faiss leaves the order among *equal* scores unspecified; the
  oracle fixes it to (score desc, index asc), which is one of the orders faiss can
  produce and the order the engine guarantees.
* ``search_one_by_one`` - the reference's per-block add/search/reset/merge loop
  (`/root/reference/src/test_HAConvDR_topiocqa.py:74-162`): id map `:110`, stable
  2-way merge with ``>=`` so the earlier block wins ties `:138`, drain loops
  `:144-149`, float64 [Q,2k] / int64 [Q,2k] outputs `:151-159` ([Q,k] when a
  single block loads `:126-128`).
* ``offsets_to_ranked_pids`` - slice to top-k, offset->pid, duplicate-pid
  suppression (`/root/reference/src/test_HAConvDR_topiocqa.py:232-255`).
* ``brute_force_fp64`` - the arbiter: float64 scores, (score desc, index asc).
"""
from __future__ import annotations

import os
import pickle

import numpy as np

NEG_FLT_MAX = np.float32(-3.4028234663852886e38)
QUERY_BLOCK = 4096   # faiss distance_compute_blas_query_bs
CORPUS_BLOCK = 1024  # faiss distance_compute_blas_database_bs


def _topk_desc_stable(scores: np.ndarray, ids: np.ndarray, k: int):
    """Row-wise top-k by (score desc, position asc) - ``ids`` must already be in
    the tie-break order (earlier column == preferred)."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, 1), np.take_along_axis(ids, order, 1)


class FlatIP:
    """Restated ``faiss.IndexFlatIP`` (see module docstring)."""

    def __init__(self, d: int):
        self.d = int(d)
        self._blocks: list[np.ndarray] = []
        self.ntotal = 0

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError("add: expected [n, %d] float32" % self.d)
        self._blocks.append(x.copy())   # faiss copies on add
        self.ntotal += x.shape[0]

    def reset(self) -> None:
        self._blocks = []
        self.ntotal = 0

    def _corpus(self) -> np.ndarray:
        if len(self._blocks) > 1:
            self._blocks = [np.concatenate(self._blocks, axis=0)]
        return self._blocks[0] if self._blocks else np.zeros((0, self.d), np.float32)

    def search(self, q, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise AssertionError("search: expected [nq, %d] float32" % self.d)
        k = int(k)
        if k <= 0:
            raise ValueError("k must be positive")
        x = self._corpus()
        nq, n = q.shape[0], x.shape[0]
        D = np.full((nq, k), NEG_FLT_MAX, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        for q0 in range(0, nq, QUERY_BLOCK):
            q1 = min(nq, q0 + QUERY_BLOCK)
            run_s = np.full((q1 - q0, k), NEG_FLT_MAX, dtype=np.float32)
            run_i = np.full((q1 - q0, k), -1, dtype=np.int64)
            filled = 0
            for j0 in range(0, n, CORPUS_BLOCK):
                j1 = min(n, j0 + CORPUS_BLOCK)
                ip = q[q0:q1] @ x[j0:j1].T            # fp32 sgemm block
                ids = np.broadcast_to(np.arange(j0, j1, dtype=np.int64), ip.shape)
                # running list first: it holds earlier indices, so a stable sort
                # reproduces "strict > against the heap minimum".
                keep = min(filled, k)
                cat_s = np.concatenate([run_s[:, :keep], ip], axis=1)
                cat_i = np.concatenate([run_i[:, :keep], ids], axis=1)
                top_s, top_i = _topk_desc_stable(cat_s, cat_i, k)
                filled = min(k, filled + (j1 - j0))
                run_s[:, :top_s.shape[1]] = top_s
                run_i[:, :top_i.shape[1]] = top_i
            D[q0:q1] = run_s
            I[q0:q1] = run_i
        return D, I


def brute_force_fp64(q, x, k: int):
    """Arbiter: float64 inner products, top-k by (score desc, index asc)."""
    q = np.asarray(q, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    s = q @ x.T
    ids = np.broadcast_to(np.arange(x.shape[0], dtype=np.int64), s.shape)
    kk = min(k, x.shape[0])
    D, I = _topk_desc_stable(s, ids, kk)
    if kk < k:
        pad = k - kk
        D = np.concatenate([D, np.full((q.shape[0], pad), float(NEG_FLT_MAX))], 1)
        I = np.concatenate([I, np.full((q.shape[0], pad), -1, np.int64)], 1)
    return D, I


def search_one_by_one(passage_block_num: int, block_dir: str, index, query_embeddings, topN: int):
    """Restatement of ``search_one_by_one_with_faiss``
    (`/root/reference/src/test_HAConvDR_topiocqa.py:74-162`) over any object with
    the faiss surface.  Vectorised, but rank-for-rank identical to the reference's
    Python loops: the merge keeps `merged[p1] >= cur[p2]` -> earlier block first."""
    merged_s = merged_i = None
    n_blocks = 0
    for block_id in range(passage_block_num):
        try:
            with open(os.path.join(block_dir, "passage_emb_block_%d.pb" % block_id), "rb") as h:
                emb = pickle.load(h)
            with open(os.path.join(block_dir, "passage_embid_block_%d.pb" % block_id), "rb") as h:
                emb2id = pickle.load(h)
        except Exception:       # reference: bare `except: break` (:94-95)
            break
        index.add(emb)
        D, I = index.search(query_embeddings, topN)
        cand_i = emb2id[I]                       # :110 (I == -1 wraps, as upstream)
        cand_s = np.asarray(D, dtype=np.float64)  # .tolist() widens fp32 -> python float
        index.reset()
        n_blocks += 1
        if merged_s is None:
            merged_s, merged_i = cand_s, np.asarray(cand_i, dtype=np.int64)
            continue
        # two-pointer merge of merged[:topN] and cur[:topN]; ties -> merged first.
        a_s, a_i = merged_s[:, :topN], merged_i[:, :topN]
        cat_s = np.concatenate([a_s, cand_s[:, :topN]], axis=1)
        cat_i = np.concatenate([a_i, cand_i[:, :topN]], axis=1)
        order = np.argsort(-cat_s, axis=1, kind="stable")
        merged_s = np.take_along_axis(cat_s, order, 1)
        merged_i = np.take_along_axis(cat_i, order, 1).astype(np.int64)
    if merged_s is None:
        raise TypeError("no passage block could be loaded")  # reference iterates None (:152)
    return merged_s, merged_i


def offsets_to_ranked_pids(retrieved_scores_mat, retrieved_pid_mat, offset2pid, top_k: int):
    """`/root/reference/src/test_HAConvDR_topiocqa.py:232-255`: per query take the
    first ``top_k`` columns, map offset->pid, skip already-seen pids; unfilled
    trailing slots stay ``(0, 0)``.  Returns list (per query) of top_k (pid, score)."""
    out = []
    for qi in range(len(retrieved_pid_mat)):
        seen = set()
        ranked = [(0, 0)] * top_k
        rank = 0
        sel_idx = retrieved_pid_mat[qi][:top_k]
        sel_score = retrieved_scores_mat[qi][:top_k].tolist()
        for idx, score in zip(sel_idx, sel_score):
            pid = offset2pid[idx]
            if pid not in seen:
                ranked[rank] = (pid, score)
                rank += 1
                seen.add(pid)
        out.append(ranked)
    return out


def trec_lines(query_ids, ranked, top_k: int):
    """Run-file lines exactly as `/root/reference/src/test_HAConvDR_topiocqa.py:282`."""
    lines = []
    for qid, passages in zip(query_ids, ranked):
        for i in range(top_k):
            pid, score = passages[i]
            lines.append(str(qid) + " Q0 " + str(pid) + " " + str(i + 1) + " "
                         + str(-i - 1 + 200) + " " + str(score) + " ance\n")
    return lines
