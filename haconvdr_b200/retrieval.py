"""Host-side mirror of the reference's retrieval loop and result mapping.

``search_one_by_one_with_faiss`` keeps the reference's signature and return contract
(`/root/reference/src/test_HAConvDR_topiocqa.py:74-162`; verbatim copies in
``test_HAConvDR_qrecc.py:74-162`` and ``test_PRJ_{topiocqa,qrecc}.py:83-171``):
for every block ``add -> search -> map ids -> reset -> merge``, returning
``merged_D`` float64 and ``merged_I`` int64 of shape ``[Q, 2*topN]`` (``[Q, topN]`` when a
single block loads), rank-for-rank what the reference's Python loops produce - the
per-query tuple lists, ``copy.deepcopy`` and two-pointer merges are replaced by one
vectorised stable merge per block (earlier block wins ties, `:138`).

``search_resident`` is the fast path the B200 enables: every block is streamed once into the
HBM-resident shard (no per-block ``reset``), ids are translated on the device, and a single
search yields the global top-k - the first ``topN`` columns of the reference result.

``rank_pids`` / ``write_trec_run`` mirror ``output_test_res`` (`:220-286`).
"""
from __future__ import annotations

import logging
import time

import numpy as np

from . import loader

logger = logging.getLogger(__name__)


def _block_num(args_or_int) -> int:
    return int(getattr(args_or_int, "passage_block_num", args_or_int))


def _stable_merge(prev_s, prev_i, cur_s, cur_i, topN):
    """Two sorted lists -> stable merge, ``prev`` first on ties; returns all 2*topN entries."""
    cat_s = np.concatenate([prev_s[:, :topN], cur_s[:, :topN]], axis=1)
    cat_i = np.concatenate([prev_i[:, :topN], cur_i[:, :topN]], axis=1)
    order = np.argsort(-cat_s, axis=1, kind="stable")
    return np.take_along_axis(cat_s, order, 1), np.take_along_axis(cat_i, order, 1)


def search_one_by_one_with_faiss(args, passge_embeddings_dir, index, query_embeddings, topN):
    """Drop-in for the reference function of the same name (same arguments, same return)."""
    merged_s = merged_i = None
    for block_id in range(_block_num(args)):
        logger.info("Loading passage block " + str(block_id))
        found = loader.find_block(passge_embeddings_dir, block_id)      # native .hacb if present, else the pickles
        if found is None:
            break                      # reference: bare `except: break` on the first missing block (:94-95)
        emb_path, load_ids = found
        passage_embedding2id = load_ids()
        impl = getattr(index, "_get", lambda: index)()
        if hasattr(impl, "_h"):
            loader.stream_block_into(impl, emb_path)
        else:
            index.add(loader.load_block_array(emb_path))
        logger.info("query embedding shape: " + str(query_embeddings.shape))
        tb = time.time()
        D, I = index.search(query_embeddings, topN)
        elapse = time.time() - tb
        logger.info({
            'time cost': elapse,
            'query num': query_embeddings.shape[0],
            'time cost per query': elapse / query_embeddings.shape[0]
        })
        cand_i = passage_embedding2id[I]                 # :110 - an unfilled slot (-1) wraps to the block's last id
        cand_s = np.asarray(D, dtype=np.float64)         # .tolist() widens fp32 to Python floats (:111)
        index.reset()
        if merged_s is None:
            merged_s, merged_i = cand_s, cand_i.astype(np.int64)
            continue
        merged_s, merged_i = _stable_merge(merged_s, merged_i, cand_s, cand_i, topN)
    if merged_s is None:
        raise FileNotFoundError("no passage block found under %s" % passge_embeddings_dir)
    logger.info(merged_i.shape)
    return merged_s, merged_i


def load_resident(index, passage_embeddings_dir, passage_block_num, row_range=None):
    """Stream the block files into ``index`` once and install the concatenated embedding2id table.

    ``row_range=(lo, hi)`` keeps only global rows [lo, hi) - the contiguous shard of one rank.
    Returns (n_blocks_loaded, rows_loaded)."""
    ids, n_blocks, seen = [], 0, 0
    impl = getattr(index, "_get", lambda: index)()
    for block_id in range(_block_num(passage_block_num)):
        found = loader.find_block(passage_embeddings_dir, block_id)
        if found is None:
            break
        emb_path, load_ids = found
        emb2id = load_ids()
        nb = emb2id.shape[0]
        if row_range is None:
            loader.stream_block_into(impl, emb_path)
            ids.append(emb2id)
        else:
            lo, hi = max(row_range[0], seen), min(row_range[1], seen + nb)
            if hi > lo:
                loader.stream_block_into(impl, emb_path, row_range=(lo - seen, hi - seen))
                ids.append(emb2id[lo - seen:hi - seen])
        seen += nb
        n_blocks += 1
    if n_blocks == 0:
        raise FileNotFoundError("no passage block found under %s" % passage_embeddings_dir)
    table = np.concatenate(ids) if ids else np.zeros(0, np.int64)
    if table.shape[0]:
        impl.set_id_table(table)
    return n_blocks, int(table.shape[0])


def search_resident(index, query_embeddings, topN):
    """One search over the HBM-resident corpus: returns (D float64 [Q, topN], I int64 [Q, topN]),
    equal to ``search_one_by_one_with_faiss(...)[.][:, :topN]``."""
    D, I = index.search(query_embeddings, topN)
    return np.asarray(D, dtype=np.float64), np.asarray(I, dtype=np.int64)


def rank_pids(retrieved_scores_mat, retrieved_pid_mat, offset2pid, top_k):
    """`output_test_res` mapping loop (:232-255): first ``top_k`` columns, offset -> pid, a pid already
    seen for the query is skipped, unfilled trailing slots stay ``(0, 0)``.  Vectorised gather, then a
    per-query first-occurrence filter."""
    table = offset2pid if isinstance(offset2pid, np.ndarray) else np.asarray(offset2pid)
    idx = np.asarray(retrieved_pid_mat)[:, :top_k]
    scores = np.asarray(retrieved_scores_mat)[:, :top_k]
    pids = table[idx]
    out = []
    for qi in range(pids.shape[0]):
        _, first = np.unique(pids[qi], return_index=True)
        keep = np.sort(first)
        ranked = [(pids[qi, j].item(), scores[qi, j].item()) for j in keep]
        ranked += [(0, 0)] * (top_k - len(ranked))
        out.append(ranked)
    return out


def write_trec_run(path, query_ids, ranked, top_k, with_score=True, tag="ance"):
    """Run file exactly as `:273-282` (``with_score=False`` gives the PRJ variant,
    `/root/reference/src/test_PRJ_topiocqa.py:298-299`).  A query id appearing twice behaves as in the
    reference (`:240-255`): one list per qid, a later occurrence overwrites its leading slots only."""
    by_qid = {}
    for qid, passages in zip(query_ids, ranked):
        if qid not in by_qid:
            by_qid[qid] = list(passages)
        else:
            n_new = sum(1 for p in passages if p != (0, 0))
            by_qid[qid][:n_new] = passages[:n_new]
    with open(path, "w") as g:
        for qid, passages in by_qid.items():
            lines = []
            for i in range(top_k):
                pid, score = passages[i]
                if with_score:
                    lines.append(str(qid) + " Q0 " + str(pid) + " " + str(i + 1) + " " + str(-i - 1 + 200)
                                 + " " + str(score) + " " + tag + "\n")
                else:
                    lines.append(str(qid) + " Q0 " + str(pid) + " " + str(i + 1) + " " + str(-i - 1 + 200)
                                 + " " + tag + "\n")
            g.write("".join(lines))
    return path
