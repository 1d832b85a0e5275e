"""Host-side mirror of what the PRJ drivers do with the search result
(`/root/reference/src/test_PRJ_topiocqa.py`, `/root/reference/src/test_PRJ_qrecc.py`).

The PRJ drivers run the same block loop as the HAConvDR ones (`:83-171`) over one query per
(turn, candidate history turn) of the *training* set, write a score-less run file (`:232-299`), evaluate
`recip_rank` per query with pytrec_eval (`:326-338`) and label a history turn as useful when its MRR beats the
base query's (`improve_judge`, topiocqa `:443-472`, qrecc `:403-452`).  Only the per-query reciprocal rank of the
first relevant passage is consumed, so here it is computed on the device straight from the search result:

    D, I = index.search(q_dev, top_n)                         # CUDA tensors, offsets
    rr, rank = reciprocal_ranks(I, offset2pid_dev, relevant)  # no run file, no Python loop over Q*k
    labels = improve_judge(sample_ids, rr.tolist())

`reciprocal_ranks` needs the CUDA library; `improve_judge*` are plain Python.
"""
from __future__ import annotations

import json

import numpy as np

from . import _lib
from .index import gather_ids_device


def reciprocal_ranks(I, offset2pid, relevant):
    """``I``: int64 CUDA tensor ``[Q, k]`` of global offsets (``index.search``), ``offset2pid``: int64 CUDA
    tensor, or ``None`` when ``I`` already holds pids; ``relevant``: per query an iterable of relevant pids
    (qrels rows with ``rel >= rel_threshold``, `:312-317`).  Returns ``(rr float32 [Q], rank int32 [Q])`` CUDA
    tensors: rank of the first relevant pid in the deduplicated ranking (0 = not retrieved), rr = 1/rank."""
    import torch
    assert I.is_cuda and I.dtype == torch.int64 and I.dim() == 2
    nq, k = I.shape
    assert len(relevant) == nq, "one relevant-pid list per query"
    pids = gather_ids_device(offset2pid, I) if offset2pid is not None else I.contiguous()
    ptr = np.zeros(nq + 1, np.int64)
    flat = []
    for i, r in enumerate(relevant):
        flat.extend(int(p) for p in r)
        ptr[i + 1] = len(flat)
    dev = I.device
    ptr_d = torch.from_numpy(ptr).to(dev)
    rel_d = torch.tensor(flat if flat else [0], dtype=torch.int64, device=dev)
    rr = torch.empty(nq, dtype=torch.float32, device=dev)
    rank = torch.empty(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().hac_reciprocal_rank_device(dev.index, pids.data_ptr(), nq, k, ptr_d.data_ptr(),
                                                     rel_d.data_ptr(), rr.data_ptr(), rank.data_ptr(), stream),
               "hac_reciprocal_rank_device")
    return rr, rank


def _sample_ids(input_query_file_or_ids):
    if isinstance(input_query_file_or_ids, (list, tuple)):
        return [str(s) for s in input_query_file_or_ids]
    with open(input_query_file_or_ids, "r") as f:
        return [json.loads(line)["id"] for line in f]


def improve_judge(input_query_file, score_list, ori_qrel_file=None):
    """``improve_judge`` of the PRJ drivers: sample ids are ``conv-turn[-...]-type``; within one (conv, turn)
    group the ``type == 0`` sample is the base query and every ``type > 0`` sample gets label 1 iff its score is
    strictly larger than the base score (turn 1 has no history and is skipped).  Returns
    ``{"conv-turn": [labels...], "conv-1": []}`` in the reference's insertion order.

    Without ``ori_qrel_file``: the TopiOCQA variant (`test_PRJ_topiocqa.py:443-472`) - a group closes when the
    next sample's turn id differs.  With it: the QReCC variant (`test_PRJ_qrecc.py:403-452`) - a group also
    closes when the conversation changes, and the empty ``conv-1`` entry is written only when that sample id
    occurs in the qrel file (jsonl with ``sample_id``).  ``input_query_file`` may be the jsonl path or the list of
    sample ids."""
    ids = _sample_ids(input_query_file)
    qrel_ids = None
    if ori_qrel_file is not None:
        with open(ori_qrel_file, "r") as f:
            qrel_ids = {json.loads(line)["sample_id"] for line in f}
    parsed = [s.split("-") for s in ids]
    rel_label, rel_list, base_score = {}, [], 0
    next_turn = next_conv = None
    for i, parts in enumerate(parsed):
        conv_id, turn_id, type_id = int(parts[0]), int(parts[1]), int(parts[-1])
        last = i + 1 == len(parsed)
        if not last:
            # the reference leaves these two at their previous values on the last sample
            next_turn, next_conv = int(parsed[i + 1][1]), int(parsed[i + 1][0])
        if turn_id > 1:
            if type_id == 0:
                base_score = score_list[i]
            elif type_id > 0:
                rel_list.append(1 if score_list[i] > base_score else 0)
        closes = last or turn_id != next_turn
        if qrel_ids is not None:
            closes = closes or (turn_id == next_turn and conv_id != next_conv)
        if closes:
            if qrel_ids is None or (str(conv_id) + "-1") in qrel_ids:
                rel_label[parts[0] + "-1"] = []
            rel_label[parts[0] + "-" + parts[1]] = rel_list
            rel_list, base_score = [], 0
    return rel_label
