"""Cost-faithful restatement of the reference's per-block search loop (TEST / BASELINE INFRASTRUCTURE ONLY).

`/root/reference/src/test_HAConvDR_topiocqa.py:74-162` does, per passage block: unpickle the block and its id
array, ``index.add``, ``index.search``, map local rows to ids, turn ``D`` / ids into per-query lists of
``(score, id)`` TUPLES, ``index.reset``, deep-copy the running lists and merge them with the block's lists by a
two-pointer walk (``>=`` keeps the earlier block first).  ``oracle/flat_ip.search_one_by_one`` reproduces the RESULT of
that loop with vectorised numpy; this module reproduces its COST PROFILE as well - Python tuples, ``copy.deepcopy``,
interpreted while-loops - so that ``bench.py --impl reference`` times what a reference user waits for when the real
module is not mounted (the GPU box).  Where `/root/reference` exists the reference's own function is used instead
(``oracle/ref_harness.py``) and ``tests/test_oracle.py`` checks that both return identical arrays.
"""
from __future__ import annotations

import copy
import os
import pickle

import numpy as np


def search_blocks_python(passage_block_num: int, block_dir: str, index, query_embeddings, topN: int):
    running = None                                   # per query: list of (score, id), best first
    for b in range(int(passage_block_num)):
        try:
            with open(os.path.join(block_dir, "passage_emb_block_%d.pb" % b), "rb") as fh:
                rows = pickle.load(fh)
            with open(os.path.join(block_dir, "passage_embid_block_%d.pb" % b), "rb") as fh:
                row_ids = pickle.load(fh)
        except Exception:                            # the reference stops at the first block it cannot load (:94-95)
            break
        index.add(rows)
        D, I = index.search(query_embeddings, topN)
        ids = row_ids[I].tolist()                    # I == -1 wraps to the block's last id, as upstream (:110)
        scores = D.tolist()
        block_lists = [list(zip(s_row, i_row)) for s_row, i_row in zip(scores, ids)]
        index.reset()
        del rows, row_ids
        if running is None:
            running = block_lists
            continue
        previous = copy.deepcopy(running)
        running = []
        for old, new in zip(previous, block_lists):
            out, a, b2 = [], 0, 0
            while a < topN and b2 < topN:
                if old[a][0] >= new[b2][0]:          # ties: the earlier block stays in front (:138)
                    out.append(old[a])
                    a += 1
                else:
                    out.append(new[b2])
                    b2 += 1
            out.extend(old[a:topN])
            out.extend(new[b2:topN])
            running.append(out)
    if running is None:
        raise TypeError("no passage block could be loaded")
    merged_D = np.array([[c[0] for c in lst] for lst in running])
    merged_I = np.array([[c[1] for c in lst] for lst in running])
    return merged_D, merged_I
