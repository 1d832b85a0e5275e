"""Shared helpers for the parity tests (fixtures <-> block pickles, data makers)."""
import os
import pickle

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    blocks = []
    while "x%d" % len(blocks) in g:
        blocks.append(np.ascontiguousarray(g["x%d" % len(blocks)], dtype=np.float32))
    g["blocks"] = blocks
    g["q"] = np.ascontiguousarray(g["q"], dtype=np.float32)
    g["k"] = int(g["k"])
    g["id_start"] = int(g.get("id_start", 0))
    return g


def write_blocks(dirname, blocks, id_start=0):
    o = id_start
    for i, b in enumerate(blocks):
        with open(os.path.join(dirname, "passage_emb_block_%d.pb" % i), "wb") as h:
            pickle.dump(np.ascontiguousarray(b, np.float32), h, protocol=4)
        with open(os.path.join(dirname, "passage_embid_block_%d.pb" % i), "wb") as h:
            pickle.dump(np.arange(o, o + len(b), dtype=np.int64), h, protocol=4)
        o += len(b)


GOLDEN_MERGE_CASES = [
    "kat_int_d768_1block", "merge_3blocks_d64", "merge_ties_across_blocks_d64",
    "short_single_block_d64", "short_two_blocks_d64", "block_num_limit_d64", "single_query_k1_d128",
    "trec_run_dedup_d64", "prj_run_d64",
]
