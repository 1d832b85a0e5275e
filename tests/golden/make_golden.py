"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

    python tests/golden/make_golden.py

Every fixture is produced by the reference's OWN functions
(`/root/reference/src/test_HAConvDR_topiocqa.py:74-162` search_one_by_one_with_faiss
and `:220-286` output_test_res), imported unmodified through oracle/ref_harness.py,
driven with the oracle index (oracle/flat_ip.py FlatIP) in place of the absent faiss.
They pin the block loop, id mapping, merge tie rule, output shapes/dtypes, the
offset2pid/dedup mapping and the TREC line format.  The faiss arithmetic itself
stays unpinned (no faiss in this image); integer-valued cases make the expected
scores exact in any summation order.
"""
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness  # noqa: E402
from oracle.flat_ip import FlatIP  # noqa: E402


def write_blocks(d, blocks, ids):
    for i, (b, e) in enumerate(zip(blocks, ids)):
        with open(os.path.join(d, "passage_emb_block_%d.pb" % i), "wb") as h:
            pickle.dump(np.ascontiguousarray(b, np.float32), h, protocol=4)
        with open(os.path.join(d, "passage_embid_block_%d.pb" % i), "wb") as h:
            pickle.dump(np.ascontiguousarray(e, np.int64), h, protocol=4)


def run_reference(mod, blocks, ids, q, k, block_num=None):
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, blocks, ids)
        args = types.SimpleNamespace(passage_block_num=block_num or (len(blocks) + 3))
        D, I = mod.search_one_by_one_with_faiss(args, d, FlatIP(q.shape[1]), q, k)
    return np.asarray(D), np.asarray(I)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-34s %8.1f KiB" % (name, os.path.getsize(path) / 1024))


def contiguous_ids(blocks, start=0):
    out, o = [], start
    for b in blocks:
        out.append(np.arange(o, o + len(b), dtype=np.int64))
        o += len(b)
    return out


def make_prj():
    """H. the PRJ drivers' copy of the loop and their score-less run file
    (`/root/reference/src/test_PRJ_topiocqa.py:83-171`, `:218-299`; qrecc twin identical)."""
    mod = ref_harness.load_reference_module("test_PRJ_topiocqa")
    mod_q = ref_harness.load_reference_module("test_PRJ_qrecc")
    rng = np.random.default_rng(777)
    blocks = [rng.standard_normal((n, 64)).astype(np.float32) for n in (140, 160, 90)]
    q = rng.standard_normal((8, 64)).astype(np.float32)
    k = 12
    D, I = run_reference(mod, blocks, contiguous_ids(blocks), q, k)
    D2, I2 = run_reference(mod_q, blocks, contiguous_ids(blocks), q, k)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    offset2pid = [int(v) for v in rng.integers(0, 80, size=390)]
    qids = ["%d-%d-%d" % (i // 4 + 1, i % 4 + 2, i % 4) for i in range(len(q))]       # conv-turn-historyturn
    with tempfile.TemporaryDirectory() as d:
        test_file = os.path.join(d, "test.json")
        with open(test_file, "w") as f:
            for s in qids:
                f.write(json.dumps({"id": s, "query": "q", "conv_id": 1, "turn_id": 1, "query_pair": "p"}) + "\n")
        mod.print_trec_res = lambda *a, **kw: {}
        args = types.SimpleNamespace(top_n=k, test_file_path=test_file, qrel_output_path=d,
                                     trec_gold_qrel_file_path="", rel_threshold=1)
        mod.output_test_res(qids, D, I, offset2pid, args)
        with open(os.path.join(d, "dev_dense_rel_res.trec")) as f:
            run_text = f.read()
    save("prj_run_d64", x0=blocks[0], x1=blocks[1], x2=blocks[2], q=q, k=np.int64(k), D=D, I=I,
         offset2pid=np.asarray(offset2pid, np.int64), qids=np.asarray(qids), run_text=np.asarray(run_text))


def make_prj_judge():
    """I. ``improve_judge`` of both PRJ drivers (`/root/reference/src/test_PRJ_topiocqa.py:443-472`,
    `/root/reference/src/test_PRJ_qrecc.py:403-452`): pure Python over the sample ids and a per-sample score list."""
    mod = ref_harness.load_reference_module("test_PRJ_topiocqa")
    mod_q = ref_harness.load_reference_module("test_PRJ_qrecc")
    rng = np.random.default_rng(4711)
    ids = []
    for conv in (1, 2, 5):
        for turn in range(1, int(rng.integers(3, 6))):
            ids.append("%d-%d-0" % (conv, turn))                       # base query of the turn
            for hist in range(1, turn):
                ids.append("%d-%d-%d" % (conv, turn, hist))            # one candidate history turn each
    # two consecutive conversations ending / starting on the same turn id (the case the qrecc variant adds)
    ids += ["7-2-0", "7-2-1", "8-2-0", "8-2-1"]
    scores = [float(v) for v in rng.choice([0.0, 0.05, 0.1, 0.2, 0.25, 1.0 / 3, 0.5, 1.0], size=len(ids))]
    qrel_ids = ["1-1", "1-2", "5-1", "8-2"]
    with tempfile.TemporaryDirectory() as d:
        qf, rf = os.path.join(d, "q.json"), os.path.join(d, "qrel.json")
        with open(qf, "w") as f:
            for s in ids:
                f.write(json.dumps({"id": s}) + "\n")
        with open(rf, "w") as f:
            for s in qrel_ids:
                f.write(json.dumps({"sample_id": s}) + "\n")
        out_t = mod.improve_judge(qf, scores)
        out_q = mod_q.improve_judge(qf, scores, rf)
    path = os.path.join(HERE, "prj_improve_judge.json")
    with open(path, "w") as f:
        json.dump({"ids": ids, "scores": scores, "qrel_ids": qrel_ids,
                   "topiocqa": list(out_t.items()), "qrecc": list(out_q.items())}, f)
    print("%-34s %8.1f KiB" % ("prj_improve_judge.json", os.path.getsize(path) / 1024))


def main():
    if "--only-prj-judge" in sys.argv:
        return make_prj_judge()
    if "--only-prj" in sys.argv:
        return make_prj()
    mod = ref_harness.load_reference_module("test_HAConvDR_topiocqa")
    mod_q = ref_harness.load_reference_module("test_HAConvDR_qrecc")
    rng = np.random.default_rng(20241018)

    # A. integer-valued known-answer set, d=768, single block: exact scores, many ties.
    x = rng.integers(-3, 4, size=(1024, 768)).astype(np.int8)
    q = rng.integers(-3, 4, size=(16, 768)).astype(np.int8)
    blocks = [x.astype(np.float32)]
    D, I = run_reference(mod, blocks, contiguous_ids(blocks), q.astype(np.float32), 100)
    save("kat_int_d768_1block", x0=x, q=q, k=np.int64(100), D=D, I=I)

    # B. random-normal, 3 ragged blocks, d=64, ids offset by 1000 (global stream offsets).
    blocks = [rng.standard_normal((n, 64)).astype(np.float32) for n in (700, 650, 300)]
    q = rng.standard_normal((32, 64)).astype(np.float32)
    ids = contiguous_ids(blocks, start=1000)
    D, I = run_reference(mod, blocks, ids, q, 50)
    D2, I2 = run_reference(mod_q, blocks, ids, q, 50)   # qrecc copy of the same loop
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    save("merge_3blocks_d64", x0=blocks[0], x1=blocks[1], x2=blocks[2], q=q, id_start=np.int64(1000),
         k=np.int64(50), D=D, I=I)

    # C. duplicated rows inside and across blocks -> exact ties decided by the merge rule.
    base = rng.integers(-4, 5, size=(40, 64)).astype(np.float32)
    b0 = np.concatenate([base, base[:20]], 0)
    b1 = np.concatenate([base[10:30], rng.integers(-4, 5, size=(50, 64)).astype(np.float32), base[:5]], 0)
    b2 = base[::-1].copy()
    blocks = [b0, b1, b2]
    q = rng.integers(-4, 5, size=(12, 64)).astype(np.float32)
    D, I = run_reference(mod, blocks, contiguous_ids(blocks), q, 25)
    save("merge_ties_across_blocks_d64", x0=b0, x1=b1, x2=b2, q=q, k=np.int64(25), D=D, I=I)

    # D. blocks shorter than k: -FLT_MAX / -1 fill, and the reference's emb2id[-1] wrap (:110).
    blocks = [rng.standard_normal((30, 64)).astype(np.float32)]
    q = rng.standard_normal((5, 64)).astype(np.float32)
    D, I = run_reference(mod, blocks, contiguous_ids(blocks, 7), q, 50)
    save("short_single_block_d64", x0=blocks[0], q=q, id_start=np.int64(7), k=np.int64(50), D=D, I=I)
    blocks = [rng.standard_normal((30, 64)).astype(np.float32), rng.standard_normal((45, 64)).astype(np.float32)]
    D, I = run_reference(mod, blocks, contiguous_ids(blocks, 7), q, 50)
    save("short_two_blocks_d64", x0=blocks[0], x1=blocks[1], q=q, id_start=np.int64(7), k=np.int64(50), D=D, I=I)

    # E. passage_block_num smaller than the number of files: only the first blocks are read (:77).
    blocks = [rng.standard_normal((n, 64)).astype(np.float32) for n in (120, 90, 200)]
    q = rng.standard_normal((9, 64)).astype(np.float32)
    D, I = run_reference(mod, blocks, contiguous_ids(blocks), q, 20, block_num=2)
    save("block_num_limit_d64", x0=blocks[0], x1=blocks[1], x2=blocks[2], q=q, k=np.int64(20),
         block_num=np.int64(2), D=D, I=I)

    # F. single query (nq=1) and k=1.
    blocks = [rng.standard_normal((n, 128)).astype(np.float32) for n in (257, 255)]
    q = rng.standard_normal((1, 128)).astype(np.float32)
    D, I = run_reference(mod, blocks, contiguous_ids(blocks), q, 1)
    save("single_query_k1_d128", x0=blocks[0], x1=blocks[1], q=q, k=np.int64(1), D=D, I=I)

    # G. offset2pid + duplicate-pid suppression + TREC run lines (output_test_res :220-286).
    blocks = [rng.standard_normal((n, 64)).astype(np.float32) for n in (150, 170)]
    q = rng.standard_normal((6, 64)).astype(np.float32)
    k = 10
    D, I = run_reference(mod, blocks, contiguous_ids(blocks), q, k)
    offset2pid = [int(v) for v in rng.integers(0, 60, size=320)]     # many duplicate pids
    qids = ["%d-%d" % (i // 3 + 1, i % 3 + 1) for i in range(len(q))]
    with tempfile.TemporaryDirectory() as d:
        test_file = os.path.join(d, "test.json")
        with open(test_file, "w") as f:
            for s in qids:
                f.write(json.dumps({"sample_id": s}) + "\n")
        mod.print_trec_res = lambda *a, **kw: {}
        args = types.SimpleNamespace(top_k=k, test_file_path=test_file, qrel_output_path=d,
                                     output_trec_file="run.trec", trec_gold_qrel_file_path="", rel_threshold=1)
        mod.output_test_res(qids, D, I, offset2pid, args)
        with open(os.path.join(d, "run.trec")) as f:
            run_text = f.read()
    save("trec_run_dedup_d64", x0=blocks[0], x1=blocks[1], q=q, k=np.int64(k), D=D, I=I,
         offset2pid=np.asarray(offset2pid, np.int64), qids=np.asarray(qids), run_text=np.asarray(run_text))
    make_prj()
    make_prj_judge()


if __name__ == "__main__":
    main()
