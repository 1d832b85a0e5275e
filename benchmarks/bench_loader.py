#!/usr/bin/env python
"""Block-loop timing: the reference's per-block add/search/reset loop over real block pickles
(`/root/reference/src/test_HAConvDR_topiocqa.py:74-162`) through the engine, phase by phase, and
the HBM-resident fast path.  Block files are written to a scratch directory first (page cache /
tmpfs on the GPU box, so the "disk" figure is a host-memory read rate).

Usage: python benchmarks/bench_loader.py [--blocks 4] [--rows-per-block 500000] [--queries 2514]"""
import argparse, json, os, pickle, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--rows-per-block", type=int, default=500_000)
    ap.add_argument("--queries", type=int, default=2514)
    ap.add_argument("--k", type=int, default=100)
    args = ap.parse_args()
    import numpy as np
    import torch
    from haconvdr_b200 import FlatIPIndex, loader, retrieval
    from haconvdr_b200 import faiss_compat as faiss
    from haconvdr_b200.index import synth_rows_device
    d = 768
    tmp = tempfile.mkdtemp(prefix="hac_blocks_")
    t0 = time.perf_counter()
    for b in range(args.blocks):
        x = synth_rows_device(args.rows_per_block, d, seed=42, row0=b * args.rows_per_block).cpu().numpy()
        with open(os.path.join(tmp, "passage_emb_block_%d.pb" % b), "wb") as h:
            pickle.dump(x, h, protocol=4)
        with open(os.path.join(tmp, "passage_embid_block_%d.pb" % b), "wb") as h:
            pickle.dump(np.arange(b * args.rows_per_block, (b + 1) * args.rows_per_block, dtype=np.int64), h, protocol=4)
    write_s = time.perf_counter() - t0
    q = synth_rows_device(args.queries, d, seed=4242).cpu().numpy()
    block_gb = args.rows_per_block * d * 4 / 1e9

    # (1) pickle.load + add, the reference's way of getting a block into the index
    idx = FlatIPIndex(d, 0)
    t0 = time.perf_counter()
    with open(os.path.join(tmp, "passage_emb_block_0.pb"), "rb") as h:
        arr = pickle.load(h)
    t_unpickle = time.perf_counter() - t0
    t0 = time.perf_counter()
    idx.add(arr)
    t_add_pageable = time.perf_counter() - t0
    idx.reset()
    del arr
    # (2) header walk + readinto pinned staging + add
    t0 = time.perf_counter()
    loader.stream_block_into(idx, os.path.join(tmp, "passage_emb_block_0.pb"))
    t_stream = time.perf_counter() - t0
    idx.reset()
    t0 = time.perf_counter()
    loader.stream_block_into(idx, os.path.join(tmp, "passage_emb_block_0.pb"))
    t_stream2 = time.perf_counter() - t0
    idx.close()
    print(json.dumps({"phase": "block -> HBM", "block_gb": block_gb, "pickle_load_s": t_unpickle,
                      "add_from_pageable_s": t_add_pageable,
                      "reference_way_gbs": block_gb / (t_unpickle + t_add_pageable),
                      "stream_pinned_s": t_stream, "stream_pinned_warm_s": t_stream2,
                      "stream_pinned_gbs": block_gb / t_stream2, "write_blocks_s": write_s}), flush=True)

    # (3) the reference loop (mirror) end to end
    index = faiss.index_cpu_to_gpu_multiple([None], [0], faiss.IndexFlatIP(d), faiss.GpuMultipleClonerOptions())
    for rep in range(2):
        t0 = time.perf_counter()
        D, I = retrieval.search_one_by_one_with_faiss(args.blocks + 2, tmp, index, q, args.k)
        t_loop = time.perf_counter() - t0
    print(json.dumps({"phase": "reference loop (add/search/reset per block + merge)", "blocks": args.blocks,
                      "rows": args.blocks * args.rows_per_block, "queries": args.queries, "seconds": t_loop,
                      "seconds_per_block": t_loop / args.blocks, "out_shape": list(D.shape)}), flush=True)

    # (4) resident fast path
    idx = FlatIPIndex(d, 0, reserve=args.blocks * args.rows_per_block)
    t0 = time.perf_counter()
    retrieval.load_resident(idx, tmp, args.blocks + 2)
    t_load = time.perf_counter() - t0
    for rep in range(3):
        t0 = time.perf_counter()
        Dr, Ir = retrieval.search_resident(idx, q, args.k)
        t_search = time.perf_counter() - t0
    assert np.array_equal(Ir, I[:, :args.k]) and np.array_equal(Dr, D[:, :args.k]), "resident path != block loop"
    print(json.dumps({"phase": "resident path", "load_s": t_load, "load_gbs": args.blocks * block_gb / t_load,
                      "search_s": t_search, "queries_per_s": args.queries / t_search,
                      "identical_to_block_loop_first_k_columns": True}), flush=True)
    idx.close()

    # (5) engine-native block files (4 KiB-aligned raw payload): buffered and O_DIRECT reads
    t0 = time.perf_counter()
    nat = loader.convert_block_to_native(tmp, 0)
    t_conv = time.perf_counter() - t0
    idx = FlatIPIndex(d, 0, reserve=args.rows_per_block)
    res = {"phase": "native block -> HBM", "block_gb": block_gb, "convert_s": t_conv}
    for name, direct in (("buffered", False), ("o_direct", True)):
        for rep in range(2):
            idx.reset()
            st = {}
            t0 = time.perf_counter()
            loader.stream_block_into(idx, nat, direct=direct, stats=st)
            dt = time.perf_counter() - t0
        res[name + "_s"] = dt
        res[name + "_gbs"] = block_gb / dt
        res[name + "_used_o_direct"] = bool(st.get("direct"))
    print(json.dumps(res), flush=True)
    idx.close()
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
