#!/usr/bin/env python
"""Round-2 A/B on one GPU (full TopiOCQA-scale corpus by default): the pipelined int8 search (scans back to back,
rescore + refresh on a side stream beside the next scan) against the chunk-synchronous schedule, and the warm-start /
chunk-growth settings, for the headline batch and for the small turn batches, alternating in one process.

One JSON line per (batch, variant).  Usage: python benchmarks/ab_pipeline.py [--rows N] [--reps R] [--only big,small]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


VARIANTS = [
    # name, options
    ("sync_b0", {"i8_pipeline": 0}),
    ("pipe_b0", {"i8_pipeline": 1}),
    ("w768k_g35", {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 35}),
    ("w768k_g25", {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 25}),
    ("w768k_g18", {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 18}),
    ("w768k_g50", {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 50}),
    ("w1536k_g35", {"i8_warm_rows": 1572864, "i8_chunk_growth_x100": 35}),
    ("w1536k_g25", {"i8_warm_rows": 1572864, "i8_chunk_growth_x100": 25}),
    ("w384k_g25", {"i8_warm_rows": 393216, "i8_chunk_growth_x100": 25}),
    ("w3072k_g25", {"i8_warm_rows": 3145728, "i8_chunk_growth_x100": 25}),
    ("pipe_w768k", {"i8_pipeline": 1, "i8_warm_rows": 786432}),
    ("sync_w0", {"i8_warm_rows": 0}),
    ("sync_w128k", {"i8_warm_rows": 131072}),
    ("sync_w256k", {"i8_warm_rows": 262144}),
    ("sync_w384k", {"i8_warm_rows": 393216}),
    ("sync_w768k", {"i8_warm_rows": 786432}),
    ("sync_w384k_g35", {"i8_warm_rows": 393216, "i8_chunk_growth_x100": 35}),
    ("sync_w384k_g100", {"i8_warm_rows": 393216, "i8_chunk_growth_x100": 100}),
    ("pipe_w384k", {"i8_pipeline": 1, "i8_warm_rows": 393216}),
    # unit order of the int8 scan behind the default warm start: tile-major (every pair streams queries AND corpus) against
    # query-stationary (a pair keeps its query tiles in shared memory, lanes of pairs walk the corpus in step)
    ("def_tm", {"i8_warm_rows": -1, "scan_tile_major": 1}),
    ("def_qs", {"i8_warm_rows": -1, "scan_tile_major": 2}),
]
DEFAULTS = {"i8_pipeline": 0, "i8_pipe_growth_x1000": 125, "i8_pipe_min_rows": 0, "i8_pipe_dist": 2,
            "i8_chunk_growth_x100": 0, "i8_warm_rows": 0, "scan_tile_major": -1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="big,small")
    ap.add_argument("--variants", default="")
    args = ap.parse_args()
    import torch
    from haconvdr_b200 import FlatIPIndex
    from haconvdr_b200.index import synth_rows_device
    d = 768
    idx = FlatIPIndex(d, 0)
    idx.reserve(args.rows)
    idx.add_synthetic(args.rows, seed=42)
    variants = [v for v in VARIANTS if not args.variants or v[0] in args.variants.split(",")]
    batches = []
    if "big" in args.only:
        batches.append(2514)
    if "small" in args.only:
        batches += [1, 4, 32, 128]
    for nq in batches:
        q = synth_rows_device(nq, d, seed=4242)
        ref = None
        acc = {name: [] for name, _ in variants}
        wall = {name: [] for name, _ in variants}
        for r in range(args.reps + 1):
            for name, opts in variants:
                k = 100
                for kk, vv in {**DEFAULTS, **opts}.items():
                    if kk == "_k":
                        k = vv                                   # pseudo-option: top-k of this variant
                    else:
                        idx.set_option(kk, vv)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                D, I = idx.search(q, k)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) * 1e3
                if ref is None and k == 100:
                    ref = (D.clone(), I.clone())
                if ref is not None:
                    assert torch.equal(I, ref[1][:, :k]) and torch.equal(D, ref[0][:, :k]), name
                if r >= 1:
                    acc[name].append(idx.stats())
                    wall[name].append(dt)
        for name, _ in variants:
            sts = acc[name]
            st = sts[-1]
            print(json.dumps({"exp": "pipeline", "Q": nq, "variant": name, "total_ms": med(s["total_ms"] for s in sts),
                              "best_ms": min(s["total_ms"] for s in sts), "wall_ms": med(wall[name]),
                              "scan_ms": med(s["scan_ms"] for s in sts), "tail_ms": med(s["tail_ms"] for s in sts),
                              "chunks": st["n_chunks"], "sync_chunks": st["n_sync_chunks"],
                              "launches": st["kernel_launches"], "emitted": med(s["candidates_emitted"] for s in sts),
                              "rescored": med(s["candidates_rescored"] for s in sts), "retries": st["retries"],
                              "path": st["path"]}), flush=True)
    for kk, vv in DEFAULTS.items():
        idx.set_option(kk, vv)
    idx.close()


if __name__ == "__main__":
    main()
