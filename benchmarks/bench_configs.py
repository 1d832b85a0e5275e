#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (configs[2..4]) on ONE GPU; one JSON line each.

  small-batch turn latency (configs[3]): Q = 1 / 4 / 32 over 25.7M x 768, HBM roofline
  k sweep (configs[4]):                  k = 1 / 10 / 100 / 1000 at Q = 2514, tensor roofline
  QReCC per-GPU shard (configs[2]):      6 821 633 rows (1/8 of 54.57M) x 8209 queries, k = 100

Usage: python benchmarks/bench_configs.py [--rows N] [--reps R] [--only small,ksweep,qrecc]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def timed(fn, reps, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="small,ksweep,qrecc")
    args = ap.parse_args()
    import torch
    from haconvdr_b200 import FlatIPIndex, HAC_PATH_GEMV, HAC_PATH_I8, HAC_PATH_MMA
    from haconvdr_b200.index import synth_rows_device
    pk, kind = peaks()
    only = set(args.only.split(","))
    d = 768
    names = {HAC_PATH_GEMV: "gemv_fp32", HAC_PATH_MMA: "mma_f16_screen", HAC_PATH_I8: "mma_i8_screen"}
    bytes_per_elem = {HAC_PATH_GEMV: 4, HAC_PATH_MMA: 2, HAC_PATH_I8: 1}

    if only & {"small", "ksweep"}:
        idx = FlatIPIndex(d, 0)               # the int8 image (rows*768 B of HBM) is built by default
        idx.reserve(args.rows)
        idx.add_synthetic(args.rows, seed=42)
        n = idx.ntotal
        if "small" in only:
            for nq, paths in ((1, (HAC_PATH_GEMV, HAC_PATH_MMA, HAC_PATH_I8)), (4, (HAC_PATH_GEMV, HAC_PATH_MMA, HAC_PATH_I8)),
                              (32, (HAC_PATH_MMA, HAC_PATH_I8))):
                q = synth_rows_device(nq, d, seed=4242)
                for path in paths:
                    med, best = timed(lambda: idx.search(q, 100, path=path), args.reps)
                    st = idx.stats()
                    # bytes the scan has to stream once: fp32 rows (GEMV, SURVEY 8d: N*768*4), the f16 or the int8 image
                    moved = n * d * bytes_per_elem[path]
                    algo_bytes = moved
                    print(json.dumps({
                        "config": "turn latency Q=%d over %dx768, k=100" % (nq, n), "path": names[path],
                        "ms_per_batch_median": med, "ms_per_batch_best": best, "scan_ms": st["scan_ms"],
                        "queries_per_s": nq / (med * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": algo_bytes / (st["scan_ms"] * 1e-3) / 1e9,
                                     "peak": pk["hbm_gbs"], "unit": "GB/s",
                                     "frac": algo_bytes / (st["scan_ms"] * 1e-3) / 1e9 / pk["hbm_gbs"],
                                     "bytes_per_batch": moved, "fp32_corpus_bytes": n * d * 4,
                                     "survey_8d_gbs": n * d * 4 / (med * 1e-3) / 1e9,   # N*768*4 B per batch / whole latency
                                     "peak_source": kind + " hbm_gbs (copy = half reads, half writes; a pure read "
                                                           "stream can exceed it)"},
                        "n_chunks": st["n_chunks"], "launches": st["kernel_launches"],
                        "candidates_rescored": st["candidates_rescored"]}), flush=True)
        if "ksweep" in only:
            q = synth_rows_device(2514, d, seed=4242)
            for k in (1, 10, 100, 1000):
                med, best = timed(lambda: idx.search(q, k), max(3, args.reps // 2))
                st = idx.stats()
                fl = 2.0 * 2514 * n * d
                peak = pk["bf16_tflops_sustained"] * (2.0 if st["path"] == HAC_PATH_I8 else 1.0)   # kind::i8 = 2 x bf16 rate
                print(json.dumps({
                    "config": "k sweep k=%d, Q=2514 over %dx768" % (k, n), "path": names[st["path"]],
                    "ms_per_search_median": med,
                    "queries_per_s": 2514 / (med * 1e-3), "scan_ms": st["scan_ms"],
                    "roofline": {"bound": "tensor", "achieved": fl / (st["scan_ms"] * 1e-3) / 1e12,
                                 "peak": peak, "unit": "TFLOP/s",
                                 "frac": fl / (st["scan_ms"] * 1e-3) / 1e12 / peak},
                    "candidates_emitted": st["candidates_emitted"], "candidates_rescored": st["candidates_rescored"],
                    "n_chunks": st["n_chunks"], "retries": st["retries"]}), flush=True)
        idx.close()
        del idx
        torch.cuda.empty_cache()

    if "qrecc" in only:
        rows = 54_573_064 // 8
        idx = FlatIPIndex(d, 0, reserve=rows)
        idx.add_synthetic(rows, seed=42)
        q = synth_rows_device(8209, d, seed=4242)
        med, best = timed(lambda: idx.search(q, 100), max(3, args.reps // 2))
        st = idx.stats()
        fl = 2.0 * 8209 * rows * d
        peak = pk["bf16_tflops_sustained"] * (2.0 if st["path"] == HAC_PATH_I8 else 1.0)
        print(json.dumps({
            "config": "QReCC per-GPU shard: %d rows (1/8 of 54 573 064) x 8209 queries, k=100" % rows,
            "path": names[st["path"]],
            "ms_per_search_median": med, "queries_per_s_this_shard": 8209 / (med * 1e-3), "scan_ms": st["scan_ms"],
            "roofline": {"bound": "tensor", "achieved": fl / (st["scan_ms"] * 1e-3) / 1e12,
                         "peak": peak, "unit": "TFLOP/s",
                         "frac": fl / (st["scan_ms"] * 1e-3) / 1e12 / peak},
            "n_chunks": st["n_chunks"], "retries": st["retries"]}), flush=True)


if __name__ == "__main__":
    main()
