#!/usr/bin/env python
"""Two single-GPU experiments in one process (full TopiOCQA-scale corpus by default):

  small : turn latency Q = 1 / 4 / 32 / 128 with the f16 screen and the int8 screen (needs build_i8), alternating
  drop  : headline search (Q = 2514, k = 100) for several (corpus, query) f16 mantissa-drop settings

One JSON line per measurement.  Usage: python benchmarks/ab_round1b.py [--only small,drop] [--rows N]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--only", default="small,drop")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--drops", default="3:0,3:2,3:3,4:0,4:2,5:0")
    args = ap.parse_args()
    import torch
    from haconvdr_b200 import FlatIPIndex, HAC_PATH_MMA, HAC_PATH_I8
    from haconvdr_b200.index import synth_rows_device
    only = set(args.only.split(","))
    d = 768
    if "small" in only:
        idx = FlatIPIndex(d, 0)
        idx.set_option("build_i8", 1)
        idx.reserve(args.rows)
        idx.add_synthetic(args.rows, seed=42)
        for nq in (1, 4, 32, 128):
            q = synth_rows_device(nq, d, seed=4242)
            ref = None
            for name, path in (("f16", HAC_PATH_MMA), ("i8", HAC_PATH_I8)):
                sts = []
                for r in range(args.reps + 2):
                    D, I = idx.search(q, 100, path=path)
                    if r >= 2:
                        sts.append(idx.stats())
                if ref is None:
                    ref = (D.clone(), I.clone())
                same = bool(torch.equal(I, ref[1]) and torch.equal(D, ref[0]))
                st = sts[-1]
                print(json.dumps({"exp": "small", "Q": nq, "path": name, "total_ms": med(s["total_ms"] for s in sts),
                                  "scan_ms": med(s["scan_ms"] for s in sts), "best_ms": min(s["total_ms"] for s in sts),
                                  "chunks": st["n_chunks"], "launches": st["kernel_launches"],
                                  "emitted": st["candidates_emitted"], "rescored": st["candidates_rescored"],
                                  "retries": st["retries"], "used_path": st["path"], "same_as_f16": same}), flush=True)
        idx.close()
        del idx
        torch.cuda.empty_cache()
    if "drop" in only:
        q = synth_rows_device(2514, d, seed=4242)
        ref = None
        for spec in args.drops.split(","):
            bx, bq = (int(v) for v in spec.split(":"))
            idx = FlatIPIndex(d, 0)
            idx.set_option("f16_drop_bits_corpus", bx)
            idx.set_option("f16_drop_bits_queries", bq)
            idx.reserve(args.rows)
            idx.add_synthetic(args.rows, seed=42)
            sts = []
            for r in range(args.reps + 2):
                D, I = idx.search(q, 100)
                if r >= 2:
                    sts.append(idx.stats())
            if ref is None:
                ref = (D.clone(), I.clone())
            same = bool(torch.equal(I, ref[1]) and torch.equal(D, ref[0]))
            st = sts[-1]
            print(json.dumps({"exp": "drop", "bits_corpus": bx, "bits_queries": bq,
                              "total_ms": med(s["total_ms"] for s in sts), "scan_ms": med(s["scan_ms"] for s in sts),
                              "best_ms": min(s["total_ms"] for s in sts), "emitted": st["candidates_emitted"],
                              "rescored": st["candidates_rescored"], "margin_max": st["margin_max"],
                              "retries": st["retries"], "same_as_first": same}), flush=True)
            idx.close()
            del idx
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
