#!/usr/bin/env python
"""Large k (BASELINE configs[4], k = 1000): the f16 screen (what HAC_PATH_AUTO ran for k > 128 until now) against the
int8 screen with a forced warm start and several chunk growths, alternating in one process on the full corpus.
All variants must return bitwise identical results.  One JSON line per (k, variant).

    python benchmarks/ab_k1000.py [--rows N] [--queries Q] [--ks 1000,500,250] [--reps R]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


VARIANTS = [
    # name, path (0 AUTO with i8_auto_max_k = 128 -> f16 for these k; 3 = int8 forced), options
    ("f16", 2, {}),
    ("i8_w0_g100", 3, {"i8_warm_rows": 0, "i8_chunk_growth_x100": 100}),
    ("i8_w768k_g100", 3, {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 100}),
    ("i8_w768k_g50", 3, {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 50}),
    ("i8_w768k_g35", 3, {"i8_warm_rows": 786432, "i8_chunk_growth_x100": 35}),
    ("i8_w1536k_g50", 3, {"i8_warm_rows": 1572864, "i8_chunk_growth_x100": 50}),
    ("i8_w3072k_g50", 3, {"i8_warm_rows": 3145728, "i8_chunk_growth_x100": 50}),
]
DEFAULTS = {"i8_chunk_growth_x100": 0, "i8_warm_rows": -1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--queries", type=int, default=2514)
    ap.add_argument("--ks", default="1000,250")
    ap.add_argument("--reps", type=int, default=3)
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
    q = synth_rows_device(args.queries, d, seed=4242)
    for k in [int(v) for v in args.ks.split(",")]:
        ref = None
        acc = {name: [] for name, _, _ in variants}
        failed = {}
        for r in range(args.reps + 1):
            for name, path, opts in variants:
                if name in failed:
                    continue
                for kk, vv in {**DEFAULTS, **opts}.items():
                    idx.set_option(kk, vv)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                try:
                    D, I = idx.search(q, k, path=path)
                except Exception as e:   # noqa: BLE001
                    failed[name] = repr(e)[:200]
                    continue
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) * 1e3
                if ref is None:
                    ref = (D.clone(), I.clone())
                same = bool(torch.equal(I, ref[1]) and torch.equal(D, ref[0]))
                st = idx.stats()
                st["wall_ms"] = dt
                st["same"] = same
                if r >= 1:
                    acc[name].append(st)
        for name, path, _ in variants:
            if name in failed:
                print(json.dumps({"exp": "k_large", "k": k, "variant": name, "failed": failed[name]}), flush=True)
                continue
            sts = acc[name]
            st = sts[-1]
            print(json.dumps({"exp": "k_large", "k": k, "Q": args.queries, "variant": name,
                              "total_ms": med(s["total_ms"] for s in sts), "best_ms": min(s["total_ms"] for s in sts),
                              "wall_ms": med(s["wall_ms"] for s in sts), "scan_ms": med(s["scan_ms"] for s in sts),
                              "chunks": st["n_chunks"], "launches": st["kernel_launches"],
                              "emitted": med(s["candidates_emitted"] for s in sts),
                              "rescored": med(s["candidates_rescored"] for s in sts), "retries": st["retries"],
                              "path": st["path"], "warm_rows": st["warm_rows"], "margin_max": st["margin_max"],
                              "identical_results": all(s["same"] for s in sts),
                              "hbm_f16_gb": st["bytes_shadow"] / 1e9}), flush=True)
    for kk, vv in DEFAULTS.items():
        idx.set_option(kk, vv)
    idx.close()


if __name__ == "__main__":
    main()
