#!/usr/bin/env python
"""A/B of the scan paths (f16 screen vs int8 screen) inside ONE process, alternating."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--queries", type=int, default=2514)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    import torch
    from haconvdr_b200 import FlatIPIndex, HAC_PATH_MMA, HAC_PATH_I8
    from haconvdr_b200.index import synth_rows_device
    idx = FlatIPIndex(768, 0)
    idx.set_option("build_i8", 1)
    idx.reserve(args.rows)
    idx.add_synthetic(args.rows, seed=42)
    q = synth_rows_device(args.queries, 768, seed=4242)
    # (name, path, options)
    variants = (("f16", HAC_PATH_MMA, {}), ("i8", HAC_PATH_I8, {}),
                ("i8_growth_1.0", HAC_PATH_I8, {"i8_chunk_growth_x100": 100}),
                ("i8_growth_0.6", HAC_PATH_I8, {"i8_chunk_growth_x100": 60}),
                ("i8_growth_0.35", HAC_PATH_I8, {"i8_chunk_growth_x100": 35}))
    defaults = {"scan_tile_major": -1, "i8_cta_group": 2, "i8_chunk_growth_x100": 0}
    res = {v[0]: [] for v in variants}
    ref = None
    for r in range(args.rounds + 1):
        for name, path, opts in variants:
            for key, val in {**defaults, **opts}.items():
                idx.set_option(key, val)
            for _ in range(3):
                D, I = idx.search(q, args.k, path=path)
                st = idx.stats()
                if r > 0:
                    res[name].append(st)
            if ref is None:
                ref = (D.clone(), I.clone())
            assert torch.equal(I, ref[1]) and torch.equal(D, ref[0]), "paths disagree"
    for name, sts in res.items():
        med = lambda key: sorted(s[key] for s in sts)[len(sts) // 2]
        print(json.dumps({"path": name, "total_ms": med("total_ms"), "scan_ms": med("scan_ms"),
                          "emitted": sts[-1]["candidates_emitted"], "rescored": sts[-1]["candidates_rescored"],
                          "chunks": sts[-1]["n_chunks"], "launches": sts[-1]["kernel_launches"],
                          "margin_max": sts[-1]["margin_max"], "screen_err_max": sts[-1]["screen_err_max"],
                          "retries": sts[-1]["retries"], "qps": args.queries / med("total_ms") * 1e3}), flush=True)


if __name__ == "__main__":
    main()
