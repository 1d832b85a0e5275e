#!/usr/bin/env python
"""A/B of scan-kernel variants inside ONE process (same GPU, same thermal state), alternating.
Usage: python benchmarks/ab_scan.py [--rows N] [--rounds R]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--queries", type=int, default=2514)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--per", type=int, default=4)
    ap.add_argument("--variants", default="cg1,cg2")
    args = ap.parse_args()
    import torch
    from haconvdr_b200 import FlatIPIndex
    from haconvdr_b200.index import synth_rows_device
    idx = FlatIPIndex(768, 0, reserve=args.rows)
    idx.add_synthetic(args.rows, seed=42)
    q = synth_rows_device(args.queries, 768, seed=4242)
    variants = {"cg1": [("mma_cta_group", 1)], "cg2": [("mma_cta_group", 2)],
                "cg1_g2": [("mma_cta_group", 1), ("chunk_growth_x100", 200)],
                "cg1_g8": [("mma_cta_group", 1), ("chunk_growth_x100", 800)],
                "cg2_g8": [("mma_cta_group", 2), ("chunk_growth_x100", 800)]}
    res = {v: [] for v in args.variants.split(",")}
    ref = None
    for r in range(args.rounds + 1):
        for v in res:
            idx.set_option("chunk_growth_x100", 400)
            for name, val in variants[v]:
                idx.set_option(name, val)
            for _ in range(args.per):
                D, I = idx.search(q, 100)
                st = idx.stats()
                if r > 0:
                    res[v].append((st["scan_ms"], st["total_ms"]))
            if ref is None:
                ref = I.clone()
            assert torch.equal(I, ref), "variant %s changed the result" % v
    fl = 2.0 * args.queries * args.rows * 768
    for v, xs in res.items():
        scan = sorted(x[0] for x in xs); tot = sorted(x[1] for x in xs)
        print(json.dumps({"variant": v, "scan_ms_median": scan[len(scan) // 2], "scan_ms_min": scan[0],
                          "total_ms_median": tot[len(tot) // 2], "tflops_median": fl / scan[len(scan) // 2] / 1e9,
                          "n": len(xs)}), flush=True)


if __name__ == "__main__":
    main()
