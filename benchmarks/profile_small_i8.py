#!/usr/bin/env python
"""Two Q=1 searches through the int8 screen over the full TopiOCQA-scale corpus, for ncu:
each search launches 7 scan kernels (`scan_mma_kernel<1, 1>`), the last one covers the final ~12.9M rows, so
`ncu -k regex:scan_mma -s 13 -c 1 --set full ...` captures the largest chunk of the second search."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--queries", type=int, default=1)
    ap.add_argument("--cta-group", type=int, default=1)
    ap.add_argument("--k", type=int, default=100)
    args = ap.parse_args()
    from haconvdr_b200 import FlatIPIndex, HAC_PATH_I8
    from haconvdr_b200.index import synth_rows_device
    idx = FlatIPIndex(768, 0)
    idx.set_option("build_i8", 1)
    idx.set_option("i8_cta_group", args.cta_group)
    idx.reserve(args.rows)
    idx.add_synthetic(args.rows, seed=42)
    q = synth_rows_device(args.queries, 768, seed=4242)
    for _ in range(2):
        D, I = idx.search(q, args.k, path=HAC_PATH_I8)
    st = idx.stats()
    print(json.dumps({"rows": args.rows, "queries": args.queries, "total_ms": st["total_ms"], "scan_ms": st["scan_ms"],
                      "chunks": st["n_chunks"], "launches": st["kernel_launches"], "path": st["path"]}))


if __name__ == "__main__":
    main()
