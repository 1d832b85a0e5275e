#!/usr/bin/env python
"""The drop-in multi-GPU path inside ONE process (what `build_faiss_index` with n_gpu > 1 becomes,
/root/reference/src/test_HAConvDR_topiocqa.py:42-66 -> faiss_compat.ShardedInProcessIndex), timed on the headline
workload and compared with the torchrun path of bench.py: synthetic 25.7M x 768 corpus split over all visible GPUs,
2514 host queries in, host results out.  One JSON line.

    python benchmarks/bench_inprocess.py [--gpus N] [--rows R] [--queries Q] [--k K] [--steps S]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--rows", type=int, default=25_700_592)
    ap.add_argument("--queries", type=int, default=2514)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-exchange", action="store_true")
    ap.add_argument("--pinned", action="store_true", help="page-locked query / result arrays (what bench.py's e2e leg uses)")
    args = ap.parse_args()
    import numpy as np
    import torch
    from haconvdr_b200 import faiss_compat as faiss
    from haconvdr_b200.index import synth_rows_device
    G = args.gpus or torch.cuda.device_count()
    d = 768
    cpu_index = faiss.IndexFlatIP(d)
    co = faiss.GpuMultipleClonerOptions()
    co.shard = True
    index = faiss.index_cpu_to_gpu_multiple([faiss.StandardGpuResources() for _ in range(G)], list(range(G)), cpu_index, co)
    assert isinstance(index, faiss.ShardedInProcessIndex) or G == 1
    sharded = isinstance(index, faiss.ShardedInProcessIndex)
    t0 = time.perf_counter()
    if sharded:
        # the same row-indexed synthetic corpus as bench.py, generated on each device (79 GB of host rows are not needed)
        bounds = [(g * args.rows) // G for g in range(G + 1)]
        for g, sh in enumerate(index.shards):
            sh.reserve(bounds[g + 1] - bounds[g])
            sh.add_synthetic(bounds[g + 1] - bounds[g], seed=42, row0=bounds[g])
            index._rows[g].append((bounds[g], bounds[g + 1] - bounds[g]))
            index._set_ids(g)
        index.ntotal = args.rows
        index.threshold_exchange = not args.no_exchange
    else:
        impl = index._get()
        impl.reserve(args.rows)
        impl.add_synthetic(args.rows, seed=42)
    for g in range(G):
        torch.cuda.synchronize(g)
    setup_s = time.perf_counter() - t0
    q = synth_rows_device(args.queries, d, seed=4242, device=0).cpu().numpy()
    out = {}
    if args.pinned:
        qp = torch.empty((args.queries, d), dtype=torch.float32).pin_memory()
        qp.copy_(torch.from_numpy(q))
        q = qp.numpy()
        out = {"D": torch.empty((args.queries, args.k), dtype=torch.float32).pin_memory().numpy(),
               "I": torch.empty((args.queries, args.k), dtype=torch.int64).pin_memory().numpy()}
    for _ in range(args.warmup):
        D, I = index.search(q, args.k, **out)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        D, I = index.search(q, args.k, **out)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    # parity of 32 queries spread over the batch against an fp64 re-scoring of the WHOLE regenerated corpus (every device
    # re-scores its own rows in slabs): returned scores within 1e-5 relative of the fp64 ones, order non-increasing, and
    # the id lists identical up to rows whose fp64 scores tie with the k-th best within 1e-5 relative
    sel = np.unique(np.linspace(0, args.queries - 1, 32).astype(np.int64))
    kk = args.k + 16
    best_s, best_i = [], []
    bounds = [(g * args.rows) // G for g in range(G + 1)] if sharded else [0, args.rows]
    for g in range(len(bounds) - 1):
        dev = torch.device("cuda", g)
        q64 = torch.from_numpy(q[sel]).to(dev).double()
        bs = torch.full((len(sel), kk), -float("inf"), dtype=torch.float64, device=dev)
        bi = torch.full((len(sel), kk), -1, dtype=torch.int64, device=dev)
        for r0 in range(bounds[g], bounds[g + 1], 1_000_000):
            nr = min(1_000_000, bounds[g + 1] - r0)
            xs = synth_rows_device(nr, d, seed=42, row0=r0, device=g)
            top = torch.topk(q64 @ xs.double().T, min(kk, nr), dim=1)
            cs, ci = torch.cat([bs, top.values], 1), torch.cat([bi, top.indices + r0], 1)
            b = torch.topk(cs, kk, dim=1)
            bs, bi = b.values, torch.gather(ci, 1, b.indices)
            del xs
        best_s.append(bs.cpu().numpy())
        best_i.append(bi.cpu().numpy())
    ext_s, ext_i = np.concatenate(best_s, 1), np.concatenate(best_i, 1)
    order = np.lexsort((ext_i, -ext_s), axis=1)[:, :kk]
    ext_s, ext_i = np.take_along_axis(ext_s, order, 1), np.take_along_axis(ext_i, order, 1)
    ok, n_identical = True, 0
    for r, qi in enumerate(sel):
        ref = dict(zip(ext_i[r].tolist(), ext_s[r].tolist()))
        got_s = np.asarray([ref.get(int(i), np.nan) for i in I[qi]])          # fp64 scores of the returned ids
        kth = ext_s[r, args.k - 1]
        ok = ok and not np.isnan(got_s).any() and len(set(I[qi].tolist())) == args.k
        ok = ok and bool(np.all(np.abs(got_s - D[qi]) <= 1e-5 * np.abs(got_s) + 1e-30)) and bool(np.all(np.diff(D[qi]) <= 0))
        ok = ok and bool(np.all(got_s >= kth - 1e-5 * abs(kth)))               # nothing better than the k-th best was missed
        n_identical += int(np.array_equal(I[qi], ext_i[r, :args.k]))
    st = index.stats() if sharded else [index.stats()]
    print(json.dumps({"what": "in-process drop-in (faiss_compat.index_cpu_to_gpu_multiple, co.shard)", "n_gpus": G, "pinned_host_arrays": bool(args.pinned),
                      "rows": args.rows, "queries": args.queries, "k": args.k, "steps": args.steps,
                      "ms_per_search_host_to_host": ms, "best_ms": 1e3 * min(times),
                      "queries_per_s": args.queries / (ms * 1e-3), "spot_check_ok": bool(ok),
                      "parity": {"queries_checked": int(len(sel)), "rows_identical_order": int(n_identical), "ok": bool(ok),
                                 "how": "fp64 re-scoring of the whole regenerated corpus on every device, scores within 1e-5 "
                                        "relative, ties within 1e-5 relative of the k-th best tolerated"},
                      "exchange": bool(sharded and index.threshold_exchange),
                      "rescored_pairs_per_shard": [int(s["candidates_rescored"]) for s in st],
                      "local_total_ms_per_shard": [round(float(s["total_ms"]), 3) for s in st],
                      "setup_s": round(setup_s, 2),
                      "host_phases_ms_last_search": getattr(index, "last_phase_ms", None)}))
    assert ok


if __name__ == "__main__":
    main()
