#!/usr/bin/env python
"""Headline benchmark: exact top-100 inner-product search, TopiOCQA-scale synthetic corpus.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): 25 700 592 x 768 fp32 random-normal passages generated on the
device (counter-based, seed 42, row-indexed so any shard count sees the same corpus), 2514 queries
(seed 4242), k = 100.  One "step" = one search of all 2514 queries over the whole corpus.
With N ranks (torchrun, one per GPU) the corpus is sharded N ways (strong scaling): local exact
top-k -> one NCCL all-gather of Q x k candidates -> device k-way merge.

Prints ONE JSON line (rank 0): `value` = queries/s with corpus, queries and results resident in HBM,
`e2e` = the same through the host-buffer API (numpy in / numpy out, H2D + D2H inside the timed
region), `roofline` for the dominant kernel (tcgen05 screen scan) measured live with CUDA events,
`cpu_baseline` = the faiss-restatement CPU port timed on this box's host cores (bounded sample).
`--impl reference` times only that CPU port (the reference's own CPU path; faiss is not installable
here) on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS = 25_700_592
N_QUERIES = 2514
DIM = 768
TOP_K = 100
METRIC = "queries/sec, exact top-100, 25.7Mx768"
WORKLOAD = "topiocqa-scale synthetic 25.7Mx768 fp32, 2514 queries, top_k=100 (BASELINE.json configs[1])"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--queries", type=int, default=N_QUERIES)
    ap.add_argument("--k", type=int, default=TOP_K)
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--workload", default="topiocqa", choices=["topiocqa", "qrecc"],
                    help="qrecc = BASELINE.json configs[2]: 54 573 064 x 768, 8209 queries (needs >= 2 GPUs)")
    args = ap.parse_args()
    if args.workload == "qrecc":
        global METRIC, WORKLOAD
        METRIC = "queries/sec, exact top-100, 54.6Mx768"
        WORKLOAD = "qrecc-scale synthetic 54.6Mx768 fp32, 8209 queries, top_k=100 (BASELINE.json configs[2])"
        if args.rows == N_ROWS:
            args.rows = 54_573_064
        if args.queries == N_QUERIES:
            args.queries = 8209
    return args


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(n_sample_rows, n_queries, k, n_total_rows, steps=1, warmup=0, x=None, q=None):
    """Time the faiss-restatement CPU port (oracle/cpu_port.py: BLAS sgemm in faiss' 4096x1024 blocking
    + C heaps) on a bounded sample and scale linearly in corpus rows to the full workload."""
    import numpy as np
    from oracle import cpu_port
    cpu_port.build()
    rng = np.random.default_rng(42)
    if x is None:
        x = rng.standard_normal((n_sample_rows, DIM), dtype=np.float32)
    if q is None:
        q = rng.standard_normal((n_queries, DIM), dtype=np.float32)
    for _ in range(warmup):
        cpu_port.search_blas(q, x, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port.search_blas(q, x, k)
    dt = (time.perf_counter() - t0) / steps
    full_time = dt * (n_total_rows / x.shape[0])
    return q.shape[0] / full_time, dt, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1; the CPU reference uses every host core (before MKL / libgomp load)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    os.environ["MKL_NUM_THREADS"] = str(os.cpu_count())
    import torch
    torch.set_num_threads(os.cpu_count())
    n_sample = min(args.cpu_sample_rows, args.rows)
    t_start = time.perf_counter()
    qps, dt, cores = cpu_port_rate(n_sample, args.queries, args.k, args.rows, steps=args.steps, warmup=args.warmup)
    sample = "%d of %d corpus rows x %d queries per step (%.2f s/step), scaled linearly in rows" % (
        n_sample, args.rows, args.queries, dt)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * (args.rows / n_sample),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "queries": args.queries, "k": args.k, "dim": DIM,
                   "note": "faiss is not installable offline; this is the faiss-restatement CPU port "
                           "(IndexFlatIP: sgemm 4096x1024 blocks + per-query heaps)"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.samples, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 7 for n, v in zip(names, s[3:7]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from haconvdr_b200 import FlatIPIndex, HAC_PATH_I8, HAC_PATH_MMA
    from haconvdr_b200.index import synth_rows_device
    from haconvdr_b200.sharded import ShardedFlatIPIndex

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; stdout must carry only the
        # one JSON line, so fd 1 points at stderr until the first collective has run.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    t_setup = time.perf_counter()

    index = ShardedFlatIPIndex(DIM, FlatIPIndex(DIM, local_rank), exchange=args.exchange)
    index.add_synthetic(args.rows, seed=42)
    shard_rows = index.local.ntotal
    q_dev = synth_rows_device(args.queries, DIM, seed=4242, device=local_rank)
    q_pinned = torch.empty(q_dev.shape, dtype=torch.float32).pin_memory()   # e2e inputs live in pinned host memory
    q_pinned.copy_(q_dev)
    q_host = q_pinned.numpy()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`) ---------------------------------------------------
    for _ in range(args.warmup):
        D, I = index.search(q_dev, args.k)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, total_ms, launches, emitted, rescored = 0.0, 0.0, 0, 0, 0
    ev0.record()
    for _ in range(args.steps):
        D, I = index.search(q_dev, args.k)
        st = index.local.stats()
        scan_ms += st["scan_ms"]
        total_ms += st["total_ms"]
        launches += st["kernel_launches"] + (1 if world > 1 else 0)
        emitted += st["candidates_emitted"]
        rescored += st["candidates_rescored"]
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # per-rank device time of the local search (diagnostic: shows how unevenly the GPUs of the box run)
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        t = torch.tensor([total_ms / args.steps], dtype=torch.float64, device=dev)
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        per_rank_ms = [float(v.item()) for v in g]
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    scan_ms_step = max_over_ranks(scan_ms / args.steps)
    qps = args.queries / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer API (`e2e`) -------------------------------------------
    # results are read back into page-locked arrays the caller owns (faiss-style D=, I= arguments)
    Dh = torch.empty((args.queries, args.k), dtype=torch.float32).pin_memory().numpy()
    Ih = torch.empty((args.queries, args.k), dtype=torch.int64).pin_memory().numpy()
    for _ in range(2):
        index.search(q_host, args.k, D=Dh, I=Ih)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        index.search(q_host, args.k, D=Dh, I=Ih)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    e2e_qps = args.queries / e2e_s
    assert np.array_equal(Ih, I.cpu().numpy()), "host and device API disagree"

    # ---- phase breakdown of the multi-GPU step (diagnostics, outside the timed regions) ----------------
    phases = None
    if world > 1:
        index.profile = True
        acc = {}
        for _ in range(5):
            index.search(q_dev, args.k)
            for kk, vv in index.last_phase_ms.items():
                acc[kk] = acc.get(kk, 0.0) + vv / 5
        index.profile = False
        phases = {kk: max_over_ranks(vv) for kk, vv in acc.items()}
        barrier()

    # ---- parity spot check run with every timing: 4 queries re-scored in fp64 with torch over the regenerated
    # corpus (independent of the engine and of oracle/): same top-k ids, scores within 1e-5 relative ---------------
    n_chk = 4
    lo = index._bases[0][1]
    chk_s = torch.full((n_chk, args.k), -float("inf"), dtype=torch.float64, device=dev)
    chk_i = torch.full((n_chk, args.k), -1, dtype=torch.int64, device=dev)
    q64 = q_dev[:n_chk].double()
    for r0 in range(0, shard_rows, 2_000_000):
        nr = min(2_000_000, shard_rows - r0)
        xs = synth_rows_device(nr, DIM, seed=42, row0=lo + r0, device=local_rank)
        sc = q64 @ xs.double().T
        top = torch.topk(sc, min(args.k, nr), dim=1)
        cat_s = torch.cat([chk_s, top.values], 1)
        cat_i = torch.cat([chk_i, top.indices + lo + r0], 1)
        best = torch.topk(cat_s, args.k, dim=1)
        chk_s, chk_i = best.values, torch.gather(cat_i, 1, best.indices)
        del xs, sc
    if world > 1:
        gs = [torch.empty_like(chk_s) for _ in range(world)]
        gi = [torch.empty_like(chk_i) for _ in range(world)]
        dist.all_gather(gs, chk_s)
        dist.all_gather(gi, chk_i)
        cat_s, cat_i = torch.cat(gs, 1), torch.cat(gi, 1)
        best = torch.topk(cat_s, args.k, dim=1)
        chk_s, chk_i = best.values, torch.gather(cat_i, 1, best.indices)
    got_s, got_i = D[:n_chk].double(), I[:n_chk]
    same_sets = all(set(chk_i[r].tolist()) == set(got_i[r].tolist()) for r in range(n_chk))
    order = torch.argsort(chk_i, dim=1)
    ref_sorted = torch.gather(chk_s, 1, order)
    got_sorted = torch.gather(got_s, 1, torch.argsort(got_i, dim=1))
    max_rel = float(((ref_sorted - got_sorted).abs() / ref_sorted.abs().clamp_min(1e-30)).max()) if same_sets else float("nan")
    parity = {"queries_checked": n_chk, "recall_at_k": 1.0 if same_sets else 0.0, "max_rel_score_err": max_rel,
              "ok": bool(same_sets and max_rel <= 1e-5), "how": "fp64 torch re-scoring of the regenerated corpus"}
    assert parity["ok"], parity

    # ---- sanity: results are plausible for the N(0,1) corpus (rank-100 score ~ 4.5 sigma) ----------
    st = index.local.stats()
    assert st["path"] in (HAC_PATH_I8, HAC_PATH_MMA) and st["retries"] == 0, st
    i8 = st["path"] == HAC_PATH_I8
    assert st["screen_err_max"] <= st["margin_max"], st

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    flops_per_launch_set = 2.0 * args.queries * shard_rows * DIM        # algorithmic, per search per rank
    achieved = flops_per_launch_set / (scan_ms_step * 1e-3) / 1e12
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    if i8:
        # the scan runs tcgen05.mma.kind::i8: the int8 dense rate of the tensor pipe is twice the bf16 rate and
        # MEASURED_PEAKS.json holds no int8 figure, so the denominator is 2 x the measured sustained bf16 number.
        # DRAM traffic from the committed `ncu --set full` capture of the largest launch
        # (profiles/r01b_ncu_scan_i8_q2514_summary.txt): 7.227 GB read + 0.024 GB written for 9 392 880 rows of 768 B.
        peak, bytes_row, traffic_row = 2.0 * bf16_peak, 768.0, 772.0
        peak_src = "2 x %s bf16_tflops_sustained (kind::i8 runs at twice the bf16 tensor-pipe rate; no int8 entry in MEASURED_PEAKS.json)" % peak_kind
        cap_note = "7.25 GB for 9.39M rows"
    else:
        # f16 screen; ncu capture profiles/r01_final_ncu_scan_mma_summary.txt: 29.70 GB read + 0.05 GB written by the
        # launch that covers 19 300 592 rows, i.e. 1539 B per corpus row against 1536 B algorithmic: read once.
        peak, bytes_row, traffic_row = bf16_peak, 1536.0, 1539.0
        peak_src = "%s bf16_tflops_sustained (f16 and bf16 share the tensor-pipe rate)" % peak_kind
        cap_note = "29.70 GB for 19.3M rows"
    traffic = traffic_row * shard_rows
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "DRAM bytes per search per GPU, scaled from the ncu capture of the largest launch "
                                "(%s); algorithmic operand bytes %.2f GB" % (cap_note, shard_rows * bytes_row / 1e9),
                "peak_source": peak_src, "frac_of_bf16_peak": achieved / bf16_peak,
                "kernel": "scan_mma_kernel<2, int8>" if i8 else "scan_mma_kernel<1, f16>",
                "kernel_ms_per_step": scan_ms_step, "kernel_share_of_step": scan_ms_step / ms_per_step}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        n_sample = min(args.cpu_sample_rows, shard_rows)
        # the same workload: the first rows of the very corpus the GPU searched
        from haconvdr_b200.index import synth_rows_device as srd
        x_s = srd(n_sample, DIM, seed=42, row0=0, device=local_rank).cpu().numpy()
        cqps, dt, cores = cpu_port_rate(n_sample, args.queries, args.k, args.rows, steps=3, warmup=1, x=x_s, q=q_host)
        cpu_baseline = {"value": cqps, "unit": "queries/s", "cores": cores, "kind": "port",
                        "sample": "first %d of %d corpus rows x %d queries (%.2f s per pass, median-free mean of 3), "
                                  "scaled linearly in rows" % (n_sample, args.rows, args.queries, dt)}

    line = {
        "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": ("int8 screen (s32 accumulate)" if i8 else "f16 screen (f32 accumulate)") + " + f32 exact rescore",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "rows_per_gpu": shard_rows, "queries": args.queries,
                   "k": args.k, "dim": DIM, "parallelism": "corpus-shard x%d" % world,
                   "exchange": ("p2p symmetric-memory merge" if index._symm is not None else "nccl all-gather + merge") if world > 1 else None,
                   "l2": "inputs larger than L2 (%.1f GB of screen operands + the rescored fp32 rows streamed per step per GPU)" % (
                       shard_rows * bytes_row / 1e9)},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(q_host.nbytes),
                "d2h_bytes_per_step": int(args.queries * args.k * 12), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "parity_check": parity,
        "stats": {"candidates_emitted_per_step": emitted // args.steps, "candidates_rescored_per_step": rescored // args.steps,
                  "margin_max": st["margin_max"], "screen_err_max": st["screen_err_max"], "n_chunks": st["n_chunks"],
                  "search_ms_per_step_device": total_ms / args.steps, "setup_s": setup_s,
                  "multi_gpu_phase_ms_max_over_ranks": phases,
                  "local_search_ms_per_rank": [round(v, 3) for v in per_rank_ms],
                  "hbm_fp32_gb": st["bytes_fp32"] / 1e9, "hbm_shadow_gb": st["bytes_shadow"] / 1e9},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
