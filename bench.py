#!/usr/bin/env python
"""Benchmarks of the exact inner-product search path (BASELINE.json configs), one JSON line each.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config topiocqa|qrecc|turn|ksweep]

  topiocqa (default, configs[1]): 25 700 592 x 768 fp32 random-normal passages generated on the device (counter-based,
      seed 42, row-indexed so any shard count sees the same corpus), 2514 queries (seed 4242), k = 100.  One "step" =
      one search of all 2514 queries over the whole corpus.  With N ranks (torchrun, one per GPU) the corpus is sharded N
      ways (strong scaling): local exact top-k with cross-shard threshold exchange over NVLink -> one P2P merge kernel.
      The line also carries `secondary`: turn latencies (batch 1 / 4 / 32) and the k sweep, measured on the same index.
  qrecc (configs[2]): 54 573 064 x 768, 8209 queries - needs >= 2 GPUs.
  turn (configs[3]): batches of 1 / 4 / 32 queries over the 25.7M corpus, latency per search, HBM roofline.
  ksweep (configs[4]): k = 1 / 10 / 100 / 1000 at 2514 queries.

`value` = the metric with corpus, queries and results resident in HBM; `e2e` = the same through the host-buffer API
(numpy in / numpy out, H2D + D2H inside the timed region); `roofline` = the dominant kernel (tcgen05 int8 screen scan)
timed live with CUDA events inside the library, against the int8 tensor peak measured by benchmarks/peak_i8.py
(profiles/r02_peak_i8.json); `cpu_baseline` = the faiss-restatement CPU port on this box's host cores (bounded sample).
Every run ends with a parity check: >= 64 queries spread over all query tiles against an fp64 re-scoring of the
regenerated corpus, ids + ORDER + scores, tolerance groups per SURVEY.md 8d (oracle.compare, used as the checker only).
`--impl reference` times only the reference's CPU path (see run_reference).
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

DIM = 768
CONFIGS = {
    "topiocqa": {"rows": 25_700_592, "queries": 2514, "k": 100, "metric": "queries/sec, exact top-100, 25.7Mx768",
                 "workload": "topiocqa-scale synthetic 25.7Mx768 fp32, 2514 queries, top_k=100 (BASELINE.json configs[1])"},
    "qrecc": {"rows": 54_573_064, "queries": 8209, "k": 100, "metric": "queries/sec, exact top-100, 54.6Mx768",
              "workload": "qrecc-scale synthetic 54.6Mx768 fp32, 8209 queries, top_k=100 (BASELINE.json configs[2])"},
    "turn": {"rows": 25_700_592, "queries": 32, "k": 100, "metric": "online turn latency, batch 1/4/32, exact top-100, 25.7Mx768",
             "workload": "online conversational turns: batches of 1/4/32 queries over 25.7Mx768 (BASELINE.json configs[3])"},
    "ksweep": {"rows": 25_700_592, "queries": 2514, "k": 100, "metric": "queries/sec, exact top-k sweep k=1/10/100/1000, 25.7Mx768",
               "workload": "top_k sweep k=1/10/100/1000 at 2514 queries over 25.7Mx768 (BASELINE.json configs[4])"},
}
TURN_BATCHES = (1, 4, 32)
SWEEP_KS = (1, 10, 100, 1000)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--workload", default=None, choices=["topiocqa", "qrecc"], help="alias of --config (round 1)")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the turn-latency / k-sweep keys of the default line")
    ap.add_argument("--parity-queries", type=int, default=64)
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--ref-blocks", type=int, default=2, help="--impl reference: on-disk blocks per step")
    ap.add_argument("--ref-block-rows", type=int, default=1_000_000, help="--impl reference: rows per on-disk block")
    args = ap.parse_args()
    args.config = args.config or args.workload or "topiocqa"
    cfg = CONFIGS[args.config]
    args.rows = args.rows or cfg["rows"]
    args.queries = args.queries or cfg["queries"]
    args.k = args.k or cfg["k"]
    args.metric, args.workload_name = cfg["metric"], cfg["workload"]
    return args


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(n_queries, k, n_total_rows, x, q, steps=1, warmup=0):
    """Time the faiss-restatement CPU port (oracle/cpu_port.py: BLAS sgemm in faiss' 4096x1024 blocking
    + C heaps) on a bounded sample and scale linearly in corpus rows to the full workload."""
    from oracle import cpu_port
    cpu_port.build()
    for _ in range(warmup):
        cpu_port.search_blas(q, x, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port.search_blas(q, x, k)
    dt = (time.perf_counter() - t0) / steps
    full_time = dt * (n_total_rows / x.shape[0])
    return n_queries / full_time, dt


class _CpuPortIndex:
    """faiss.IndexFlatIP surface over the CPU port, with the time spent inside it recorded."""

    def __init__(self, d):
        self.d, self._x, self.t_index = d, None, 0.0

    def add(self, x):
        import numpy as np
        t0 = time.perf_counter()
        x = np.ascontiguousarray(x, dtype=np.float32)            # faiss copies on add
        self._x = x.copy() if self._x is None else np.concatenate([self._x, x], 0)
        self.t_index += time.perf_counter() - t0

    def search(self, q, k):
        from oracle import cpu_port
        t0 = time.perf_counter()
        out = cpu_port.search_blas(q, self._x, k)
        self.t_index += time.perf_counter() - t0
        return out

    def reset(self):
        self._x = None


def run_reference(args):
    """The reference's CPU path, timed through the reference's own per-block loop.

    faiss is not installable offline, so `index` is the faiss-restatement CPU port
    (oracle/cpu_port.py: sgemm 4096x1024 blocks + C heaps, all host threads).  It is driven by the reference's own
    `search_one_by_one_with_faiss` (/root/reference/src/test_HAConvDR_topiocqa.py:74-162, imported unmodified through
    oracle/ref_harness.py) when the reference tree is mounted, else by its cost-faithful restatement
    oracle/ref_loop.py (tuples, deepcopy, two-pointer merges) - so block unpickling, add, search, reset, tuple
    materialisation and the Python merge are all inside the step.  One step = that loop over `--ref-blocks` on-disk
    pickle blocks of `--ref-block-rows` rows (a bounded sample); `ms_per_step` is the MEASURED sample time, `value` the
    projection to the full workload (rows-proportional part scaled by rows, per-block Python part by the reference's
    block count), labelled `extrapolated`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count()
    os.environ["OMP_NUM_THREADS"] = str(n_threads)     # torchrun pins OMP_NUM_THREADS=1; the CPU path uses every core
    os.environ["MKL_NUM_THREADS"] = str(n_threads)
    import pickle
    import tempfile
    import types
    import numpy as np
    import torch
    torch.set_num_threads(n_threads)
    from oracle import cpu_port, ref_harness
    cpu_port.build()
    t_start = time.perf_counter()
    rng = np.random.default_rng(42)
    q = rng.standard_normal((args.queries, DIM), dtype=np.float32)
    n_blocks, n_b = args.ref_blocks, min(args.ref_block_rows, args.rows)
    if ref_harness.available():
        loop = ref_harness.load_reference_module().search_one_by_one_with_faiss
        loop_name = "reference's own search_one_by_one_with_faiss (src/test_HAConvDR_topiocqa.py:74-162)"
        call = lambda d, idx: loop(types.SimpleNamespace(passage_block_num=n_blocks), d, idx, q, args.k)   # noqa: E731
    else:
        from oracle.ref_loop import search_blocks_python
        loop_name = "oracle/ref_loop.py (cost-faithful restatement of the reference loop; /root/reference not mounted)"
        call = lambda d, idx: search_blocks_python(n_blocks, d, idx, q, args.k)   # noqa: E731
    with tempfile.TemporaryDirectory(prefix="hac_ref_blocks_") as tmp:
        o = 0
        for b in range(n_blocks):                       # written with the reference writer's protocol (gen_doc_embeddings.py:127-155)
            x = rng.standard_normal((n_b, DIM), dtype=np.float32)
            with open(os.path.join(tmp, "passage_emb_block_%d.pb" % b), "wb") as h:
                pickle.dump(x, h, protocol=4)
            with open(os.path.join(tmp, "passage_embid_block_%d.pb" % b), "wb") as h:
                pickle.dump(np.arange(o, o + n_b, dtype=np.int64), h, protocol=4)
            o += n_b
            del x
        t0 = time.perf_counter()
        with open(os.path.join(tmp, "passage_emb_block_0.pb"), "rb") as h:
            pickle.load(h)
        t_load_block = time.perf_counter() - t0         # page-cache read + unpickle of one block
        index = _CpuPortIndex(DIM)
        for _ in range(args.warmup):
            call(tmp, index)
        index.t_index = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            D, I = call(tmp, index)
        dt = (time.perf_counter() - t0) / args.steps
        t_index = index.t_index / args.steps
    assert D.shape == (args.queries, 2 * args.k if n_blocks > 1 else args.k)
    # projection to the full workload: add + search + unpickling scale with rows; tuple lists and the Python merge are
    # a per-block cost (Q x k work), and the reference cuts TopiOCQA into 11 blocks, QReCC into 22
    sample_rows = n_blocks * n_b
    ref_blocks_full = max(1, round(args.rows / 2_336_418))
    t_rows = t_index + n_blocks * t_load_block
    t_python_per_block = max(0.0, dt - t_rows) / n_blocks
    full_s = t_rows * (args.rows / sample_rows) + t_python_per_block * ref_blocks_full
    qps = args.queries / full_s
    sample = ("%d blocks x %d rows (of %d) x %d queries per step through %s: %.2f s/step measured "
              "(index add+search %.2f s, unpickling %.2f s, Python tuples+merge %.2f s)" % (
                  n_blocks, n_b, args.rows, args.queries, loop_name, dt, t_index, n_blocks * t_load_block,
                  t_python_per_block * n_blocks))
    line = {
        "impl": "reference", "metric": args.metric, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "extrapolated": True, "sample_ms_per_step": dt * 1e3, "projected_full_workload_ms_per_step": full_s * 1e3,
        "projection": "rows-proportional time x %.2f + per-block Python time x %d reference blocks" % (
            args.rows / sample_rows, ref_blocks_full),
        "config": {"workload": args.workload_name, "rows": args.rows, "queries": args.queries, "k": args.k, "dim": DIM,
                   "sample_rows_per_step": sample_rows,
                   "note": "faiss is not installable offline; the index is the faiss-restatement CPU port (IndexFlatIP: "
                           "sgemm 4096x1024 blocks + per-query heaps); `ms_per_step` is the measured sample step, "
                           "`value` the projection to the full workload"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample, "extrapolated": True},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML in-process every 10 ms
    (nvidia_ml_py), falling back to one `nvidia-smi` query per 100 ms when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu, self.samples, self._stop, self._t = gpu_index, [], threading.Event(), None
        self.source, self._nvml, self._h, self._max = "nvidia-smi", None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        try:
            watts = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        except Exception:
            watts = 0.0
        return [sm, self._max, watts] + [bool(mask & self.BITS[k]) for k in self.NAMES]

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        v = [x.strip() for x in out.split(",")]
        if len(v) < 7 or not v[0].replace(".", "").isdigit():
            return None
        watts = float(v[2]) if v[2].replace(".", "").isdigit() else 0.0
        return [float(v[0]), float(v[1]), watts] + [x == "Active" for x in v[3:7]]

    def _run(self):
        while not self._stop.is_set():
            try:
                smp = self._sample_nvml() if self._nvml is not None else self._sample_smi()
                if smp:
                    self.samples.append(smp)
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples if s[1]]
        reasons = sorted({n for s in self.samples for n, v in zip(self.NAMES, s[3:7]) if v})
        watts = [s[2] for s in self.samples if s[2]]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(watts) if watts else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


def measured_peaks():
    """(peaks dict, source): HBM / bf16 from the driver-written MEASURED_PEAKS.json, the int8 tensor peak from this
    repo's own measurement (benchmarks/peak_i8.py -> profiles/r02_peak_i8.json)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            peaks, kind = json.load(f), "measured"
    else:
        peaks, kind = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"
    p8 = os.path.join(ROOT, "profiles", "r02_peak_i8.json")
    if os.path.exists(p8):
        with open(p8) as f:
            j = json.load(f)
        if j.get("i8_tops_sustained"):
            peaks["i8_tops_sustained"] = float(j["i8_tops_sustained"])
            peaks["i8_tops_burst"] = float(j.get("i8_tops_burst") or 0.0)
    return peaks, kind


def step_traffic_note():
    """Whole-step DRAM traffic from the committed ncu capture of one full search (profiles/r02_step_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_step_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


class Harness:
    """Index + queries + the distributed plumbing shared by every config."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from haconvdr_b200 import FlatIPIndex
        from haconvdr_b200.index import synth_rows_device
        from haconvdr_b200.sharded import ShardedFlatIPIndex
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner on stdout at communicator creation; stdout must carry only the
            # one JSON line, so fd 1 points at stderr until the first collective has run.
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
        t0 = time.perf_counter()
        self.index = ShardedFlatIPIndex(DIM, FlatIPIndex(DIM, self.local_rank), exchange=args.exchange)
        self.index.add_synthetic(args.rows, seed=42)
        self.shard_rows = self.index.local.ntotal
        self.row_lo = self.index._bases[0][1]
        self.q_dev = synth_rows_device(args.queries, DIM, seed=4242, device=self.local_rank)
        self.q_pinned = torch.empty(self.q_dev.shape, dtype=torch.float32).pin_memory()   # e2e inputs: pinned host memory
        self.q_pinned.copy_(self.q_dev)
        self.q_host = self.q_pinned.numpy()
        torch.cuda.synchronize()
        self.setup_s = time.perf_counter() - t0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    # -- timing ------------------------------------------------------------------------------------
    def time_device(self, q_dev, k, steps, warmup):
        """`steps` searches with device-resident queries / results: (ms per step max over ranks, stats sums, D, I)."""
        torch = self.torch
        for _ in range(warmup):
            D, I = self.index.search(q_dev, k)
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {"scan_ms": 0.0, "total_ms": 0.0, "tail_ms": 0.0, "kernel_launches": 0, "candidates_emitted": 0,
               "candidates_rescored": 0}
        ev0.record()
        for _ in range(steps):
            D, I = self.index.search(q_dev, k)
            st = self.index.local.stats()
            for kk in acc:
                acc[kk] += st[kk]
            acc["kernel_launches"] += 1 if self.world > 1 else 0      # the cross-shard merge kernel
        ev1.record()
        self.barrier()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1) / steps)
        return ms, {kk: vv / steps for kk, vv in acc.items()}, D, I

    def time_host(self, q_host, k, steps, warmup=2):
        """The host-buffer API (numpy in, page-locked numpy out): seconds per step, max over ranks."""
        torch = self.torch
        nq = q_host.shape[0]
        Dh = torch.empty((nq, k), dtype=torch.float32).pin_memory().numpy()
        Ih = torch.empty((nq, k), dtype=torch.int64).pin_memory().numpy()
        for _ in range(warmup):
            self.index.search(q_host, k, D=Dh, I=Ih)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.index.search(q_host, k, D=Dh, I=Ih)
        torch.cuda.synchronize()
        s = self.max_over_ranks((time.perf_counter() - t0) / steps)
        self.barrier()
        return s, Dh, Ih

    # -- parity ------------------------------------------------------------------------------------
    def parity(self, q_dev, D, I, k, n_chk):
        """ids + order + scores of `n_chk` queries spread over every query tile against an fp64 re-scoring of the
        regenerated corpus (torch, independent of the engine's search kernels), merged over the ranks; compared with
        the tolerance-group comparator of the test-suite (oracle.compare - the checker, not the product)."""
        import numpy as np
        from haconvdr_b200.index import synth_rows_device
        from oracle.compare import compare_topk
        torch, dist = self.torch, self.dist
        nq = q_dev.shape[0]
        sel = np.unique(np.linspace(0, nq - 1, min(n_chk, nq)).astype(np.int64))
        sel_t = torch.from_numpy(sel).to(self.dev)
        kk = k + 16                                          # a few ranks past k: boundary substitutions within tolerance
        best_s = torch.full((len(sel), kk), -float("inf"), dtype=torch.float64, device=self.dev)
        best_i = torch.full((len(sel), kk), -1, dtype=torch.int64, device=self.dev)
        q64 = q_dev[sel_t].double()
        slab = 1_000_000
        for r0 in range(0, self.shard_rows, slab):
            nr = min(slab, self.shard_rows - r0)
            xs = synth_rows_device(nr, DIM, seed=42, row0=self.row_lo + r0, device=self.local_rank)
            sc = q64 @ xs.double().T
            top = torch.topk(sc, min(kk, nr), dim=1)
            cat_s = torch.cat([best_s, top.values], 1)
            cat_i = torch.cat([best_i, top.indices + self.row_lo + r0], 1)
            best = torch.topk(cat_s, kk, dim=1)
            best_s, best_i = best.values, torch.gather(cat_i, 1, best.indices)
            del xs, sc
        if self.world > 1:
            gs = [torch.empty_like(best_s) for _ in range(self.world)]
            gi = [torch.empty_like(best_i) for _ in range(self.world)]
            dist.all_gather(gs, best_s)
            dist.all_gather(gi, best_i)
            best_s, best_i = torch.cat(gs, 1), torch.cat(gi, 1)
        ext_D, ext_I = best_s.cpu().numpy(), best_i.cpu().numpy()
        order = np.lexsort((ext_I, -ext_D), axis=1)[:, :kk]           # (score desc, id asc)
        ext_D, ext_I = np.take_along_axis(ext_D, order, 1), np.take_along_axis(ext_I, order, 1)
        lookup = [dict(zip(ext_I[r].tolist(), ext_D[r].tolist())) for r in range(len(sel))]
        scores_of = lambda qi, ids: np.asarray([lookup[qi].get(int(i), -np.inf) for i in ids])   # noqa: E731
        got_D, got_I = D[sel_t].cpu().numpy(), I[sel_t].cpu().numpy()
        rep = compare_topk(ext_D[:, :k], ext_I[:, :k], got_D, got_I, rtol=1e-5, ref_scores_of=scores_of)
        out = {"queries_checked": int(len(sel)), "query_tiles_covered": int(len(set((sel // 128).tolist()))),
               "order_checked": True, "rows_identical_order": rep.n_exact_rows, "rows_equal_up_to_tolerance_groups":
               rep.n_tolerated_rows, "recall_at_k": rep.recall, "max_rel_score_err": rep.max_rel_score_err,
               "ok": bool(rep.ok and rep.recall == 1.0), "failures": rep.failures[:4],
               "how": "fp64 torch re-scoring of the regenerated corpus, merged over %d rank(s); oracle.compare.compare_topk "
                      "(ids, order, scores; tolerance groups at 1e-5 relative)" % self.world}
        assert out["ok"], out
        return out


def roofline_tensor(h, stats, ms_per_step, queries, i8):
    peaks, peak_kind = measured_peaks()
    flops = 2.0 * queries * h.shard_rows * DIM                    # algorithmic, per search per rank
    scan_ms = h.max_over_ranks(stats["scan_ms"])
    achieved = flops / (scan_ms * 1e-3) / 1e12
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    if i8:
        # ncu capture of the largest scan launch (profiles/r01b_ncu_scan_i8_q2514_summary.txt): 7.227 GB read + 0.024 GB
        # written for 9 392 880 rows of 768 B -> 772 B per corpus row
        bytes_row, traffic_row, cap_note = 768.0, 772.0, "7.25 GB for 9.39M rows"
        if "i8_tops_sustained" in peaks:
            peak = peaks["i8_tops_sustained"]
            peak_src = ("measured tcgen05.mma.kind::i8 cta_group::2 sustained rate of this repo's own peak kernel "
                        "(benchmarks/peak_i8.cu -> profiles/r02_peak_i8.json; burst %.0f)" % peaks.get("i8_tops_burst", 0.0))
        else:
            peak = 2.0 * bf16_peak
            peak_src = "2 x %s bf16_tflops_sustained (profiles/r02_peak_i8.json missing)" % peak_kind
    else:
        # f16 screen; ncu capture profiles/r01_final_ncu_scan_mma_summary.txt: 1539 B per corpus row against 1536 B
        peak, bytes_row, traffic_row, cap_note = bf16_peak, 1536.0, 1539.0, "29.70 GB for 19.3M rows"
        peak_src = "%s bf16_tflops_sustained (f16 and bf16 share the tensor-pipe rate)" % peak_kind
    step = step_traffic_note()
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic_row * h.shard_rows,
            "traffic_note": "DRAM bytes of the scan launches of one search per GPU, scaled from the ncu capture of the "
                            "largest launch (%s); algorithmic operand bytes %.2f GB" % (cap_note, h.shard_rows * bytes_row / 1e9),
            "traffic_step": step, "peak_source": peak_src, "frac_of_bf16_peak": achieved / bf16_peak,
            "frac_of_2x_bf16_sustained": achieved / (2.0 * bf16_peak),
            "kernel": "scan_mma_kernel<2, int8>" if i8 else "scan_mma_kernel<1, f16>",
            "kernel_ms_per_step": scan_ms, "kernel_share_of_step": scan_ms / ms_per_step,
            "step_frac_of_peak": flops / (ms_per_step * 1e-3) / 1e12 / peak}


def turn_latencies(h, k, steps, warmup, with_host=True):
    """Batches of 1 / 4 / 32 queries: per-search latency (device API, CUDA events, max over ranks) and through the host
    API, with the HBM figures of the scan: bytes actually streamed (the int8 image, rows x 768 B) over the scan kernel
    time, and SURVEY 8d's definition N x 768 x 4 B over the whole search."""
    peaks, _ = measured_peaks()
    out = {}
    for nq in TURN_BATCHES:
        q = h.q_dev[:nq].contiguous()
        ms, st, D, I = h.time_device(q, k, steps, warmup)
        scan_ms = h.max_over_ranks(st["scan_ms"])
        streamed = h.shard_rows * DIM * 1.0                      # int8 image bytes read by the scan of one search
        ent = {"ms_per_search": ms, "scan_ms": scan_ms, "tail_ms": st["tail_ms"], "launches": st["kernel_launches"],
               "hbm_gbs_streamed": streamed / (scan_ms * 1e-3) / 1e9,
               "hbm_frac_of_measured_copy_peak": streamed / (scan_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]),
               "hbm_gbs_8d_definition": h.shard_rows * DIM * 4.0 / (ms * 1e-3) / 1e9,
               "frac_8d_of_8tbs": h.shard_rows * DIM * 4.0 / (ms * 1e-3) / 8e12}
        if with_host:
            s, _, Ih = h.time_host(h.q_host[:nq], k, steps)
            ent["e2e_ms_per_search"] = s * 1e3
        out[str(nq)] = ent
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from haconvdr_b200 import HAC_PATH_I8, HAC_PATH_MMA
    h = Harness(args)
    rank, world = h.rank, h.world
    index = h.index
    line = {"metric": args.metric, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic"}
    config = {"workload": args.workload_name, "rows": args.rows, "rows_per_gpu": h.shard_rows, "queries": args.queries,
              "k": args.k, "dim": DIM, "parallelism": "corpus-shard x%d" % world}
    sampler = ClockSampler(h.local_rank)

    if args.config in ("topiocqa", "qrecc"):
        for _ in range(args.warmup):
            index.search(h.q_dev, args.k)
        h.barrier()
        if rank == 0:
            sampler.start()
        ms_per_step, st_avg, D, I = h.time_device(h.q_dev, args.k, args.steps, 0)
        clocks = sampler.stop() if rank == 0 else None
        per_rank_ms = [st_avg["total_ms"]]
        if world > 1:
            t = torch.tensor([st_avg["total_ms"]], dtype=torch.float64, device=h.dev)
            g = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            per_rank_ms = [float(v.item()) for v in g]
        qps = args.queries / (ms_per_step * 1e-3)
        e2e_s, Dh, Ih = h.time_host(h.q_host, args.k, args.steps)
        assert np.array_equal(Ih, I.cpu().numpy()), "host and device API disagree"
        phases = None
        if world > 1:                                   # phase breakdown of the multi-GPU step (outside the timed regions)
            index.profile = True
            acc = {}
            for _ in range(5):
                index.search(h.q_dev, args.k)
                for kk, vv in index.last_phase_ms.items():
                    acc[kk] = acc.get(kk, 0.0) + vv / 5
            index.profile = False
            phases = {kk: h.max_over_ranks(vv) for kk, vv in acc.items()}
            h.barrier()
        parity = h.parity(h.q_dev, D, I, args.k, args.parity_queries)
        st = index.local.stats()
        assert st["path"] in (HAC_PATH_I8, HAC_PATH_MMA) and st["retries"] == 0, st
        assert st["screen_err_max"] <= st["margin_max"], st
        i8 = st["path"] == HAC_PATH_I8
        pairs_all = h.sum_over_ranks(st_avg["candidates_rescored"])
        secondary = None
        if args.config == "topiocqa" and not args.no_secondary:
            secondary = {"turn_latency": turn_latencies(h, args.k, steps=20, warmup=3),
                         "note": "configs[3] / configs[4] on the same resident index, outside the headline timed region; "
                                 "`python bench.py --config turn|ksweep` gives each its own line"}
            sweep = {}
            for k in SWEEP_KS:
                if k == args.k:
                    sweep[str(k)] = {"ms_per_step": ms_per_step, "queries_per_s": qps, "path": st["path"]}
                    continue
                ms_k, st_k, Dk, Ik = h.time_device(h.q_dev, k, 3, 1)
                assert torch.equal(Ik[:, :min(k, args.k)], I[:, :min(k, args.k)]), "k sweep: prefix mismatch at k=%d" % k
                sweep[str(k)] = {"ms_per_step": ms_k, "queries_per_s": args.queries / (ms_k * 1e-3),
                                 "path": index.local.stats()["path"], "scan_ms": st_k["scan_ms"],
                                 "rescored_pairs": st_k["candidates_rescored"]}
            secondary["ksweep"] = sweep
        roofline = roofline_tensor(h, st_avg, ms_per_step, args.queries, i8)     # (max over ranks: every rank takes part)
        if rank != 0:
            if world > 1:
                dist.destroy_process_group()
            return
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            from haconvdr_b200.index import synth_rows_device as srd
            n_sample = min(args.cpu_sample_rows, h.shard_rows)
            x_s = srd(n_sample, DIM, seed=42, row0=0, device=h.local_rank).cpu().numpy()   # the corpus the GPU searched
            cqps, dt = cpu_port_rate(args.queries, args.k, args.rows, x_s, h.q_host, steps=3, warmup=1)
            cpu_baseline = {"value": cqps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                            "extrapolated": True,
                            "sample": "first %d of %d corpus rows x %d queries (%.2f s per pass, mean of 3), "
                                      "scaled linearly in rows" % (n_sample, args.rows, args.queries, dt)}
        config.update({"exchange": ("p2p symmetric-memory merge" if index._symm is not None else "nccl all-gather + merge") if world > 1 else None,
                       "l2": "inputs larger than L2 (%.1f GB of screen operands + the rescored fp32 rows streamed per step per GPU)" % (
                           h.shard_rows * (768.0 if i8 else 1536.0) / 1e9)})
        line.update({
            "value": qps, "ms_per_step": ms_per_step,
            "dtype": ("int8 screen (s32 accumulate)" if i8 else "f16 screen (f32 accumulate)") + " + f32 exact rescore",
            "config": config,
            "e2e": {"value": args.queries / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(h.q_host.nbytes),
                    "d2h_bytes_per_step": int(args.queries * args.k * 12), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(round(st_avg["kernel_launches"] * args.steps)), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "clocks": clocks, "parity_check": parity,
            "stats": {"candidates_emitted_per_step": st_avg["candidates_emitted"],
                      "candidates_rescored_per_step": st_avg["candidates_rescored"],
                      "rescored_pairs_all_ranks": pairs_all, "rescored_pairs_per_query": pairs_all / args.queries,
                      "margin_max": st["margin_max"], "screen_err_max": st["screen_err_max"], "n_chunks": st["n_chunks"],
                      "n_sync_chunks": st["n_sync_chunks"], "pipelined": st["pipelined"], "tail_ms": st_avg["tail_ms"],
                      "search_ms_per_step_device": st_avg["total_ms"], "setup_s": h.setup_s,
                      "multi_gpu_phase_ms_max_over_ranks": phases,
                      "local_search_ms_per_rank": [round(v, 3) for v in per_rank_ms],
                      "hbm_fp32_gb": st["bytes_fp32"] / 1e9, "hbm_f16_gb": index.local.stats()["bytes_shadow"] / 1e9,
                      "hbm_int8_gb": st["bytes_i8"] / 1e9},
            "secondary": secondary})

    elif args.config == "turn":
        if rank == 0:
            sampler.start()
        batches = turn_latencies(h, args.k, args.steps, max(args.warmup, 3))
        clocks = sampler.stop() if rank == 0 else None
        D, I = index.search(h.q_dev, args.k)
        parity = h.parity(h.q_dev, D, I, args.k, 32)
        for nq in TURN_BATCHES[:-1]:                     # smaller batches return the rows of the largest one, bitwise
            Ds, Is = index.search(h.q_dev[:nq].contiguous(), args.k)
            assert torch.equal(Is, I[:nq]) and torch.equal(Ds, D[:nq])
        st = index.local.stats()
        if rank != 0:
            if world > 1:
                dist.destroy_process_group()
            return
        peaks, peak_kind = measured_peaks()
        b1 = batches["1"]
        config.update({"batches": list(TURN_BATCHES), "l2": "inputs larger than L2 (%.1f GB int8 image streamed per search per GPU)" % (
            h.shard_rows * 768.0 / 1e9)})
        line.update({
            "value": b1["ms_per_search"], "unit": "ms", "higher_is_better": False, "ms_per_step": b1["ms_per_search"],
            "dtype": "int8 screen (s32 accumulate) + f32 exact rescore", "config": config, "batches": batches,
            "e2e": {"value": b1["e2e_ms_per_search"], "unit": "ms", "h2d_bytes_per_step": DIM * 4, "d2h_bytes_per_step": args.k * 12},
            "gpu_launches": int(b1["launches"] * args.steps),
            "roofline": {"bound": "hbm", "achieved": b1["hbm_gbs_streamed"], "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                         "frac": b1["hbm_gbs_streamed"] / float(peaks["hbm_gbs"]),
                         "traffic": h.shard_rows * 768.0,
                         "traffic_note": "algorithmic bytes of the kernel = the int8 image, rows x 768 B (ncu: dram read = "
                                         "the image once, profiles/); SURVEY 8d counts the fp32 rows (N x 768 x 4 B / whole search): "
                                         "%.0f GB/s = %.2f of the nominal 8 TB/s" % (b1["hbm_gbs_8d_definition"], b1["frac_8d_of_8tbs"]),
                         "peak_source": "%s hbm_gbs (copy bandwidth)" % peak_kind, "kernel": "scan_mma_kernel<*, int8> (one query tile)",
                         "kernel_ms_per_step": b1["scan_ms"], "kernel_share_of_step": b1["scan_ms"] / b1["ms_per_search"]},
            "cpu_baseline": None, "clocks": clocks, "parity_check": parity,
            "stats": {"n_chunks": st["n_chunks"], "n_sync_chunks": st["n_sync_chunks"], "setup_s": h.setup_s}})

    else:   # ksweep
        sweep, clocks = {}, None
        ref = None
        for k in SWEEP_KS:
            if rank == 0 and k == 100:
                sampler.start()
            ms_k, st_k, Dk, Ik = h.time_device(h.q_dev, k, args.steps, max(args.warmup, 3))
            if rank == 0 and k == 100:
                clocks = sampler.stop()
            e2e_s, _, _ = h.time_host(h.q_host, k, max(2, args.steps // 2))
            stl = index.local.stats()
            par = h.parity(h.q_dev, Dk, Ik, k, 16)
            if ref is not None:
                kk = min(k, ref[0].shape[1])
                assert torch.equal(Ik[:, :kk], ref[1][:, :kk]) and torch.equal(Dk[:, :kk], ref[0][:, :kk]), "prefix mismatch"
            if ref is None or k > ref[0].shape[1]:
                ref = (Dk, Ik)
            sweep[str(k)] = {"ms_per_step": ms_k, "queries_per_s": args.queries / (ms_k * 1e-3),
                             "e2e_queries_per_s": args.queries / e2e_s, "path": stl["path"], "scan_ms": st_k["scan_ms"],
                             "rescored_pairs": st_k["candidates_rescored"], "launches": st_k["kernel_launches"],
                             "parity_ok": par["ok"], "d2h_bytes": args.queries * k * 12}
        s100 = sweep["100"]
        roofline = roofline_tensor(h, {"scan_ms": s100["scan_ms"]}, s100["ms_per_step"], args.queries, True)   # all ranks
        if rank != 0:
            if world > 1:
                dist.destroy_process_group()
            return
        config.update({"ks": list(SWEEP_KS), "l2": "inputs larger than L2"})
        line.update({
            "value": s100["queries_per_s"], "ms_per_step": s100["ms_per_step"],
            "dtype": "int8 screen (s32 accumulate; f16 warm slab sized by k) + f32 exact rescore", "config": config, "sweep": sweep,
            "e2e": {"value": s100["e2e_queries_per_s"], "unit": "queries/s", "h2d_bytes_per_step": int(h.q_host.nbytes),
                    "d2h_bytes_per_step": int(args.queries * 100 * 12)},
            "gpu_launches": int(s100["launches"] * args.steps),
            "roofline": roofline,
            "cpu_baseline": None, "clocks": clocks,
            "parity_check": {"ok": all(v["parity_ok"] for v in sweep.values()), "queries_checked": 16 * len(SWEEP_KS),
                             "order_checked": True, "how": "per k: fp64 re-scoring + prefix consistency across k"},
            "stats": {"setup_s": h.setup_s, "hbm_f16_gb": index.local.stats()["bytes_shadow"] / 1e9}})

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
