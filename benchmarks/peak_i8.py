#!/usr/bin/env python
"""Measure the int8 tensor-pipe peak of this GPU with the scan's own instruction (benchmarks/peak_i8.cu) and record
the clocks next to it.  Writes the summary bench.py uses as the roofline denominator of the int8 screen:

    python benchmarks/peak_i8.py [--seconds 3] [--out profiles/r02_peak_i8.json]

burst = one ~5 ms launch after 2 s of idle (best of 5), sustained = back-to-back launches for >= `seconds`.
The clock / power / throttle samples (nvidia-smi, 10 Hz) are split per variant at the "sustained_begin" marks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "benchmarks")
BIN = os.path.join(HERE, "peak_i8")
Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def build():
    """nvcc for sm_100a, in-tree (the binary travels to the GPU box with the snapshot)."""
    src = os.path.join(HERE, "peak_i8.cu")
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(src):
        return BIN
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", BIN, src])
    return BIN


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_peak_i8.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    build()
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    samples.append((time.time(), [v.strip() for v in out.split(",")]))
            except Exception:
                pass
            stop.wait(0.1)

    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    cmd = [BIN, "--seconds", str(args.seconds)] + (["--only", args.only] if args.only else [])
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True)
    events = []
    for line in proc.stdout:
        line = line.strip()
        if not line:
            continue
        ev = json.loads(line)
        ev["t"] = time.time()
        events.append(ev)
        print(line, flush=True)
    rc = proc.wait()
    stop.set()
    th.join(timeout=6)
    if rc != 0:
        sys.exit("peak_i8 exited with %d" % rc)
    results = []
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for i, ev in enumerate(events):
        if ev.get("event") != "sustained_begin":
            continue
        res = next(e for e in events[i + 1:] if e.get("event") == "result")
        win = [s for t, s in samples if ev["t"] <= t <= res["t"]]
        sm = sorted(float(s[0]) for s in win if s[0].replace(".", "").isdigit())
        pw = [float(s[2]) for s in win if s[2].replace(".", "").isdigit()]
        res = dict(res)
        res.pop("t", None)
        res.pop("event", None)
        res["clocks_sustained"] = {
            "sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
            "sm_max_mhz": max((float(s[1]) for s in win if s[1].replace(".", "").isdigit()), default=None),
            "power_w_max": max(pw) if pw else None, "samples": len(win),
            "reasons": sorted({n for s in win if len(s) >= 7 for n, v in zip(names, s[3:7]) if v == "Active"})}
        results.append(res)
    dev = next((e for e in events if e.get("event") == "device"), {})
    main_res = next((r for r in results if r["cta_group"] == 2 and r["data"] == "gauss"), results[0] if results else {})
    summary = {
        "what": "tcgen05.mma.kind::i8 dense peak measured with benchmarks/peak_i8.cu (no operand loads, no epilogue)",
        "gpu": dev.get("name"), "sm_count": dev.get("sm_count"), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "i8_tops_burst": main_res.get("i8_tops_burst"), "i8_tops_sustained": main_res.get("i8_tops_sustained"),
        "i8_tops_steady_tail": main_res.get("i8_tops_steady_tail"),
        "headline_variant": "cta_group::2, int8 operands ~ N(0, 34^2) (the scan's operand statistics)",
        "variants": results,
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({"written": args.out, "i8_tops_burst": summary["i8_tops_burst"],
                      "i8_tops_sustained": summary["i8_tops_sustained"]}))


if __name__ == "__main__":
    main()
