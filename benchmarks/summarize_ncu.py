#!/usr/bin/env python
"""Turn ncu CSV output into the short summaries kept under profiles/.

  launches <csv>            `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list ->
                            per-kernel time and share of the LAST complete search in the list (a search starts at
                            init_search_kernel and ends at final_select_kernel)
  traffic <csv> [out.json] the same launch list taken with `--metrics gpu__time_duration.sum,dram__bytes_read.sum,
                            dram__bytes_write.sum`: DRAM bytes read / written per kernel over the LAST complete search
                            (what bench.py reports as `roofline.traffic_step`), printed and optionally written as JSON
  raw <csv> [kernel-regex]  `ncu -i x.ncu-rep --page raw --csv` -> the roofline-relevant metrics of the first
                            matching kernel, one `name [unit] = value` line each
"""
import csv
import re
import sys

KEEP = ("sm__pipe_tensor_cycles_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__inst_executed.sum.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_active.avg")


def rows_of(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("hac::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def launches(path):
    rows = [r for r in rows_of(path) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [short(r["Kernel Name"]) for r in rows]
    ends = [i for i, n in enumerate(names) if n.startswith("final_select_kernel")]
    if not ends:
        sys.exit("no final_select_kernel in the launch list")
    end = ends[-1]
    start = max(i for i, n in enumerate(names[:end]) if n.startswith("init_search_kernel"))
    agg, order = {}, []
    for r, n in zip(rows[start:end + 1], names[start:end + 1]):
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        if n not in agg:
            agg[n] = [0.0, 0]
            order.append(n)
        agg[n][0] += ns
        agg[n][1] += 1
    total = sum(v[0] for v in agg.values())
    for n in order:
        print("%-44s x%-3d %12.1f us  %5.1f%%" % (n[:44], agg[n][1], agg[n][0] / 1e3, 100.0 * agg[n][0] / total))
    print("%-44s      %12.1f us" % ("total (one search, launches %d..%d)" % (start, end), total / 1e3))


def _value(r):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3,
             "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
    return v * scale.get(u, 1.0)


def traffic(path, out_json=None):
    import json
    per_launch = {}
    for r in rows_of(path):
        key = int(r["ID"])
        ent = per_launch.setdefault(key, {"name": short(r["Kernel Name"])})
        ent[r["Metric Name"]] = _value(r)
    ids = sorted(per_launch)
    names = [per_launch[i]["name"] for i in ids]
    ends = [i for i, n in enumerate(names) if n.startswith("final_select_kernel")]
    if not ends:
        sys.exit("no final_select_kernel in the launch list")
    end = ends[-1]
    start = max(i for i, n in enumerate(names[:end]) if n.startswith("init_search_kernel"))
    agg, order = {}, []
    for i in ids[start:end + 1]:
        e = per_launch[i]
        n = e["name"]
        if n not in agg:
            agg[n] = {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0}
            order.append(n)
        agg[n]["launches"] += 1
        agg[n]["time_us"] += e.get("gpu__time_duration.sum", 0.0) / 1e3
        agg[n]["dram_read_bytes"] += e.get("dram__bytes_read.sum", 0.0)
        agg[n]["dram_write_bytes"] += e.get("dram__bytes_write.sum", 0.0)
    tot_r = sum(v["dram_read_bytes"] for v in agg.values())
    tot_w = sum(v["dram_write_bytes"] for v in agg.values())
    tot_t = sum(v["time_us"] for v in agg.values())
    for n in order:
        v = agg[n]
        print("%-44s x%-3d %10.1f us  read %9.3f GB  write %8.3f GB" % (n[:44], v["launches"], v["time_us"],
                                                                       v["dram_read_bytes"] / 1e9, v["dram_write_bytes"] / 1e9))
    print("%-44s      %10.1f us  read %9.3f GB  write %8.3f GB" % ("total (one search)", tot_t, tot_r / 1e9, tot_w / 1e9))
    if out_json:
        with open(out_json, "w") as f:
            json.dump({"what": "DRAM bytes of every kernel of ONE search (ncu --metrics dram__bytes_read.sum,"
                               "dram__bytes_write.sum,gpu__time_duration.sum --clock-control none; serialised launches)",
                       "dram_read_bytes": tot_r, "dram_write_bytes": tot_w, "dram_bytes": tot_r + tot_w,
                       "kernel_time_us_serialised": tot_t, "per_kernel": {n: agg[n] for n in order}}, f, indent=1)


def raw(path, pattern):
    rows = rows_of(path)
    # --page raw --csv: one row per launch, one column per metric; row 0 after the header holds the units
    units = rows[0] if rows and not rows[0].get("ID", "").strip().isdigit() else {}
    for r in rows:
        if not r.get("ID", "").strip().isdigit() or not re.search(pattern, r.get("Kernel Name", "")):
            continue
        print("Kernel Name [] = %s" % short(r["Kernel Name"]))
        for k in sorted(r):
            if any(k.startswith(p) for p in KEEP) and r[k] not in ("", "n/a"):
                print("%s [%s] = %s" % (k, units.get(k, ""), r[k]))
        return
    sys.exit("no kernel matching %r" % pattern)


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "raw", "traffic"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ".")
