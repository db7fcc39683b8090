#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/ncu_summarise.py launches gpurun_out/launches.csv  > profiles/rNN_launches_summary.csv
    python tools/ncu_summarise.py full gpurun_out/a.ncu-rep [b.ncu-rep ...]  > profiles/rNN_ncu_full_summary.csv
    python tools/ncu_summarise.py traffic gpurun_out/profile_info.json name=report.ncu-rep:chunk[:coarse|fine] ...
                                                                        > profiles/traffic.json   (read by bench.py)
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic"]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("nf::", "").replace("(int)", "").replace("(bool)", "")
    return name.strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    n = 0
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= vi or "gpu__time_duration" not in ",".join(r):
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1; a[1] += v; n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {n} launches captured, {tot:.1f} ms total (cold-cache, serialised under the profiler: compare SHARES)")
    print("kernel,launches,total_ms,share")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{c},{ms:.3f},{ms / tot:.4f}')


def full(paths):
    print("report,kernel," + ",".join(METRICS))
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for v in rows[2:]:
            d = dict(zip(hdr, v)); u = dict(zip(hdr, units))
            print(p.split("/")[-1] + ',"' + short(d["Kernel Name"]) + '",' +
                  ",".join(f"{d.get(m, '')} {u.get(m, '')}".strip().replace(",", "") for m in METRICS))


def traffic(info_path, specs):
    import json
    info = json.load(open(info_path))
    out = {}
    for spec in specs:
        name, rest = spec.split("=")
        parts = rest.split(":")
        rep, chunk = parts[0], int(parts[1])
        which = parts[2] if len(parts) > 2 else "fine"
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
        ch = info["chunks"][chunk]
        out[name] = {"kernel": short(d["Kernel Name"]), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                     "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]), "chunk": chunk, "rays": ch["rays"],
                     "rows": ch["rows_" + which], "report": rep.split("/")[-1]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3:])
    else:
        full(sys.argv[2:])
