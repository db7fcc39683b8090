#!/usr/bin/env python
"""Per-source-line summary of an ncu report: samples, instructions executed, shared-memory excess wavefronts.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) > 10:
        d = dict(zip(hdr[4:], r[4:]))
        if not d.get("# Samples", "").isdigit():      # source text with embedded quotes: align from the right
            d = dict(zip(hdr[::-1], r[::-1]))
        if d.get("# Samples", "").isdigit() and d.get("Instructions Executed", "").isdigit():
            lines.append((cur_file, int(r[0]), r[1].strip(), d))
tot_s = sum(int(d["# Samples"]) for *_, d in lines) or 1
tot_i = sum(int(d["Instructions Executed"]) for *_, d in lines) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(d.get(k, 0) or 0) for *_, d in lines) for k in stall_keys}
print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
lines.sort(key=lambda t: -int(t[3]["# Samples"]))
for f, ln, src, d in lines[:top]:
    st = sorted(((int(d.get(k, 0) or 0), k[6:]) for k in stall_keys), reverse=True)[:2]
    print(f"{100 * int(d['# Samples']) / tot_s:5.1f}% smp {100 * int(d['Instructions Executed']) / tot_i:5.1f}% ins "
          f"xs={d.get('L1 Wavefronts Shared Excessive', '0'):>10} {f}:{ln:<4} [{st[0][1]},{st[1][1]}] {src[:110]}")
