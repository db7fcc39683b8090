#!/usr/bin/env python
"""Share of stall samples / executed instructions per code section of nf_render.cu in an ncu report.

    python tools/ncu_sections.py gpurun_out/x.ncu-rep
"""
import csv, subprocess, sys, os
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; L = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0].isdigit() and len(r) > 10:
        d = dict(zip(hdr[4:], r[4:]))
        if not d.get("# Samples", "").isdigit(): d = dict(zip(hdr[::-1], r[::-1]))
        if d.get("# Samples", "").isdigit(): L.append((cur, int(r[0]), int(d["# Samples"]), int(d["Instructions Executed"])))
ts = sum(x[2] for x in L) or 1; ti = sum(x[3] for x in L) or 1
src = open(os.path.join(ROOT, "neurofluid_b200/csrc/nf_render.cu")).read().split("\n")
def find(t):
    return next(i + 1 for i, l in enumerate(src) if t in l)
marks = sorted([("search_stream", find("int search_stream(")), ("scs_head", find("int search_scs(")), ("scs_gather", find("// ---- 1. gather")),
         ("scs_sweep", find("// ---- 2 + 3. walk")), ("scs_enumerate", find("if (fresh || lst_word < 0) { lst_len = 0; lst_word = 0; }")),
         ("group_head", find("void ray_query_group(")), ("geometry", find("// ---- per-lane local geometry")),
         ("record", find("const bool full = in &&")), ("composite", find("void ray_composite(")), ("q0", find("// stage Q0")),
         ("mid_head", find("// stage MID")), ("pdf", find("// ---------------- sample_pdf")), ("invcdf", find("// inverse CDF")),
         ("merge", find("// rank merge")), ("fin", find("// stage FIN"))], key=lambda m: m[1])
agg = {}
for f, ln, s, i in L:
    key = "other:" + f
    if f == "nf_render.cu":
        key = "pre"
        for name, start in marks:
            if ln >= start: key = name
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += i
print(f"total samples {ts}, instructions {ti}")
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:34s} samples {100*s/ts:5.1f}%  instr {100*i/ti:5.1f}%")
