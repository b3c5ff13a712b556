#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list of ONE forward pass of tools/prof_one.py (128 ten-second
utterances = one chunk of 131072 activation rows):

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,\
l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
      -s <launches of the warm-up passes> -c <launches of one pass> --csv --log-file launches.csv \
      python tools/prof_one.py

usage: make_traffic.py launches.csv precision rows_per_pass out.json   (merges into out.json)"""
import csv
import json
import os
import sys


def main():
    path, prec, rows, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    lines = [l for l in open(path) if l.startswith('"')]
    rd = list(csv.DictReader(lines))
    per = {}
    for r in rd:
        full = r["Kernel Name"].split("(")[0]
        k = full.split("<")[0].split("::")[-1]
        if k == "gemm_kernel" and "<" in full:             # gemm_kernel<KIND, CG, GRAN, MODE>: MODE 1 = the output
            targs = full.split("<", 1)[1].rstrip(">").split(",")   # layer fused with LogSoftmax + prior + argmax
            if len(targs) >= 4 and targs[3].strip() == "1":
                k = "gemm_kernel_lsm"
        d = per.setdefault((r["ID"], k), {})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3, "%": 1.0}.get(unit, 1.0)
        d[r["Metric Name"]] = v * mult
    cats = {}
    for (_, k), d in per.items():
        c = cats.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "pipe": []})
        c["launches"] += 1
        c["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        c["us"] += d.get("gpu__time_duration.sum", 0.0)
        p = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        if p:
            c["pipe"].append((d.get("gpu__time_duration.sum", 0.0), p))
    tr = json.load(open(out)) if os.path.exists(out) else {}
    total = sum(c["dram_bytes"] for c in cats.values())
    tr["step_%s_dram_bytes_per_row" % prec] = round(total / rows, 1)
    tr["kernels_%s" % prec] = {k: {"launches": c["launches"], "dram_bytes": int(c["dram_bytes"]), "us": round(c["us"], 1)}
                               for k, c in sorted(cats.items())}
    # the fbank kernel is bound by issue slots and shared-memory wavefronts, not by HBM
    for (_, k), d in per.items():
        if k == "fbank_kernel":
            tr["fbank_issue_active_pct"] = round(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0), 2)
            tr["fbank_l1tex_throughput_pct"] = round(d.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 0.0), 2)
            tr["fbank_us_per_launch"] = round(d.get("gpu__time_duration.sum", 0.0), 1)
    g = cats.get("gemm_kernel")
    if g:
        tr["gemm_%s_dram_bytes_per_row" % prec] = round(g["dram_bytes"] / g["launches"] / rows, 1)
        tr["gemm_%s_dram_bytes_per_launch" % prec] = int(g["dram_bytes"] / g["launches"])
        t = sum(x for x, _ in g["pipe"])
        if t > 0:   # time-weighted over the launches of the pass
            tr["gemm_%s_tensor_pipe_active_pct" % prec] = round(sum(x * p for x, p in g["pipe"]) / t, 2)
    tr["note"] = ("ncu (gpu__time_duration, dram__bytes_read + dram__bytes_write, sm__pipe_tensor_cycles_active) of every "
                  "launch of one forward pass over 128 ten-second utterances (131072 activation rows, one chunk; "
                  "tools/prof_one.py, tools/make_traffic.py); per-launch times are cold-cache and serialised. "
                  "bench.py scales the per-row figures to the rows of its own chunking.")
    tr["rows_per_pass"] = rows
    json.dump(tr, open(out, "w"), indent=1)
    print(json.dumps(tr, indent=1))


if __name__ == "__main__":
    main()
