#!/usr/bin/env python
"""Top warp-stall sampling sites per kernel from `ncu -i rep --page source --csv`."""
import csv
import subprocess
import sys


def main():
    rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for k in kernels:
        ix = {h: i for i, h in enumerate(k["hdr"])}
        smp = [int(r[ix["# Samples"]] or 0) for r in k["data"]]
        print("== %s: %d SASS lines, %d samples" % (k["name"][:80], len(smp), sum(smp)))
        order = sorted(range(len(smp)), key=lambda i: -smp[i])[:n]
        for i in sorted(order):
            print("  %5d %6d  %s" % (i, smp[i], k["data"][i][ix["Source"]][:100]))


if __name__ == "__main__":
    main()
