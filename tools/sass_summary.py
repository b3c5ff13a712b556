#!/usr/bin/env python
"""SASS opcode counts per kernel of catears_b200/libce_gpu.so (cuobjdump -sass): the evidence that the GEMMs are
tcgen05 / TMEM / TMA code (UTCIMMA, UTCHMMA, LDTM, UTMALDG, UTMASTG) and that nothing uses mma.sync.
usage: sass_summary.py [lib.so] > profiles/<tag>_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

WANT = ["UTCIMMA", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "IMMA",
        "FADD2", "FFMA2", "FMUL2", "MUFU", "I2FP", "F2FP", "FMNMX3", "LDG", "STG", "LDS", "STS", "SHFL", "RED", "REDUX", "LDL", "STL"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "catears_b200/libce_gpu.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, per = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            per[name]["_n"] += 1
            per[name][m.group(1)] += 1
            if ".2CTA" in m.group(2):
                per[name]["_2cta"] += 1
    print("SASS opcode counts per kernel of %s (cuobjdump -sass, sm_100a)." % lib)
    print("tcgen05 = UTCIMMA (kind::i8) / UTCHMMA (kind::f16, kind::tf32), .2CTA forms in the cta_group::2 instantiations;")
    print("TMEM = LDTM; TMA = UTMALDG / UTMASTG; mbarrier = SYNCS; commit = UTCBAR; packed fp32 = FADD2 / FFMA2.")
    print("gemm_kernel<KIND, CG, GRAN, LSM>: LSM = the output layer fused with LogSoftmax + prior + argmax.\n")
    for k, c in per.items():
        parts = ["%s %d" % (w, c[w]) for w in WANT if c[w]]
        if c["_2cta"]:
            parts.append(".2CTA forms %d" % c["_2cta"])
        print("%s\n   %d instructions; %s" % (k, c["_n"], ", ".join(parts)))


if __name__ == "__main__":
    main()
