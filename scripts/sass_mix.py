#!/usr/bin/env python
"""Static SASS instruction mix of one kernel, attributed to source lines (needs -lineinfo).

  python scripts/sass_mix.py <file.o|.so|.cubin> <kernel-name-substring> [first_line last_line [source-file-substring]]

Prints the opcode histogram of the instructions whose line-info falls inside [first_line, last_line] (whole kernel
when omitted) and the split between the two half-rate integer pipes of sm_100 (B300_MICROARCH.md: IMAD / IDP on the
FMA-heavy pipe, IADD3 / LEA / SHF / LOP3 / PRMT / I2IP / ISETP / SEL on the ALU pipe).  Used to compare kernel
variants without a GPU; the numbers quoted in DESIGN.md come from it.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

FMA = {"IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
ALU = {"IADD3", "IADD", "LEA", "SHF", "LOP3", "PRMT", "I2IP", "ISETP", "SEL", "IMNMX", "VIMNMX", "VIMNMX3", "MOV", "IABS", "FMNMX", "PLOP3", "FSETP", "VIADD", "SGXT", "BMSK", "FLO", "POPC"}


def cubin_of(path):
    if path.endswith(".cubin"):
        return path
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, stdout=subprocess.DEVNULL)
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cubin")]


def main():
    path, kern = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1 << 30)
    srcsub = sys.argv[5] if len(sys.argv) > 5 else None
    cubins = cubin_of(path)
    cubins = cubins if isinstance(cubins, list) else [cubins]
    for cb in cubins:
        dis = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        cur_fn, cur_line, cur_file = None, 0, ""
        hist = collections.Counter()
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_file, cur_line = m.group(1), int(m.group(2))
                continue
            if cur_fn is None or kern not in cur_fn:
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
            if not m:
                continue
            if lo <= cur_line <= hi and (srcsub is None or srcsub in cur_file):
                hist[m.group(1)] += 1
        if not hist:
            continue
        tot = sum(hist.values())
        fma = sum(v for k, v in hist.items() if k in FMA)
        alu = sum(v for k, v in hist.items() if k in ALU)
        print("%s: %d instructions, FMA-pipe %d, ALU-pipe %d, other %d" % (os.path.basename(cb), tot, fma, alu, tot - fma - alu))
        print("  " + "  ".join("%s %d" % kv for kv in hist.most_common(24)))


if __name__ == "__main__":
    main()
