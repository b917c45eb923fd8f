#!/usr/bin/env python
"""Opcode histogram per kernel of libsnacb.so (cuobjdump -sass), the evidence that the contractions are tcgen05 / TMEM /
TMA code and not mma.sync in disguise:   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tts_inference_b200", "libsnacb.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU", "HFMA2", "FFMA", "FFMA2", "LDS",
       "STS", "LDG", "STG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("snacb::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    tot = collections.Counter()
    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)} -- SASS opcodes per kernel (tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM,")
    print("cp.async.bulk.tensor = UTMALDG/UTMASTG, cp.async.bulk = UBLKCP, mbarrier = SYNCS; HMMA would be mma.sync: there is none)")
    print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEY))
    for name, c in kernels.items():
        n = sum(c.values())
        print(f"{name[:70]:70s} {n:7d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEY))
        tot.update(c)
    print(f"{'TOTAL':70s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(k, 0):7d}" for k in KEY))


if __name__ == "__main__":
    main()
