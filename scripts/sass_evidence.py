#!/usr/bin/env python
"""SASS opcode evidence for the shipped library: per kernel of cafexp_b200/libcafe_b200.so, how many instructions of the
families that prove the sm_100a features used — DMMA (FP64 tensor pipe), UBLKCP (1-D bulk copy through the TMA engine),
SYNCS (mbarrier), LDTM / STTM (tensor-memory loads / stores), UTCATOMSWS / UVIRTCOUNT (tensor-memory allocation),
USETMAXREG (warpgroup register re-allocation), BAR (named barriers).  Writes profiles/<round>_sass_opcodes.txt.

    python scripts/sass_evidence.py r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = ["DMMA", "UBLKCP", "SYNCS", "LDTM", "STTM", "USETMAXREG", "UTCATOMSWS", "UVIRTCOUNT", "BAR", "DMUL", "DSETP", "DADD", "LDL", "STL",
            "LDS", "STS", "LDG", "STG", "REDUX", "SHFL", "MUFU"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lib = os.path.join(ROOT, "cafexp_b200", "libcafe_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = name.replace("cafe::", "").replace("(cafe::PruneParams)", "").replace("(cafe::PupkoParams)", "")
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            per[name]["total"] += 1
            for fam in FAMILIES:
                if op.startswith(fam):
                    per[name][fam] += 1
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.txt")
    with open(out, "w") as fh:
        fh.write("cuobjdump -sass cafexp_b200/libcafe_b200.so (sm_100a), static instruction counts per kernel\n")
        fh.write("DMMA = mma.sync.m8n8k4.f64 (FP64 tensor pipe; tcgen05.mma has no FP64 kind), UBLKCP = cp.async.bulk (TMA engine, 1-D),\n")
        fh.write("SYNCS = mbarrier, LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTCATOMSWS / UVIRTCOUNT = tcgen05.alloc / dealloc,\n")
        fh.write("USETMAXREG = setmaxnreg (warpgroup register re-allocation), LDL / STL = local-memory spills\n\n")
        cols = ["total"] + FAMILIES
        fh.write(f"{'kernel':58s}" + "".join(f"{c[:7]:>8s}" for c in cols) + "\n")
        for k, cnt in per.items():
            short = k if len(k) <= 57 else k[:54] + "..."
            fh.write(f"{short:58s}" + "".join(f"{cnt.get(c, 0):8d}" for c in cols) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
