#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump (SASS view): stall samples per opcode and the hottest instructions.

    ncu -i gpurun_out/x.ncu-rep --page source --csv > /tmp/x.csv;  python scripts/ncu_hotspots.py /tmp/x.csv [top]
"""
import csv
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    samp = lambda r: int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    total = sum(samp(r) for r in body)
    print("instructions", len(body), "samples", total)
    reasons = Counter()
    for r in body:
        for s in stall_cols:
            reasons[s] += int(r[col[s]] or 0)
    print("by reason:", ", ".join(f"{k[6:]}={v} ({100*v/total:.1f}%)" for k, v in reasons.most_common(10)))
    byop = Counter(); cnt = Counter(); execd = Counter()
    for r in body:
        op = r[col["Source"]].split()[0] if not r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[1]
        op = op.split(".")[0]
        byop[op] += samp(r); cnt[op] += 1; execd[op] += int(r[col["Instructions Executed"]] or 0)
    print("by opcode (samples, static count, executed):")
    for op, v in byop.most_common(16):
        print(f"  {op:10s} {v:8d} {100*v/total:5.1f}%  n={cnt[op]:4d} exec={execd[op]}")
    print("hottest instructions:")
    order = sorted(range(len(body)), key=lambda i: -samp(body[i]))[:top]
    for i in sorted(order):
        r = body[i]
        why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:3]
        print(f"  [{i:5d}] {samp(r):7d} {100*samp(r)/total:5.1f}%  {r[col['Source']].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in why if c))


if __name__ == "__main__":
    main()
