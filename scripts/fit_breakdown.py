#!/usr/bin/env python
"""Where does a fit's wall time go?  Runs the reference's optimizer over the CUDA drop-in (oracle/_ref/ref_harness_cuda)
on the mammal set and on a config-5 slice and prints the harness' time breakdown (first evaluation, bind, library staging /
enqueue / wait, device time).  Usage: python scripts/fit_breakdown.py [--devices 0,1] [--slice 65536]"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from cafexp_b200 import hostio, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0", help='device lists separated by ";", e.g. "0;0,1;0,1,2,3"')
    ap.add_argument("--slice", type=int, default=65536)
    ap.add_argument("--quick", action="store_true", help="only the mammal gamma fit and the config-5 slice fit")
    args = ap.parse_args()
    device_lists = [[int(d) for d in part.split(",")] for part in args.devices.split(";")]
    inp = json.load(open(os.path.join(ROOT, "tests", "golden", "mammal_inputs.json")))
    with tempfile.TemporaryDirectory() as tmp:
        tpath, fpath = os.path.join(tmp, "t.txt"), os.path.join(tmp, "f.txt")
        open(tpath, "w").write(inp["tree"] + "\n")
        flat = hostio.flatten_tree(hostio.parse_newick(inp["tree"]))
        z = np.load(os.path.join(ROOT, "tests", "golden", "mammal_counts.npz"))
        hostio.write_gene_families(fpath, flat, [str(i) for i in z["ids"]], z["counts"])
        cases = (("mammal gamma k=4", {"k": 4}),) if args.quick else (
            ("mammal single lambda", {}), ("mammal single lambda (again)", {}), ("mammal gamma k=4", {"k": 4}), ("mammal lambda+epsilon", {"esterr": 1}))
        t5 = f5 = None
        if args.slice > 0:
            tree5, counts5, newick5 = synth.config5(1_000_000, first=0, last=args.slice)
            t5, f5 = os.path.join(tmp, "t5.txt"), os.path.join(tmp, "f5.txt")
            open(t5, "w").write(newick5 + "\n")
            bench.write_family_table(f5, tree5, counts5)
        for devices in device_lists:
            for name, kw in cases:
                r = bench.harness_fit(True, devices, 600, tree=tpath, fam=fpath, **kw)
                print(json.dumps({"fit": name, "n_devices": len(devices), **{k: v for k, v in r.items() if k not in ("node_order",)}}), flush=True)
            if t5:
                r = bench.harness_fit(True, devices, 900, tree=t5, fam=f5, k=4, filter=0, maxfam=bench.MF, maxroot=bench.MRF)
                print(json.dumps({"fit": f"config-5 slice ({args.slice} families) gamma k=4", "n_devices": len(devices),
                                  **{k: v for k, v in r.items() if k not in ("node_order",)}}), flush=True)


if __name__ == "__main__":
    main()
