#!/usr/bin/env python
"""Pruning-kernel geometry sweep on the config-5 workload: one context per CAFE_B200_GEOM setting ("groups,chunks per
stage,producer warps[,ring stages]"), kernel time from the library's CUDA events, fraction of the FP64 DMMA peak.

    python scripts/geom_sweep.py --families 262144 --geoms "3,2,2 2,2,2 2,4,2"
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cafexp_b200 import engine, synth  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=262144)
    ap.add_argument("--geoms", default="default")
    ap.add_argument("--evals", type=int, default=3)
    ap.add_argument("--peak", type=float, default=37.085)
    args = ap.parse_args()
    tree, counts, _ = synth.config5(args.families, bench.N_LEAVES, bench.SEED, bench.LAMBDA, first=0, last=args.families)
    counts = counts.astype(np.uint8)
    freq, rate, prior = bench.gamma_parameters()
    lams = np.ascontiguousarray(rate[:, None] * np.array([[bench.LAMBDA]]))
    flops = bench.algorithmic_flops_per_family_category(tree, bench.MF, bench.MRF) * bench.K * args.families
    for spec in args.geoms.split():
        geom = spec
        if geom == "default":
            os.environ.pop("CAFE_B200_GEOM", None)
        else:
            os.environ["CAFE_B200_GEOM"] = geom
        try:
            with engine.Engine(tree, counts, bench.MF, bench.MRF) as eng:
                best = 1e30
                for _ in range(args.evals):
                    res = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM, want_family=False, want_cat=False)
                    tm = eng.last_timings_ms()
                    best = min(best, tm["prune"])
                print(json.dumps({"geom": spec, "prune_ms": round(best, 3), "build_ms": round(tm["matrix_build"], 3), "tflops": round(flops / best / 1e9, 3),
                                  "frac": round(flops / best / 1e9 / args.peak, 4), "neg_lnl": res["score"], "describe": eng.describe()}), flush=True)
        except Exception as e:   # noqa: BLE001
            print(json.dumps({"geom": spec, "error": str(e)}), flush=True)


if __name__ == "__main__":
    main()
