#!/usr/bin/env python
"""Small driver for ncu: one config-5-shaped evaluation (+ optional reconstruction) so every kernel launches once or twice.

    ncu --set full --import-source on -k regex:prune_kernel -c 1 python scripts/profile_run.py --families 131072
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cafexp_b200 import engine, synth  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=131072)
    ap.add_argument("--evals", type=int, default=2)
    ap.add_argument("--recon", type=int, default=0, help="families to reconstruct (0 = skip)")
    args = ap.parse_args()
    n_gen = max(args.families, args.recon)
    tree, counts, _ = synth.config5(n_gen, bench.N_LEAVES, bench.SEED, bench.LAMBDA, first=0, last=n_gen)
    counts = counts.astype(np.uint8)
    freq, rate, prior = bench.gamma_parameters()
    lams = np.ascontiguousarray(rate[:, None] * np.array([[bench.LAMBDA]]))
    with engine.Engine(tree, counts[: args.families], bench.MF, bench.MRF) as eng:
        for _ in range(args.evals):
            res = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM, want_family=False, want_cat=False)
        print("score", res["score"], eng.last_timings_ms())
    if args.recon:
        from cafexp_b200 import params
        prior_sz = params.prior_uniform(bench.MRF, None, min(bench.MF, bench.MRF) + 1)
        with engine.Engine(tree, counts[: args.recon], bench.MF, bench.MRF) as eng:
            for _ in range(2):
                eng.reconstruct(lams, prior_sz)
            t = eng.last_timings_ms()["reconstruct"]
            pairs = bench.pupko_pairs_per_family_category(tree, bench.MF) * args.recon * bench.K
            print("reconstruct", args.recon, "families:", round(t, 3), "ms,", round(pairs / t / 1e9, 3), "T pairs/s")


if __name__ == "__main__":
    main()
