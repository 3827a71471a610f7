#!/usr/bin/env python
"""Golden vectors of BASELINE.json config 5 from the compiled, UNMODIFIED reference (oracle/_ref/ref_harness):

    tests/golden/config5_slice.npz
        cat_lk   [10000][4] float64   gamma_model::_category_likelihoods of the first 10 000 synthetic families
                                      (lambda 0.005, alpha 0.7, k = 4, Nmax 150, root max 125, uniform prior)
        score    float64              -lnL the reference returns for that slice
        states   [1000][4][99] uint8  reconstruct_ancestral_states of the first 1 000 of them (Pupko, per category)
        meta     json                 generator arguments and the reference's node order

The inputs are not stored: cafexp_b200.synth.config5 regenerates them from the seed (any first/last range of the same
global data set).  Run in the build container (needs /root/reference compiled by oracle/Makefile); takes a few minutes
of CPU.  SURVEY section 8d asks for exactly this slice.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cafexp_b200 import hostio, synth  # noqa: E402
from oracle import binding as orc  # noqa: E402

N_EVAL, N_RECON = 10000, 1000
LAMBDA, ALPHA, K = 0.005, 0.7, 4
MF, MRF = synth.CONFIG5_MAX_FAMILY_SIZE, synth.CONFIG5_MAX_ROOT_FAMILY_SIZE


def main():
    assert orc.have_ref(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    tree, counts, newick = synth.config5(1_000_000, first=0, last=N_EVAL)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        tpath, fpath, rpath = (os.path.join(tmp, n) for n in ("tree.txt", "fam.txt", "fam_recon.txt"))
        open(tpath, "w").write(newick + "\n")
        hostio.write_gene_families(fpath, tree, [str(i) for i in range(N_EVAL)], counts)
        hostio.write_gene_families(rpath, tree, [str(i) for i in range(N_RECON)], counts[:N_RECON])
        common = dict(tree=tpath, filter=0, k=K, alpha=ALPHA, maxfam=MF, maxroot=MRF, **{"lambda": LAMBDA})
        t0 = time.time()
        dump = os.path.join(tmp, "cat.bin")
        r = orc.run_ref("eval", fam=fpath, dump=dump, **common)
        cat = np.fromfile(dump, np.float64).reshape(N_EVAL, K)
        print(f"eval: {time.time() - t0:.1f} s, score {r['score']!r}, threads {r['threads']}", flush=True)
        t0 = time.time()
        dump_r = os.path.join(tmp, "rec.bin")
        rr = orc.run_ref("eval", fam=rpath, recon=True, dumprecon=dump_r, dump=os.path.join(tmp, "cat2.bin"), **common)
        n_internal = len(rr["internal_order"])
        states = np.fromfile(dump_r, np.int32).reshape(N_RECON, K, n_internal)
        print(f"reconstruction: {time.time() - t0:.1f} s ({rr['recon_seconds']:.1f} s in reconstruct_ancestral_states)", flush=True)
        assert states.min() >= 0 and states.max() <= 255
        # the reference's node order must be the flattened tree's
        flat_internal = tree.internal_names
        out = dict(cat_lk=cat, score=np.float64(float(r["score"])), states=states.astype(np.uint8),
                   meta=json.dumps({"n_eval": N_EVAL, "n_recon": N_RECON, "lambda": LAMBDA, "alpha": ALPHA, "k": K, "max_family_size": MF,
                                    "max_root_family_size": MRF, "seed": 12345, "families_total": 1_000_000, "multipliers": r["multipliers"],
                                    "cat_probs": r["cat_probs"], "internal_order": rr["internal_order"], "node_order": r["node_order"],
                                    "flat_internal_names": flat_internal, "generator": "scripts/make_golden_config5.py"}))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "config5_slice.npz"), **out)
    print("wrote tests/golden/config5_slice.npz")


if __name__ == "__main__":
    main()
