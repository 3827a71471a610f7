#!/usr/bin/env python
"""Generate tests/golden/* by running the compiled, UNMODIFIED reference (oracle/_ref/ref_harness,
built by `make -C oracle ref` from /root/reference) on the reference's own example inputs and on
the fixtures of its unit tests (test.cpp).  Runs only in the build container (needs /root/reference);
the outputs are committed so the GPU box and the CPU test-suite never need the reference.

    python scripts/make_golden.py            # everything except the long fits
    python scripts/make_golden.py --fits     # also the seed-10 Nelder-Mead fits (minutes..)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cafexp_b200 import hostio  # noqa: E402
from oracle import binding as orc  # noqa: E402

REF = "/root/reference"
EX = os.path.join(REF, "examples")
GOLD = os.path.join(ROOT, "tests", "golden")


def read_bin(path, dtype, shape=None):
    a = np.fromfile(path, dtype=dtype)
    return a.reshape(shape) if shape is not None else a


def mammal_inputs():
    """Inputs of BASELINE.json configs 1-4, stored compactly (data, not code)."""
    tree_text = open(os.path.join(EX, "mammals_tree.txt")).readline().strip()
    ltree_text = open(os.path.join(EX, "chimphuman_separate_lambda.txt")).readline().strip()
    err_text = open(os.path.join(EX, "errormodel_0.1.txt")).read()
    rootdist_text = open(os.path.join(EX, "poisson_root_dist_1000.txt")).read()
    flat = hostio.flatten_tree(hostio.parse_newick(tree_text))
    ids, counts = hostio.read_gene_families(os.path.join(EX, "mammal_gene_families.txt"), flat)
    assert counts.max() < 256
    with open(os.path.join(GOLD, "mammal_inputs.json"), "w") as fh:
        json.dump({"tree": tree_text, "lambda_tree": ltree_text, "error_model": err_text, "rootdist": rootdist_text,
                   "leaf_names": flat.leaf_names,
                   "source": "examples/{mammals_tree,chimphuman_separate_lambda,errormodel_0.1,poisson_root_dist_1000}.txt"}, fh, indent=1)
    np.savez_compressed(os.path.join(GOLD, "mammal_counts.npz"), counts=counts.astype(np.uint8),
                        ids=np.asarray([int(i) for i in ids], np.int32))
    return flat, ids, counts


def scalars():
    out = {"bd": [], "bdlog": [], "gamma": [], "poisson": [], "matrix_small": []}
    # Probability/probability_of_some_values (test.cpp:601), the_probability_of_going... (:641) and more
    for lam, t, s, c in [(0.05, 5, 5, 9), (0.05, 5, 5, 10), (0.05, 5, 5, 5), (0.05, 1, 10, 8), (0.05, 5, 5, 8),
                         (0.006335, 68.7105, 5, 5), (0.006335, 68, 5, 5), (0.002, 96.435575, 40, 46), (0.01, 132.0, 90, 3),
                         (0.5, 3.0, 4, 4), (0.2, 2.5, 3, 7), (0.0, 10.0, 3, 3), (0.001, 0.0005, 2, 2)]:
        r = orc.run_ref("bd", **{"lambda": float(lam), "t": float(t), "s": s, "c": c})
        out["bd"].append({"lambda": lam, "t": t, "s": s, "c": c, "p": r["p"]})
    # Inference/birthdeath_rate_with_log_alpha (test.cpp:1287)
    for s, c, la, co in [(40, 42, -1.37, 0.5), (41, 34, -1.262, 0.4), (40, 42, -1.37, 0.5), (5, 5, -0.3, 0.1), (140, 140, -2.0, 0.7),
                         (1, 0, -1.0, 0.3), (3, 150, -0.7, 0.2)]:
        r = orc.run_ref("bdlog", s=s, c=c, logalpha=float(la), coeff=float(co))
        out["bdlog"].append({"s": s, "c": c, "logalpha": la, "coeff": co, "p": r["p"]})
    for k in (2, 3, 4, 5, 8):
        for alpha in (0.05, 0.25, 0.425, 0.481515908358985, 0.5, 0.7, 1.0, 2.5, 10.0, 55.5):
            r = orc.run_ref("gamma", alpha=float(alpha), k=k)
            out["gamma"].append({"k": k, "alpha": alpha, "rate": r["rate"], "freq": r["freq"]})
    for lp, n in [(10.0, 30), (0.75, 112), (50.1, 125)]:
        r = orc.run_ref("poisson", **{"lambda": float(lp), "n": n})
        out["poisson"].append({"lambda": lp, "n": n, "prior": r["prior"]})
    # Probability/probability_of_matrix (test.cpp:646) 5x5 and friends, incl. a saturated one
    for n, lam, t in [(5, 0.05, 5.0), (8, 0.01, 1.0), (12, 0.3, 2.0), (6, 0.3, 4.0)]:
        r = orc.run_ref("matrix", n=n, **{"lambda": float(lam), "t": float(t)})
        out["matrix_small"].append({"n": n, "lambda": lam, "t": t, "lambda_q": r["lambda_q"], "t_q": r["t_q"],
                                    "saturated": r["saturated"], "m": r["m"]})
    with open(os.path.join(GOLD, "scalars.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def matrices(tmp):
    """Full-size transition matrices (mammal N=141 and synthetic N=151) incl. the quantisation case."""
    # (n, lambda, t, rows kept) — rows=None keeps the full matrix; the rest keep a few rows to stay small
    cases = [(141, 0.006335, 68.7105, None), (151, 0.005, 12.345, None), (31, 0.01, 1.0, None),
             (141, 0.002, 68.710507, [0, 1, 7, 70, 140]), (141, 0.00354641825220246 * 3.2, 96.435575, [1, 33, 90, 139]),
             (151, 0.0139, 35.9, [1, 2, 75, 150]), (151, 0.05, 12.0, [0, 1, 150])]
    store = {}
    meta = []
    for i, (n, lam, t, rows) in enumerate(cases):
        path = os.path.join(tmp, f"m{i}.bin")
        r = orc.run_ref("matrix", n=n, dump=path, **{"lambda": float(lam), "t": float(t)})
        full = read_bin(path, np.float64, (n, n))
        store[f"m{i}"] = full if rows is None else full[rows]
        meta.append({"key": f"m{i}", "n": n, "lambda": lam, "t": t, "rows": rows, "lambda_q": r["lambda_q"], "t_q": r["t_q"],
                     "saturated": r["saturated"], "total": float(full.sum())})
    np.savez_compressed(os.path.join(GOLD, "matrices.npz"), meta=json.dumps(meta), **store)


def write_fixture(tmp, name, newick, species, rows, ltree=None):
    tpath = os.path.join(tmp, name + "_tree.txt")
    fpath = os.path.join(tmp, name + "_fam.txt")
    open(tpath, "w").write(newick + "\n")
    with open(fpath, "w") as fh:
        fh.write("Desc\tFamily ID\t" + "\t".join(species) + "\n")
        for i, row in enumerate(rows):
            fh.write("(null)\t" + str(i) + "\t" + "\t".join(str(v) for v in row) + "\n")
    out = {"tree": tpath, "fam": fpath}
    if ltree:
        lpath = os.path.join(tmp, name + "_ltree.txt")
        open(lpath, "w").write(ltree + "\n")
        out["ltree"] = lpath
    return out


def unit_fixtures(tmp):
    """The reference's own unit-test fixtures (test.cpp), re-run through the compiled reference at 17 digits."""
    out = []

    def add(name, newick, species, rows, cmd, ltree=None, err_text=None, **kw):
        paths = write_fixture(tmp, name, newick, species, rows, ltree)
        if err_text:
            epath = os.path.join(tmp, name + "_err.txt")
            open(epath, "w").write(err_text)
            paths["err"] = epath
        dump = os.path.join(tmp, name + ".bin")
        r = orc.run_ref(cmd, dump=dump, filter=0, **paths, **kw)
        rec = {"name": name, "cmd": cmd, "newick": newick, "lambda_tree": ltree, "species": species, "rows": rows,
               "error_model": err_text, "args": {k: v for k, v in kw.items() if k != "dumprecon"},
               "max_family_size": kw.get("maxfam", r["max_family_size"]), "max_root_family_size": kw.get("maxroot", r["max_root_family_size"]),
               "node_order": r["node_order"]}
        if cmd == "prune":
            rec["root"] = read_bin(dump, np.float64).reshape(len(rows), -1).tolist()
        else:
            rec["score"] = r["score"]
            data = read_bin(dump, np.float64)
            if r["model"] == "Gamma":
                rec["cat_lk"] = data.reshape(len(rows), -1).tolist()
                rec["multipliers"] = r["multipliers"]
                rec["cat_probs"] = r["cat_probs"]
            else:
                rec["family_lnl"] = data.tolist()
            if "recon" in kw:
                rec["internal_order"] = r["internal_order"]
                rec["states"] = read_bin(dump + ".rec", np.int32).reshape(len(rows), -1).tolist()
        out.append(rec)

    ab = "(A:1,B:1);"
    abcd = "((A:1,B:1):1,(C:1,D:1):1);"
    rec4 = "((A:1,B:3):7,(C:11,D:17):23);"
    # Inference/prune (test.cpp:1642): ((A,B),(C,D)) lambda 0.03, multiplier 1.5, mf=20 mrf=20, counts 3,6,? ...
    add("prune_abcd", abcd, ["A", "B", "C", "D"], [[3, 6, 6, 3], [1, 2, 0, 4]], "prune", **{"lambda": 0.03, "mult": 1.5, "maxfam": 20, "maxroot": 20})
    add("prune_ab", ab, ["A", "B"], [[1, 2], [0, 5], [7, 7]], "prune", **{"lambda": 0.05, "mult": 1.0, "maxfam": 10, "maxroot": 8})
    # mrf > mf exercises the N = max(mrf, mf) + 1 rule (src/base_model.cpp:77)
    add("prune_ab_bigroot", ab, ["A", "B"], [[1, 2], [3, 3]], "prune", **{"lambda": 0.05, "mult": 1.0, "maxfam": 6, "maxroot": 9})
    # n-ary (3 children) node
    add("prune_tri", "((A:2,B:1.5,C:3):1,D:4);", ["A", "B", "C", "D"], [[2, 3, 1, 2], [0, 0, 1, 1]], "prune", **{"lambda": 0.02, "mult": 1.0, "maxfam": 15, "maxroot": 12})
    # Inference/infer_processes (test.cpp:519): 4 families on (A:1,B:1), lambda 0.01, mf 56, mrf 30
    add("infer_processes", ab, ["A", "B"], [[1, 2], [2, 1], [3, 6], [6, 3]], "eval", **{"lambda": 0.01, "maxfam": 56, "maxroot": 30})
    # Inference/gamma_lambda_optimizer-like evaluation (test.cpp:2240): k=4 alpha=0.25 lambda=0.01
    add("gamma_eval", ab, ["A", "B"], [[1, 2], [2, 1], [3, 6], [6, 3]], "eval", **{"lambda": 0.01, "k": 4, "alpha": 0.25, "maxfam": 56, "maxroot": 30})
    # Inference/gamma_lambda_optimizer (test.cpp:2240): fixture family A=1,B=2, mf=mrf=10 -> 6.4168
    add("gamma_lambda_optimizer", ab, ["A", "B"], [[1, 2]], "eval", **{"lambda": 0.01, "k": 4, "alpha": 0.25, "maxfam": 10, "maxroot": 10})
    add("gamma_eval_abcd", abcd, ["A", "B", "C", "D"], [[3, 6, 6, 3], [1, 2, 0, 4], [5, 5, 5, 5]], "eval", **{"lambda": 0.03, "k": 3, "alpha": 0.7, "maxfam": 25, "maxroot": 20})
    # leaf error model (test.cpp:1745 uses {0.2,0.6,0.2}); here through a file
    err = "maxcnt:10\ncntdiff -1 0 1\n0 0.0 0.8 0.2\n1 0.2 0.6 0.2\n"
    add("errmodel_eval", abcd, ["A", "B", "C", "D"], [[3, 6, 6, 3], [1, 2, 0, 4], [0, 0, 1, 1]], "eval", err_text=err, **{"lambda": 0.03, "maxfam": 25, "maxroot": 20})
    add("errmodel_prune", ab, ["A", "B"], [[1, 2], [0, 5]], "prune", err_text=err, **{"lambda": 0.05, "mult": 1.0, "maxfam": 10, "maxroot": 8})
    # multiple lambdas
    add("multilambda_eval", abcd, ["A", "B", "C", "D"], [[3, 6, 6, 3], [1, 2, 0, 4]], "eval", ltree="((A:1,B:1):1,(C:2,D:2):2);", **{"lambda": "0.03,0.09", "maxfam": 25, "maxroot": 20})
    # Reconstruction fixture (test.cpp:865-887, :1040): ((A:1,B:3):7,(C:11,D:17):23), lambda 0.05
    add("reconstruct_rec4", rec4, ["A", "B", "C", "D"], [[11, 2, 5, 6], [3, 3, 3, 3], [0, 1, 9, 2], [10, 10, 1, 1]], "eval",
        **{"lambda": 0.05, "maxfam": 30, "maxroot": 25, "recon": True, "dumprecon": os.path.join(tmp, "reconstruct_rec4.bin.rec")})
    add("reconstruct_gamma", rec4, ["A", "B", "C", "D"], [[11, 2, 5, 6], [3, 3, 3, 3], [0, 1, 9, 2]], "eval",
        **{"lambda": 0.01, "k": 3, "alpha": 0.6, "maxfam": 30, "maxroot": 25, "recon": True, "dumprecon": os.path.join(tmp, "reconstruct_gamma.bin.rec")})
    add("reconstruct_tri", "((A:2,B:1.5,C:3):1,D:4);", ["A", "B", "C", "D"], [[2, 3, 1, 2], [0, 0, 1, 1], [9, 1, 1, 4]], "eval",
        **{"lambda": 0.02, "maxfam": 15, "maxroot": 12, "recon": True, "dumprecon": os.path.join(tmp, "reconstruct_tri.bin.rec")})
    with open(os.path.join(GOLD, "unit_fixtures.json"), "w") as fh:
        json.dump(out, fh)


def mammal_outputs(tmp, flat, ids, counts):
    E = {"tree": os.path.join(EX, "mammals_tree.txt"), "fam": os.path.join(EX, "mammal_gene_families.txt")}
    err = os.path.join(EX, "errormodel_0.1.txt")
    ltree = os.path.join(EX, "chimphuman_separate_lambda.txt")
    rootdist = os.path.join(EX, "poisson_root_dist_1000.txt")
    keep = hostio.exists_at_root(flat, counts)
    F = int(keep.sum())
    meta = {}
    store = {}

    def run(name, cmd="eval", gamma_k=0, **kw):
        dump = os.path.join(tmp, name + ".bin")
        r = orc.run_ref(cmd, dump=dump, **E, **kw)
        assert r["n_families"] == F, (r["n_families"], F)
        meta[name] = {"args": {k: (os.path.basename(v) if isinstance(v, str) and v.startswith("/") else v) for k, v in kw.items()},
                      "score": r["score"], "max_family_size": r["max_family_size"], "max_root_family_size": r["max_root_family_size"],
                      "node_order": r["node_order"]}
        data = read_bin(dump, np.float64)
        if gamma_k:
            meta[name]["multipliers"] = r["multipliers"]
            meta[name]["cat_probs"] = r["cat_probs"]
            store[name + "_cat_lk"] = data.reshape(F, gamma_k)
        else:
            store[name + "_lnl"] = data
        if "recon" in kw:
            meta[name]["internal_order"] = r["internal_order"]
            rec = read_bin(kw["dumprecon"], np.int32)
            store[name + "_states"] = rec.reshape(F, -1).astype(np.int16)
        print(name, r["score"], flush=True)
        return r

    run("base_l002", **{"lambda": 0.002})
    run("base_err_l002", err=err, **{"lambda": 0.002})
    run("base_err_l01_recon", err=err, recon=True, dumprecon=os.path.join(tmp, "r1.rec"), **{"lambda": 0.01})
    run("base_poisson_l002", poisson=10.0, **{"lambda": 0.002})
    run("gamma4_fit", gamma_k=4, k=4, alpha=0.481515908358985, **{"lambda": 0.00354641825220246})
    run("gamma4_fail", gamma_k=4, k=4, alpha=0.5, **{"lambda": 0.002})
    run("gamma3_recon", gamma_k=3, k=3, alpha=0.425, recon=True, dumprecon=os.path.join(tmp, "r2.rec"), **{"lambda": 0.002})
    run("multi_rootdist", ltree=ltree, rootdist=rootdist, **{"lambda": "0.01,0.05"})
    run("multi_poisson_recon", ltree=ltree, poisson=12.5, recon=True, dumprecon=os.path.join(tmp, "r3.rec"), **{"lambda": "0.01,0.05"})
    np.savez_compressed(os.path.join(GOLD, "mammal_outputs.npz"), meta=json.dumps(meta), keep=np.flatnonzero(keep).astype(np.int32), **store)


def pvalues(tmp):
    """compute_pvalues (src/probability.cpp:411-444) on the first 300 root-filtered mammal families, 50 simulations per
    root size, seed 10: the simulated leaf counts, their likelihoods (unsorted conditional distributions), the observed
    families' likelihoods and the reference's p-values."""
    E = {"tree": os.path.join(EX, "mammals_tree.txt"), "fam": os.path.join(EX, "mammal_gene_families.txt")}
    limit, nsim, seed, lam = 300, 50, 10, 0.002
    dump = os.path.join(tmp, "pv.bin")
    r = orc.run_ref("pvalues", dump=dump, limit=limit, nsim=nsim, seed=seed, **E, **{"lambda": lam})
    assert r["replay_matches"] is True
    mrf, nl, F = r["max_root_family_size"], r["n_leaves"], r["n_families"]
    raw = open(dump, "rb").read()
    o = 0
    sim = np.frombuffer(raw, np.int32, mrf * nsim * nl, o).reshape(mrf * nsim, nl); o += sim.nbytes
    cond = np.frombuffer(raw, np.float64, mrf * nsim, o).reshape(mrf, nsim); o += cond.nbytes
    obs = np.frombuffer(raw, np.float64, F, o); o += obs.nbytes
    pv = np.frombuffer(raw, np.float64, F, o); o += pv.nbytes
    assert o == len(raw) and sim.max() < 256
    meta = {"limit": limit, "nsim": nsim, "seed": seed, "lambda": lam, "max_family_size": r["max_family_size"],
            "max_root_family_size": mrf, "leaf_order": r["leaf_order"]}
    np.savez_compressed(os.path.join(GOLD, "mammal_pvalues.npz"), meta=json.dumps(meta), sim_counts=sim.astype(np.uint8), cond=cond,
                        observed=obs, pvalues=pv)
    print("pvalues", F, "families,", mrf * nsim, "simulated;", np.unique(pv).size, "distinct p-values", flush=True)


def viterbi(tmp):
    """compute_viterbi_sum (src/gene_family_reconstructor.cpp:361-400) for every node of the first 400 root-filtered mammal
    families after a base-model reconstruction at lambda = 0.002: reconstructed sizes of all nodes and the probabilities."""
    E = {"tree": os.path.join(EX, "mammals_tree.txt"), "fam": os.path.join(EX, "mammal_gene_families.txt")}
    limit, lam = 400, 0.002
    dump = os.path.join(tmp, "vit.bin")
    r = orc.run_ref("eval", limit=limit, recon=True, viterbi=True, dumprecon=os.path.join(tmp, "vit.rec"), dumpviterbi=dump, **E, **{"lambda": lam})
    F, nn = r["n_families"], len(r["node_order"])
    raw = open(dump, "rb").read()
    sizes = np.frombuffer(raw, np.int32, F * nn, 0).reshape(F, nn)
    probs = np.frombuffer(raw, np.float64, F * nn, sizes.nbytes).reshape(F, nn)
    assert sizes.nbytes + probs.nbytes == len(raw) and sizes.max() < 32768
    meta = {"limit": limit, "lambda": lam, "max_family_size": r["max_family_size"], "max_root_family_size": r["max_root_family_size"],
            "node_order": r["node_order"]}
    np.savez_compressed(os.path.join(GOLD, "mammal_viterbi.npz"), meta=json.dumps(meta), node_sizes=sizes.astype(np.int16), probabilities=probs)
    print("viterbi", F, "families,", int((probs >= 0).sum()), "valid branch probabilities", flush=True)


def fits():
    E = {"tree": os.path.join(EX, "mammals_tree.txt"), "fam": os.path.join(EX, "mammal_gene_families.txt")}
    path = os.path.join(GOLD, "fits.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    jobs = {
        "single_lambda_seed10": dict(seed=10),
        "lambda_epsilon_seed10": dict(seed=10, err=os.path.join(EX, "errormodel_0.1.txt")),
        "two_lambda_seed10": dict(seed=10, ltree=os.path.join(EX, "chimphuman_separate_lambda.txt")),
        "lambda_estimated_epsilon_seed10": dict(seed=10, esterr=1),
        "gamma4_lambda_alpha_seed10": dict(seed=10, k=4),
    }
    for name, kw in jobs.items():
        if name in out:
            continue
        r = orc.run_ref("fit", **E, **kw)
        r.pop("node_order", None)
        out[name] = r
        print(name, r, flush=True)
        json.dump(out, open(path, "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fits", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if not orc.have_ref():
        raise SystemExit("oracle/_ref/ref_harness missing: run `make -C oracle ref` (needs /root/reference)")
    os.makedirs(GOLD, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        steps = args.only.split(",") if args.only else ["inputs", "scalars", "matrices", "unit", "mammal", "pvalues", "viterbi"]
        flat = ids = counts = None
        if "inputs" in steps or "mammal" in steps:
            flat, ids, counts = mammal_inputs()
        if "scalars" in steps:
            scalars()
        if "matrices" in steps:
            matrices(tmp)
        if "unit" in steps:
            unit_fixtures(tmp)
        if "mammal" in steps:
            mammal_outputs(tmp, flat, ids, counts)
        if "pvalues" in steps:
            pvalues(tmp)
        if "viterbi" in steps:
            viterbi(tmp)
        if args.fits:
            fits()


if __name__ == "__main__":
    main()
