"""Drop-in parity: the reference's own host code — readers, lambda containers, scorers, Nelder-Mead optimizer —
driving libcafe_b200.so through the CUDA-backed model subclasses of integration/cuda_models.cpp.

oracle/_ref/ref_harness_cuda is the reference's unmodified objects + oracle/ref_harness.cpp + integration/
(built by `make -C oracle refcuda` in the build container; the binary travels to the GPU box).  With
--cuda 1 the model objects are cuda_base_model / cuda_gamma_model, everything else is the reference.

Bars (north_star): scores / per-family values 1e-9 relative (asserted tighter), fitted lambda / alpha / epsilon
1e-6 relative against the seed-10 fits of the unmodified reference (tests/golden/fits.json), reconstructed
counts exact.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, fnum, load_json
from cafexp_b200 import hostio

pytestmark = pytest.mark.gpu

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness_cuda")


def run_harness(cmd, **kw):
    argv = [HARNESS, cmd, "--cuda", "1"]
    for key, val in kw.items():
        if val is None or val is False:
            continue
        argv.append("--" + key)
        if val is not True:
            argv.append(repr(val) if isinstance(val, float) else str(val))
    res = subprocess.run(argv, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads([l for l in res.stdout.splitlines() if l.startswith("{\"")][-1])


@pytest.fixture(scope="module")
def files(mammal, tmp_path_factory):
    if not os.path.exists(HARNESS):
        pytest.fail("oracle/_ref/ref_harness_cuda is missing: run `make -C oracle refcuda` in the build container")
    d = tmp_path_factory.mktemp("dropin")
    inp = mammal["inputs"]
    paths = {k: str(d / f"{k}.txt") for k in ("tree", "fam", "err", "ltree", "rootdist")}
    open(paths["tree"], "w").write(inp["tree"] + "\n")
    open(paths["ltree"], "w").write(inp["lambda_tree"] + "\n")
    open(paths["err"], "w").write(inp["error_model"])
    open(paths["rootdist"], "w").write(inp["rootdist"])
    ids = np.load(os.path.join(GOLD, "mammal_counts.npz"))["ids"]
    hostio.write_gene_families(paths["fam"], mammal["tree"], [str(i) for i in ids], mammal["counts_all"])
    paths["dir"] = str(d)
    return paths


def rel(a, b):
    return abs(a - b) / abs(b)


CASES = {
    # name in tests/golden/mammal_outputs.npz -> harness arguments (as scripts/make_golden.py ran the reference)
    "base_l002": dict(lam=0.002),
    "base_err_l002": dict(lam=0.002, err=True),
    "base_err_l01_recon": dict(lam=0.01, err=True, recon=True),
    "base_poisson_l002": dict(lam=0.002, poisson=10.0),
    "gamma4_fit": dict(lam=0.00354641825220246, k=4, alpha=0.481515908358985),
    "gamma4_fail": dict(lam=0.002, k=4, alpha=0.5),
    "gamma3_recon": dict(lam=0.002, k=3, alpha=0.425, recon=True),
    "multi_rootdist": dict(lam="0.01,0.05", ltree=True, rootdist=True),
    "multi_poisson_recon": dict(lam="0.01,0.05", ltree=True, poisson=12.5, recon=True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_reference_host_with_cuda_models_matches_reference(mammal, files, name):
    """model::infer_family_likelihoods / reconstruct_ancestral_states through the subclass overrides,
    BASELINE.json configs 1-4, against the unmodified reference's dumps."""
    c = CASES[name]
    gold, meta = mammal["gold"], mammal["meta"][name]
    dump = os.path.join(files["dir"], name + ".bin")
    rec = os.path.join(files["dir"], name + ".rec")
    kw = {"tree": files["tree"], "fam": files["fam"], "lambda": c["lam"], "dump": dump}
    if c.get("err"):
        kw["err"] = files["err"]
    if c.get("ltree"):
        kw["ltree"] = files["ltree"]
    if c.get("rootdist"):
        kw["rootdist"] = files["rootdist"]
    for key in ("poisson", "k", "alpha"):
        if key in c:
            kw[key] = c[key]
    if c.get("recon"):
        kw.update(recon=True, dumprecon=rec)
    r = run_harness("eval", **kw)
    F = len(mammal["counts"])
    assert r["n_families"] == F and r["max_family_size"] == meta["max_family_size"]
    want = fnum(meta["score"])
    got = fnum(r["score"])
    if np.isinf(want):
        assert got == want
    else:
        assert rel(got, want) < 1e-11
        data = np.fromfile(dump, np.float64)
        if "k" in c:
            assert np.allclose(data.reshape(F, c["k"]), gold[name + "_cat_lk"], rtol=1e-11, atol=0)
        else:
            assert np.allclose(data, gold[name + "_lnl"], rtol=1e-11, atol=0)
    if c.get("recon"):
        states = np.fromfile(rec, np.int32).reshape(F, -1)
        assert np.array_equal(states, gold[name + "_states"])


FITS = {
    # name in tests/golden/fits.json -> harness arguments
    "single_lambda_seed10": dict(),
    "lambda_epsilon_seed10": dict(err=True),
    "lambda_estimated_epsilon_seed10": dict(esterr=1),
    "two_lambda_seed10": dict(ltree=True),
    "gamma4_lambda_alpha_seed10": dict(k=4),
}


@pytest.mark.parametrize("name", list(FITS))
def test_fit_through_reference_optimizer(files, name):
    """optimizer::optimize (src/optimizer.cpp:539) over the scorer the CUDA-backed model hands out: same seed,
    same Nelder-Mead path, parameters within 1e-6 relative of the unmodified reference's fit."""
    gold = load_json("fits.json")
    if name not in gold:
        pytest.skip(f"{name}: no golden fit committed (scripts/make_golden.py --fits)")
    want = gold[name]
    c = FITS[name]
    kw = {"tree": files["tree"], "fam": files["fam"], "seed": 10}
    if c.get("err"):
        kw["err"] = files["err"]
    if c.get("ltree"):
        kw["ltree"] = files["ltree"]
    for key in ("k", "esterr"):
        if key in c:
            kw[key] = c[key]
    r = run_harness("fit", **kw)
    assert r["model"] == want["model"]
    assert len(r["values"]) == len(want["values"])
    for g, w in zip(r["values"], want["values"]):
        assert rel(g, w) < 1e-6, (r["values"], want["values"])
    assert rel(fnum(r["score"]), fnum(want["score"])) < 1e-9
    assert r["evaluations"] == want["evaluations"] and r["iterations"] == want["iterations"]
