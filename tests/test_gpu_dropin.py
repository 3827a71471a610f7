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


def run_harness(cmd, devices=None, **kw):
    argv = [HARNESS, cmd, "--cuda", "1"]
    env = dict(os.environ)
    env.pop("CAFE_B200_DEVICES", None)
    if devices:
        env["CAFE_B200_DEVICES"] = devices
    for key, val in kw.items():
        if val is None or val is False:
            continue
        argv.append("--" + key)
        if val is not True:
            argv.append(repr(val) if isinstance(val, float) else str(val))
    res = subprocess.run(argv, capture_output=True, text=True, env=env)
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


def _device_count():
    from cafexp_b200 import engine
    return engine.device_count()


def test_lazy_results_are_what_the_reference_writes(files, mammal):
    """model::results is materialised on demand by the drop-in (write_family_likelihoods); the harness dumps it after
    several evaluations in a row: it must hold the LAST evaluation's values (reps = 3 here)."""
    dump = os.path.join(files["dir"], "lazy.bin")
    r = run_harness("eval", tree=files["tree"], fam=files["fam"], dump=dump, reps=3, **{"lambda": 0.002})
    assert rel(fnum(r["score"]), fnum(mammal["meta"]["base_l002"]["score"])) < 1e-11
    assert np.allclose(np.fromfile(dump, np.float64), mammal["gold"]["base_l002_lnl"], rtol=1e-11, atol=0)


def test_drop_in_on_two_devices(files, mammal):
    """CAFE_B200_DEVICES=0,1: the reference's single host thread drives both devices through one context
    (cafe_b200_create_multi).  Evaluation with error model + reconstruction, and the joint lambda/alpha fit: identical
    optimizer path (evaluations, iterations), parameters equal to the one-device fit."""
    if _device_count() < 2:
        pytest.skip("needs two visible CUDA devices")
    name = "base_err_l01_recon"
    gold = mammal["gold"]
    dump, rec = os.path.join(files["dir"], "two.bin"), os.path.join(files["dir"], "two.rec")
    r = run_harness("eval", devices="0,1", tree=files["tree"], fam=files["fam"], err=files["err"], dump=dump, recon=True, dumprecon=rec,
                    **{"lambda": 0.01})
    F = len(mammal["counts"])
    assert rel(fnum(r["score"]), fnum(mammal["meta"][name]["score"])) < 1e-11
    assert np.allclose(np.fromfile(dump, np.float64), gold[name + "_lnl"], rtol=1e-11, atol=0)
    assert np.array_equal(np.fromfile(rec, np.int32).reshape(F, -1), gold[name + "_states"])
    kw = {"tree": files["tree"], "fam": files["fam"], "seed": 10, "k": 4}
    one = run_harness("fit", **kw)
    two = run_harness("fit", devices="0,1", **kw)
    assert one["devices"] == 1 and two["devices"] == 2
    assert two["evaluations"] == one["evaluations"] and two["iterations"] == one["iterations"]
    for a, b in zip(two["values"], one["values"]):
        assert rel(a, b) < 1e-9
    assert rel(fnum(two["score"]), fnum(one["score"])) < 1e-12
