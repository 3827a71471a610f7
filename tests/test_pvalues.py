"""Family p-values (compute_pvalues, src/probability.cpp:411-444): SURVEY section 8(f) rank 1.

CPU part: the oracle's root_max / pvalues restatement against the golden dump of the compiled reference
(tests/golden/mammal_pvalues.npz, scripts/make_golden.py --only pvalues).
GPU part (marked gpu): cafe_b200_root_max / cafe_b200_pvalues through the C ABI against oracle and golden, and the
reference's own host code (random stream included) driving the CUDA path through compute_pvalues_cuda.

The simulated likelihoods only matter through their ORDER relative to the observed ones, so p-values are exact
(multiples of 1/n_sim) unless an observed likelihood and a simulated one agree to ~1e-11 relative; the tests
assert exact equality and, for the DMMA path, allow a flip of one rank at such near-ties.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT
from cafexp_b200 import hostio
from oracle import binding as orc


@pytest.fixture(scope="module")
def pv(mammal):
    z = np.load(os.path.join(GOLD, "mammal_pvalues.npz"))
    meta = json.loads(str(z["meta"]))
    tree = mammal["tree"]
    # columns of the golden dump follow the reference's reverse-level-order leaves; map them onto the count-matrix columns
    col = [meta["leaf_order"].index(n) for n in tree.leaf_names]
    sim = z["sim_counts"].astype(np.int32)[:, col]
    return {"meta": meta, "sim": np.ascontiguousarray(sim), "cond": z["cond"], "observed": z["observed"], "pvalues": z["pvalues"],
            "families": np.ascontiguousarray(mammal["counts"][: meta["limit"]]), "tree": tree}


def test_oracle_root_max_matches_reference(pv):
    m = pv["meta"]
    got = orc.root_max(pv["tree"], pv["families"], [m["lambda"]], m["max_family_size"], m["max_root_family_size"])
    assert np.allclose(got, pv["observed"], rtol=1e-12, atol=0)
    # a slice of the simulated families (one per root size) against their golden likelihoods
    idx = np.arange(0, len(pv["sim"]), m["nsim"])
    got = orc.root_max(pv["tree"], pv["sim"][idx], [m["lambda"]], m["max_family_size"], m["max_root_family_size"])
    assert np.allclose(got, pv["cond"].ravel()[idx], rtol=1e-12, atol=0)


def test_oracle_pvalues_match_reference(pv):
    got = orc.pvalues(pv["cond"], pv["observed"])
    assert np.array_equal(got, pv["pvalues"])


def test_pvalue_edge_semantics():
    """src/probability.cpp:379-389: idx = count of simulated values <= v, except size-1 when nothing is greater."""
    sorted_row = np.array([0.1, 0.2, 0.2, 0.5])
    L = orc.lib()
    import ctypes as C
    p = sorted_row.ctypes.data_as(C.POINTER(C.c_double))
    assert L.orc_pvalue(0.05, p, 4) == 0.0
    assert L.orc_pvalue(0.1, p, 4) == 0.25
    assert L.orc_pvalue(0.2, p, 4) == 0.75
    assert L.orc_pvalue(0.5, p, 4) == 0.75      # nothing greater: size-1, not size
    assert L.orc_pvalue(9.0, p, 4) == 0.75


# ------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_root_max_against_oracle_and_golden(pv):
    from cafexp_b200 import engine
    m = pv["meta"]
    with engine.Engine(pv["tree"], pv["families"], m["max_family_size"], m["max_root_family_size"]) as eng:
        got = eng.root_max([m["lambda"]])
    assert np.allclose(got, pv["observed"], rtol=1e-11, atol=0)
    with engine.Engine(pv["tree"], pv["sim"], m["max_family_size"], m["max_root_family_size"]) as eng:
        got = eng.root_max([m["lambda"]])
    assert np.allclose(got, pv["cond"].ravel(), rtol=1e-11, atol=0)
    assert np.array_equal(got == 0, pv["cond"].ravel() == 0)


@pytest.mark.gpu
def test_gpu_pvalues_kernel_exact(pv):
    from cafexp_b200 import engine
    got = engine.pvalues(pv["cond"], pv["observed"])
    assert np.array_equal(got, pv["pvalues"])
    # sizes that are not a power of two, ties, values above and below every simulated one
    rng = np.random.default_rng(5)
    for n_root, n_sim, F in ((1, 1, 7), (3, 1000, 5000), (17, 37, 1), (5, 4096, 300)):
        cond = rng.choice(rng.random(max(4, n_sim // 3)), size=(n_root, n_sim))
        obs = np.concatenate([rng.choice(cond.ravel(), size=F // 2 + 1), rng.random(F) * 1.2 - 0.1])[:F]
        assert np.array_equal(engine.pvalues(cond, obs), orc.pvalues(cond, obs)), (n_root, n_sim, F)


@pytest.mark.gpu
def test_gpu_pvalues_end_to_end(pv):
    """simulated + observed likelihoods from the pruning kernel, p-values from the device: equals the reference's
    p-values except where an observed likelihood ties a simulated one to rounding."""
    from cafexp_b200 import engine
    m = pv["meta"]
    with engine.Engine(pv["tree"], pv["sim"], m["max_family_size"], m["max_root_family_size"]) as eng:
        cond = eng.root_max([m["lambda"]]).reshape(m["max_root_family_size"], m["nsim"])
    with engine.Engine(pv["tree"], pv["families"], m["max_family_size"], m["max_root_family_size"]) as eng:
        obs = eng.root_max([m["lambda"]])
    got = engine.pvalues(cond, obs)
    diff = np.abs(got - pv["pvalues"])
    assert diff.max() <= 1.0 / m["nsim"] + 1e-15
    assert (diff > 0).mean() < 0.02


@pytest.mark.gpu
def test_reference_host_code_with_cuda_pvalues(pv, mammal, tmp_path):
    """ref_harness_cuda pvalues --cuda 1: the reference's generator (same seed, same order of draws) feeds
    compute_pvalues_cuda (integration/cuda_models.cpp); result vs the reference's own compute_pvalues (golden)."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness_cuda")
    if not os.path.exists(harness):
        pytest.fail("oracle/_ref/ref_harness_cuda is missing: run `make -C oracle refcuda` in the build container")
    m = pv["meta"]
    tree_path, fam_path = tmp_path / "tree.txt", tmp_path / "fam.txt"
    tree_path.write_text(mammal["inputs"]["tree"] + "\n")
    ids = np.load(os.path.join(GOLD, "mammal_counts.npz"))["ids"]
    hostio.write_gene_families(str(fam_path), mammal["tree"], [str(i) for i in ids], mammal["counts_all"])
    dump = tmp_path / "pv.bin"
    res = subprocess.run([harness, "pvalues", "--cuda", "1", "--tree", str(tree_path), "--fam", str(fam_path), "--lambda", repr(m["lambda"]),
                          "--limit", str(m["limit"]), "--nsim", str(m["nsim"]), "--seed", str(m["seed"]), "--dump", str(dump)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    raw = dump.read_bytes()
    F = m["limit"]
    got = np.frombuffer(raw, np.float64, F, len(raw) - 8 * F)
    diff = np.abs(got - pv["pvalues"])
    assert diff.max() <= 1.0 / m["nsim"] + 1e-15
    assert (diff > 0).mean() < 0.02


@pytest.mark.gpu
def test_gpu_root_max_sharded_over_two_devices(pv):
    """cafe_b200_root_max through a two-device context: the same values, family by family, as on one device."""
    from cafexp_b200 import engine
    if engine.device_count() < 2:
        pytest.skip("needs two visible CUDA devices")
    m = pv["meta"]
    for rows in (pv["families"], pv["sim"]):
        with engine.Engine(pv["tree"], rows, m["max_family_size"], m["max_root_family_size"], device=0) as eng:
            one = eng.root_max([m["lambda"]])
        with engine.Engine(pv["tree"], rows, m["max_family_size"], m["max_root_family_size"], device=[0, 1]) as eng:
            assert "devices=2" in eng.describe()
            two = eng.root_max([m["lambda"]])
        assert np.array_equal(one, two)
