"""Pins the CPU oracle (oracle/cafe_oracle.c) against golden vectors produced by the compiled,
unmodified reference (scripts/make_golden.py -> tests/golden/).  CPU only.

Tolerances: the oracle follows the reference operation-for-operation with the same libm, so the
expectation is bit-identity; 1e-13 relative is asserted to stay robust to libm differences between
machines.  Integer outputs (reconstructed states) must match exactly.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, fnum, load_json
from cafexp_b200 import hostio
from oracle import binding as orc

RTOL = 1e-13


def close(a, b, rtol=RTOL):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.allclose(a, b, rtol=rtol, atol=0.0, equal_nan=True)


def test_bd_probabilities():
    g = load_json("scalars.json")
    for case in g["bd"]:
        got = orc.bd_probability(case["lambda"], case["t"], case["s"], case["c"])
        assert close(got, case["p"]), case
    for case in g["bdlog"]:
        got = orc.lib().orc_birthdeath_rate_with_log_alpha(case["s"], case["c"], case["logalpha"], case["coeff"])
        assert close(got, case["p"]), case


def test_reference_unit_test_literals():
    """Literals of the reference's own tests (test.cpp:601-660, loose tolerances as written there)."""
    assert abs(orc.bd_probability(0.05, 5, 5, 9) - 0.0152237) < 1e-5
    assert abs(orc.bd_probability(0.05, 5, 10, 9) - 0.17573) < 1e-5
    assert abs(orc.bd_probability(0.05, 5, 10, 10) - 0.182728) < 1e-5
    assert abs(orc.bd_probability(0.05, 1, 10, 10) - 0.465565) < 1e-5
    assert abs(orc.bd_probability(0.006335, 68.7105, 5, 5) - 0.194661) < 1e-5
    assert abs(orc.build_matrix(141, 0.006335, 68.0)[5, 5] - 0.195791) < 1e-5
    want = np.asarray([[1, 0, 0, 0, 0], [0.2, 0.64, 0.128, 0.0256, 0.00512], [0.04, 0.256, 0.4608, 0.17408, 0.0512],
                       [0.008, 0.0768, 0.26112, 0.36352, 0.187392], [0.0016, 0.02048, 0.1024, 0.249856, 0.305562]])
    assert np.abs(orc.build_matrix(5, 0.05, 5) - want).max() < 1e-5


def test_small_matrices():
    for case in load_json("scalars.json")["matrix_small"]:
        n = case["n"]
        got = orc.build_matrix(n, case["lambda"], case["t"])
        assert close(got.ravel(), case["m"]), case
        assert orc.lib().orc_quantise_lambda(case["lambda"]) == case["lambda_q"]
        assert orc.lib().orc_quantise_branch(case["t"]) == case["t_q"]
        assert orc.lib().orc_is_saturated(case["t_q"], case["lambda_q"]) == case["saturated"]
        if case["saturated"]:
            assert got[0, 0] == 1.0 and got.sum() == 1.0


def test_full_matrices():
    z = np.load(os.path.join(GOLD, "matrices.npz"))
    for meta in json.loads(str(z["meta"])):
        got = orc.build_matrix(meta["n"], meta["lambda"], meta["t"])
        want = z[meta["key"]]
        if meta["rows"] is not None:
            assert close(got[meta["rows"]], want), meta
        else:
            assert close(got, want), meta
        assert close(got.sum(), meta["total"], 1e-12)


def test_key_quantisation_matters():
    """Probability/matrices_take_fractional_branch_lengths_into_account (test.cpp:631) + src/matrix_cache.h:47-60:
    t is truncated to 3 decimals, lambda to 9."""
    a = orc.build_matrix(20, 0.006335, 68.7105)
    b = orc.build_matrix(20, 0.006335, 68.71059)
    c = orc.build_matrix(20, 0.006335, 68.0)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, c)
    assert np.array_equal(orc.build_matrix(20, 0.0063350004, 10.0), orc.build_matrix(20, 0.0063350009, 10.0))


def test_discrete_gamma():
    for case in load_json("scalars.json")["gamma"]:
        freq, rate = orc.get_gamma(case["k"], case["alpha"])
        assert close(rate, case["rate"]), case
        assert close(freq, case["freq"]), case


def test_poisson_prior():
    for case in load_json("scalars.json")["poisson"]:
        n = case["n"]
        got = orc.prior_poisson(case["lambda"], n, None, n + 2)
        assert np.array_equal(got, np.asarray(case["prior"])), case


def test_uniform_prior_semantics():
    """Inference/uniform_distribution (test.cpp:549): list[val] / sum(list) through float."""
    p = orc.prior_uniform(10)
    assert np.all(p == np.float64(np.float32(1.0) / np.float32(10)))
    p = orc.prior_uniform(5, {1: 2, 2: 1, 3: 1}, 6)
    want = [np.float32(v) / np.float32(7) for v in (1, 1, 2, 3)] + [0, 0]
    assert np.array_equal(p, np.asarray(want, np.float64))


def _fixture_tree(rec):
    root = hostio.parse_newick(rec["newick"])
    ltree = hostio.parse_newick(rec["lambda_tree"], True) if rec.get("lambda_tree") else None
    flat = hostio.flatten_tree(root, ltree)
    assert flat.names == rec["node_order"]
    col = {name: i for i, name in enumerate(flat.leaf_names)}
    counts = np.zeros((len(rec["rows"]), flat.n_leaves), np.int32)
    for j, sp in enumerate(rec["species"]):
        counts[:, col[sp]] = [row[j] for row in rec["rows"]]
    return flat, counts


def _fixture_err(rec, mf):
    if not rec.get("error_model"):
        return None
    path = os.path.join(GOLD, "_tmp_err.txt")
    with open(path, "w") as fh:
        fh.write(rec["error_model"])
    try:
        em = hostio.read_error_model(path)
    finally:
        os.remove(path)
    return em.dense(mf + 1)


def _lambdas(arg):
    if isinstance(arg, str):
        return [float(v) for v in arg.split(",")]
    return [float(arg)]


@pytest.mark.parametrize("rec", load_json("unit_fixtures.json"), ids=lambda r: r["name"])
def test_unit_fixtures(rec):
    flat, counts = _fixture_tree(rec)
    mf, mrf = rec["max_family_size"], rec["max_root_family_size"]
    err = _fixture_err(rec, mf)
    lam = np.asarray(_lambdas(rec["args"]["lambda"]))
    if rec["cmd"] == "prune":
        mult = rec["args"].get("mult", 1.0)
        for row, want in zip(counts, rec["root"]):
            got = orc.inference_prune(flat, row, lam * mult, mf, mrf, err)
            assert close(got, want), rec["name"]
        return
    prior = orc.prior_uniform(mrf, None, max(mrf, mf) + 1)
    if "cat_lk" in rec:
        mults = np.asarray(rec["multipliers"])
        freq, rate = orc.get_gamma(len(mults), rec["args"]["alpha"])
        assert close(rate, mults)
        lams = mults[:, None] * lam[None, :]
        res = orc.infer(flat, counts, lams, rec["cat_probs"], prior, mf, mrf, orc.GAMMA_LINSUM, err)
        assert close(res["cat_lk"], rec["cat_lk"])
    else:
        lams = lam[None, :]
        res = orc.infer(flat, counts, lams, [1.0], prior, mf, mrf, orc.BASE_LOGMAX, err)
        assert close(res["family_lnl"], rec["family_lnl"])
    assert close(res["score"], fnum(rec["score"]))
    if "states" in rec:
        assert flat.internal_names == rec["internal_order"]
        states = orc.reconstruct(flat, counts, lams, prior, mf, mrf)
        assert np.array_equal(states.reshape(len(counts), -1), np.asarray(rec["states"]))


def test_reference_literals_infer_processes():
    """Inference/infer_processes (test.cpp:519) expects 41.7504 at 1e-3; gamma_lambda_optimizer (:2240) 6.4168."""
    recs = {r["name"]: r for r in load_json("unit_fixtures.json")}
    assert abs(recs["infer_processes"]["score"] - 41.7504) < 1e-3
    assert abs(recs["gamma_lambda_optimizer"]["score"] - 6.4168) < 1e-4


def test_error_model_reader_and_epsilon(mammal):
    path = os.path.join(GOLD, "_tmp_err2.txt")
    with open(path, "w") as fh:
        fh.write(mammal["inputs"]["error_model"])
    try:
        em = hostio.read_error_model(path)
    finally:
        os.remove(path)
    assert em.max_family_size == 90 and em.deviations == [-1, 0, 1]
    assert em.get_probs(0) == [0.0, 0.95, 0.05] and em.get_probs(17) == [0.05, 0.9, 0.05]
    assert em.epsilons() == [0.05]
    dense = em.dense(91)
    assert orc.lib().orc_error_model_replace_epsilon(dense.ctypes.data_as(orc._dp), 91, 0.05, 0.1) == 0
    assert np.allclose(dense[0], [0, 0.9, 0.1]) and np.allclose(dense[5], [0.1, 0.8, 0.1])


def _check_mammal(mammal, name, tree_key="tree", err=False, prior="uniform", prior_arg=None, rootdist=False, k=0, recon=False,
                  stride=1):
    meta = mammal["meta"][name]
    flat = mammal[tree_key]
    assert flat.names == meta["node_order"]
    mf, mrf = mammal["mf"], mammal["mrf"]
    assert (mf, mrf) == (meta["max_family_size"], meta["max_root_family_size"])
    counts = mammal["counts"]
    sel = slice(None, None, stride)
    errtab = None
    if err:
        path = os.path.join(GOLD, "_tmp_err3.txt")
        with open(path, "w") as fh:
            fh.write(mammal["inputs"]["error_model"])
        try:
            errtab = hostio.read_error_model(path).dense(mf + 1)
        finally:
            os.remove(path)
    rd = None
    if rootdist:
        rd = {int(a): int(b) for a, b in (line.split() for line in mammal["inputs"]["rootdist"].splitlines() if line.strip())}
    n_prior = max(mf, mrf) + 1
    pr = orc.prior_uniform(mrf, rd, n_prior) if prior == "uniform" else orc.prior_poisson(prior_arg, mrf, rd, n_prior)
    lam = np.asarray(_lambdas(meta["args"]["lambda"]))
    gold = mammal["gold"]
    if k:
        mults = np.asarray(meta["multipliers"])
        lams = mults[:, None] * lam[None, :]
        res = orc.infer(flat, counts[sel], lams, meta["cat_probs"], pr, mf, mrf, orc.GAMMA_LINSUM, errtab)
        want = gold[name + "_cat_lk"][sel]
        ok = ~np.isnan(want).any(axis=1)
        assert np.array_equal(res["failed"] == 0, ok)
        assert close(res["cat_lk"][ok], want[ok])
    else:
        lams = lam[None, :]
        res = orc.infer(flat, counts[sel], lams, [1.0], pr, mf, mrf, orc.BASE_LOGMAX, errtab)
        assert close(res["family_lnl"], gold[name + "_lnl"][sel])
    if stride == 1:
        assert close(res["score"], fnum(meta["score"]), 1e-12)
    if recon:
        rsel = slice(None, None, max(stride, 8))
        states = orc.reconstruct(flat, counts[rsel], lams, pr, mf, mrf)
        want = gold[name + "_states"][rsel]
        assert np.array_equal(states.reshape(states.shape[0], -1), want)
    return res


def test_mammal_base(mammal):
    """BASELINE.json config 1 at lambda = 0.002: -lnL 164876.196089535."""
    res = _check_mammal(mammal, "base_l002")
    assert abs(res["score"] - 164876.196089535) < 1e-6


def test_mammal_error_model(mammal):
    res = _check_mammal(mammal, "base_err_l002", err=True)
    assert abs(res["score"] - 158170.965028356) < 1e-6


def test_mammal_poisson_prior(mammal):
    _check_mammal(mammal, "base_poisson_l002", prior="poisson", prior_arg=10.0, stride=4)


@pytest.mark.slow
def test_mammal_gamma(mammal):
    """Config 2 at the reference's fitted optimum: -lnL 154787.038270984."""
    res = _check_mammal(mammal, "gamma4_fit", k=4, stride=1)
    assert abs(res["score"] - 154787.038270984) < 1e-5


def test_mammal_gamma_failure_path(mammal):
    """Config 2 at (0.002, 0.5): the lowest category underflows for some families -> +inf."""
    meta = mammal["meta"]["gamma4_fail"]
    assert fnum(meta["score"]) == float("inf")
    _check_mammal(mammal, "gamma4_fail", k=4, stride=16)


def test_mammal_reconstruction(mammal):
    """Config 3: error model (ignored by Pupko, as in the reference) at lambda = 0.01."""
    _check_mammal(mammal, "base_err_l01_recon", err=True, recon=True, stride=8)


def test_mammal_gamma_reconstruction(mammal):
    _check_mammal(mammal, "gamma3_recon", k=3, recon=True, stride=32)


def test_mammal_multilambda(mammal):
    """Config 4: two lambdas + rootdist map (uniform over the expanded list) and Poisson prior."""
    _check_mammal(mammal, "multi_rootdist", tree_key="tree2", rootdist=True, stride=4)
    _check_mammal(mammal, "multi_poisson_recon", tree_key="tree2", prior="poisson", prior_arg=12.5, recon=True, stride=16)
