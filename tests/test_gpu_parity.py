"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors of the
compiled reference.  Needs a B200: every test is marked gpu.

Tolerances
  * transition-matrix entries: 4e-15 relative — inputs of exp() are bit-identical to the reference's, only
    CUDA's exp() (<1 ulp) vs glibc's can differ, and the sum has only positive terms.
  * root vectors / per-family lnL / scores: 1e-9 relative is north_star's bar; we assert 1e-11 (DMMA sums in a
    different order than the reference's serial loop).
  * reconstructed ancestral counts, failure flags: exact.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, fnum, load_json
from cafexp_b200 import engine, hostio, synth
from oracle import binding as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-11


def err_table(text, rows):
    path = os.path.join(GOLD, "_tmp_err_gpu.txt")
    with open(path, "w") as fh:
        fh.write(text)
    try:
        return hostio.read_error_model(path).dense(rows)
    finally:
        os.remove(path)


def rel_err(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / np.abs(b)
    r[(a == b)] = 0.0
    return np.nanmax(r) if r.size else 0.0


def test_matrix_builder_against_golden_matrices():
    """Kernel 1 vs matrices dumped from the reference's matrix_cache (tests/golden/matrices.npz)."""
    z = np.load(os.path.join(GOLD, "matrices.npz"))
    tree = hostio.flatten_tree(hostio.parse_newick("(A:1,B:1);"))
    for meta in json.loads(str(z["meta"])):
        n = meta["n"]
        t2 = hostio.flatten_tree(hostio.parse_newick(f"(A:{meta['t']!r},B:1);"))
        with engine.Engine(t2, np.array([[1, 1]], np.int32), n - 1, n - 1) as eng:
            assert eng.matrix_size == n
            got = eng.build_matrices([[meta["lambda"]]])[0, list(t2.names).index("A")]
        want = z[meta["key"]]
        if meta["rows"] is not None:
            got = got[meta["rows"]]
        assert got.shape == want.shape
        assert rel_err(got, want) < 4e-15, meta
        assert np.array_equal(got == 0, want == 0)


def test_matrix_builder_saturated_and_degenerate_keys():
    tree = hostio.flatten_tree(hostio.parse_newick("(A:4,B:0.0004);"))
    with engine.Engine(tree, np.array([[1, 1]], np.int32), 11, 8) as eng:
        m = eng.build_matrices([[0.3]])[0]
    a = list(tree.names).index("A")
    b = list(tree.names).index("B")
    assert m[a, 0, 0] == 1.0 and m[a].sum() == 1.0          # saturated: 1 - 2 alpha < 0  (src/matrix_cache.cpp:152)
    assert np.array_equal(m[a], orc.build_matrix(12, 0.3, 4.0))
    assert np.array_equal(m[b], orc.build_matrix(12, 0.3, 0.0004))   # t quantises to 0 -> coeff == 1 -> zeros


def _fixture(rec):
    root = hostio.parse_newick(rec["newick"])
    ltree = hostio.parse_newick(rec["lambda_tree"], True) if rec.get("lambda_tree") else None
    flat = hostio.flatten_tree(root, ltree)
    col = {name: i for i, name in enumerate(flat.leaf_names)}
    counts = np.zeros((len(rec["rows"]), flat.n_leaves), np.int32)
    for j, sp in enumerate(rec["species"]):
        counts[:, col[sp]] = [row[j] for row in rec["rows"]]
    lam = rec["args"]["lambda"]
    lam = np.asarray([float(v) for v in lam.split(",")] if isinstance(lam, str) else [float(lam)])
    return flat, counts, lam


@pytest.mark.parametrize("rec", load_json("unit_fixtures.json"), ids=lambda r: r["name"])
def test_reference_unit_fixtures(rec):
    """The reference's own unit-test fixtures (test.cpp), expectations from the compiled reference at 17 digits."""
    flat, counts, lam = _fixture(rec)
    mf, mrf = rec["max_family_size"], rec["max_root_family_size"]
    with engine.Engine(flat, counts, mf, mrf) as eng:
        if rec.get("error_model"):
            eng.set_error_model(err_table(rec["error_model"], mf + 1))
        if rec["cmd"] == "prune":
            got = eng.prune_roots(lam[None, :] * rec["args"].get("mult", 1.0))[:, 0, :]
            assert rel_err(got, np.asarray(rec["root"])) < RTOL
            return
        prior = orc.prior_uniform(mrf, None, max(mrf, mf) + 1)
        if "cat_lk" in rec:
            lams = np.asarray(rec["multipliers"])[:, None] * lam[None, :]
            res = eng.infer(lams, prior, rec["cat_probs"], engine.GAMMA_LINSUM)
            assert rel_err(res["cat_lk"], np.asarray(rec["cat_lk"])) < RTOL
        else:
            lams = lam[None, :]
            res = eng.infer(lams, prior, None, engine.BASE_LOGMAX)
            assert rel_err(res["family_lnl"], np.asarray(rec["family_lnl"])) < RTOL
        want_score = fnum(rec["score"])
        assert res["score"] == want_score or abs(res["score"] - want_score) <= RTOL * abs(want_score)
        if "states" in rec:
            states = eng.reconstruct(lams, prior)
            assert np.array_equal(states.reshape(len(counts), -1), np.asarray(rec["states"]))


def _mammal_case(mammal, name, tree_key="tree", err=False, prior="uniform", prior_arg=None, rootdist=False, k=0, recon=False,
                 rescale=False):
    meta = mammal["meta"][name]
    flat = mammal[tree_key]
    mf, mrf = mammal["mf"], mammal["mrf"]
    counts = mammal["counts"]
    rd = None
    if rootdist:
        rd = {int(a): int(b) for a, b in (line.split() for line in mammal["inputs"]["rootdist"].splitlines() if line.strip())}
    n_prior = max(mf, mrf) + 1
    pr = orc.prior_uniform(mrf, rd, n_prior) if prior == "uniform" else orc.prior_poisson(prior_arg, mrf, rd, n_prior)
    lam = meta["args"]["lambda"]
    lam = np.asarray([float(v) for v in lam.split(",")] if isinstance(lam, str) else [float(lam)])
    gold = mammal["gold"]
    with engine.Engine(flat, counts, mf, mrf) as eng:
        eng.set_rescale(rescale)
        if err:
            eng.set_error_model(err_table(mammal["inputs"]["error_model"], mf + 1))
        if k:
            lams = np.asarray(meta["multipliers"])[:, None] * lam[None, :]
            res = eng.infer(lams, pr, meta["cat_probs"], engine.GAMMA_LINSUM, failed_cap=20000)
            want = gold[name + "_cat_lk"]
            failed = np.isnan(want).any(axis=1)
            assert res["n_failed"] == int(failed.sum())
            assert np.array_equal(res["failed_idx"], np.flatnonzero(failed))
            assert np.array_equal(np.isnan(res["family_lnl"]), failed)
            ok = ~failed
            # a failed family stops at its first underflowing category in the reference; compare complete rows
            assert rel_err(res["cat_lk"][ok], want[ok]) < RTOL
        else:
            lams = lam[None, :]
            res = eng.infer(lams, pr, None, engine.BASE_LOGMAX)
            assert rel_err(res["family_lnl"], gold[name + "_lnl"]) < RTOL
        want_score = fnum(meta["score"])
        if np.isinf(want_score):
            assert res["score"] == want_score
        else:
            assert abs(res["score"] - want_score) <= RTOL * abs(want_score)
        if recon:
            states = eng.reconstruct(lams, pr)
            want_states = gold[name + "_states"]
            got = states.reshape(states.shape[0], -1)
            assert got.shape == want_states.shape
            assert np.array_equal(got, want_states), f"{int((got != want_states).sum())} reconstructed counts differ"
    return res


def test_config1_mammal_base(mammal):
    """BASELINE.json config 1: single lambda, uniform prior, 10 956 families; reference -lnL 164876.196089535."""
    res = _mammal_case(mammal, "base_l002")
    assert abs(res["score"] - 164876.196089535) < 1e-5


def test_config2_mammal_gamma(mammal):
    """Config 2 at the reference's fitted (lambda, alpha): -lnL 154787.038270984."""
    res = _mammal_case(mammal, "gamma4_fit", k=4)
    assert abs(res["score"] - 154787.038270984) < 1e-5


def test_config2_mammal_gamma_failure_path(mammal):
    """Config 2 at (0.002, 0.5): categories underflow for some families -> +inf and the same failed families."""
    res = _mammal_case(mammal, "gamma4_fail", k=4)
    assert res["score"] == float("inf") and res["n_failed"] > 0


def test_config3_mammal_error_model_and_reconstruction(mammal):
    _mammal_case(mammal, "base_err_l002", err=True)
    _mammal_case(mammal, "base_err_l01_recon", err=True, recon=True)


def test_config3_gamma_reconstruction(mammal):
    _mammal_case(mammal, "gamma3_recon", k=3, recon=True)


def test_config4_mammal_multilambda_rootdist_poisson(mammal):
    _mammal_case(mammal, "multi_rootdist", tree_key="tree2", rootdist=True)
    _mammal_case(mammal, "multi_poisson_recon", tree_key="tree2", prior="poisson", prior_arg=12.5, recon=True)
    _mammal_case(mammal, "base_poisson_l002", prior="poisson", prior_arg=10.0)


def test_rescaling_is_exact_where_nothing_underflows(mammal):
    """Power-of-two rescaling must not change a single bit of the result when the reference does not underflow,
    and must reproduce the failure verdicts when it does."""
    a = _mammal_case(mammal, "base_l002", rescale=False)
    b = _mammal_case(mammal, "base_l002", rescale=True)
    assert rel_err(b["family_lnl"], a["family_lnl"]) < 1e-15
    c = _mammal_case(mammal, "gamma4_fail", k=4, rescale=True)
    assert c["score"] == float("inf")


@pytest.fixture(scope="module")
def config5_small():
    tree, counts, newick = synth.config5(1500)
    return tree, counts


def test_config5_slice_against_oracle(config5_small):
    """Config 5 shape (100 taxa, N = 151, gamma k = 4) on a slice the CPU oracle finishes in seconds."""
    tree, counts = config5_small
    mf, mrf = synth.CONFIG5_MAX_FAMILY_SIZE, synth.CONFIG5_MAX_ROOT_FAMILY_SIZE
    freq, rate = orc.get_gamma(4, 0.7)
    lams = rate[:, None] * np.array([[0.005]])
    prior = orc.prior_uniform(mrf, None, mf + 1)
    want = orc.infer(tree, counts, lams, freq, prior, mf, mrf, orc.GAMMA_LINSUM)
    with engine.Engine(tree, counts, mf, mrf) as eng:
        got = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
        assert got["n_failed"] == want["n_failed"]
        assert rel_err(got["cat_lk"], want["cat_lk"]) < RTOL
        assert rel_err(got["family_lnl"], want["family_lnl"]) < RTOL
        assert abs(got["score"] - want["score"]) <= RTOL * abs(want["score"])
        roots = eng.prune_roots(lams[:1])[:64, 0]
        for i in range(0, 64, 16):
            assert rel_err(roots[i], orc.inference_prune(tree, counts[i], lams[0], mf, mrf)) < RTOL
        sub = slice(0, 96)
        states = eng.reconstruct(lams, prior)[sub]
        assert np.array_equal(states, orc.reconstruct(tree, counts[sub], lams, prior, mf, mrf))
        # forced spilling (2 and 3 slots) must not change anything
        for slots in (3, 2):
            eng._check(eng._lib.cafe_b200_set_option(eng._h, 2, slots), "set_option")
            again = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
            assert np.array_equal(again["cat_lk"], got["cat_lk"])


def test_config5_properties_at_scale():
    """Size-independent properties on 40 000 families (too many for the oracle): permutation equivariance,
    duplication = 2x, tile-boundary independence, sum of per-family lnL == score."""
    tree, counts, _ = synth.config5(40000, seed=777)
    mf, mrf = synth.CONFIG5_MAX_FAMILY_SIZE, synth.CONFIG5_MAX_ROOT_FAMILY_SIZE
    freq, rate = orc.get_gamma(4, 0.7)
    lams = rate[:, None] * np.array([[0.005]])
    prior = orc.prior_uniform(mrf, None, mf + 1)
    with engine.Engine(tree, counts, mf, mrf) as eng:
        a = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
    assert a["n_failed"] == 0
    assert abs(-np.sum(a["family_lnl"]) - a["score"]) <= 1e-12 * abs(a["score"])
    perm = np.random.default_rng(5).permutation(len(counts))
    with engine.Engine(tree, counts[perm], mf, mrf) as eng:
        b = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
    assert np.array_equal(b["family_lnl"], a["family_lnl"][perm])       # bit-identical per family wherever it sits
    odd = counts[:10007]                                                 # not a multiple of the tile size
    with engine.Engine(tree, np.concatenate([odd, odd]), mf, mrf) as eng:
        c = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
    assert np.array_equal(c["family_lnl"][:10007], a["family_lnl"][:10007])
    assert np.array_equal(c["family_lnl"][10007:], a["family_lnl"][:10007])
    assert abs(c["score"] - 2 * (-np.sum(a["family_lnl"][:10007]))) <= 1e-12 * abs(c["score"])


def test_edge_cases():
    tree = hostio.flatten_tree(hostio.parse_newick("((A:1,B:1):1,(C:1,D:1):1);"))
    prior = orc.prior_uniform(20, None, 26)
    # empty family set
    with engine.Engine(tree, np.zeros((0, 4), np.int32), 25, 20) as eng:
        res = eng.infer([[0.03]], prior)
        assert res["score"] == 0.0 and res["family_lnl"].shape == (0,)
    # one family; all-zero family (extinct everywhere): the reference gives lnL = log(M..) values, no crash
    counts = np.array([[0, 0, 0, 0], [25, 25, 25, 25], [0, 25, 0, 25]], np.int32)
    want = orc.infer(tree, counts, [[0.03]], [1.0], prior, 25, 20, orc.BASE_LOGMAX)
    with engine.Engine(tree, counts, 25, 20) as eng:
        got = eng.infer([[0.03]], prior)
        assert np.allclose(got["family_lnl"], want["family_lnl"], rtol=RTOL, atol=0, equal_nan=True)
        assert (got["score"] == want["score"]) or abs(got["score"] - want["score"]) <= RTOL * abs(want["score"])
        # count above max_family_size is rejected like an out-of-range index would be
        with pytest.raises(engine.CafeB200Error, match="COUNT_RANGE"):
            eng.set_families(np.array([[0, 0, 0, 26]] * 3, np.int32))
        # the rejected matrix is on the device (the check runs there): evaluations are refused until a valid one is set
        with pytest.raises(engine.CafeB200Error, match="COUNT_RANGE"):
            eng.infer([[0.03]], prior)
        with pytest.raises(engine.CafeB200Error, match="COUNT_RANGE"):
            eng.set_families(np.array([[0, 0, 0, -1]] * 3, np.int32))
        with pytest.raises(engine.CafeB200Error, match="COUNT_RANGE"):
            eng.infer([[0.03]], prior)
        eng.set_families(counts)
        again = eng.infer([[0.03]], prior)
        assert np.array_equal(again["family_lnl"], got["family_lnl"], equal_nan=True)
    with pytest.raises(engine.CafeB200Error, match="COUNT_RANGE"):
        engine.Engine(tree, np.array([[0, 0, 0, 26]], np.int32), 25, 20)
    # mrf > mf exercises N = max(mrf, mf) + 1
    want = orc.infer(tree, counts[:1] + 1, [[0.05]], [1.0], orc.prior_uniform(30), 12, 30, orc.BASE_LOGMAX)
    with engine.Engine(tree, counts[:1] + 1, 12, 30) as eng:
        got = eng.infer([[0.05]], orc.prior_uniform(30))
        assert rel_err(got["family_lnl"], want["family_lnl"]) < RTOL


def test_set_families_and_launch_accounting(mammal):
    flat, counts = mammal["tree"], mammal["counts"][:1000]
    prior = orc.prior_uniform(mammal["mrf"])
    with engine.Engine(flat, counts, mammal["mf"], mammal["mrf"]) as eng:
        assert eng.launches == 1                      # create: the device-side ingest (narrowing + range check) of the counts
        a = eng.infer([[0.002]], prior)
        assert eng.launches == 5                      # + matrix build, prune, finalize, final sum
        eng.set_families(counts[::-1].copy())
        b = eng.infer([[0.002]], prior)
        assert eng.launches == 10                     # + ingest of the new counts + the four kernels of an evaluation
        eng.set_families(counts[::-1].astype(np.uint8))     # counts may travel as one byte each
        c = eng.infer([[0.002]], prior)
        assert np.array_equal(c["family_lnl"], b["family_lnl"])
        assert np.array_equal(b["family_lnl"], a["family_lnl"][::-1])
        t = eng.last_timings_ms()
        assert t["prune"] > 0 and t["matrix_build"] > 0


def test_two_and_three_group_layouts_give_identical_results(mammal, monkeypatch):
    """The 32-family tile layout (two consumer groups, what matrix sizes above 160 use) computes the same numbers as
    the default 48-family one: same MMA tiles and the same order of every product."""
    flat, counts = mammal["tree"], mammal["counts"][:1500]
    mf, mrf = mammal["mf"], mammal["mrf"]
    freq, rate = orc.get_gamma(3, 0.6)
    lams = rate[:, None] * np.array([[0.0025]])
    prior = orc.prior_uniform(mrf)
    with engine.Engine(flat, counts, mf, mrf) as eng:
        two = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
        roots2 = eng.prune_roots(lams[:1])
        assert "groups=3" in eng.describe()
    monkeypatch.setenv("CAFE_B200_GEOM", "2,2,2")
    with engine.Engine(flat, counts, mf, mrf) as eng:
        assert "groups=2" in eng.describe()
        three = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
        roots3 = eng.prune_roots(lams[:1])
    assert np.array_equal(two["cat_lk"], three["cat_lk"], equal_nan=True)
    assert np.array_equal(roots2, roots3)
    assert two["score"] == three["score"] and two["n_failed"] == three["n_failed"]


def test_rescale_survives_subnormal_and_zero_intermediates():
    """Round-1 advisor finding: the renormalisation multiplied by 2^-e, which overflows when the row maximum is
    subnormal, and scaled dead rows.  A 70-leaf caterpillar with large counts drives the un-rescaled partial likelihoods
    through the subnormal range to zero (the reference arithmetic gives -inf); with CAFE_B200_OPT_RESCALE the likelihood
    must come out finite and equal to the same recursion carried out with explicit exponents in numpy."""
    n_leaves, mf, mrf, lam = 70, 60, 40, 0.0004
    newick = "L0:3"
    for i in range(1, n_leaves):
        newick = f"({newick},L{i}:{3 + i % 5}):2"
    tree = hostio.flatten_tree(hostio.parse_newick(newick + ";"))
    rng = np.random.default_rng(11)
    counts = rng.integers(25, 45, size=(37, tree.n_leaves)).astype(np.int32)
    prior = orc.prior_uniform(mrf, None, max(mf, mrf) + 1)
    n = max(mf, mrf) + 1
    mats = {v: orc.build_matrix(n, lam, tree.branch[v])[:, :mf + 1] for v in range(tree.n_nodes - 1)}

    def scaled_lnl(row):
        vec, exp = {}, {}
        for v in range(tree.n_nodes):
            kids = tree.child_list[tree.child_offset[v]:tree.child_offset[v + 1]]
            if len(kids) == 0:
                continue
            acc, e = np.ones(n), 0
            for c in kids:
                if tree.leaf_col[c] >= 0:
                    f = mats[c][:, row[tree.leaf_col[c]]]
                else:
                    f = mats[c] @ vec[c][:mf + 1]
                    e += exp[c]
                acc = acc * f
                if acc.max() > 0:
                    _, ex = np.frexp(acc.max())
                    acc, e = np.ldexp(acc, -ex), e + int(ex)     # exact power-of-two renormalisation after every factor
            vec[v], exp[v] = acc, e
        root = tree.n_nodes - 1
        with np.errstate(divide="ignore"):
            return float(np.max(np.log(vec[root][1:mrf + 1]) + np.log(prior[:mrf])) + exp[root] * np.log(2.0))

    want = np.array([scaled_lnl(r) for r in counts])
    assert np.isfinite(want).all() and want.min() < -750, "the case must leave the double range without rescaling"
    with engine.Engine(tree, counts, mf, mrf) as eng:
        plain = eng.infer([[lam]], prior)
        assert not np.isfinite(plain["family_lnl"]).all()          # reference arithmetic: underflow to zero, lnL = -inf
        eng.set_rescale(True)
        got = eng.infer([[lam]], prior)
        assert np.isfinite(got["family_lnl"]).all()
        assert rel_err(got["family_lnl"], want) < 1e-11
        eng.set_max_slots(2)                                      # parked products in device scratch instead of tensor memory
        again = eng.infer([[lam]], prior)
        assert np.array_equal(again["family_lnl"], got["family_lnl"])


def test_rescale_with_a_subnormal_row_maximum():
    """One node whose vector is already subnormal when it completes (a 34-leaf star with alternating counts 10 / 20:
    no parent size explains both, the largest product is ~1e-315): the renormalisation must scale it by ldexp (2^-e itself would
    overflow) and the likelihood must stay finite and close to the exactly-scaled recursion (the subnormal entries have
    lost low bits, so this is not a 1e-11 comparison)."""
    star = "(" + ",".join(f"S{i}:5" for i in range(34)) + "):3"
    tree = hostio.flatten_tree(hostio.parse_newick(f"({star},(A:4,B:4):2);"))
    mf, mrf, lam = 60, 40, 0.0005
    n = max(mf, mrf) + 1
    prior = orc.prior_uniform(mrf, None, n)
    names = tree.leaf_names
    base = np.array([[(10 if int(nm[1:]) % 2 == 0 else 20) if nm.startswith("S") else 15 for nm in names]], np.int32)
    counts = np.repeat(base, 20, axis=0)
    counts[:, [names.index("A"), names.index("B")]] += np.arange(20)[:, None] % 7
    mats = {v: orc.build_matrix(n, lam, tree.branch[v])[:, :mf + 1] for v in range(tree.n_nodes - 1)}
    star_node = tree.parent[[v for v in range(tree.n_nodes) if tree.names[v] == "S0"][0]]
    plain_star = np.ones(n)
    for c in tree.child_list[tree.child_offset[star_node]:tree.child_offset[star_node + 1]]:
        plain_star = plain_star * mats[c][:, counts[0, tree.leaf_col[c]]]
    assert 0 < plain_star.max() < 2.3e-308, f"the star's vector must be subnormal ({plain_star.max()})"

    def scaled_lnl(row):
        vec, exp = {}, {}
        for v in range(tree.n_nodes):
            kids = tree.child_list[tree.child_offset[v]:tree.child_offset[v + 1]]
            if len(kids) == 0:
                continue
            acc, e = np.ones(n), 0
            for c in kids:
                f = mats[c][:, row[tree.leaf_col[c]]] if tree.leaf_col[c] >= 0 else mats[c] @ vec[c][:mf + 1]
                e += 0 if tree.leaf_col[c] >= 0 else exp[c]
                acc = acc * f
                _, ex = np.frexp(acc.max())
                acc, e = np.ldexp(acc, -ex), e + int(ex)
            vec[v], exp[v] = acc, e
        root = tree.n_nodes - 1
        with np.errstate(divide="ignore"):
            return float(np.max(np.log(vec[root][1:mrf + 1]) + np.log(prior[:mrf])) + exp[root] * np.log(2.0))

    want = np.array([scaled_lnl(r) for r in counts])
    with engine.Engine(tree, counts, mf, mrf) as eng:
        eng.set_rescale(True)
        got = eng.infer([[lam]], prior)["family_lnl"]
    assert np.isfinite(got).all(), got
    assert np.max(np.abs(got - want)) < 1e-3 * np.max(np.abs(want)), (got[:4], want[:4])
