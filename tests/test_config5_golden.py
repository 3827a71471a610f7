"""BASELINE.json config 5 against the compiled reference itself: tests/golden/config5_slice.npz holds
the reference's category likelihoods of the first 10 000 synthetic families and its Pupko reconstruction of the first
1 000 (scripts/make_golden_config5.py; the inputs are regenerated from the seed).

CPU half: the C restatement (oracle) against those vectors on a subset.  GPU half: the CUDA path, every family, through
the C ABI — per-family likelihoods at 1e-11 relative, reconstructed counts exact — on one device and, when two are
visible, sharded over two (cafe_b200_create_multi)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from cafexp_b200 import engine, synth
from oracle import binding as orc

LAMBDA, ALPHA, K = 0.005, 0.7, 4
MF, MRF = synth.CONFIG5_MAX_FAMILY_SIZE, synth.CONFIG5_MAX_ROOT_FAMILY_SIZE


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLD, "config5_slice.npz"))
    meta = json.loads(str(z["meta"]))
    tree, counts, _ = synth.config5(meta["families_total"], first=0, last=meta["n_eval"])
    assert tree.internal_names == meta["internal_order"], "the reference visits internal nodes in the flattened tree's order"
    freq, rate = orc.get_gamma(K, ALPHA)
    np.testing.assert_allclose(rate, meta["multipliers"], rtol=1e-15)
    return {"tree": tree, "counts": counts, "cat_lk": z["cat_lk"], "score": float(z["score"]), "states": z["states"].astype(np.int32),
            "lams": rate[:, None] * np.array([[LAMBDA]]), "freq": freq, "meta": meta}


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def test_golden_slice_is_self_consistent(gold):
    cat = gold["cat_lk"]
    assert cat.shape == (10000, K) and np.isfinite(cat).all() and (cat > 0).all()
    assert abs(-np.log(cat.sum(axis=1)).sum() - gold["score"]) <= 1e-12 * gold["score"]
    assert gold["states"].shape == (1000, K, 99)


def test_oracle_matches_reference_on_config5(gold):
    """The plain-C restatement against the reference on the config-5 shape (every 40th family: seconds on the CPU)."""
    sel = np.arange(0, 10000, 40)
    prior = orc.prior_uniform(MRF, None, MF + 1)
    got = orc.infer(gold["tree"], gold["counts"][sel], gold["lams"], gold["freq"], prior, MF, MRF, orc.GAMMA_LINSUM)
    assert got["n_failed"] == 0
    assert rel(got["cat_lk"], gold["cat_lk"][sel]) < 1e-11
    sub = np.arange(0, 1000, 50)
    states = orc.reconstruct(gold["tree"], gold["counts"][sub], gold["lams"], prior, MF, MRF)
    assert np.array_equal(states, gold["states"][sub])


def _check_device_path(gold, device):
    prior = orc.prior_uniform(MRF, None, MF + 1)
    counts = gold["counts"]
    with engine.Engine(gold["tree"], counts.astype(np.uint8), MF, MRF, device=device) as eng:
        got = eng.infer(gold["lams"], prior, gold["freq"], engine.GAMMA_LINSUM)
        assert got["n_failed"] == 0
        assert rel(got["cat_lk"], gold["cat_lk"]) < 1e-11                              # all 10 000 families x 4 categories
        assert rel(got["family_lnl"], np.log(gold["cat_lk"].sum(axis=1))) < 1e-11
        assert abs(got["score"] - gold["score"]) <= 1e-11 * gold["score"]
        assert np.array_equal(eng.fetch_category_likelihoods(K), got["cat_lk"])
        desc = eng.describe()
    with engine.Engine(gold["tree"], counts[:1000], MF, MRF, device=device) as eng:
        states = eng.reconstruct(gold["lams"], prior)
        assert np.array_equal(states, gold["states"])                                   # 1 000 families x 4 x 99 nodes, exact
    return got, desc


@pytest.mark.gpu
def test_cuda_matches_reference_on_config5(gold):
    got, desc = _check_device_path(gold, 0)
    assert "groups=3" in desc and "tmem_entries=4" in desc and "spill_entries=0" in desc


@pytest.mark.gpu
def test_cuda_two_devices_match_reference_and_one_device(gold):
    """The same 10 000 families sharded over two devices behind ONE context and one host thread."""
    if engine.device_count() < 2:
        pytest.skip("needs two visible CUDA devices")
    one, _ = _check_device_path(gold, 0)
    two, desc = _check_device_path(gold, [0, 1])
    assert "devices=2" in desc
    assert np.array_equal(one["cat_lk"], two["cat_lk"]) and np.array_equal(one["family_lnl"], two["family_lnl"])
    assert abs(one["score"] - two["score"]) <= 1e-13 * abs(one["score"])                # two partial sums instead of one
