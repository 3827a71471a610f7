"""Randomised differential test of the CUDA path against the CPU oracle: random tree shapes (binary and n-ary,
balanced and caterpillar), matrix sizes across every row-block count (N = 6 .. 250, and 261 .. 512 for the likelihood), rate categories, multiple
lambdas, error models with 3 and 5 deviations, shared-memory slot limits that force spills, ragged family counts,
power-of-two rescaling.  Every case checks evaluation, root vectors, the p-value statistic and the Pupko
reconstruction.  Needs a B200."""
import os

import numpy as np
import pytest

from cafexp_b200 import engine, hostio
from oracle import binding as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def random_newick(rng, n_leaves, shape, max_children):
    """Random rooted tree with named leaves; branch lengths with <= 3 decimals (quantisation is a no-op) or raw."""
    nodes = [f"L{i}" for i in range(n_leaves)]
    rng.shuffle(nodes)

    def bl():
        t = float(rng.choice([rng.uniform(0.5, 40.0), rng.integers(1, 60), rng.uniform(0.001, 2.0)]))
        return f"{t:.3f}" if rng.random() < 0.7 else repr(t)

    while len(nodes) > 1:
        k = int(rng.integers(2, max_children + 1)) if max_children > 2 else 2
        k = min(k, len(nodes))
        if shape == "caterpillar":
            pick = [len(nodes) - 1] + list(range(k - 1))                 # keep extending the same lineage
        else:
            pick = list(rng.choice(len(nodes), size=k, replace=False))
        kids = [nodes[i] for i in pick]
        for i in sorted(pick, reverse=True):
            nodes.pop(i)
        nodes.append("(" + ",".join(f"{c}:{bl()}" for c in kids) + ")")
    return nodes[0] + ";"


def make_case(seed, big=False):
    rng = np.random.default_rng(seed)
    n_leaves = int(rng.choice([250, 400])) if big else int(rng.choice([2, 3, 5, 8, 13, 24, 40]))
    shape = str(rng.choice(["random", "caterpillar"]))
    max_children = int(rng.choice([2, 2, 3, 4]))
    flat = hostio.flatten_tree(hostio.parse_newick(random_newick(rng, n_leaves, shape, max_children)))
    mf = int(rng.choice([12, 31, 40])) if big else int(rng.choice([5, 20, 31, 32, 63, 64, 100, 127, 159, 160, 200, 249]))
    mrf = int(np.clip(mf + rng.integers(-mf // 2, 6), 1, 249))
    n_lambdas = int(rng.choice([1, 1, 2, 3]))
    flat.lambda_index[:] = rng.integers(0, n_lambdas, size=flat.n_nodes)
    F = int(rng.choice([5, 33, 70])) if big else int(rng.choice([1, 7, 16, 33, 100, 257]))
    ndev = int(rng.choice([0, 0, 3, 5]))
    hi = max(1, mf - (ndev - 1) // 2 - 1 if ndev else mf)
    mean_hi = 2.0 if big else min(30, hi)          # many leaves: keep the likelihood inside the double range
    counts = np.minimum(rng.poisson(rng.uniform(0.5, mean_hi), size=(F, flat.n_leaves)), hi).astype(np.int32)
    if rng.random() < 0.3:
        counts[rng.integers(0, F)] = 0                                    # an all-zero family
    if rng.random() < 0.3:
        counts[rng.integers(0, F)] = hi                                   # the largest allowed counts
    k = int(rng.choice([1, 1, 2, 4, 7]))
    lam = rng.uniform(0.0005, 0.004 if big else 0.03, size=n_lambdas)
    if k > 1:
        freq, rate = orc.get_gamma(k, float(rng.uniform(0.2, 2.0)))
    else:
        freq, rate = np.ones(1), np.ones(1)
    lams = rate[:, None] * lam[None, :]
    err = None
    if ndev:
        rows = int(counts.max()) + 1
        err = rng.uniform(0.0, 1.0, size=(rows, ndev))
        err /= err.sum(axis=1, keepdims=True)
        err[0, : (ndev - 1) // 2] = 0.0                                   # no deviation below zero
    return dict(flat=flat, mf=mf, mrf=mrf, counts=counts, k=k, lams=lams, freq=freq, err=err,
                slots=int(rng.choice([0, 0, 2, 3])), rescale=bool(rng.random() < 0.3), n_leaves=n_leaves, shape=shape)


@pytest.mark.parametrize("seed", range(1000, 1006))
def test_random_big_tree_against_oracle(seed):
    """250 / 400 leaves: the per-tile counts no longer fit the shared-memory staging area (counts read from global
    memory), the op list has thousands of entries, the schedule spills."""
    check_case(make_case(seed, big=True))


def make_large_matrix_case(seed):
    """Matrix sizes 257 .. 512 (round 1 stopped at 256): the likelihood runs with eight warps per group and 64-row
    blocks, leaf counts above 255 travel and live on the device as uint16, lgamma arguments pass the reference's
    1024-entry table (src/probability.cpp:58-64: libm beyond it)."""
    rng = np.random.default_rng(seed)
    n_leaves = int(rng.choice([3, 4, 6]))
    flat = hostio.flatten_tree(hostio.parse_newick(random_newick(rng, n_leaves, str(rng.choice(["random", "caterpillar"])), int(rng.choice([2, 2, 3])))))
    mf = int(rng.choice([260, 299, 319, 320, 383, 400, 447, 480, 511]))
    mrf = int(np.clip(mf + rng.integers(-mf // 2, 1), 1, 511))
    n_lambdas = int(rng.choice([1, 2]))
    flat.lambda_index[:] = rng.integers(0, n_lambdas, size=flat.n_nodes)
    F = int(rng.choice([3, 17, 33]))
    ndev = int(rng.choice([0, 0, 3]))
    hi = mf - 2
    counts = np.minimum(rng.poisson(rng.uniform(5.0, 120.0), size=(F, flat.n_leaves)), hi).astype(np.int32)
    counts[rng.integers(0, F), rng.integers(0, flat.n_leaves)] = hi            # a count above 255: two bytes per count
    k = int(rng.choice([1, 2, 3]))
    lam = rng.uniform(0.0005, 0.01, size=n_lambdas)
    freq, rate = orc.get_gamma(k, float(rng.uniform(0.3, 2.0))) if k > 1 else (np.ones(1), np.ones(1))
    err = None
    if ndev:
        err = rng.uniform(0.0, 1.0, size=(int(counts.max()) + 1, ndev))
        err /= err.sum(axis=1, keepdims=True)
        err[0, : (ndev - 1) // 2] = 0.0
    return dict(flat=flat, mf=mf, mrf=mrf, counts=counts, k=k, lams=rate[:, None] * lam[None, :], freq=freq, err=err, slots=int(rng.choice([0, 2])),
                rescale=bool(rng.random() < 0.3), n_leaves=n_leaves, shape="large-matrix")


@pytest.mark.parametrize("seed", range(2000, 2006))
def test_matrix_sizes_above_256_against_oracle(seed):
    c = make_large_matrix_case(seed)
    check_case(c)
    with engine.Engine(c["flat"], c["counts"], c["mf"], c["mrf"]) as eng:
        assert "gw=8" in eng.describe() and "count_bytes=2" in eng.describe()


@pytest.mark.parametrize("seed", range(int(os.environ.get("CAFE_B200_FUZZ_SEEDS", "40"))))
def test_random_case_against_oracle(seed):
    check_case(make_case(seed))


def check_case(c):
    flat, mf, mrf, counts, k, lams = c["flat"], c["mf"], c["mrf"], c["counts"], c["k"], c["lams"]
    prior = orc.prior_uniform(mrf, None, max(mf, mrf) + 1)
    mode = orc.GAMMA_LINSUM if k > 1 else orc.BASE_LOGMAX
    want = orc.infer(flat, counts, lams, c["freq"], prior, mf, mrf, mode, err=c["err"])
    with engine.Engine(flat, counts, mf, mrf) as eng:
        if c["slots"]:
            eng.set_max_slots(c["slots"])
        eng.set_rescale(c["rescale"])
        if c["err"] is not None:
            eng.set_error_model(c["err"])
        got = eng.infer(lams, prior, c["freq"], engine.GAMMA_LINSUM if k > 1 else engine.BASE_LOGMAX, failed_cap=len(counts))
        if not c["rescale"]:
            # without rescaling underflow happens exactly where the reference's does: same failures, same NaNs
            assert got["n_failed"] == want["n_failed"]
            assert np.array_equal(np.isnan(got["family_lnl"]), np.isnan(want["family_lnl"]))
        ok = ~np.isnan(want["family_lnl"]) & ~np.isnan(got["family_lnl"]) & np.isfinite(want["family_lnl"])
        assert np.allclose(got["family_lnl"][ok], want["family_lnl"][ok], rtol=RTOL, atol=0), c
        if k > 1 and not c["rescale"]:
            okc = ~np.isnan(want["cat_lk"]).any(axis=1)
            assert np.allclose(got["cat_lk"][okc], want["cat_lk"][okc], rtol=RTOL, atol=0)
        if np.isfinite(want["score"]) and not c["rescale"]:
            assert abs(got["score"] - want["score"]) <= RTOL * abs(want["score"])
        if c["err"] is None:
            # root vectors of the first category, and the p-value statistic under the first lambda set
            roots = eng.prune_roots(lams[:1])[:, 0, :]
            n_ref = 12 if max(mf, mrf) < 256 else 2        # the oracle rebuilds every matrix per call: O(N^3) each above 256
            ref = np.stack([orc.inference_prune(flat, counts[i], lams[0], mf, mrf) for i in range(min(len(counts), n_ref))])
            assert np.allclose(roots[: len(ref)], ref, rtol=RTOL, atol=1e-300)
            assert np.allclose(eng.root_max(lams[0]), orc.root_max(flat, counts, lams[0], mf, mrf), rtol=RTOL, atol=1e-300)
            # Pupko reconstruction (ignores the error model in the reference too): exact
            prior_sz = orc.prior_uniform(mrf, None, min(mf, mrf) + 1)
            if max(mf, mrf) + 1 <= 256:
                assert np.array_equal(eng.reconstruct(lams, prior_sz), orc.reconstruct(flat, counts, lams, prior_sz, mf, mrf)), c
            else:
                # the reconstruction kernel keeps 32-family tiles and one-byte argmax tables: matrix sizes up to 256
                with pytest.raises(engine.CafeB200Error, match="ERR_LIMIT"):
                    eng.reconstruct(lams, prior_sz)
