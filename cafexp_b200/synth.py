"""Deterministic synthetic inputs of BASELINE.json config 5: a random ultrametric binary species tree
and gene families simulated forward along it with the linear birth-death process.

Only a DATA generator for benchmarks and parity tests (numpy, float64 log-space formula); it is not
part of the likelihood path and is never compared against anything.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.special import gammaln

from . import hostio


def random_ultrametric_newick(n_leaves: int, seed: int = 12345, height: float = 100.0) -> str:
    """Yule-like ultrametric binary tree; node heights are whole thousandths so branch lengths have 3 decimals."""
    rng = np.random.default_rng(seed)
    # coalescent construction backwards in time: lineages merge at increasing heights
    waits = rng.exponential(1.0 / np.arange(n_leaves, 1, -1))
    heights = np.cumsum(waits)
    heights = heights / heights[-1] * height
    h_int = np.maximum(np.round(heights * 1000).astype(np.int64), 1)
    for i in range(1, len(h_int)):                      # strictly increasing, at least 5 thousandths apart
        h_int[i] = max(h_int[i], h_int[i - 1] + 5)
    h_int[0] = max(h_int[0], 5)
    nodes = [(f"sp{i}", 0) for i in range(n_leaves)]    # (newick text, height in thousandths)
    for h in h_int:
        i, j = sorted(rng.choice(len(nodes), size=2, replace=False))
        a, b = nodes[i], nodes[j]
        text = f"({a[0]}:{(h - a[1]) / 1000.0:.3f},{b[0]}:{(h - b[1]) / 1000.0:.3f})"
        nodes = [n for idx, n in enumerate(nodes) if idx not in (i, j)] + [(text, int(h))]
    return nodes[0][0] + ";"


_LOGC_CACHE = {}


def _log_binomial_terms(n: int):
    """lnC(s,j) + lnC(s+c-1-j, s-1) and the exponent (s+c-2j) for s, c, j in 0..n-1 (independent of lambda, t)."""
    if n not in _LOGC_CACHE:
        s = np.arange(n)[:, None, None].astype(float)
        c = np.arange(n)[None, :, None].astype(float)
        j = np.arange(n)[None, None, :].astype(float)
        valid = (j <= np.minimum(s, c)) & (s >= 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            logc = (gammaln(s + 1) - gammaln(j + 1) - gammaln(s - j + 1) + gammaln(s + c - j) - gammaln(s) - gammaln(c - j + 1))
        logc = np.where(valid, logc, -np.inf)
        _LOGC_CACHE[n] = (logc, s + c - 2 * j, j)
    return _LOGC_CACHE[n]


def bd_transition_matrix(lam: float, t: float, n: int) -> np.ndarray:
    """P(s -> c), s, c in 0..n-1, rows renormalised over the truncated range (generator use only)."""
    alpha = lam * t / (1.0 + lam * t)
    coeff = 1.0 - 2.0 * alpha
    logc, power, j = _log_binomial_terms(n)
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = np.exp(logc + power * np.log(alpha) + j * np.log(abs(coeff)))
    if coeff < 0:
        terms = np.where(j % 2 == 1, -terms, terms)
    m = np.clip(terms.sum(axis=2), 0.0, 1.0)
    m[0, :] = 0.0
    m[0, 0] = 1.0
    m /= m.sum(axis=1, keepdims=True)
    return m


GEN_CHUNK = 65536


def _branch_cdfs(tree: hostio.FlatTree, lam: float, state_cap: int):
    mats = {}
    for v in range(tree.n_nodes - 1):
        key = round(float(tree.branch[v]), 6)
        if key not in mats:
            cdf = np.cumsum(bd_transition_matrix(lam, key, state_cap), axis=1)
            cdf[:, -1] = 1.0
            # one sorted array for all rows: row s occupies (2s, 2s+1]
            mats[key] = (cdf + 2.0 * np.arange(state_cap)[:, None]).ravel()
    return mats


def _simulate_chunk(tree, mats, rng, want: int, root_mean: float, max_count: int, state_cap: int) -> np.ndarray:
    n = tree.n_nodes
    leaves = np.flatnonzero(tree.leaf_col >= 0)
    out = np.empty((want, tree.n_leaves), np.int32)
    filled = 0
    while filled < want:
        m = max(1024, int((want - filled) * 1.1) + 16)
        state = np.zeros((n, m), np.int32)
        state[n - 1] = np.minimum(1 + rng.poisson(root_mean, m), max_count)
        for v in range(n - 2, -1, -1):
            parent = state[tree.parent[v]]
            flat = mats[round(float(tree.branch[v]), 6)]
            idx = np.searchsorted(flat, rng.random(m) + 2.0 * parent, side="left")
            state[v] = np.minimum(idx - parent * state_cap, state_cap - 1)
        counts = np.empty((m, tree.n_leaves), np.int32)
        counts[:, tree.leaf_col[leaves]] = state[leaves].T
        ok = (counts.max(axis=1) <= max_count) & hostio.exists_at_root(tree, counts)
        counts = counts[ok]
        take = min(len(counts), want - filled)
        out[filled:filled + take] = counts[:take]
        filled += take
    return out


def simulate_families(tree: hostio.FlatTree, n_families: int, lam: float = 0.005, seed: int = 12345, root_mean: float = 10.0,
                      max_count: int = 100, state_cap: int = 151, first: int = 0, last: int = -1) -> np.ndarray:
    """int32 counts [last-first, n_leaves] = families first..last-1 of the n_families-family data set.

    Every leaf count <= max_count and every family is present on both sides of the root (the reference's
    default filter, src/cafexp.cpp:189-199).  Families are generated in independent chunks of GEN_CHUNK
    (seeded by (seed, chunk index)), so any rank can produce exactly its shard of the same global data set."""
    if last < 0:
        last = n_families
    mats = _branch_cdfs(tree, lam, state_cap)
    parts = []
    for chunk in range(first // GEN_CHUNK, (max(last, first + 1) - 1) // GEN_CHUNK + 1):
        lo, hi = chunk * GEN_CHUNK, min((chunk + 1) * GEN_CHUNK, n_families)
        rng = np.random.default_rng([seed, chunk])
        block = _simulate_chunk(tree, mats, rng, hi - lo, root_mean, max_count, state_cap)
        parts.append(block[max(first, lo) - lo:min(last, hi) - lo])
    return np.concatenate(parts) if parts else np.zeros((0, tree.n_leaves), np.int32)


def config5(n_families: int, n_leaves: int = 100, seed: int = 12345, lam: float = 0.005, first: int = 0, last: int = -1
            ) -> Tuple[hostio.FlatTree, np.ndarray, str]:
    """(tree, counts[first:last], newick) of the synthetic benchmark; max_family_size=150, max_root_family_size=125."""
    newick = random_ultrametric_newick(n_leaves, seed)
    tree = hostio.flatten_tree(hostio.parse_newick(newick))
    counts = simulate_families(tree, n_families, lam=lam, seed=seed, first=first, last=last)
    return tree, counts, newick


CONFIG5_MAX_FAMILY_SIZE = 150
CONFIG5_MAX_ROOT_FAMILY_SIZE = 125
