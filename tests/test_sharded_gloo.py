"""Host logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups.

The per-shard evaluator is the CPU oracle here (test infrastructure standing in for the CUDA engine, which has
no CPU fallback); what is under test is cafexp_b200/sharded.py: family ranges, the 2-double allreduce, the
+inf-if-any-shard-failed rule and the ordered gather of per-family outputs.
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cafexp_b200 import hostio, sharded


def test_shard_ranges_partition_the_families():
    for F in (0, 1, 7, 64, 1000003):
        for W in (1, 2, 3, 8):
            spans = [sharded.shard_range(F, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    tree = hostio.flatten_tree(hostio.parse_newick("((A:1,B:3):7,(C:11,(D:17,E:2.5):4):23);"))
    rng = np.random.default_rng(3)
    counts = rng.integers(0, 9, size=(37, tree.n_leaves)).astype(np.int32)
    counts[:, 0] = np.maximum(counts[:, 0], 1)
    counts[:, 2] = np.maximum(counts[:, 2], 1)
    return tree, counts, 30, 24


def _worker(rank, world, port, lam, q):
    from oracle import binding as orc
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        tree, counts, mf, mrf = _problem()
        lo, hi = sharded.shard_range(len(counts), rank, world)
        freq, rate = orc.get_gamma(3, 0.6)
        lams = rate[:, None] * np.array([[lam]])
        prior = orc.prior_uniform(mrf)
        local = {}

        def local_eval(lambdas, pr, cat_probs, mode, result):
            res = orc.infer(tree, counts[lo:hi], lambdas, cat_probs, pr, mf, mrf, mode)
            local["res"] = res
            ok = ~np.isnan(res["family_lnl"])
            result[0] = float(res["family_lnl"][ok].sum())
            result[1] = float(res["n_failed"])

        sl = sharded.ShardedLikelihood(local_eval, torch.zeros(2, dtype=torch.float64))
        score = sl.score(lams, prior, freq, orc.GAMMA_LINSUM)
        fam = sl.gather_family_values(local["res"]["family_lnl"], len(counts))
        cat = sl.gather_family_values(local["res"]["cat_lk"], len(counts))
        q.put((rank, score, None if fam is None else fam.tolist(), None if cat is None else cat.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,lam", [(2, 0.02), (3, 0.02), (2, 0.4)])
def test_sharded_score_equals_unsharded(world, lam):
    """Sum over shards == the single-process evaluation; every rank gets the same score; lam = 0.4 saturates
    (1 - 2*alpha < 0 on the long branches) so families fail and every rank must report +inf."""
    from oracle import binding as orc
    tree, counts, mf, mrf = _problem()
    freq, rate = orc.get_gamma(3, 0.6)
    lams = rate[:, None] * np.array([[lam]])
    want = orc.infer(tree, counts, lams, freq, orc.prior_uniform(mrf), mf, mrf, orc.GAMMA_LINSUM)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lam, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    for rank, score, fam, cat in got:
        if math.isinf(want["score"]):
            assert score == want["score"]
        else:
            assert abs(score - want["score"]) <= 1e-12 * abs(want["score"])
        if rank == 0:
            np.testing.assert_array_equal(np.asarray(fam), want["family_lnl"])
            np.testing.assert_array_equal(np.asarray(cat), want["cat_lk"])
        else:
            assert fam is None and cat is None
    if lam == 0.4:
        assert want["n_failed"] > 0
