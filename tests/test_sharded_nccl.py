"""The multi-process path on hardware: two ranks, one GPU each, NCCL — cafexp_b200/sharded.py over the CUDA engine.

What is asserted (the CPU gloo tests cover the host logic with the oracle as evaluator): the all-reduced score of the
sharded evaluation equals the single-GPU score of the same families, WITHOUT the caller binding the engine to a stream
by hand (engine_local_eval binds it to torch's current stream, which is the stream NCCL orders against)."""
import os
import socket

import numpy as np
import pytest

from cafexp_b200 import engine, sharded, synth

pytestmark = pytest.mark.gpu

LAMBDA, ALPHA, K = 0.005, 0.7, 4
MF, MRF = synth.CONFIG5_MAX_FAMILY_SIZE, synth.CONFIG5_MAX_ROOT_FAMILY_SIZE
N_FAMILIES = 20000


def _inputs():
    from cafexp_b200 import params
    freq, rate = params.get_gamma(K, ALPHA)
    return np.ascontiguousarray(rate[:, None] * np.array([[LAMBDA]])), freq, params.prior_uniform(MRF, None, MRF)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        lo, hi = sharded.shard_range(N_FAMILIES, rank, world)
        tree, counts, _ = synth.config5(N_FAMILIES, first=lo, last=hi)
        lams, freq, prior = _inputs()
        result = torch.zeros(2, dtype=torch.float64, device=f"cuda:{rank}")
        with engine.Engine(tree, counts.astype(np.uint8), MF, MRF, device=rank) as eng:
            job = sharded.ShardedLikelihood(sharded.engine_local_eval(eng), result)
            scores = []
            for distributed_build in (False, True):
                # replicated matrix build, then: each rank builds half of the matrices, one NCCL all-gather hands them over
                eng.set_build_partition(rank, world, sharded.nccl_matrix_gather()) if distributed_build else eng.set_build_partition(0, 1)
                for stream in (torch.cuda.current_stream(), torch.cuda.Stream()):       # default stream, then a side stream
                    with torch.cuda.stream(stream):
                        for _ in range(3):
                            scores.append(job.score(lams, prior, freq, engine.GAMMA_LINSUM))
        q.put((rank, scores))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_score_equals_one_gpu():
    if engine.device_count() < 2:
        pytest.skip("needs two visible CUDA devices")
    import torch.multiprocessing as mp
    tree, counts, _ = synth.config5(N_FAMILIES)
    lams, freq, prior = _inputs()
    with engine.Engine(tree, counts, MF, MRF, device=0) as eng:
        want = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)["score"]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scores in got:
        for sc in scores:
            assert abs(sc - want) <= 1e-12 * abs(want), (rank, sc, want)
