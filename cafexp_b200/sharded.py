"""Families sharded over the GPUs of one box: one process per GPU, one 2-double allreduce per evaluation.

Gene families are independent given (tree, lambda, alpha, epsilon, prior) — the reference exploits exactly this
with ``#pragma omp parallel for`` over families (src/base_model.cpp:81,89, src/gamma_core.cpp:201) — so the
family axis is the only one that is partitioned: rank r owns the contiguous range
``[r*F//W, (r+1)*F//W)`` of the count matrix for the lifetime of the context.  Every rank rebuilds the (small)
transition matrices redundantly, prunes its shard, and leaves ``[sum_i lnL_i, n_failed]`` in device memory
(``cafe_b200_eval_device``); one sum-allreduce of those two doubles over NCCL (NVLink 5 / NVSwitch) gives every
rank the score the optimizer needs: ``-sum`` or ``+inf`` if any family anywhere failed
(src/gamma_core.cpp:227-236).  There is no exchange step inside the tree recursion, hence no other collective.
Per-family outputs stay sharded and are gathered only when the caller asks (final compute / reconstruction).

The local evaluator is injectable so that the host logic (ranges, reduction, +inf semantics, gathers) is covered
by world_size-2 ``gloo`` tests on CPU; the product path is :class:`cafexp_b200.engine.Engine` on ``cuda:LOCAL_RANK``.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n_families: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous family range of ``rank``; ranges are disjoint, ordered and cover [0, n_families)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * n_families // world, (rank + 1) * n_families // world


class ShardedLikelihood:
    """One evaluation = local shard evaluation + one allreduce of ``[sum lnL, n_failed]``.

    ``local_eval(lambdas, prior, cat_probs, mode, result)`` must leave the shard's ``[sum_i lnL_i over non-failed
    families, number of failed families]`` in the 2-element float64 tensor ``result`` (asynchronously, on the
    current stream, for the CUDA engine).
    """

    def __init__(self, local_eval: Callable, result_tensor, group=None):
        import torch.distributed as dist
        self._dist = dist
        self._local_eval = local_eval
        self.result = result_tensor
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def enqueue(self, lambdas, prior, cat_probs, mode) -> None:
        """Asynchronous: after this, ``self.result`` holds the global pair on every rank (stream-ordered on CUDA)."""
        self._local_eval(lambdas, prior, cat_probs, mode, self.result)
        if self.world > 1:
            nvtx = getattr(getattr(__import__("torch"), "cuda", None), "nvtx", None) if self.result.is_cuda else None
            if nvtx:
                nvtx.range_push("cafe_b200: allreduce [sum lnL, n_failed]")
            self._dist.all_reduce(self.result, op=self._dist.ReduceOp.SUM, group=self.group)
            if nvtx:
                nvtx.range_pop()

    def score(self, lambdas, prior, cat_probs, mode) -> float:
        """-lnL of ALL families, identical on every rank; +inf if any family on any rank failed."""
        self.enqueue(lambdas, prior, cat_probs, mode)
        total, failed = self.result.cpu().tolist()         # synchronises
        return math.inf if failed > 0 or math.isnan(total) else -total

    def gather_family_values(self, local_values: np.ndarray, n_families: int, dst: int = 0) -> Optional[np.ndarray]:
        """Per-family outputs (lnL, category likelihoods, reconstructed states) of all shards in family order on
        rank ``dst``; ``None`` elsewhere.  Shards have different lengths, so this is a gather of padded rows."""
        import torch
        if self.world == 1:
            return np.asarray(local_values)
        local = np.ascontiguousarray(local_values)
        width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
        longest = max(shard_range(n_families, r, self.world)[1] - shard_range(n_families, r, self.world)[0] for r in range(self.world))
        dev = self.result.device
        t_local = torch.from_numpy(local.reshape(len(local), width))
        buf = torch.zeros((longest, width), dtype=t_local.dtype, device=dev)
        if len(local):
            buf[: len(local)] = t_local.to(dev)
        parts = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == dst else None
        self._dist.gather(buf, parts, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        rows = []
        for r, part in enumerate(parts):
            lo, hi = shard_range(n_families, r, self.world)
            rows.append(part[: hi - lo].cpu().numpy())
        out = np.concatenate(rows, axis=0)
        return out.reshape((n_families,) + local.shape[1:])


class _DevicePointer:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def nccl_matrix_gather(group=None) -> Callable:
    """The exchange step of the distributed matrix build (cafe_b200_set_build_partition): every rank has built one slab
    of the evaluation's transition matrices into its own buffer; ONE in-place NCCL all-gather over NVLink / NVSwitch hands
    every slab to every rank (config 5, 8 ranks: 32 MB sent per rank instead of 2.4 ms of redundant FP64 work on every
    GPU).  Runs on torch's current stream, which is the stream the engine is bound to."""
    import torch
    import torch.distributed as dist
    views = {}

    def gather(matrices, slab_bytes, n_parts, stream):
        if torch.cuda.current_stream().cuda_stream != stream:
            raise RuntimeError("the engine is not bound to torch's current stream")
        rank = dist.get_rank(group)
        if dist.get_world_size(group) != n_parts:
            raise RuntimeError("build partition and process group disagree")
        key = (matrices, slab_bytes, n_parts)
        if key not in views:
            views[key] = torch.as_tensor(_DevicePointer(matrices, slab_bytes * n_parts), device=torch.device("cuda", torch.cuda.current_device()))
        full = views[key]
        dist.all_gather_into_tensor(full, full[rank * slab_bytes:(rank + 1) * slab_bytes], group=group)

    return gather


def engine_local_eval(eng) -> Callable:
    """The product path: ``cafe_b200_eval_device``, result left in device memory.

    The engine's kernels must be ordered against the NCCL all-reduce that follows, and NCCL orders against torch's
    CURRENT stream only.  So the engine is bound to that stream before every evaluation (a no-op when it already is):
    the library's kernels, the collective and the caller's timing events then share one stream and no event plumbing
    is needed."""
    bound = {"stream": None}

    def run(lambdas, prior, cat_probs, mode, result):
        import torch
        if not result.is_cuda:
            raise RuntimeError("the CUDA engine writes its result to device memory")
        stream = torch.cuda.current_stream(result.device).cuda_stream
        if bound["stream"] != stream:
            eng.set_stream(stream)
            bound["stream"] = stream
        eng.infer_device(lambdas, prior, cat_probs, mode, result.data_ptr())
    return run
