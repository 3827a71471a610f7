#!/usr/bin/env python
"""Benchmark of the family-likelihood hot path on B200 (BASELINE.json metric: family-likelihood evals/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--families F] [--impl reference]

Workload (config 5 of BASELINE.json): synthetic F = 1 000 000 gene families on a 100-taxon ultrametric tree,
max_family_size 150 (Nmax), max_root_family_size 125, gamma k = 4 (alpha 0.7), lambda 0.005, uniform root prior.
A "step" is ONE likelihood evaluation of all F families under one parameter point, i.e. what the Nelder-Mead
optimiser asks for per simplex vertex: transition matrices for every (branch, lambda, category) are rebuilt,
every family is pruned for every category, root prior / category weights applied, sum reduced (and summed
across ranks with one 2-double NCCL allreduce when N > 1).  Families are sharded across ranks (F/N each), so
total work is fixed: "scaling": "strong".

value      families/s with the count matrix already resident in HBM (device-timed, max over ranks)
e2e        families/s through the C ABI with HOST buffers: every step uploads the count matrix from pinned host
           memory (cafe_b200_set_families), evaluates (cafe_b200_eval) and reads back the score and all per-family
           log-likelihoods and category likelihoods.
roofline   FP64 tensor (DMMA) roofline of the pruning kernel: algorithmic FLOPs (internal edges only, leaf edges
           count 0) / its CUDA-event duration, against the FP64 peak measured on this box
           by scripts/fp64_peak.cu (MEASURED_PEAKS.json carries no FP64 figure).
cpu_baseline  the reference's own CPU path (oracle/_ref/ref_harness = unmodified reference objects; falls back to
           the C port oracle/liboracle.so) timed on this host on a bounded slice of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LAMBDA = 0.005
ALPHA = 0.7
K = 4
MF = 150
MRF = 125
N_LEAVES = 100
SEED = 12345


def algorithmic_flops_per_family_category(tree, mf, mrf):
    """2*rows*(mf+1) per edge whose child is internal; rows = mf+1 under a non-root parent, mrf under the root."""
    flops = 0
    root = tree.n_nodes - 1
    for v in range(tree.n_nodes - 1):
        if tree.leaf_col[v] >= 0:
            continue
        rows = mrf if tree.parent[v] == root else mf + 1
        flops += 2 * rows * (mf + 1)
    return flops


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.dev = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.dev)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8 or not (t0 - 0.1 <= ts <= t1 + 0.3):
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def gamma_parameters():
    """Discrete-gamma multipliers, category probabilities and the uniform root prior: host-side scalar math of the
    reference (get_gamma, src/gamma.cpp:225; uniform_distribution, src/root_equilibrium_distribution.cpp:20-32),
    from the product's own host mirror."""
    from cafexp_b200 import params
    freq, rate = params.get_gamma(K, ALPHA)
    prior = params.prior_uniform(MRF, None, MRF)
    return freq, rate, prior


def cpu_reference_rate(tree, newick, counts, target_seconds=15.0, threads=None, n_full=None, repeats=1, timed=1):
    """families/s of the reference CPU implementation on a bounded slice (kind, cores, sample, value)."""
    from cafexp_b200 import hostio
    from oracle import binding as orc
    cores = threads or os.cpu_count() or 1
    freq, rate, prior = gamma_parameters()

    def run_ref(n):
        with tempfile.TemporaryDirectory() as tmp:
            tpath = os.path.join(tmp, "tree.txt")
            fpath = os.path.join(tmp, "fam.txt")
            open(tpath, "w").write(newick + "\n")
            hostio.write_gene_families(fpath, tree, [str(i) for i in range(n)], counts[:n])
            r = orc.run_ref("eval", threads=cores, tree=tpath, fam=fpath, filter=0, k=K, alpha=ALPHA, maxfam=MF, maxroot=MRF, reps=1,
                            **{"lambda": LAMBDA})
            return r["seconds_best"], r["score"]

    def run_port(n):
        os.environ["OMP_NUM_THREADS"] = str(cores)
        t0 = time.perf_counter()
        res = orc.infer(tree, counts[:n], rate[:, None] * np.array([[LAMBDA]]), freq, prior, MF, MRF, orc.GAMMA_LINSUM)
        return time.perf_counter() - t0, res["score"]

    kind, run = ("reference", run_ref) if orc.have_ref() else ("port", run_port)
    n_small = min(len(counts), 32)
    try:
        t_small, _ = run(n_small)
    except Exception:                              # harness missing libs etc.: fall back to the port
        kind, run = "port", run_port
        t_small, _ = run(n_small)
    # The reference rebuilds all k x edges transition matrices inside every evaluation (src/gamma_core.cpp:196-197):
    # time(F) = a (matrices) + b * F (pruning).  Two sample sizes separate a and b; the figure reported is the
    # throughput that model gives for the FULL workload, families / (a + b * families).
    n_mid = min(len(counts), max(8 * n_small, 16 * cores))
    t_mid, _ = run(n_mid)
    b = max((t_mid - t_small) / max(n_mid - n_small, 1), 1e-9)
    a = max(t_small - b * n_small, 0.0)
    n_big = int(min(len(counts), 8 * n_mid, max(n_mid, (target_seconds - a) / b)))
    full = n_full if n_full else len(counts)
    values, t_big, score = [], t_mid, None
    for _ in range(max(1, repeats)):
        if n_big > n_mid:
            t_big, score = run(n_big)
            b = max((t_big - t_small) / (n_big - n_small), 1e-9)
            a = max(t_small - b * n_small, 0.0)
        values.append(full / (a + b * full))
    return {"value": float(np.mean(values[-max(1, timed):])), "unit": "families/s", "cores": cores, "kind": kind,
            "sample": f"infer_family_likelihoods timed on the first {n_small} and {max(n_big, n_mid)} config-5 families (k={K}): {t_small:.2f} s and {t_big:.2f} s "
                      f"=> {a:.2f} s per evaluation for the {K}x198 transition matrices + {b * 1e3:.3f} ms per family; value = {full} / (a + b*{full}), "
                      f"i.e. linear extrapolation to the full workload (families are independent)",
            "score": score, "fixed_s": a, "per_family_s": b, "n_sample": max(n_big, n_mid), "values": values}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cafexp_b200 import synth
    tree, counts, newick = synth.config5(args.families, N_LEAVES, SEED, LAMBDA, first=0, last=min(args.families, 8192))
    # calibration (two small runs) once, then W untimed + K timed runs of the bounded sample
    base = cpu_reference_rate(tree, newick, counts, target_seconds=8.0, n_full=args.families, repeats=args.warmup + args.steps, timed=args.steps)
    value = base["value"]
    n_sample = args.families
    line = {"impl": "reference", "metric": "family-likelihood evals/sec", "value": value, "unit": "families/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_sample / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": "families/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["value"] = value
    emit(line)


def workload_config(args, world):
    return {"workload": f"BASELINE.json configs[4]: synthetic {args.families} families x {N_LEAVES}-taxon ultrametric tree, Nmax={MF}, "
                        f"max_root_family_size={MRF}, gamma k={K} alpha={ALPHA}, lambda={LAMBDA}, uniform root prior",
            "families": args.families, "taxa": N_LEAVES, "matrix_size": MF + 1, "gamma_categories": K,
            "parallelism": f"families sharded over {world} GPU(s), one 2-double NCCL allreduce per evaluation" if world > 1 else "1 GPU",
            "l2": "L2 flushed (256 MiB write) between timed steps", "seed": SEED}


def measure_fp64_peak():
    """FP64 DMMA / cuBLAS DGEMM peak on this box (scripts/fp64_peak.cu); falls back to the committed measurement."""
    exe = os.path.join(ROOT, "scripts", "bin", "fp64_peak")
    try:
        out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
        d = json.loads(out.strip().splitlines()[-1])
        src = "measured now by scripts/fp64_peak.cu"
    except Exception:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
        src = "profiles/r01_fp64_peaks.json (earlier measurement on this pool)"
    dmma = max(v for k, v in d.items() if k.startswith("dmma_") and k.endswith(("w8", "w16")))
    return dmma, d.get("cublas_dgemm_8192_sustained_tflops"), src


def measured_traffic(families_per_launch):
    """DRAM bytes per launch of the pruning kernel from the committed ncu capture (profiles/prune_traffic.json),
    scaled per family; None when no capture is committed."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "prune_traffic.json")))
        return float(d["bytes_per_family"]) * families_per_launch, d["source"]
    except Exception:
        return None, None


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, torchrun notices), so
    file descriptor 1 is pointed at stderr for the whole run and the result line goes to a private copy of the original."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--families", type=int, default=1_000_000)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    from cafexp_b200 import engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if engine.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: cafexp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    F = args.families
    first, last = rank * F // world, (rank + 1) * F // world          # == sharded.shard_range(F, rank, world)
    t_gen = time.time()
    tree, counts, newick = synth.config5(F, N_LEAVES, SEED, LAMBDA, first=first, last=last)
    t_gen = time.time() - t_gen
    freq, rate, prior = gamma_parameters()
    lams = np.ascontiguousarray(rate[:, None] * np.array([[LAMBDA]]))
    flops_fc = algorithmic_flops_per_family_category(tree, MF, MRF)

    pinned = torch.from_numpy(counts).pin_memory()
    counts_pinned = pinned.numpy()
    eng = engine.Engine(tree, counts_pinned, MF, MRF, device=local)
    stream = torch.cuda.Stream(device=dev)          # the library, NCCL and the timing events all use this stream
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    result = torch.zeros(2, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    from cafexp_b200 import sharded
    job = sharded.ShardedLikelihood(sharded.engine_local_eval(eng), result)     # shard evaluation + one 2-double allreduce

    def step():
        job.enqueue(lams, prior, freq, engine.GAMMA_LINSUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream; L2 flushed in between
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    prune_ms = []
    build_ms = []
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        if i == args.steps - 1 or True:
            torch.cuda.synchronize()
            tm = eng.last_timings_ms()
            prune_ms.append(tm["prune"])
            build_ms.append(tm["matrix_build"])
    barrier()
    t1 = time.time()
    launches = eng.launches - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    prune_avg = torch.tensor([float(np.mean(prune_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(prune_avg, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    score = result.cpu().numpy()
    neg_lnl = float("inf") if score[1] > 0 else -float(score[0])

    # ---- end to end through the C ABI with host buffers (upload counts, evaluate, read everything back)
    e2e = None
    if not args.no_e2e:
        n_e2e = max(2, min(args.steps, 3))
        eng.set_families(counts_pinned)
        eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
        barrier()
        t_e = time.perf_counter()
        for _ in range(n_e2e):
            eng.set_families(counts_pinned)
            res = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM)
            if world > 1:
                part = torch.tensor([0.0 if res["n_failed"] else -res["score"], float(res["n_failed"])], dtype=torch.float64, device=dev)
                dist.all_reduce(part)
        barrier()
        dt = torch.tensor([time.perf_counter() - t_e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": F * n_e2e / float(dt.item()), "unit": "families/s",
               "h2d_bytes_per_step": int(counts_pinned.nbytes + lams.nbytes + prior.nbytes + freq.nbytes),
               "d2h_bytes_per_step": int(len(counts) * 8 * (1 + K) + 16), "steps": n_e2e,
               "note": "per rank: cafe_b200_set_families from pinned host memory + cafe_b200_eval returning score, per-family lnL and category likelihoods"}

    if rank == 0:
        peak_dmma, peak_cublas, peak_src = measure_fp64_peak()
        fam_local = last - first
        prune_s = float(prune_avg.item()) / 1e3
        achieved = flops_fc * K * fam_local / prune_s / 1e12 if prune_s > 0 else 0.0
        traffic, traffic_src = measured_traffic(fam_local)
        line = {
            "metric": "family-likelihood evals/sec", "value": F * args.steps / (total_ms / 1e3), "unit": "families/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "family_category_evals_per_s": F * K * args.steps / (total_ms / 1e3),
            "neg_lnl": neg_lnl, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": {"bound": "tensor", "kernel": "cafe::prune_kernel<5> (FP64 DMMA pruning)", "achieved": achieved, "peak": peak_dmma,
                         "unit": "TFLOP/s", "frac": achieved / peak_dmma if peak_dmma else None, "traffic": traffic, "traffic_unit": "bytes/launch",
                         "traffic_source": traffic_src, "algorithmic_bytes": (N_LEAVES * 4 + 8 * (K + 1)) * fam_local,
                         "peak_source": f"FP64 mma.sync peak, {peak_src}; cuBLAS DGEMM 8192^3 sustained = {peak_cublas} TFLOP/s; MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_family_category": flops_fc, "families_per_launch": fam_local, "categories": K,
                         "kernel_ms": prune_s * 1e3, "matrix_build_ms": float(np.mean(build_ms)),
                         "kernel_share_of_step": prune_s * 1e3 / (total_ms / args.steps)},
            "data_generation_s": t_gen,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_rate(tree, newick, counts, n_full=F)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
