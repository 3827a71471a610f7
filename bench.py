#!/usr/bin/env python
"""Benchmark of the family-likelihood hot path on B200 (BASELINE.json metric: family-likelihood evals/sec; full ML fit wall time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--families F] [--impl reference]

Workload (config 5 of BASELINE.json): synthetic F = 1 000 000 gene families on a 100-taxon ultrametric tree,
max_family_size 150 (Nmax), max_root_family_size 125, gamma k = 4 (alpha 0.7), lambda 0.005, uniform root prior.
A "step" is ONE likelihood evaluation of all F families under one parameter point, i.e. what the Nelder-Mead
optimiser asks for per simplex vertex: transition matrices for every (branch, lambda, category) are rebuilt,
every family is pruned for every category, root prior / category weights applied, sum reduced (and summed
across ranks with one 2-double NCCL allreduce when N > 1).  Families are sharded across ranks (F/N each), so
total work is fixed: "scaling": "strong".

value      families/s with the count matrix already resident in HBM: K steps enqueued back to back, each bracketed by
           CUDA events on the launching stream (L2 flushed in between, outside the brackets), max over ranks
e2e        families/s through the C ABI with HOST buffers, one synchronous call sequence per step as the optimizer
           issues it: upload the count matrix from pinned host memory (cafe_b200_set_families_ex, one byte per count),
           evaluate (cafe_b200_eval), read back the score and all per-family log-likelihoods and category likelihoods
roofline   FP64 tensor (DMMA) roofline of the pruning kernel: algorithmic FLOPs (internal edges only, leaf edges
           count 0) / its CUDA-event duration inside the timed steps, against the FP64 DMMA
           peak measured on this box by scripts/fp64_peak.cu (builder-measured: MEASURED_PEAKS.json has no FP64 figure)
fit        wall time of the reference's own optimizer::optimize (seed 10) over the CUDA drop-in (integration/
           cuda_models.cpp, all N devices through cafe_b200_create_multi): mammal lambda+alpha k=4, and a config-5 slice
reconstruct  Pupko reconstruction throughput on a bounded sample of each rank's shard
cpu_baseline  the reference's own CPU path (oracle/_ref/ref_harness = unmodified reference objects; falls back to
           the C port oracle/liboracle.so) timed on this host on a bounded slice (>= 10 000 families) of the workload
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
CPU_SLICE = 10240            # families of the CPU-reference slice (>= 10 000)
FIT_SLICE = 65536            # families of the config-5 fit
RECON_SAMPLE = 131072        # families reconstructed per rank (bounded: the states array is 1.6 KB per family)
SCORE_FILE = os.path.join(ROOT, "profiles", "r02_score_1M_1gpu.json")


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


def pupko_pairs_per_family_category(tree, mf):
    """(multiply, compare) pairs of the max-product recursion: (mf+1)^2 per internal non-root edge."""
    internal_children = sum(1 for v in range(tree.n_nodes - 1) if tree.leaf_col[v] < 0)
    return internal_children * (mf + 1) * (mf + 1)


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


def write_family_table(path, tree, counts):
    """CAFE tab format, vectorised (hostio.write_gene_families is a per-row Python loop)."""
    ids = np.arange(len(counts))
    body = np.column_stack([ids, counts]).astype(np.int64)
    with open(path, "w") as fh:
        fh.write("Desc\tFamily ID\t" + "\t".join(tree.leaf_names) + "\n")
        np.savetxt(fh, body, fmt="(null)\t%d" + "\t%d" * counts.shape[1])


# ------------------------------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref = the unmodified reference; the one place besides tests that executes oracle/)
# ------------------------------------------------------------------------------------------------------------------

class CpuReference:
    """families/s of the reference CPU implementation on a bounded slice of the config-5 workload."""

    def __init__(self, tree, newick, counts, threads=None):
        from oracle import binding as orc
        self.orc = orc
        self.tree, self.newick, self.counts = tree, newick, counts
        self.cores = threads or os.cpu_count() or 1
        self.freq, self.rate, self.prior = gamma_parameters()
        self.kind = "reference" if orc.have_ref() else "port"

    def run(self, n):
        """(seconds, score) of ONE infer_family_likelihoods over the first n families, all host threads."""
        if self.kind == "reference":
            try:
                with tempfile.TemporaryDirectory() as tmp:
                    tpath = os.path.join(tmp, "tree.txt")
                    fpath = os.path.join(tmp, "fam.txt")
                    open(tpath, "w").write(self.newick + "\n")
                    write_family_table(fpath, self.tree, self.counts[:n])
                    r = self.orc.run_ref("eval", threads=self.cores, tree=tpath, fam=fpath, filter=0, k=K, alpha=ALPHA, maxfam=MF, maxroot=MRF,
                                         reps=1, **{"lambda": LAMBDA})
                    return r["seconds_best"], float(r["score"])
            except Exception:                      # harness missing libs etc.: fall back to the port
                self.kind = "port"
        os.environ["OMP_NUM_THREADS"] = str(self.cores)
        t0 = time.perf_counter()
        res = self.orc.infer(self.tree, self.counts[:n], self.rate[:, None] * np.array([[LAMBDA]]), self.freq, self.prior, MF, MRF, self.orc.GAMMA_LINSUM)
        return time.perf_counter() - t0, res["score"]

    def measure(self, n_full, n_slice=CPU_SLICE, timed=1, warm=0, budget_s=150.0):
        """The reference rebuilds all k x edges transition matrices inside every evaluation (src/gamma_core.cpp:196-197):
        time(F) = a (matrices) + b * F (pruning).  A 32-family run gives a; `timed` runs of the n_slice-family slice give
        b; the value reported is the throughput that gives for the FULL workload, F / (a + b F) — a linear extrapolation
        in F (families are independent, the gamma path has no de-duplication), flagged as such.  At most budget_s seconds
        of slice runs: with many requested steps the later ones repeat the mean of those that were timed (said in `sample`)."""
        n_small = min(len(self.counts), 32)
        n_slice = min(len(self.counts), n_slice)
        t_small, _ = self.run(n_small)
        for _ in range(warm):
            self.run(n_small)
        values, t_big, score = [], None, None
        a = b = 0.0
        t_start = time.perf_counter()
        for i in range(max(1, timed)):
            if i > 0 and time.perf_counter() - t_start + (t_big or 0.0) > budget_s:
                break
            t_big, score = self.run(n_slice)
            b = max((t_big - t_small) / max(n_slice - n_small, 1), 1e-9)
            a = max(t_small - b * n_small, 0.0)
            values.append(n_full / (a + b * n_full))
        return {"value": float(np.mean(values)), "unit": "families/s", "cores": self.cores, "kind": self.kind, "extrapolated": True,
                "sample": f"infer_family_likelihoods timed on the first {n_small} and {n_slice} config-5 families (k={K}): {t_small:.2f} s and {t_big:.2f} s "
                          f"=> {a:.2f} s per evaluation for the {K}x198 transition matrices + {b * 1e3:.3f} ms per family; value = {n_full} / (a + b*{n_full}), "
                          f"i.e. linear extrapolation to the full workload (families are independent); measured on the slice alone: {n_slice / t_big:.1f} families/s; "
                          f"{len(values)} timed slice run(s) for {max(1, timed)} requested step(s)",
                "slice_families": n_slice, "slice_seconds": t_big, "slice_families_per_s": n_slice / t_big, "score_of_slice": score,
                "fixed_s": a, "per_family_s": b, "values": values}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cafexp_b200 import synth
    tree, counts, newick = synth.config5(args.families, N_LEAVES, SEED, LAMBDA, first=0, last=min(args.families, CPU_SLICE))
    ref = CpuReference(tree, newick, counts)
    base = ref.measure(args.families, timed=args.steps, warm=args.warmup)
    value = base["value"]
    line = {"impl": "reference", "metric": "family-likelihood evals/sec", "value": value, "unit": "families/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.families / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated",
                                                                                         "slice_families", "slice_families_per_s")},
            "e2e": {"value": value, "unit": "families/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, world):
    return {"workload": f"BASELINE.json configs[4]: synthetic {args.families} families x {N_LEAVES}-taxon ultrametric tree, Nmax={MF}, "
                        f"max_root_family_size={MRF}, gamma k={K} alpha={ALPHA}, lambda={LAMBDA}, uniform root prior",
            "families": args.families, "taxa": N_LEAVES, "matrix_size": MF + 1, "gamma_categories": K,
            "parallelism": (f"families sharded over {world} GPU(s); per evaluation: transition matrices built once (1/{world} per rank) + NCCL all-gather over "
                            f"NVLink, then one 2-double NCCL allreduce") if world > 1 else "1 GPU",
            "l2": "L2 flushed (256 MiB write) between timed steps", "seed": SEED}


def measure_fp64_peak():
    """FP64 DMMA / cuBLAS DGEMM peak on this box (scripts/fp64_peak.cu); falls back to the committed measurement."""
    exe = os.path.join(ROOT, "scripts", "bin", "fp64_peak")
    try:
        out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
        d = json.loads(out.strip().splitlines()[-1])
        src = "measured now by scripts/fp64_peak.cu (builder-measured; the driver's MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
        src = "profiles/r01_fp64_peaks.json (earlier builder measurement on this pool)"
    dmma = max(v for k, v in d.items() if k.startswith("dmma_") and k.endswith(("w8", "w16")))
    return dmma, d.get("cublas_dgemm_8192_sustained_tflops"), src


def measured_traffic(families_per_launch):
    """DRAM bytes per launch of the pruning kernel from the committed ncu capture (profiles/prune_traffic.json):
    a fixed part (the matrices, read once) plus a per-family part; None when no capture is committed."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "prune_traffic.json")))
        if "fixed_bytes" in d:
            return float(d["fixed_bytes"]) + float(d["bytes_per_family"]) * families_per_launch, d["source"]
        return float(d["bytes_per_family"]) * families_per_launch, d["source"]
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------------------------
# fit wall time: the reference's optimizer over the CUDA drop-in (and over its own CPU models beside it)
# ------------------------------------------------------------------------------------------------------------------

def harness_fit(cuda, devices, timeout, **kw):
    """One `ref_harness[_cuda] fit` process; returns its JSON plus the process wall time, or {"error": ...}."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness_cuda" if cuda else "ref_harness")
    if not os.path.exists(exe):
        return {"error": f"{os.path.relpath(exe, ROOT)} not built"}
    argv = [exe, "fit", "--seed", "10"]
    if cuda:
        argv += ["--cuda", "1"]
    for key, val in kw.items():
        if val is None:
            continue
        argv += ["--" + key, repr(val) if isinstance(val, float) else str(val)]
    env = dict(os.environ)
    # The process sees only the devices it uses: cuInit initialises every VISIBLE device (about 0.7 s each on an 8-GPU B200
    # node — 5.6 s before the first kernel of a 1-device fit otherwise).  Ordinals inside the process are 0 .. n-1.
    visible = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    env["CUDA_VISIBLE_DEVICES"] = ",".join(visible[d] if d < len(visible) else str(d) for d in devices)
    env["CAFE_B200_DEVICES"] = ",".join(str(i) for i in range(len(devices)))
    env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)        # torchrun pins its ranks to one OpenMP thread; this process has the host to itself
    env.pop("CAFE_B200_GEOM", None)
    t0 = time.perf_counter()
    try:
        res = subprocess.run(argv, capture_output=True, text=True, env=env, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout} s"}
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        return {"error": res.stderr.strip()[-300:]}
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{\"")][-1])
    out["process_seconds"] = wall
    return out


def summarize_fit(r):
    if "error" in r:
        return r
    keep = {k: r[k] for k in ("values", "score", "iterations", "evaluations", "seconds", "seconds_in_score", "first_evaluation_seconds",
                              "process_seconds", "devices", "device_evaluations", "device_seconds", "bind_seconds", "staging_seconds",
                              "enqueue_seconds", "wait_seconds", "n_families", "threads") if k in r}
    if "device_seconds" in r and r.get("device_evaluations", 0) > 1 and r.get("evaluations", 0) > 1:
        # the first evaluation carries the one-time set-up (CUDA context, flattening + de-duplicating the families, upload);
        # evaluations with invalid parameters never reach the device, so device time is averaged over those that did
        n, nd = r["evaluations"], r["device_evaluations"]
        keep["device_ms_per_evaluation"] = 1e3 * r["device_seconds"] / nd
        steady = (r["seconds_in_score"] - r["first_evaluation_seconds"]) / (n - 1)
        device_share = r["device_seconds"] * (nd - 1) / nd / (n - 1)
        keep["steady_state_ms_per_evaluation"] = 1e3 * steady
        keep["host_overhead_us_per_evaluation"] = 1e6 * (steady - device_share)
        keep["overhead_note"] = ("per evaluation after the first, above the CUDA-event device time: the reference's scorer and lambda/prior bookkeeping, "
                                 "the drop-in's tables, the library's staging + launches, one synchronisation")
    return keep


def fit_benchmarks(world, tree5, newick5, counts5_slice, with_cpu=True):
    """BASELINE.json metric, second half.  (1) configs[1]: the mammal set, gamma k = 4, lambda and alpha fitted jointly,
    randomizer_engine.seed(10), through the reference's optimizer::optimize over the CUDA models on `world` devices; the
    reference's CPU time for the same fit is the recorded 757 s (tests/golden/fits.json, 8 threads; 16 minutes is beyond a
    benchmark run) and the single-lambda fit is timed live on both.  (2) the same joint fit on a config-5 slice."""
    from cafexp_b200 import hostio
    devices = list(range(world))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fits.json")))
    inp = json.load(open(os.path.join(ROOT, "tests", "golden", "mammal_inputs.json")))
    out = {"optimizer": "the reference's own optimizer::optimize (Nelder-Mead, src/optimizer.cpp:539), seed 10, driving integration/cuda_models.cpp",
           "devices": world}
    with tempfile.TemporaryDirectory() as tmp:
        tpath, fpath = os.path.join(tmp, "mammal_tree.txt"), os.path.join(tmp, "mammal_fam.txt")
        open(tpath, "w").write(inp["tree"] + "\n")
        flat = hostio.flatten_tree(hostio.parse_newick(inp["tree"]))
        z = np.load(os.path.join(ROOT, "tests", "golden", "mammal_counts.npz"))
        hostio.write_gene_families(fpath, flat, [str(i) for i in z["ids"]] if "ids" in z.files else [str(i) for i in range(len(z["counts"]))], z["counts"])
        harness_fit(True, devices, 120, tree=tpath, fam=fpath)                        # warm the process image / CUDA driver once
        g = summarize_fit(harness_fit(True, devices, 300, tree=tpath, fam=fpath, k=4))
        g["workload"] = "BASELINE.json configs[1]: mammal_gene_families x mammals_tree (10 956 families after the root filter, N=141), gamma k=4, lambda and alpha fitted"
        rec = gold["gamma4_lambda_alpha_seed10"]
        g["cpu_reference"] = {"seconds": rec["seconds"], "evaluations": rec["evaluations"], "threads": rec["threads"], "values": rec["values"],
                              "source": "tests/golden/fits.json: the unmodified reference in the build container (recorded; 12.6 minutes)"}
        if "seconds" in g:
            g["speedup_vs_recorded_cpu"] = rec["seconds"] / g["seconds"]
            g["same_evaluations_as_reference"] = g.get("evaluations") == rec["evaluations"]
        out["mammal_gamma4"] = g
        b = summarize_fit(harness_fit(True, devices, 300, tree=tpath, fam=fpath))
        b["workload"] = "BASELINE.json configs[0]: same data, single lambda"
        if with_cpu:
            c = harness_fit(False, devices, 600, tree=tpath, fam=fpath)
            b["cpu_reference"] = summarize_fit(c) if "error" in c else {"seconds": c["seconds"], "evaluations": c["evaluations"], "threads": c["threads"],
                                                                       "values": c["values"], "source": "oracle/_ref/ref_harness on this host, live"}
            if "seconds" in b and "seconds" in b["cpu_reference"]:
                b["speedup_vs_cpu"] = b["cpu_reference"]["seconds"] / b["seconds"]
        out["mammal_single_lambda"] = b
        # config-5 slice: the families shard over the devices inside ONE process (cafe_b200_create_multi)
        t5, f5 = os.path.join(tmp, "tree5.txt"), os.path.join(tmp, "fam5.txt")
        open(t5, "w").write(newick5 + "\n")
        write_family_table(f5, tree5, counts5_slice)
        s = summarize_fit(harness_fit(True, devices, 600, tree=t5, fam=f5, k=K, filter=0, maxfam=MF, maxroot=MRF))
        s["workload"] = f"config-5 slice: first {len(counts5_slice)} synthetic families x {N_LEAVES} taxa, N={MF + 1}, gamma k={K}, lambda and alpha fitted"
        out["config5_slice_gamma4"] = s
    return out


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
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-reconstruct", action="store_true")
    ap.add_argument("--replicated-build", action="store_true", help="N > 1: every rank builds all transition matrices itself (round-1 behaviour)")
    ap.add_argument("--record-score", action="store_true", help="write the 1-GPU score of this workload to profiles/ (the N > 1 parity reference)")
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
        host_group = dist.new_group(backend="gloo")       # host-side waits (ranks idle on the CPU, not in an NCCL kernel)
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
    fam_local = last - first

    # counts travel and live on the device as one byte each (max_family_size <= 255)
    pinned = torch.from_numpy(counts.astype(np.uint8)).pin_memory()
    counts_pinned = pinned.numpy()
    eng = engine.Engine(tree, counts_pinned, MF, MRF, device=local)
    stream = torch.cuda.Stream(device=dev)          # the library, NCCL and the timing events all use this stream
    torch.cuda.set_stream(stream)
    result = torch.zeros(2, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    from cafexp_b200 import sharded
    job = sharded.ShardedLikelihood(sharded.engine_local_eval(eng), result)     # shard evaluation + one 2-double allreduce
    if world > 1 and not args.replicated_build:
        # every transition matrix is built once, by one rank, and all-gathered over NVLink (instead of N redundant builds)
        eng.set_build_partition(rank, world, sharded.nccl_matrix_gather())

    def step():
        job.enqueue(lams, prior, freq, engine.GAMMA_LINSUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    # ---- timed region: K steps enqueued back to back, each bracketed by CUDA events on the launching stream; the L2
    #      flush runs between the brackets; no host synchronisation inside
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
    barrier()
    t1 = time.time()
    launches = eng.launches - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    hist = eng.timing_history_ms(min(args.steps, 64))                      # per-step kernel durations, newest first
    prune_ms, build_ms = hist[:, 1], hist[:, 0]
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    prune_avg = torch.tensor([float(np.mean(prune_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(prune_avg, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    score = result.cpu().numpy()
    neg_lnl = float("inf") if score[1] > 0 else -float(score[0])

    # ---- parity of the sharded score with the 1-GPU score of the same workload (recorded by --record-score at N = 1)
    parity = None
    if rank == 0:
        if world == 1 and args.record_score:
            json.dump({"families": F, "seed": SEED, "neg_lnl": neg_lnl, "n_gpus": 1, "describe": eng.describe()}, open(SCORE_FILE, "w"))
        try:
            rec = json.load(open(SCORE_FILE))
            if rec["families"] == F and rec["seed"] == SEED:
                rel = abs(neg_lnl - rec["neg_lnl"]) / abs(rec["neg_lnl"])
                parity = {"expected": rec["neg_lnl"], "got": neg_lnl, "rel_err": rel, "tolerance": 1e-12, "ok": bool(rel <= 1e-12),
                          "source": os.path.relpath(SCORE_FILE, ROOT)}
        except Exception:
            parity = None

    # ---- end to end through the C ABI with host buffers: upload counts, evaluate, read everything back, every step
    e2e = None
    if not args.no_e2e:
        n_e2e = max(3, args.steps)
        out_family = torch.empty(max(fam_local, 1), dtype=torch.float64).pin_memory().numpy()
        out_cat = torch.empty((max(fam_local, 1), K), dtype=torch.float64).pin_memory().numpy()
        eng.set_families(counts_pinned)
        eng.infer(lams, prior, freq, engine.GAMMA_LINSUM, out_family=out_family, out_cat=out_cat)
        barrier()
        t_e = time.perf_counter()
        for _ in range(n_e2e):
            eng.set_families(counts_pinned)
            res = eng.infer(lams, prior, freq, engine.GAMMA_LINSUM, out_family=out_family, out_cat=out_cat)
            if world > 1:
                part = torch.tensor([0.0 if res["n_failed"] else -res["score"], float(res["n_failed"])], dtype=torch.float64, device=dev)
                dist.all_reduce(part)
        barrier()
        dt = torch.tensor([time.perf_counter() - t_e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": F * n_e2e / float(dt.item()), "unit": "families/s",
               "h2d_bytes_per_step": int(counts_pinned.nbytes + lams.nbytes + prior.nbytes + freq.nbytes),
               "d2h_bytes_per_step": int(fam_local * 8 * (1 + K) + 16), "steps": n_e2e,
               "note": "per rank and step: cafe_b200_set_families_ex (uint8 counts from pinned host memory) + cafe_b200_eval returning the score, "
                       "per-family lnL and category likelihoods into pinned host memory; synchronous, as the optimizer calls it"}

    # ---- Pupko reconstruction throughput on a bounded sample of this rank's shard
    recon = None
    if not args.no_reconstruct:
        from cafexp_b200 import params
        n_rec = min(fam_local, RECON_SAMPLE)
        prior_sz = params.prior_uniform(MRF, None, min(MF, MRF) + 1)
        with engine.Engine(tree, counts_pinned[:n_rec], MF, MRF, device=local) as reng:
            reng.reconstruct(lams, prior_sz)
            barrier()
            reng.reconstruct(lams, prior_sz)
            rec_ms = torch.tensor([reng.last_timings_ms()["reconstruct"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(rec_ms, op=dist.ReduceOp.MAX)
        rec_s = float(rec_ms.item()) / 1e3
        pairs = pupko_pairs_per_family_category(tree, MF)
        recon = {"families_per_rank": n_rec, "categories": K, "kernel_ms": rec_s * 1e3, "families_per_s": n_rec * world / rec_s,
                 "family_categories_per_s": n_rec * world * K / rec_s, "multiply_compare_pairs_per_s": n_rec * world * K * pairs / rec_s,
                 "unit": "families/s (whole job, kernel time; each family = k max-product traversals + traceback of 99 internal nodes)"}

    eng.close()
    del flush
    torch.cuda.empty_cache()

    # ---- fit wall time through the reference's optimizer (rank 0 drives all N devices from one process)
    fit = None
    if not args.no_fit:
        barrier()
        if world > 1:
            dist.barrier(group=host_group)
        if rank == 0:
            try:
                tree5, counts5, newick5 = synth.config5(F, N_LEAVES, SEED, LAMBDA, first=0, last=min(F, FIT_SLICE))
                fit = fit_benchmarks(world, tree5, newick5, counts5, with_cpu=(world == 1 and not args.no_cpu_baseline))
            except Exception as e:      # noqa: BLE001 — the fit is an extra object; the headline line must still print
                fit = {"error": repr(e)}
        if world > 1:
            dist.barrier(group=host_group)

    if rank == 0:
        peak_dmma, peak_cublas, peak_src = measure_fp64_peak()
        prune_s = float(prune_avg.item()) / 1e3
        achieved = flops_fc * K * fam_local / prune_s / 1e12 if prune_s > 0 else 0.0
        traffic, traffic_src = measured_traffic(fam_local)
        line = {
            "metric": "family-likelihood evals/sec", "value": F * args.steps / (total_ms / 1e3), "unit": "families/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "family_category_evals_per_s": F * K * args.steps / (total_ms / 1e3),
            "neg_lnl": neg_lnl, "parity_vs_1gpu": parity, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": {"bound": "tensor", "kernel": "cafe::prune_kernel<5,4,3,...> (FP64 DMMA pruning, three consumer groups, partial products parked in tensor memory)",
                         "achieved": achieved, "peak": peak_dmma,
                         "unit": "TFLOP/s", "frac": achieved / peak_dmma if peak_dmma else None, "traffic": traffic, "traffic_unit": "bytes/launch",
                         "traffic_source": traffic_src, "algorithmic_bytes": (N_LEAVES * 1 + 8 * (K + 1)) * fam_local,
                         "algorithmic_bytes_note": "per family: 100 one-byte leaf counts in, k category likelihoods + lnL out",
                         "peak_source": f"FP64 mma.sync (DMMA) peak, {peak_src}; cuBLAS DGEMM 8192^3 sustained = {peak_cublas} TFLOP/s",
                         "flops_per_family_category": flops_fc, "families_per_launch": fam_local, "categories": K,
                         "kernel_ms": prune_s * 1e3, "matrix_build_ms": float(np.mean(build_ms)),
                         "kernel_share_of_step": prune_s * 1e3 / (total_ms / args.steps)},
            "reconstruct": recon, "fit": fit, "data_generation_s": t_gen,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = CpuReference(tree, newick, counts).measure(F)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "slice_families", "slice_families_per_s")}
        emit(line)
        if parity is not None and not parity["ok"]:
            sys.stderr.write(f"bench.py: the score at {world} GPU(s) differs from the recorded 1-GPU score: {parity}\n")
            if world > 1:
                dist.destroy_process_group()
            sys.exit(3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
