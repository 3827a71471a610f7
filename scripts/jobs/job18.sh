mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_1M_8gpu.json 2> gpurun_out/r02_bench_1M_8gpu.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r02_bench_1M_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 --no-fit > gpurun_out/r02_bench_1M_4gpu.json 2> gpurun_out/r02_bench_1M_4gpu.err; echo "bench4 rc=$?"
python scripts/fit_breakdown.py --quick --devices "0;0,1;0,1,2,3;0,1,2,3,4,5,6,7" > gpurun_out/r02_fit_walltime_by_devices.jsonl 2> gpurun_out/r02_fit.err; echo "fit rc=$?"
python - <<'PY'
import json
for n in (8,4):
    d=json.load(open(f'gpurun_out/r02_bench_1M_{n}gpu.json'))
    print(n, {k:d[k] for k in ('value','ms_per_step','neg_lnl')}, d['parity_vs_1gpu']['ok'], 'e2e', d['e2e']['value'], 'kernel', d['roofline']['kernel_ms'], 'build', d['roofline']['matrix_build_ms'])
for l in open('gpurun_out/r02_fit_walltime_by_devices.jsonl'):
    r=json.loads(l); print(r['fit'], r['n_devices'], {k:r.get(k) for k in ('seconds','first_evaluation_seconds','evaluations','iterations','device_seconds','values')})
PY
