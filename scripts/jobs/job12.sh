mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/t12_bench_8gpu.json 2> gpurun_out/t12_bench_8gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/t12_bench_8gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/t12_bench_8gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','neg_lnl','gpu_launches','parity_vs_1gpu')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['matrix_build_ms'])
print('recon', d['reconstruct'])
for k,v in d['fit'].items():
    if isinstance(v, dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('seconds','evaluations','first_evaluation_seconds','device_seconds','steady_state_ms_per_evaluation','host_overhead_us_per_evaluation','process_seconds','error','devices','values')})
PY
