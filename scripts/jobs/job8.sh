mkdir -p gpurun_out
timeout 900 python bench.py --record-score > gpurun_out/t8_bench_1M.json 2> gpurun_out/t8_bench_1M.err; echo "bench rc=$?"; tail -3 gpurun_out/t8_bench_1M.err
cp profiles/r02_score_1M_1gpu.json gpurun_out/ 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/t8_bench_1M.json'))
print({k:d[k] for k in ('value','ms_per_step','neg_lnl','gpu_launches')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['matrix_build_ms'])
print('recon', d['reconstruct'])
print('cpu', d.get('cpu_baseline'))
for k,v in d['fit'].items():
    if isinstance(v, dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('seconds','evaluations','first_evaluation_seconds','device_seconds','host_overhead_us_per_evaluation','process_seconds','error')})
PY
