mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t17_pytest.log); tail -4 gpurun_out/t17_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/t17_bench_1M.json 2> gpurun_out/t17_bench_1M.err; echo "bench rc=$?"; tail -2 gpurun_out/t17_bench_1M.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t17_bench_1M.json'))
print({k:d[k] for k in ('value','ms_per_step','neg_lnl','gpu_launches')}, d['parity_vs_1gpu']['ok'])
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['traffic'])
print('recon', d['reconstruct']['kernel_ms'], d['reconstruct']['families_per_s'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['slice_families_per_s'], d['cpu_baseline']['cores'])
for k,v in d['fit'].items():
    if isinstance(v, dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('seconds','evaluations','first_evaluation_seconds','steady_state_ms_per_evaluation','host_overhead_us_per_evaluation','error','speedup_vs_cpu','speedup_vs_recorded_cpu')})
PY
