mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sharded_nccl.py tests/test_config5_golden.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/t11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t11_pytest.log); tail -8 gpurun_out/t11_pytest.log
for mode in "" "--replicated-build"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-fit --no-reconstruct $mode > gpurun_out/t11_bench_2gpu$mode.json 2> gpurun_out/t11_bench_2gpu$mode.err; echo "bench rc=$?"; tail -2 gpurun_out/t11_bench_2gpu$mode.err
python - <<PY
import json
d=json.load(open('gpurun_out/t11_bench_2gpu$mode.json'))
print('$mode', {k:d[k] for k in ('value','ms_per_step','neg_lnl','gpu_launches','parity_vs_1gpu')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['matrix_build_ms'])
PY
done
