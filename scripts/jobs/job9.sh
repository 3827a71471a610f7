mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_config5_golden.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t9_pytest.log); tail -5 gpurun_out/t9_pytest.log
python scripts/profile_run.py --families 4096 --evals 1 --recon 131072 2>&1 | tail -2
python scripts/fit_breakdown.py --slice 0 2>&1 | tail -4 | cut -c1-900
