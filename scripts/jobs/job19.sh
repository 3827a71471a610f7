mkdir -p gpurun_out
timeout 200 python scripts/geom_sweep.py --families 262144 --geoms "default 3,2,2" 2>&1 | cut -c1-120
(timeout 900 python -m pytest tests/test_config5_golden.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/t19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t19_pytest.log); tail -3 gpurun_out/t19_pytest.log
