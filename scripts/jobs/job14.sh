mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k "above_256 or big_tree" > gpurun_out/t14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t14_pytest.log); tail -30 gpurun_out/t14_pytest.log
