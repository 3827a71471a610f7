mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_pvalues.py tests/test_viterbi.py tests/test_config5_golden.py tests/test_sharded_nccl.py -m gpu -x -q > gpurun_out/t20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t20_pytest.log); tail -4 gpurun_out/t20_pytest.log
(timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "two_devices or lazy" > gpurun_out/t20_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t20_pytest2.log); tail -3 gpurun_out/t20_pytest2.log
