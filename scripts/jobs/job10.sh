mkdir -p gpurun_out
python scripts/profile_run.py --families 4096 --evals 1 --recon 131072 2>&1 | tail -1
(timeout 900 python -m pytest tests/test_config5_golden.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t10_pytest.log); tail -3 gpurun_out/t10_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pupko_kernel -c 1 -o gpurun_out/t10_pupko -f python scripts/profile_run.py --families 4096 --evals 1 --recon 16384 > gpurun_out/t10_ncu.log 2>&1
tail -1 gpurun_out/t10_ncu.log
