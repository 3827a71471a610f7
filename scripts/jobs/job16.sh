mkdir -p gpurun_out
for w in 8 16; do CAFE_B200_PUPKO_WARPS=$w python scripts/profile_run.py --families 4096 --evals 1 --recon 131072 2>&1 | tail -1; done
(CAFE_B200_PUPKO_WARPS=16 timeout 900 python -m pytest tests/test_config5_golden.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/t16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t16_pytest.log); tail -5 gpurun_out/t16_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -c 1 -o gpurun_out/r02b_prune_1M -f python scripts/profile_run.py --families 1000000 --evals 1 > gpurun_out/r02b_ncu_prune.log 2>&1; echo "prune capture rc=$?"
