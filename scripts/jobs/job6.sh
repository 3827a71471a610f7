mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t6_pytest.log); tail -15 gpurun_out/t6_pytest.log
timeout 600 python bench.py --families 262144 --steps 3 --warmup 3 > gpurun_out/t6_bench_262144.json 2> gpurun_out/t6_bench_262144.err; echo "bench rc=$?"; tail -5 gpurun_out/t6_bench_262144.err; cat gpurun_out/t6_bench_262144.json | cut -c1-3000
