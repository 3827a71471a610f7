mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-fit --no-reconstruct --no-cpu-baseline > gpurun_out/r02_bench_1M_plain.json 2>/dev/null; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_1M.csv python bench.py --steps 3 --warmup 3 --no-fit --no-reconstruct --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -c 1 -o gpurun_out/r02c_prune_1M -f python scripts/profile_run.py --families 1000000 --evals 1 > gpurun_out/r02c_ncu_prune.log 2>&1; echo "prune capture rc=$?"
