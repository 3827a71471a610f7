mkdir -p gpurun_out
timeout 400 python scripts/geom_sweep.py --families 262144 --geoms "default default:1 3,3,2 3,3,2,2 3,4,2 3,4,1 3,4,2:1 3,2,1:1" > gpurun_out/t5_sweep.log 2>&1
cat gpurun_out/t5_sweep.log | cut -c1-330
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -c 1 -o gpurun_out/t5_prune_default -f python scripts/profile_run.py --families 65536 --evals 1 > gpurun_out/t5_ncu.log 2>&1
tail -2 gpurun_out/t5_ncu.log
