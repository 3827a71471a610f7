mkdir -p gpurun_out
timeout 400 python scripts/geom_sweep.py --families 262144 --geoms "default 2,2,2 3,2,1 3,1,2 2,4,2 2,2,1 2,2,2,4 3,2,2,3" > gpurun_out/t3_sweep.log 2>&1
cat gpurun_out/t3_sweep.log
for g in 3,2,2 2,2,2; do
  CAFE_B200_GEOM=$g timeout 300 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -c 1 -o gpurun_out/t3_prune_${g//,/_} -f python scripts/profile_run.py --families 65536 --evals 1 > gpurun_out/t3_ncu_${g//,/_}.log 2>&1
  tail -2 gpurun_out/t3_ncu_${g//,/_}.log
done
ls -la gpurun_out/*.ncu-rep
