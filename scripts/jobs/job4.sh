mkdir -p gpurun_out
timeout 400 python scripts/geom_sweep.py --families 262144 --geoms "default 3,2,2,3 3,2,1 3,2,1,3 3,1,2 3,1,2,6 2,2,2 2,2,2,4 2,4,2" > gpurun_out/t4_sweep.log 2>&1
cat gpurun_out/t4_sweep.log
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t4_pytest.log); tail -12 gpurun_out/t4_pytest.log
