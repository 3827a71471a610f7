mkdir -p gpurun_out
python scripts/fit_breakdown.py --quick --devices "0;0,1;0,1,2,3;0,1,2,3,4,5,6,7" > gpurun_out/r02_fit_walltime_by_devices.jsonl 2> gpurun_out/r02_fit.err; echo "fit rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_fit_walltime_by_devices.jsonl'):
    r=json.loads(l); print(r['fit'], r['n_devices'], {k:r.get(k) for k in ('seconds','first_evaluation_seconds','evaluations','iterations','device_seconds','process_seconds')})
PY
