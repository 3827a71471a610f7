#!/bin/bash
# usage: scripts/sweep.sh VAR "v1 v2 ..." [families]  — bench the pruning kernel under an env-var sweep
var=$1; vals=$2; fam=${3:-262144}
for v in $vals; do
  env $var=$v python bench.py --families $fam --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/sweep.err | python -c "
import sys,json
s=sys.stdin.read()
try:
    d=json.loads(s); print('$var', '$v', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'nlnl', d['neg_lnl'])
except Exception as e:
    print('$var', '$v', 'FAILED', s[:200])
"
done
