mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t13_pytest.log); tail -4 gpurun_out/t13_pytest.log
for e in "" "CAFE_B200_NO_TAIL_SPLIT=1"; do
env $e python bench.py --families 125000 --steps 5 --warmup 3 --no-fit --no-reconstruct --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$e 125000 families:', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['neg_lnl'])"
done
python bench.py --steps 5 --warmup 3 --no-fit --no-reconstruct --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('1M families:', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['neg_lnl'], d['parity_vs_1gpu']['ok'], d['e2e']['value'])"
