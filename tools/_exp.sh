timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
timeout 200 python tools/trace_step.py 32 > /tmp/tr.log 2>&1
grep -B1 -A 12 "^kernel  " /tmp/tr.log
for i in 1 2; do timeout 200 python bench.py --steps 200 --warmup 10 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('us/step', round(d['ms_per_step']*1e3,2), 'fwd', round(d['roofline']['kernel_us'],2), 'frac', round(d['roofline']['frac'],3))"; done
GOLDQ_NO_GATHER_SIDE=1 timeout 200 python bench.py --steps 200 --warmup 10 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old gather: us/step', round(d['ms_per_step']*1e3,2))"
timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K=20: us/step', round(d['ms_per_step']*1e3,2))"
