timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python tools/trace_step.py 32 > /tmp/tr.log 2>&1
grep -B1 -A 12 "^kernel  " /tmp/tr.log
grep -A 10 "^TD/dZ2: marks" /tmp/tr.log
for i in 1 2; do timeout 200 python bench.py --steps 200 --warmup 10 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']/1e6,1))"; done
