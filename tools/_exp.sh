timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_umma.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python tools/trace_step.py 48 2>&1 | tail -52
timeout 200 python bench.py --steps 200 --warmup 10 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('us/step', round(d['ms_per_step']*1e3,2))"
