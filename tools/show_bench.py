"""Print the headline fields of bench.py JSON lines (diagnostic helper)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    print(f, 'N=%d' % j['n_gpus'], 'value %.1fM' % (j['value'] / 1e6), 'us/step %.2f' % (j['ms_per_step'] * 1e3),
          'e2e %.1fM' % (j['e2e']['value'] / 1e6), 'launches', j['gpu_launches'], 'warmup', j['warmup'], 'clocks', j.get('clocks'))
    r = j['roofline']
    print('  roof:', r['kernel'], '%.2f us' % r['kernel_us'], 'frac %.3f' % r['frac'],
          ('step frac %.3f' % r['step']['frac_of_sustained_peak']) if 'step' in r else '')
    print('  kernels:', r['kernels_us'])
    for k, v in (j.get('configs') or {}).items():
        print('  ', k, '%.1fM' % (v['value'] / 1e6), '%.2f us/step' % v['us_per_step'], ('e2e %.1fM' % (v['e2e']['value'] / 1e6)) if 'e2e' in v else '',
              'floor', v.get('roofline', {}).get('launch_floor_us'))
    if j.get('cpu_baseline'):
        print('  cpu:', j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
