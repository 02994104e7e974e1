import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'launches', d.get('gpu_launches'),
              'cpu', round(d['cpu_baseline']['value'], 1) if d.get('cpu_baseline') and d['cpu_baseline'].get('value') else None)
        r = d['roofline']
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k not in ('stage_ms_per_step', 'kernel', 'peak_source')})
        print({k: round(v, 3) for k, v in r['stage_ms_per_step'].items()}, d['clocks'])
    elif not l.startswith('[gpurun] sending'):
        print(l)
