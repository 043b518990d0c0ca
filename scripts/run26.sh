mkdir -p gpurun_out
timeout 600 python bench.py --config c4 --steps 4 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s26_c4.json 2> gpurun_out/s26_c4.err; python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s26_c4.json').read().strip().splitlines()[-1])
    r = l['roofline']
    print('c4', l['ms_per_step'], {k: r.get(k) for k in ('bound','achieved','peak','frac','kernel','avg_launch_ms','traffic','algorithmic_mbyte_per_launch','kernel_share_of_step')})
    for f, t in r['pipeline_families'].items(): print('   ', f, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()})
    print('   tensor', r.get('tensor'))
except Exception as e:
    print('c4 failed', e); print(open('gpurun_out/s26_c4.err').read()[-2500:])
PY
( time timeout 1200 python bench.py > gpurun_out/s26_default.json 2> gpurun_out/s26_default.err ) 2>&1 | tail -3; python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s26_default.json').read().strip().splitlines()[-1])
    print('default', l['ms_per_step'], l['value'], l['e2e']['value'], l.get('value_fp32'), l['roofline']['frac'], l['clocks'])
    for k, o in (l.get('other_configs') or {}).items(): print('   other', k, o.get('ms_per_step'), o.get('value'), (o.get('roofline') or {}).get('kernel'), (o.get('roofline') or {}).get('frac'), o.get('error'))
except Exception as e:
    print('default failed', e); print(open('gpurun_out/s26_default.err').read()[-2500:])
PY
tail -3 gpurun_out/s26_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -c 600
