# closing session: whole GPU suite, smoke, final launch lists of c3 / c4, default bench (c2 + companions)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/s36_alltests.log; tail -3 gpurun_out/s36_alltests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/s36_launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s36_ncu_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s36_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s36_ncu_c3.log 2>&1
for c in c3 c4; do
timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extras > gpurun_out/s36_$c.json 2> gpurun_out/s36_$c.err; python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/s36_{c}.json').read().strip().splitlines()[-1])
    pb = l['phase_breakdown_ms']; r = l['roofline']
    print(c, l['ms_per_step'], l['value'], l['e2e']['value'], {k: round(v, 3) for k, v in pb.items() if k.endswith('_ms')}, r['kernel'][:24], round(r['frac'], 3), (l.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print(c, 'failed', e); print(open(f'gpurun_out/s36_{c}.err').read()[-1500:])
PY
done
timeout 1200 python bench.py > gpurun_out/s36_default.json 2> gpurun_out/s36_default.err; python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s36_default.json').read().strip().splitlines()[-1])
    print('default', l['ms_per_step'], l['value'], l['e2e']['value'], (l.get('value_fp32') or {}).get('value'), l['roofline']['frac'], l['clocks'], (l.get('cpu_baseline') or {}).get('value'))
    for k, o in (l.get('other_configs') or {}).items(): print('   other', k, o.get('ms_per_step'), o.get('value'), (o.get('roofline') or {}).get('frac'), o.get('error'))
except Exception as e:
    print('default failed', e); print(open('gpurun_out/s36_default.err').read()[-2500:])
PY
