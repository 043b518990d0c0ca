mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/s25_alltests.log; tail -4 gpurun_out/s25_alltests.log
for c in c3 c4; do
timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extras > gpurun_out/s25_$c.json 2> gpurun_out/s25_$c.err; python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/s25_{c}.json').read().strip().splitlines()[-1])
    pb = l['phase_breakdown_ms']
    print(c, l['ms_per_step'], l['value'], l['e2e']['value'], {k: v for k, v in pb.items() if k.endswith('_ms')}, l['roofline']['avg_launch_ms'], l.get('cpu_baseline'))
except Exception as e:
    print(c, 'failed', e); print(open(f'gpurun_out/s25_{c}.err').read()[-1500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/s25_launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s25_ncu_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s25_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s25_ncu_c3.log 2>&1
