mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/s35_alltests.log; tail -12 gpurun_out/s35_alltests.log
for c in c3 c4; do
timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s35_$c.json 2> gpurun_out/s35_$c.err; python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/s35_{c}.json').read().strip().splitlines()[-1])
    pb = l['phase_breakdown_ms']
    print(c, l['ms_per_step'], l['value'], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in pb.items() if k.endswith('_ms')}, l['train_info_last'])
except Exception as e:
    print(c, 'failed', e); print(open(f'gpurun_out/s35_{c}.err').read()[-1500:])
PY
done
