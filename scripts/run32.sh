mkdir -p gpurun_out
for c in c3 c4; do
timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown > gpurun_out/s32_$c.json 2> gpurun_out/s32_$c.err; python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/s32_{c}.json').read().strip().splitlines()[-1])
    print(c, l['ms_per_step'], l['roofline']['seq_step_cycles'])
except Exception as e:
    print(c, 'failed', e); print(open(f'gpurun_out/s32_{c}.err').read()[-1500:])
PY
done
