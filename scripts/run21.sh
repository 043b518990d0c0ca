mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tensorcore.py -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/s21_tc.log; tail -45 gpurun_out/s21_tc.log
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s21_c3.json 2> gpurun_out/s21_c3.err; python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s21_c3.json').read().strip().splitlines()[-1])
    print('c3', l['ms_per_step'], l['value'], l['phase_breakdown_ms'], l['roofline']['avg_launch_ms'], l['train_info_last'])
except Exception as e:
    print('c3 failed', e); print(open('gpurun_out/s21_c3.err').read()[-1500:])
PY
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s21_c4.json 2> gpurun_out/s21_c4.err; python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s21_c4.json').read().strip().splitlines()[-1])
    print('c4', l['ms_per_step'], l['value'], l['phase_breakdown_ms'], l['roofline']['avg_launch_ms'], l['train_info_last'])
except Exception as e:
    print('c4 failed', e); print(open('gpurun_out/s21_c4.err').read()[-1500:])
PY
