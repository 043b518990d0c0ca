mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -q -s -k "gru or naive_rnn" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/s23_gru.log; grep -n "tf32\|passed\|failed\|FAIL\|Error" gpurun_out/s23_gru.log | head -30
timeout 600 python scripts/diag_gru_tc.py c4_gru_smac > gpurun_out/s23_diag_c4.log 2>&1; grep -v "step " gpurun_out/s23_diag_c4.log | tail -60
for c in c3 c4; do
timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s23_$c.json 2> gpurun_out/s23_$c.err; python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/s23_{c}.json').read().strip().splitlines()[-1])
    pb = l['phase_breakdown_ms']
    print(c, l['ms_per_step'], l['value'], {k: v for k, v in pb.items() if k.endswith('_ms')}, l['roofline']['avg_launch_ms'], l['train_info_last'])
except Exception as e:
    print(c, 'failed', e); print(open(f'gpurun_out/s23_{c}.err').read()[-1500:])
PY
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
