mkdir -p gpurun_out
MAPPO_B200_INLINE_PACK=1 timeout 900 python -m pytest tests/test_gpu_tensorcore.py -q -k "c1_mlp or c2_mlp or several_tiles" 2>&1 | tail -4
for v in 0 1 0 1; do
MAPPO_B200_INLINE_PACK=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --cpu-iters 0 > gpurun_out/s28_c2_$v.json 2> gpurun_out/s28_c2_$v.err; python - $v <<'PY'
import json, sys
v = sys.argv[1]
l = json.loads(open(f'gpurun_out/s28_c2_{v}.json').read().strip().splitlines()[-1])
pb = l['phase_breakdown_ms']
print('inline_pack', v, l['ms_per_step'], l['value'], l['e2e']['value'], pb['train_ms'], pb['tc_tile_cycles_warm']['setup'], pb['tc_tile_cycles_warm']['S1'], pb['tc_tile_cycles_warm']['total'], l['gpu_launches'])
PY
done
