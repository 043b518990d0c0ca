# 2 GPUs: multi-rank parity test, c4 (strong scaling, tcgen05 GRU pipeline + gradient all-reduce), c2 (weak scaling)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 > gpurun_out/s29_multi.log; cat gpurun_out/s29_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config c4 --steps 4 --warmup 3 --no-extras > gpurun_out/s29_c4_n2.json 2> gpurun_out/s29_c4_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s29_c4_n2.json').read().strip().splitlines()[-1])
    print('c4 n2', l['ms_per_step'], l['value'], l['e2e']['value'], l['config'].get('collective'), l['config'].get('replica_checksums_identical_across_ranks'), l.get('phase_breakdown_ms'))
except Exception as e:
    print('c4 n2 failed', e); print(open('gpurun_out/s29_c4_n2.err').read()[-2000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s29_c2_n2.json 2> gpurun_out/s29_c2_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s29_c2_n2.json').read().strip().splitlines()[-1])
    print('c2 n2', l['ms_per_step'], l['value'], l['e2e']['value'], l['config'].get('replica_checksums_identical_across_ranks'))
except Exception as e:
    print('c2 n2 failed', e); print(open('gpurun_out/s29_c2_n2.err').read()[-2000:])
PY
