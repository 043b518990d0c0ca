mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --config c4 --steps 4 --warmup 3 --no-extras > gpurun_out/s37_c4_n2.json 2> gpurun_out/s37_c4_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s37_c4_n2.json').read().strip().splitlines()[-1])
    pb = l.get('phase_breakdown_ms') or {}
    print('c4 n2', l['ms_per_step'], l['value'], l['e2e']['value'], l['config'].get('replica_checksums_identical_across_ranks'), {k: round(v, 3) for k, v in pb.items() if k.endswith('_ms')})
except Exception as e:
    print('c4 n2 failed', e); print(open('gpurun_out/s37_c4_n2.err').read()[-2000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config c3 --steps 10 --warmup 3 --no-extras > gpurun_out/s37_c3_n2.json 2> gpurun_out/s37_c3_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/s37_c3_n2.json').read().strip().splitlines()[-1])
    print('c3 n2', l['ms_per_step'], l['value'], l['e2e']['value'], l['config'].get('replica_checksums_identical_across_ranks'))
except Exception as e:
    print('c3 n2 failed', e); print(open('gpurun_out/s37_c3_n2.err').read()[-2000:])
PY
