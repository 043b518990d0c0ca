mkdir -p gpurun_out
MAPPO_B200_PAIR_LIN=1 timeout 600 python -m pytest tests/test_gpu_bignet.py -m gpu -q --tb=short 2>&1 | tail -5
MAPPO_B200_PAIR_LIN=1 timeout 600 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s12_c5bench_pairlin.json 2> gpurun_out/s12_c5bench_pairlin.err; tail -c 900 gpurun_out/s12_c5bench_pairlin.json; tail -3 gpurun_out/s12_c5bench_pairlin.err
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --tb=short 2>&1 | tail -5
bash scripts/gpu_session.sh fusedbench 2>&1 | tail -8 | cut -c1-400
