mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bignet.py tests/test_gpu_tensorcore.py -m gpu -q --tb=short 2>&1 | tail -25 > gpurun_out/s9_tests.log; tail -25 gpurun_out/s9_tests.log
timeout 600 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s9_c5bench.json 2> gpurun_out/s9_c5bench.err; tail -c 1200 gpurun_out/s9_c5bench.json; tail -3 gpurun_out/s9_c5bench.err
MAPPO_B200_PAIR=0 timeout 600 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s9_c5bench_nopair.json 2> gpurun_out/s9_c5bench_nopair.err; tail -c 700 gpurun_out/s9_c5bench_nopair.json
