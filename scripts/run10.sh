mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bignet.py tests/test_gpu_tensorcore.py -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/s10_tests.log; tail -12 gpurun_out/s10_tests.log
timeout 600 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s10_c5bench.json 2> gpurun_out/s10_c5bench.err; tail -c 1300 gpurun_out/s10_c5bench.json; tail -3 gpurun_out/s10_c5bench.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_lin -s 1 -c 2 -o gpurun_out/r2c_big_lin_pair python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s10_ncu2.log 2>&1
MAPPO_B200_FUSED_TAIL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_tail -s 6 -c 2 -o gpurun_out/r2c_tc_tail python bench.py --steps 3 --warmup 3 --no-extras --no-breakdown --cpu-iters 0 > gpurun_out/s10_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
