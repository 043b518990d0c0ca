mkdir -p gpurun_out
MAPPO_B200_PDL=1 timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | tail -6
for pdl in 0 1; do
  MAPPO_B200_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 0 > gpurun_out/s15_bench_pdl$pdl.json 2> gpurun_out/s15_bench_pdl$pdl.err
  tail -c 400 gpurun_out/s15_bench_pdl$pdl.json; tail -2 gpurun_out/s15_bench_pdl$pdl.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 0 --no-prefetch > gpurun_out/s15_bench_noprefetch.json 2> gpurun_out/s15_bench_noprefetch.err
