mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v '^  *\[' | head -c 200000 > gpurun_out/s3_alltests.log; tail -30 gpurun_out/s3_alltests.log
timeout 900 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s3_c5bench.json 2> gpurun_out/s3_c5bench.err; tail -c 1800 gpurun_out/s3_c5bench.json; tail -3 gpurun_out/s3_c5bench.err
bash scripts/gpu_session.sh fusedbench 2>&1 | tail -20
