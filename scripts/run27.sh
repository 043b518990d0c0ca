mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_persistent -c 1 -o gpurun_out/s27_rollout_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s27_ncu.log 2>&1
tail -2 gpurun_out/s27_ncu.log; ls -la gpurun_out/s27_rollout_c4.ncu-rep
