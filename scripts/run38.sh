mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_gru_fast -c 1 -o gpurun_out/s38_rollout_gru_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s38_ncu.log 2>&1
tail -1 gpurun_out/s38_ncu.log; ls -la gpurun_out/s38_rollout_gru_c4.ncu-rep
