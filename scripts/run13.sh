mkdir -p gpurun_out
bash scripts/gpu_session.sh fusedbench 2>&1 | tail -8 | cut -c1-300
MAPPO_B200_FUSED_TAIL=1 timeout 600 nsys --version 2>/dev/null | head -1
MAPPO_B200_FUSED_TAIL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_c2_fused.csv -s 300 -c 200 python bench.py --steps 2 --warmup 3 --no-extras --no-breakdown --cpu-iters 0 --eager > gpurun_out/s13_ncu.log 2>&1; tail -3 gpurun_out/s13_ncu.log
