mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/s24_launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s24_ncu_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s24_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s24_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gru_tc|update_mlp_tc" -c 6 -o gpurun_out/s24_gru_tc_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s24_ncu_full.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/s24_alltests.log; tail -6 gpurun_out/s24_alltests.log
