# session 2 opening: whole GPU suite at HEAD, c3 / c4 bench lines with phase breakdown, ncu launch lists of the GRU path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/s18_alltests.log; tail -5 gpurun_out/s18_alltests.log
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s18_c3.json 2> gpurun_out/s18_c3.err; tail -c 2500 gpurun_out/s18_c3.json; tail -3 gpurun_out/s18_c3.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-extras --cpu-iters 0 > gpurun_out/s18_c4.json 2> gpurun_out/s18_c4.err; tail -c 2500 gpurun_out/s18_c4.json; tail -3 gpurun_out/s18_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s18_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s18_ncu_c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s18_launches_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-extras --cpu-iters 0 --no-breakdown --eager > gpurun_out/s18_ncu_c4.log 2>&1
tail -2 gpurun_out/s18_ncu_c3.log gpurun_out/s18_ncu_c4.log
