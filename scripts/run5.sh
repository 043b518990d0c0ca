mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/s5_alltests.log; tail -25 gpurun_out/s5_alltests.log
timeout 900 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s5_c5bench.json 2> gpurun_out/s5_c5bench.err; tail -c 1500 gpurun_out/s5_c5bench.json; tail -3 gpurun_out/s5_c5bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_c5_train.csv python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s5_ncu1.log 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_lin -c 5 -o gpurun_out/r2b_big_lin python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s5_ncu2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_grad -s 1 -c 1 -o gpurun_out/r2b_big_grad python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s5_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
