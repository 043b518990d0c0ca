mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/s16_alltests.log; tail -8 gpurun_out/s16_alltests.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_c5_train.csv python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s16_ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_grad_pair -s 1 -c 1 -o gpurun_out/r2d_big_grad_pair python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s16_ncu2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_unfold_kernel -s 1 -c 1 -o gpurun_out/r2d_big_unfold python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s16_ncu3.log 2>&1
ls -la gpurun_out/r2d*
bash scripts/gpu_session.sh sanitizer 2>&1 | tail -20
