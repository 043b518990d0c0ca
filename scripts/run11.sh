mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tests/cuda/umma_rate.cu -o /tmp/umma_rate && /tmp/umma_rate > gpurun_out/r2_umma_rate.log 2>&1; cat gpurun_out/r2_umma_rate.log
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/s11_alltests.log; tail -15 gpurun_out/s11_alltests.log
MAPPO_B200_PAIR_LIN=1 timeout 600 python -m pytest tests/test_gpu_bignet.py -m gpu -q --tb=short 2>&1 | tail -5
bash scripts/gpu_session.sh bench c4bench 2>&1 | tail -12 | cut -c1-1500
