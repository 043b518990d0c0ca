mkdir -p gpurun_out
for a in "h256_relu_l0 128 tf32 critic" "h256_relu_l0 128 fp32 critic" "h512_relu 300 tf32 actor" "h512_relu 5000 fp32 actor" "h512_relu 300 fp32 actor" "h128_tanh_multi 333 tf32 actor"; do
  python scripts/diag_big.py $a 2>&1 | tail -40
done > gpurun_out/s4_diag_big.log 2>&1
tail -60 gpurun_out/s4_diag_big.log
python -m pytest tests/test_gpu_bignet.py -m gpu -q --tb=line 2>&1 | tail -12
