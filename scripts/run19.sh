# tcgen05 GRU pipeline: first parity run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -q -x -s -k "gru or naive_rnn" 2>&1 | tail -60 > gpurun_out/s19_gru.log; tail -40 gpurun_out/s19_gru.log
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -q -k "not gru and not naive_rnn" 2>&1 | tail -15 > gpurun_out/s19_tc_rest.log; tail -8 gpurun_out/s19_tc_rest.log
