mkdir -p gpurun_out
timeout 600 python scripts/diag_gru_tc.py c3_gru_multidiscrete > gpurun_out/s20_diag.log 2>&1; tail -80 gpurun_out/s20_diag.log
