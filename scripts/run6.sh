mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v '^  *\[' | head -c 300000 > gpurun_out/s6_alltests.log; tail -12 gpurun_out/s6_alltests.log
bash scripts/gpu_session.sh fusedbench 2>&1 | tail -8 | cut -c1-600
python scripts/diag_tail.py > gpurun_out/s6_diag_tail.log 2>&1; tail -8 gpurun_out/s6_diag_tail.log
