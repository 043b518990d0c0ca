mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python -m pytest -q -x -m gpu \
   "tests/test_gpu_tensorcore.py::test_tf32_first_update_gradients" "tests/test_gpu_bignet.py::test_ppo_update_gradients_match_oracle" -k "c1_mlp or h128 or h256" 2>&1 | grep -v "Host Frame" | head -c 60000 > gpurun_out/r2_sanitizer_racecheck_full.log
grep -c "Race reported\|hazard" gpurun_out/r2_sanitizer_racecheck_full.log; grep "Race reported\|hazard detected\| in \|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_racecheck_full.log | sort | uniq -c | sort -rn | head -40
