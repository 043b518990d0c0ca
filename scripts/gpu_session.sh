#!/bin/bash
# One GPU session of round 2: new-path tests -> c5 bench -> ncu launch list + full capture -> whole GPU suite -> default bench.
# Everything lands in gpurun_out/.  Usage (through gpurun): bash scripts/gpu_session.sh [stage...]
set -u
mkdir -p gpurun_out
STAGES="${@:-newtests c5bench ncu alltests bench}"
for st in $STAGES; do
  echo "=== stage $st $(date +%T)"
  case $st in
    newtests)
      timeout 900 python -m pytest tests -m gpu -q -k "c5_h512 or c2_mlp_n128 or bignet or separated or tensorcore or clip_adam or reference_scenario" -s 2>&1 | tail -120 > gpurun_out/s_newtests.log
      tail -40 gpurun_out/s_newtests.log ;;
    c5bench)
      timeout 900 python bench.py --config c5 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s_c5bench.json 2> gpurun_out/s_c5bench.err
      tail -c 3000 gpurun_out/s_c5bench.json; tail -5 gpurun_out/s_c5bench.err ;;
    ncu)
      timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c5_train.csv \
          python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s_ncu1.log 2>&1
      tail -3 gpurun_out/s_ncu1.log
      timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_lin -c 5 -o gpurun_out/r2_big_lin \
          python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s_ncu2.log 2>&1
      timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_grad -c 2 -o gpurun_out/r2_big_grad \
          python scripts/profile_big.py --threads 1024 --epochs 1 > gpurun_out/s_ncu3.log 2>&1
      tail -3 gpurun_out/s_ncu2.log gpurun_out/s_ncu3.log; ls -la gpurun_out/*.ncu-rep ;;
    alltests)
      timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/s_alltests.log
      tail -30 gpurun_out/s_alltests.log ;;
    bench)
      timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
      tail -c 2500 gpurun_out/s_bench.json; tail -5 gpurun_out/s_bench.err ;;
    fusedbench)
      MAPPO_B200_FUSED_TAIL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 0 > gpurun_out/s_bench_fused.json 2> gpurun_out/s_bench_fused.err
      tail -c 1500 gpurun_out/s_bench_fused.json; tail -3 gpurun_out/s_bench_fused.err
      MAPPO_B200_FUSED_TAIL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 0 > gpurun_out/s_bench_unfused.json 2> gpurun_out/s_bench_unfused.err
      tail -c 1500 gpurun_out/s_bench_unfused.json; tail -3 gpurun_out/s_bench_unfused.err ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ;;
    sanitizer)
      # compute-sanitizer over the hand-rolled synchronisation: mbarrier / TMA / tcgen05 pipelines, cluster kernels, peer flags
      timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest -q -x -m gpu \
          "tests/test_gpu_bignet.py::test_big_lin_kernel_matches_matmul" "tests/test_gpu_bignet.py::test_big_grad_kernel_matches_matmul" \
          "tests/test_gpu_bignet.py::test_ppo_update_gradients_match_oracle" "tests/test_gpu_tensorcore.py::test_fused_optimiser_tail_matches_the_separate_launches" \
          "tests/test_gpu_mpe_env.py" -k "not 40000" 2>&1 | tail -25 > gpurun_out/r2_sanitizer_memcheck.log
      tail -8 gpurun_out/r2_sanitizer_memcheck.log
      timeout 1200 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest -q -x -m gpu \
          "tests/test_gpu_tensorcore.py::test_tf32_first_update_gradients" "tests/test_gpu_tensorcore.py::test_fused_optimiser_tail_matches_the_separate_launches" \
          "tests/test_gpu_bignet.py::test_ppo_update_gradients_match_oracle" -k "c1_mlp or c2_mlp or h128 or h256" 2>&1 | tail -25 > gpurun_out/r2_sanitizer_racecheck.log
      tail -8 gpurun_out/r2_sanitizer_racecheck.log ;;
    c4bench)
      timeout 900 python bench.py --config c4 --steps 3 --no-extras --cpu-iters 0 > gpurun_out/s_c4bench.json 2> gpurun_out/s_c4bench.err
      tail -c 1500 gpurun_out/s_c4bench.json; tail -3 gpurun_out/s_c4bench.err ;;
  esac
done
echo "=== done $(date +%T)"
