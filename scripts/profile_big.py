#!/usr/bin/env python
"""Profiling driver for the hidden >= 128 GEMM pipeline (c5 nets): builds ONE c5 policy at `--threads` rollout threads, fills the
rollout storage with one eager collect, then runs `--epochs` PPO epochs of train() between cudaProfilerStart / Stop so that

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python scripts/profile_big.py --threads 1024 --epochs 1
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:big_lin -c 4 -o prof \
        python scripts/profile_big.py --threads 1024 --epochs 1

see exactly one train() of the pipeline (the launch list of `--collect` adds the per-step rollout launches)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "on-policy_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=1024)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--gemm", default="tf32")
    ap.add_argument("--collect", action="store_true", help="profile the collect phase too")
    ap.add_argument("--config", default="c5")
    a = ap.parse_args()
    os.environ["MAPPO_B200_GEMM"] = a.gemm
    import torch
    import bench
    w = bench.workload(a.config, 1, threads=a.threads)
    w["cfg"].ppo_epoch = a.epochs
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ns = argparse.Namespace(env="staged", gemm=a.gemm)
    job = bench.Job(w, 0, dev, 0, ns)
    eng = job.eng
    eng.launch_iteration()                      # warm: allocations, attributes, storage filled by a real collect
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    if a.collect:
        eng.launch_iteration()
    else:
        eng._epoch_i = 0
        job.trainer.launch_train(job.buf, True, eng._draw_perm, eng.loss_out, allreduce=None)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled", "iteration" if a.collect else "train()", "launches per iteration:", eng.launches_per_iteration,
          "loss sums:", eng.loss_out.tolist())


if __name__ == "__main__":
    main()
