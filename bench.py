#!/usr/bin/env python
"""bench.py -- env-steps/s through collect -> GAE -> ppo_update on synthetic MPE-shaped rollouts.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): MPE simple_spread-shaped, 3 agents, 128 rollout threads PER GPU,
episode_length 25, shared MLP policy (Tanh, hidden 64, layer_N 1), ppo_epoch 10, num_mini_batch 1 -- the effective
hyper-parameters of train_mpe_spread.sh.  One "step" = one full iteration:
    25 x (policy forward + sample + insert) -> get_values + compute_returns -> 10 x (actor + critic update) -> after_update
and processes N_threads * T = 3200 env steps per GPU.  Weak scaling over rollout threads.

`value`  : iterations replayed from inputs already resident in HBM, device-timed with CUDA events.
`e2e`    : the same iteration driven from HOST buffers (pinned env outputs -> H2D, D2H of train_info) per step.
`--impl reference` : the CPU restatement of the reference path (oracle/, see its header) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "on-policy_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "env-steps/sec through collect->GAE->ppo_update (MAPPO, MPE simple_spread-shaped c2)"
UNIT = "env-steps/s"


def c2_config():
    from oracle import mappo_oracle as O
    return O.PathConfig(episode_length=25, n_rollout_threads=128, num_agents=3, obs_dim=18, share_obs_dim=54,
                        act_dims=(5,), use_ReLU=False, ppo_epoch=10, num_mini_batch=1, lr=7e-4, critic_lr=7e-4)


def workload_dict(cfg, n_gpus):
    return {"workload": "c2: MPE simple_spread-shaped, 3 agents x 128 rollout threads per GPU x 25 steps, shared MLP "
                        "(tanh, H=64), ppo_epoch 10, 1 minibatch",
            "rollout_threads_per_gpu": cfg.n_rollout_threads, "global_rollout_threads": cfg.n_rollout_threads * n_gpus,
            "episode_length": cfg.episode_length, "num_agents": cfg.num_agents, "ppo_epoch": cfg.ppo_epoch,
            "parallelism": f"dp{n_gpus} over rollout threads", "l2_flush_between_steps": True}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_iteration_rate(cfg, iters, warmup, threads):
    import torch
    from oracle import mappo_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    learner = O.Learner(cfg, O.init_params(cfg, False, seed=1), O.init_params(cfg, True, seed=2))
    store = O.RolloutStore(cfg)
    feed = O.make_feed(cfg, seed=0)
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.run_iteration(cfg, learner, store, feed)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return cfg.n_rollout_threads * cfg.episode_length * len(times) / tot, tot / len(times)


def best_cpu_threads(cfg):
    """The port (like the reference) is many tiny torch ops: more intra-op threads is not faster.  Probe a few
    counts with two iterations each and keep the fastest ("all the host threads it can USE")."""
    cores = os.cpu_count() or 1
    best, best_rate = 1, 0.0
    for th in sorted({1, min(4, cores), min(8, cores), min(16, cores), min(32, cores)}):
        rate, _ = cpu_iteration_rate(cfg, 2, 1, th)
        if rate > best_rate:
            best, best_rate = th, rate
    return best


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = c2_config()
    cores = best_cpu_threads(cfg)
    rate, per = cpu_iteration_rate(cfg, a.steps, a.warmup, cores)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_dict(cfg, 1),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{a.steps} full iterations of c2 (3200 env steps each) on {cores} torch threads (fastest of 1/4/8/16/32 on a {os.cpu_count()}-core host); "
                                       "oracle/mappo_oracle.py = CPU restatement of the reference (Python reference "
                                       "cannot travel to the GPU box)"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def update_flops(cfg):
    """Algorithmic GEMM FLOPs of one optimiser step per net (3F - 2*in*H per row)."""
    H, L, B = cfg.hidden_size, cfg.layer_N, cfg.episode_length * cfg.n_rollout_threads * cfg.num_agents
    fa = 2 * (cfg.obs_dim * H + L * H * H + H * sum(cfg.act_dims))
    fc = 2 * (cfg.share_obs_dim * H + L * H * H + H)
    return B * (3 * fa - 2 * cfg.obs_dim * H), B * (3 * fc - 2 * cfg.share_obs_dim * H)


def run_gpu(a):
    if os.environ.get("BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    import torch
    import torch.distributed as dist
    from oracle import mappo_oracle as O            # synthetic feed generator + (rank 0) cpu_baseline only
    from argsutil import make_args, make_spaces
    from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy
    from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO
    from onpolicy.utils.shared_buffer import SharedReplayBuffer
    from mappo_b200.engine import RolloutEngine
    from mappo_b200 import core

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = c2_config()
    args = make_args(cfg)
    obs_s, share_s, act_s = make_spaces(cfg)
    os.environ["MAPPO_B200_GEMM"] = a.gemm               # update-kernel GEMM engine: tcgen05 tf32 or exact fp32 FFMA
    torch.manual_seed(1)                                   # identical replicas on every rank
    policy = R_MAPPOPolicy(args, obs_s, share_s, act_s, device=dev)
    trainer = R_MAPPO(args, policy, device=dev)
    buf = SharedReplayBuffer(args, cfg.num_agents, obs_s, share_s, act_s)
    feed = O.make_feed(cfg, seed=100 + rank)               # each rank owns its own 128 rollout threads
    if a.env == "device":
        # closed loop: the simple_spread worlds are stepped on the GPU between policy_step and insert
        from mappo_b200.mpe_env import DeviceSpreadEnv
        env = DeviceSpreadEnv(cfg.n_rollout_threads, cfg.num_agents, 3, cfg.episode_length, device=dev, seed=100 + rank)
        eng = RolloutEngine(args, policy, trainer, buf, rng="device", seed=1 + rank, device_env=env)
        eng.reset_env()
    else:
        eng = RolloutEngine(args, policy, trainer, buf, rng="device", seed=1 + rank, share_obs_from_obs=True)
        eng.stage_feed(feed)
        eng.upload()
    torch.cuda.synchronize()
    graph_ok = not a.eager
    try:
        if a.eager:
            raise RuntimeError("--eager")
        eng.capture(warmup=2)
    except Exception as e:                                 # e.g. a collective that refuses capture: run eagerly
        graph_ok = False
        eng.graph = None
        torch.cuda.synchronize()
        if rank == 0 and not a.eager:
            print(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); running eager", file=sys.stderr)

    if world > 1:                                          # every rank must run the same mode
        ok = torch.tensor([1 if graph_ok else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            graph_ok, eng.graph = False, None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed resident loop ----
    for _ in range(max(a.warmup, 3)):
        eng.step_resident()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    t_wall0 = time.perf_counter()
    for s, e in evs:
        flush.zero_()                                      # L2 flush between timed iterations (outside the events)
        s.record()
        eng.step_resident()
        e.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms = sum(s.elapsed_time(e) for s, e in evs)
    launches = eng.launches_per_iteration * a.steps

    # ---- end to end from host buffers ----
    for _ in range(3):
        eng.step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        info = eng.step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    breakdown = eng.phase_breakdown() if world == 1 else None
    if breakdown is not None and a.gemm == "tf32":
        import ctypes as C
        from mappo_b200 import _lib
        t = (C.c_int64 * 16)()
        torch.cuda.synchronize()
        _lib.load().mappo_debug_tc_timing(t)
        names = ["setup", "S1", "fc1_mma", "S3", "fc2_mma", "S5", "head_mma", "S7_loss", "dx2_Gh_mma", "S9", "dx1_G2_mma",
                 "S11", "dump_G2", "G1_wait", "tail"]
        breakdown["tc_tile_cycles_warm"] = {nm: int(t[i + 1] - t[i]) for i, nm in enumerate(names)}
        breakdown["tc_tile_cycles_warm"]["total"] = int(t[15] - t[0])
    # ---- the dominant kernel, timed live with CUDA events on its own stream (eager pass, one train()) ----
    kt = time_update_kernel(eng, cfg, flush)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t.tolist()
    steps_env = cfg.n_rollout_threads * cfg.episode_length * world
    value = steps_env * a.steps / (ms_max * 1e-3)
    e2e_value = steps_env * a.steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        fa, fc = update_flops(cfg)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("update_mlp_tc_kernel" if a.gemm == "tf32" else "update_mlp_kernel")
        ach = (fa + fc) / 2 / (kt["avg_ms"] * 1e-3) / 1e12
        cores = best_cpu_threads(cfg) if world == 1 and a.cpu_iters > 0 else 1
        cpu_rate, cpu_per = cpu_iteration_rate(cfg, a.cpu_iters, 2, cores) if world == 1 and a.cpu_iters > 0 else (None, None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32 tensor-core GEMMs (fp32 accumulate / fp32 elsewhere)" if a.gemm == "tf32" else "f32",
                "data": "synthetic", "config": {**workload_dict(cfg, world), "cuda_graph": graph_ok, "gemm": a.gemm,
                                                "collective": ("none (1 GPU)" if world == 1 else
                                                               "one-shot peer-memory all-reduce kernel over NVLink, in-graph"
                                                               if getattr(trainer, "_p2p", None) is not None else
                                                               "NCCL all-reduce via torch.distributed between graph segments"),
                                                                "rng": "device (Philox sampling, Feistel permutations)",
                                                "rollout": "persistent kernel, one launch per iteration" if (eng.persistent_rollout or eng.closed_persistent) else "one launch per env step",
                                                "env": ("device-side simple_spread, closed loop " + ("inside one persistent rollout kernel" if eng.closed_persistent else "(policy_step -> env step -> insert per step)")) if a.env == "device" else "synthetic staged env outputs",
                                                "h2d": "obs, rewards, dones (share_obs = concat of the thread's agents' obs is formed on the device)" if eng.share_from_obs else "obs, share_obs, rewards, dones"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eng.h2d_bytes(), "d2h_bytes_per_step": 48,
                        "ms_per_step": e2e_ms_max / a.steps},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                             "traffic": traffic, "traffic_unit": "bytes/launch (ncu --set full, profiles/)",
                             "kernel": ("update_mlp_tc_kernel (fused fwd+loss+bwd, tcgen05 kind::tf32 + TMEM; launch incl. "
                                        "its 1-CTA weight-pack kernel)") if a.gemm == "tf32" else
                                       "update_mlp_kernel (fused fwd+loss+bwd, fp32 FFMA tiles)",
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590",
                             "avg_launch_ms": kt["avg_ms"], "launches_timed": kt["n"],
                             "algorithmic_gflop_per_launch": (fa + fc) / 2 / 1e9,
                             # the actor and the critic chain run concurrently (two graph branches), each with one
                             # update kernel per optimiser step: share of the step's critical path = one chain's kernels
                             "kernel_share_of_step": kt["avg_ms"] * cfg.ppo_epoch * cfg.num_mini_batch / (ms_max / a.steps),
                             "kernel_time_sum_over_step": kt["avg_ms"] * 2 * cfg.ppo_epoch * cfg.num_mini_batch / (ms_max / a.steps)},
                "clocks": clocks, "wall_s_timed_region": t_wall, "phase_breakdown_ms": breakdown,
                "train_info_last": info}
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{a.cpu_iters} full c2 iterations (3200 env steps each) of "
                                              f"oracle/mappo_oracle.py on {cores} torch threads (fastest of 1/4/8/16/32; host has {os.cpu_count()} cores), "
                                              f"{cpu_per*1e3:.0f} ms each"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_update_kernel(eng, cfg, flush):
    """Average duration of the fused update kernel, CUDA events around each launch on the launching stream."""
    import torch
    from mappo_b200 import core
    orig = core._lib.load().mappo_update_fwd_bwd
    pairs = []

    class Timed:
        def __call__(self, *args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = orig(*args)
            e.record()
            pairs.append((s, e))
            return rc

    lib = core._lib.load()
    saved_graph, saved_overlap = eng.graph, eng.trainer.overlap_nets
    eng.graph = None
    eng.trainer.overlap_nets = False          # one launch at a time: the events see this kernel alone on the stream
    try:
        lib.mappo_update_fwd_bwd = Timed()
        flush.zero_()
        eng.step_resident()
        torch.cuda.synchronize()
    finally:
        lib.mappo_update_fwd_bwd = orig
        eng.graph, eng.trainer.overlap_nets = saved_graph, saved_overlap
    ms = [s.elapsed_time(e) for s, e in pairs]
    return {"avg_ms": sum(ms) / max(len(ms), 1), "n": len(ms)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gemm", default=os.environ.get("MAPPO_B200_GEMM", "tf32"), choices=["tf32", "fp32"],
                    help="GEMM engine of the update kernels (tf32 = tcgen05 tensor cores)")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (for per-kernel profiling under ncu)")
    ap.add_argument("--env", default="staged", choices=["staged", "device"],
                    help="staged: synthetic env outputs uploaded per iteration (the BASELINE metric: the path only); "
                         "device: closed loop with the device-side simple_spread environment")
    ap.add_argument("--cpu-iters", type=int, default=100, help="oracle iterations for cpu_baseline (rank 0, N=1)")
    a = ap.parse_args()
    if a.impl == "reference":
        if a.steps > 60:
            a.steps = 60
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
