"""Diagnostic: run-to-run determinism of the hidden-64 tcgen05 train() and fused-vs-separate tail differences per tensor."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "on-policy_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ["MAPPO_B200_GEMM"] = "tf32"
from helpers import Golden
import test_gpu_parity as TP


def run(fused, epochs=None):
    os.environ["MAPPO_B200_FUSED_TAIL"] = fused
    g = Golden("c2_mlp_n128")
    cfg = g.cfg
    if epochs is not None:
        cfg.ppo_epoch = epochs
    args, policy, trainer, buf = TP.build(cfg, g)
    feed = g.feed(0)
    TP.warm(buf, feed)
    os.environ["MAPPO_B200_GEMM"] = "fp32"
    TP.collect_and_returns(cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    os.environ["MAPPO_B200_GEMM"] = "tf32"
    perms = TP.FakeRandperm(g.get("it0/perms"))
    orig = torch.randperm
    torch.randperm = perms
    try:
        info = trainer.train(buf)
    finally:
        torch.randperm = orig
    st = {}
    for nm, net, opt in (("actor", policy.actor, policy.actor_optimizer), ("critic", policy.critic, policy.critic_optimizer)):
        st[nm + "/flat"] = net.flat.cpu().numpy().copy(); st[nm + "/grad"] = net.grad.cpu().numpy().copy()
        st[nm + "/m"] = opt.exp_avg.cpu().numpy().copy(); st[nm + "/v"] = opt.exp_avg_sq.cpu().numpy().copy()
    return st, info


def diff(a, b, tag):
    for k in a:
        n = int((a[k] != b[k]).sum())
        print(f"  {tag} {k}: {n} / {a[k].size} differ" + (f" max {np.abs(a[k]-b[k]).max():.3e}" if n else ""))


for ep in (1, 2, 10):
    print("epochs", ep)
    u0, i0 = run("0", ep); u1, i1 = run("0", ep); f0, j0 = run("1", ep); f1, j1 = run("1", ep)
    diff(u0, u1, "unfused vs unfused")
    diff(f0, f1, "fused vs fused")
    diff(u0, f0, "unfused vs fused")
