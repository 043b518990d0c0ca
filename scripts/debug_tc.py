"""Compare the tcgen05 update kernel against the fp32 FFMA kernel on identical inputs (debug aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "on-policy_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import Golden
import test_gpu_parity as TP
from oracle import mappo_oracle as O
from mappo_b200 import _lib

g = Golden("c1_mlp_discrete")
cfg = g.cfg
one = O.PathConfig(**{**cfg.to_dict(), "ppo_epoch": 1, "act_dims": tuple(cfg.act_dims)})
res = {}
for mode in ("fp32", "tf32"):
    os.environ["MAPPO_B200_GEMM"] = mode
    args, policy, trainer, buf = TP.build(one, g)
    feed = g.feed(0)
    TP.warm(buf, feed)
    TP.collect_and_returns(cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    real = torch.randperm
    torch.randperm = TP.FakeRandperm([g.get("it0/perms")[0]])
    try:
        info = trainer.train(buf)
    finally:
        torch.randperm = real
    res[mode] = (info, {("actor", k): v.cpu().numpy().copy() for k, v in policy.actor.named_grads().items()} |
                 {("critic", k): v.cpu().numpy().copy() for k, v in policy.critic.named_grads().items()})
for k in res["fp32"][0]:
    print(f"{k:18s} fp32 {res['fp32'][0][k]: .6e}  tf32 {res['tf32'][0][k]: .6e}")
for key in res["fp32"][1]:
    a, b = res["tf32"][1][key], res["fp32"][1][key]
    scale = np.abs(b).max() + 1e-30
    print(f"{key[0]:6s} {key[1]:34s} shape {str(b.shape):10s} scale {scale:.3e} maxerr/scale {np.abs(a-b).max()/scale:.3e}  "
          f"corr {np.corrcoef(a.ravel(), b.ravel())[0,1] if a.size > 1 else float('nan'):.5f}")

# ---- phase timing of the last tcgen05 launch (critic of the last update) ----
import ctypes as C
lib = _lib.load()
buf16 = (C.c_int64 * 16)()
torch.cuda.synchronize()
lib.mappo_debug_tc_timing(buf16)
t = list(buf16)
names = ["setup", "S1 gather+LN0", "fc1 mma", "S3 epi", "fc2 mma", "S5 epi", "head mma", "S7 loss", "dx2+Gh mma", "S9 bwd",
         "dx1+G2 mma", "S11 bwd", "dump G2", "G1 wait", "tail"]
print("TC kernel phase cycles (CTA 0):", {n: t[i + 1] - t[i] for i, n in enumerate(names)}, "total", t[15] - t[0])
