"""Diagnostics for the tcgen05 GRU pipeline: one ppo_update of a GRU golden in the fp32 (update_gru.cu) and tf32 (update_gru_tc.cu)
builds on identical rollouts; compares every workspace plane the two pipelines share and every gradient tensor (no asserts)."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "on-policy_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import Golden  # noqa: E402
import test_gpu_parity as TP  # noqa: E402
from oracle import mappo_oracle as O  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3_gru_multidiscrete"
g = Golden(name)
cfg = g.cfg
one = O.PathConfig(**{**cfg.to_dict(), "ppo_epoch": 1, "act_dims": tuple(cfg.act_dims)})


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


res = {}
for mode in ("fp32", "tf32"):
    os.environ["MAPPO_B200_GEMM"] = mode
    args, policy, trainer, buf = TP.build(one, g)
    feed = g.feed(0)
    TP.warm(buf, feed)
    os.environ["MAPPO_B200_GEMM"] = "fp32"
    TP.collect_and_returns(cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    os.environ["MAPPO_B200_GEMM"] = mode
    torch.randperm = TP.FakeRandperm([g.get("it0/perms")[0]])
    info = trainer.train(buf)
    torch.cuda.synchronize()
    (key, (ws_a, ws_c)), = trainer._ws.items()
    planes = {}
    for nm, ws in (("actor", ws_a), ("critic", ws_c)):
        w = ws.workspace.cpu().numpy()
        P = key
        if mode == "fp32":
            pl = w[:8 * P * 64].reshape(8, P, 64)
            planes[nm] = dict(zip(["FEAT", "HM", "R", "Z", "N", "GHN", "DHH", "DFEAT"], pl))
        else:
            T = (P + 127) // 128                 # planes are tiled [p / 128][16 chunks][128 rows][4] (tc64.cuh: pl_off)
            pl = w[-10 * T * 8192:].reshape(10, T, 16, 128, 4).transpose(0, 1, 3, 2, 4).reshape(10, T * 128, 64)[:, :P]
            planes[nm] = dict(zip(["X", "R", "Z", "N", "GHN", "H", "DFEAT", "DR", "DZ", "DN"], pl))
    grads = {("actor", k): v.cpu().numpy().copy() for k, v in policy.actor.named_grads().items()}
    grads.update({("critic", k): v.cpu().numpy().copy() for k, v in policy.critic.named_grads().items()})
    res[mode] = (info, planes, grads, key)
    print(mode, "rows", key, {k: round(float(v), 6) for k, v in info.items()})

P = res["fp32"][3]
Nc = P // cfg.data_chunk_length if cfg.use_recurrent_policy else None
for nm in ("actor", "critic"):
    a, b = res["tf32"][1][nm], res["fp32"][1][nm]
    print(f"--- {nm}: planes, relative L2 (tf32 pipeline vs fp32 pipeline)")
    print("   X (xhat2) vs FEAT (y2; equal while gamma = 1, beta = 0):", rel(a["X"], b["FEAT"]))
    for k in ("R", "Z", "N", "GHN"):
        print(f"   {k}:", rel(a[k], b[k]))
    if Nc:
        print("   H[l] vs HM[l+1] (where mask = 1):", rel(a["H"][:-Nc], b["HM"][Nc:]))
    print("   DFEAT:", rel(a["DFEAT"], b["DFEAT"]))
    for l in range(P // Nc if Nc else 0):
        sl = slice(l * Nc, (l + 1) * Nc)
        print(f"     step {l}: R {rel(a['R'][sl], b['R'][sl]):.2e} N {rel(a['N'][sl], b['N'][sl]):.2e} DFEAT {rel(a['DFEAT'][sl], b['DFEAT'][sl]):.2e}")
print("--- gradients, relative L2 / max err over scale")
for k, want in res["fp32"][2].items():
    got = res["tf32"][2][k]
    print(f"   {k[0]:6s} {k[1]:40s} relL2 {rel(got, want):.3e}   max|err|/scale {np.abs(got - want).max() / (np.abs(want).max() + 1e-300):.3e}  |want| {np.abs(want).max():.2e}")
