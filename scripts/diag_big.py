"""Diagnostic for the hidden >= 128 GEMM pipeline: after one ppo_update, re-derive stored intermediates of the backward pass in
float64 from the pipeline's OWN stored inputs (activations, row statistics, upstream gradient, packed weights) and report where
they disagree.  usage: python scripts/diag_big.py <config> <n_rows> <fp32|tf32> [actor|critic]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "on-policy_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    name, n_rows, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    which = sys.argv[4] if len(sys.argv) > 4 else "critic"
    os.environ["MAPPO_B200_GEMM"] = mode
    from oracle import mappo_oracle as O
    import test_gpu_parity as TP
    import test_gpu_bignet as TB
    from mappo_b200 import _lib
    cfg = O.PathConfig(episode_length=4, n_rollout_threads=4, num_agents=2, use_max_grad_norm=False, entropy_coef=0.015,
                       lr=7e-4, critic_lr=1e-3, **TB.CONFIGS[name])
    torch.manual_seed(11)
    args, policy, trainer, buf = TP.build(cfg)
    TB._perturb(policy, 5)
    trainer.value_normalizer.state.copy_(torch.tensor([0.3e-4, 1.7e-4, 1.2e-4]))
    sample = TB._sample(cfg, n_rows, 17)
    trainer.ppo_update(sample)
    torch.cuda.synchronize()
    k = 0 if which == "actor" else 1
    net = policy.actor if k == 0 else policy.critic
    ws = trainer._workspaces(n_rows)[k].workspace.cpu().numpy().astype(np.float64)
    off = (C.c_int64 * 64)()
    _lib.check(_lib.load().mappo_debug_big_plan(C.byref(net.desc), n_rows, off))
    off = list(off)
    H, Hx, Lh, K0p = off[0:4]
    R = (n_rows + 127) // 128 * 128
    relu = bool(cfg.use_ReLU)
    g = lambda o, r, c: ws[o:o + r * c].reshape(r, c)
    x0 = g(off[4], R, K0p)[:n_rows]
    P = [g(off[5], R, H)[:n_rows], g(off[6], R, H)[:n_rows]]
    Ph = g(off[7], R, 32)[:n_rows]
    whf, whft = g(off[10], 32, H), g(off[11], H, 32)
    cvh = ws[off[12]:off[12] + 64]
    act = {l: g(off[16 + 4 * l], R, Hx)[:n_rows] for l in range(1, Lh + 1)}
    stats = {l: g(off[17 + 4 * l], R, 2)[:n_rows] for l in range(1, Lh + 1)}
    mpr = {l: g(off[18 + 4 * l], R, 2)[:n_rows] for l in range(1, Lh + 1)}
    wf = {i: g(off[40 + 3 * i], H, K0p if i == 0 else H) for i in range(Lh)}
    wft = {i: g(off[41 + 3 * i], H, H) for i in range(1, Lh)}
    cv = {i: ws[off[42 + 3 * i]:off[42 + 3 * i] + 2 * H] for i in range(Lh)}
    print(f"{name} {which} n={n_rows} {mode}: H {H} Lh {Lh} K0p {K0p}")

    def rep(tag, got, want):
        err = np.abs(got - want)
        sc = np.abs(want).max() + 1e-300
        r = err.max(axis=1) if err.ndim == 2 else err
        worst = np.argsort(-r)[:6]
        print(f"  {tag}: max err {err.max():.3e} / scale {sc:.3e} = {err.max() / sc:.3e}; rows over 1e-3 of scale: {(r > 1e-3 * sc).sum()} / {len(r)}; worst rows {worst.tolist()}")
        if err.ndim == 2 and err.max() > 1e-3 * sc:
            rr = worst[0]
            cols = np.argsort(-err[rr])[:8]
            print(f"    row {rr}: worst cols {cols.tolist()} got {got[rr, cols]} want {want[rr, cols]}")
            print(f"    bad columns histogram (32-col chunks): {[(int((err[:, c:c + 32].max(axis=1) > 1e-3 * sc).sum())) for c in range(0, err.shape[1], 32)]}")
        return err.max() / sc

    # forward consistency: stats of stored activations, extension columns
    for l in range(1, Lh + 1):
        a = act[l][:, :H]
        mu = a.mean(1); sig = np.sqrt(a.var(1) + 1e-5)
        rep(f"stats[{l}].mu", stats[l][:, 0], mu); rep(f"stats[{l}].rs", stats[l][:, 1], 1.0 / sig)
        rep(f"act[{l}] ext mu", act[l][:, H], mu); rep(f"act[{l}] ext sigma", act[l][:, H + 1], sig)
        A_in = x0 if l == 1 else act[l - 1][:, :H]
        acc = A_in @ wf[l - 1].T
        if l == 1:
            z = acc + cv[0][H:]
        else:
            z = stats[l - 1][:, 1:2] * (acc - stats[l - 1][:, 0:1] * cv[l - 1][:H]) + cv[l - 1][H:]
        rep(f"act[{l}] from its inputs", a, np.maximum(z, 0) if relu else np.tanh(z))
    # head
    aL = act[Lh][:, :H]
    zh = stats[Lh][:, 1:2] * (aL @ whf.T - stats[Lh][:, 0:1] * cvh[:32]) + cvh[32:]
    m1 = (Ph * cvh[:32]).sum(1) / H
    m2 = (Ph * (zh - cvh[32:])).sum(1) / H
    rep(f"mprime[{Lh}].m1", mpr[Lh][:, 0], m1); rep(f"mprime[{Lh}].m2", mpr[Lh][:, 1], m2)
    # backward levels whose upstream gradient still sits in a buffer
    for l in ([Lh] if Lh == 1 else [1, 2] if Lh >= 2 else []):
        if l == Lh:
            up, Wt = Ph, whft                      # acc[row][k] = sum_j Ph[row][j] Wh'[j][k]
        else:
            up, Wt = P[(l + 1) & 1], wft[l]        # acc[row][k] = sum_o P_{l+1}[row][o] W'_l[o][k];  wft[l][k][o]
        if l == 2 and Lh >= 3:
            continue                               # P_3 has been overwritten by P_1
        acc = up @ Wt.T
        a = act[l][:, :H]
        xh = (a - stats[l][:, 0:1]) * stats[l][:, 1:2]
        dA = acc - (mpr[l][:, 0:1] + xh * mpr[l][:, 1:2])
        d = (a > 0).astype(np.float64) if relu else 1.0 - a * a
        rs_prev = stats[l - 1][:, 1:2] if l > 1 else 1.0
        want = dA * d * rs_prev
        rep(f"P_{l}", P[l & 1], want)
        if l > 1:
            zin = a if relu else np.arctanh(np.clip(a, -0.99999994, 0.99999994))
            rep(f"mprime[{l - 1}].m1", mpr[l - 1][:, 0], (P[l & 1] * cv[l - 1][:H]).sum(1) / H)
            rep(f"mprime[{l - 1}].m2", mpr[l - 1][:, 1], (P[l & 1] * (zin - cv[l - 1][H:])).sum(1) / H)
    # weight gradient of matrix 0 from the stored P_1 and x0 (the last thing left in the partial / gsum buffers)
    G0 = P[1].T @ x0
    print("  (G0 = P_1^T x0 recomputed: scale %.3e)" % np.abs(G0).max())


if __name__ == "__main__":
    main()
