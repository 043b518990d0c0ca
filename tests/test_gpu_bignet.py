"""Hidden >= 128 MLP nets (BASELINE config c5: hidden 512, layer_N 2) -- the layer-by-layer GEMM pipeline of
csrc/big_*.cu -- and the folded-LayerNorm algebra of BOTH tensor-core builds, against the CPU oracle.

One random minibatch, every LayerNorm weight / bias perturbed away from its (1, 0) initial value (the golden fixtures
start from the reference's initial weights, where every LayerNorm bias is zero and the `db' beta^T` term of the
folding chain rule vanishes), one `R_MAPPO.ppo_update`: first-update gradients, loss scalars, gradient norms.

Tolerances: tests/helpers.py grad_agreement (element-wise for the smooth tanh nets; relative L2 + outlier fraction for ReLU nets, whose
on / off ties make a few rows of dW differ by one sample's contribution); losses 1e-4 (fp32 build) / 2e-3 (tcgen05 build).
"""
import numpy as np
import pytest
import torch

from oracle import mappo_oracle as O
from helpers import assert_close, grad_agreement
import test_gpu_parity as TP

pytestmark = pytest.mark.gpu

CONFIGS = {
    # the fused single-kernel path (hidden 64): tcgen05 build of update_mlp_tc.cu / FFMA build of update_mlp.cu
    "h64_tanh": dict(obs_dim=18, share_obs_dim=54, act_dims=(5,), hidden_size=64, layer_N=1, use_ReLU=False),
    # the GEMM pipeline: one N tile of 128, tanh (inverse activation in the backward row statistics), MultiDiscrete heads
    "h128_tanh_multi": dict(obs_dim=37, share_obs_dim=70, act_dims=(5, 7), multi_discrete=True, hidden_size=128, layer_N=1,
                            use_ReLU=False),
    # c5 widths: two N tiles of 256 per row, three hidden matrices, K = 660 / 785 inputs
    "h512_relu": dict(obs_dim=660, share_obs_dim=785, act_dims=(20,), hidden_size=512, layer_N=2, use_ReLU=True),
    # one N tile of 256, no hidden H x H matrix
    "h256_relu_l0": dict(obs_dim=50, share_obs_dim=90, act_dims=(9,), hidden_size=256, layer_N=0, use_ReLU=True),
}


def _perturb(policy, seed):
    g = torch.Generator().manual_seed(seed)
    for net in (policy.actor, policy.critic):
        sd = net.state_dict()
        for k, v in sd.items():
            if "feature_norm" in k or ".2." in k or "norm" in k:           # LayerNorm affine parameters
                if k.endswith("weight"):
                    v.copy_((1.0 + 0.3 * torch.randn(v.shape, generator=g)).to(v.device))
                else:
                    v.copy_((0.3 * torch.randn(v.shape, generator=g)).to(v.device))
            elif k.endswith("bias"):
                v.copy_((0.1 * torch.randn(v.shape, generator=g)).to(v.device))
            elif "action_out" in k or "v_out" in k:                         # gain-0.01 heads: make the logits matter
                v.mul_(20.0)


def _sample(cfg, n, seed):
    rng = np.random.RandomState(seed)
    obs = (rng.randn(n, cfg.obs_dim) * 1.5 + 0.3).astype(np.float32)
    cent = (rng.randn(n, cfg.share_obs_dim) * 1.5 - 0.2).astype(np.float32)
    h = np.zeros((n, 1, cfg.hidden_size), np.float32)
    masks = np.ones((n, 1), np.float32)
    active = (rng.rand(n, 1) > 0.2).astype(np.float32)
    acts = np.stack([rng.randint(0, a, size=n) for a in cfg.act_dims], 1).astype(np.float32)
    avail = None
    if cfg.has_avail:
        avail = (rng.rand(n, cfg.act_dims[0]) > 0.3).astype(np.float32)
        avail[np.arange(n), acts[:, 0].astype(int)] = 1.0
    v_old = rng.randn(n, 1).astype(np.float32)
    ret = (rng.randn(n, 1) * 2 + 0.5).astype(np.float32)
    lp_old = (-1.5 + 0.3 * rng.randn(n, cfg.act_shape)).astype(np.float32)
    adv = rng.randn(n, 1).astype(np.float32)
    return (cent, obs, h, h, acts, v_old, ret, masks, active, lp_old, adv, avail)


EXT = 64          # csrc/big_epi.cuh kExt: extra columns of a stored activation row (mean, sigma, zeros)


def _tf32_values(shape, rng, scale=1.0):
    """Random fp32 values exactly representable in tf32 (so a tf32 GEMM with fp32 accumulation is exact up to summation order)."""
    return (np.round(rng.randn(*shape) * 64 * scale) / 64).astype(np.float32)


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
@pytest.mark.parametrize("rows,K,N", [(128, 32, 128), (300, 64, 256), (128, 672, 512), (5000, 512, 512), (1000, 800, 1024), (77, 32, 32),
                                      (40000, 512, 512)])          # 313 row blocks: CTA pairs walking several 256-row blocks
def test_big_lin_kernel_matches_matmul(rows, K, N, mode):
    """The K-major GEMM kernel (TMA SWIZZLE_128B boxes -> tcgen05.mma kind::tf32 -> TMEM -> epilogue -> TMA store) in isolation."""
    import ctypes as C
    from mappo_b200 import _lib
    from mappo_b200._lib import check, ptr
    lib = _lib.load()
    rng = np.random.RandomState(rows + K + N)
    dev = torch.device("cuda:0")
    A = torch.from_numpy(_tf32_values((rows, K), rng)).to(dev)
    W = torch.from_numpy(_tf32_values((N, K), rng, 0.25)).to(dev)
    bias = torch.from_numpy(rng.randn(N).astype(np.float32)).to(dev)
    colvec = torch.cat([torch.zeros(N, device=dev), bias]).contiguous()
    out = torch.full((rows, N + EXT), float("nan"), device=dev)
    stats = torch.zeros(rows, 2, device=dev)
    scratch = torch.zeros(rows, N, device=dev)
    check(lib.mappo_debug_big_lin(ptr(A), K, ptr(W), K, ptr(out), ptr(stats), ptr(colvec), ptr(scratch), rows, K, N,
                                  1 if mode == "tf32" else 0, None))
    torch.cuda.synchronize()
    want = torch.relu(A.double() @ W.double().T + bias.double())
    assert_close(out[:, :N].cpu().numpy(), want.cpu().numpy(), 1e-5, 1e-4, f"relu(A W^T + b) {rows}x{K}x{N}")
    mu = want.mean(1)
    sigma = torch.sqrt(want.var(1, unbiased=False) + 1e-5)
    assert_close(stats[:, 0].cpu().numpy(), mu.cpu().numpy(), 1e-4, 1e-4, "row mean")
    assert_close(stats[:, 1].cpu().numpy(), (1 / sigma).cpu().numpy(), 1e-3, 1e-5, "row 1/sigma")
    assert_close(out[:, N].cpu().numpy(), mu.cpu().numpy(), 1e-4, 1e-4, "mean column")
    assert_close(out[:, N + 1].cpu().numpy(), sigma.cpu().numpy(), 1e-3, 1e-5, "sigma column")
    assert torch.all(out[:, N + 2:] == 0)


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
@pytest.mark.parametrize("rows,Pw,M,Qw", [(128, 128, 128, 32), (300, 512, 512, 544), (5000, 512, 512, 672), (2000, 576, 576, 32),
                                          (700, 256, 256, 800), (40000, 512, 512, 544),
                                          # CTA-pair kernel (P in whole 256-column tiles, Q tiles multiples of 64): 256 + 320, 256 + 256 + 192, one 192
                                          (300, 512, 512, 576), (5000, 512, 512, 704), (700, 256, 256, 192), (40000, 512, 512, 576),
                                          (1000, 1024, 1024, 1088)])
def test_big_grad_kernel_matches_matmul(rows, Pw, M, Qw, mode):
    """The MN-major GEMM kernel (contraction over the rows of two row-major matrices: TMA 128B-swizzle / 32B-atom boxes ->
    tcgen05.mma with SWIZZLE_128B_BASE32B descriptors, row-split partials + slot sum) in isolation."""
    from mappo_b200 import _lib
    from mappo_b200._lib import check, ptr
    lib = _lib.load()
    rng = np.random.RandomState(rows + Pw + Qw)
    dev = torch.device("cuda:0")
    P = torch.from_numpy(_tf32_values((rows, Pw), rng, 0.25)).to(dev)
    Q = torch.from_numpy(_tf32_values((rows, Qw), rng, 0.25)).to(dev)
    splits = int(lib.mappo_debug_big_grad_splits(rows, M, Pw, Qw))
    assert splits >= 1
    partial = torch.full((splits, M, Qw), float("nan"), device=dev)
    gsum = torch.full((M, Qw), float("nan"), device=dev)
    check(lib.mappo_debug_big_grad(ptr(P), Pw, Pw, M, ptr(Q), Qw, Qw, rows, ptr(partial), ptr(gsum), 1 if mode == "tf32" else 0, None))
    torch.cuda.synchronize()
    want = (P[:, :M].double().T @ Q.double()).cpu().numpy()
    assert_close(gsum.cpu().numpy(), want, 1e-5, 2e-3 * np.sqrt(rows / 128), f"P^T Q {rows}: {M}x{Qw} ({splits} row splits)")


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
@pytest.mark.parametrize("name,n_rows", [("h64_tanh", 600), ("h128_tanh_multi", 333), ("h512_relu", 300), ("h256_relu_l0", 128),
                                         ("h512_relu", 5000)])
def test_ppo_update_gradients_match_oracle(name, n_rows, mode, monkeypatch):
    monkeypatch.setenv("MAPPO_B200_GEMM", mode)
    cfg = O.PathConfig(episode_length=4, n_rollout_threads=4, num_agents=2, use_max_grad_norm=False, entropy_coef=0.015,
                       lr=7e-4, critic_lr=1e-3, **CONFIGS[name])
    torch.manual_seed(11)
    args, policy, trainer, buf = TP.build(cfg)
    _perturb(policy, 5)
    trainer.value_normalizer.state.copy_(torch.tensor([0.3e-4, 1.7e-4, 1.2e-4]))
    sample = _sample(cfg, n_rows, 17)
    learner = O.Learner(cfg, {k: v.detach().cpu().clone() for k, v in policy.actor.state_dict().items()},
                        {k: v.detach().cpu().clone() for k, v in policy.critic.state_dict().items()})
    learner.vn.load([0.3e-4, 1.7e-4, 1.2e-4])
    ref = learner.ppo_update(sample, keep_grads=True)
    out = trainer.ppo_update(sample)
    torch.cuda.synchronize()
    vl, cgn, pl, ent, agn, ratio = [float(x.reshape(-1)[0]) if torch.is_tensor(x) else float(x) for x in out]
    ltol = 1e-4 if mode == "fp32" else 2e-3
    assert_close(vl, ref["value_loss"], ltol, 1e-6, "value_loss")
    assert_close(pl, ref["policy_loss"], ltol, 2e-5 if mode == "tf32" else 1e-6, "policy_loss")
    assert_close(ent, ref["dist_entropy"], ltol, 1e-6, "dist_entropy")
    assert_close(ratio, ref["ratio"], ltol, 1e-6, "ratio")
    assert_close([agn, cgn], [ref["actor_grad_norm"], ref["critic_grad_norm"]], 1e-3 if mode == "fp32" else 5e-3, 1e-7, "grad norms")
    bad, report = [], []
    smooth = not cfg.use_ReLU
    for net, key in ((policy.actor, "actor_grads"), (policy.critic, "critic_grads")):
        gv, rv = [], []
        for k, v in net.named_grads().items():
            want = ref[key][k].numpy().astype(np.float64)
            got = v.cpu().numpy().astype(np.float64)
            gv.append(got.reshape(-1)); rv.append(want.reshape(-1))
            ok, _ = grad_agreement(got, want, f"{key} {k}", smooth, mode == "tf32", report)
            if not ok:
                bad.append(k)
        gv, rv = np.concatenate(gv), np.concatenate(rv)
        l2 = float(np.linalg.norm(gv - rv) / np.linalg.norm(rv))
        cos = float(gv @ rv / (np.linalg.norm(gv) * np.linalg.norm(rv)))
        report.append(f"{key}: whole-gradient relative L2 error {l2:.3e}, cosine {cos:.8f}")
        # (ReLU, 300 rows: one on / off tie moves a row of dW by ~8 % of its scale; measured up to 3.7e-2 on the critic)
        lim = (3e-3 if smooth else 5e-2) if mode == "tf32" else (1e-4 if smooth else 1e-2)
        assert l2 <= lim and cos >= 0.999, "\n".join(report)
    print(f"\n[{mode}] {name} n={n_rows}:\n  " + "\n  ".join(report))
    assert not bad, "\n".join(report)


def test_big_net_evaluate_actions_matches_oracle(monkeypatch):
    monkeypatch.setenv("MAPPO_B200_GEMM", "fp32")
    cfg = O.PathConfig(episode_length=4, n_rollout_threads=4, num_agents=2, **CONFIGS["h512_relu"])
    torch.manual_seed(3)
    args, policy, trainer, buf = TP.build(cfg)
    _perturb(policy, 9)
    (cent, obs, h, _, acts, _, _, masks, active, _, _, avail) = _sample(cfg, 200, 4)
    values, logp, ent = policy.evaluate_actions(cent, obs, h, h, acts, masks, avail, active)
    t = torch.from_numpy
    pa = {k: v.detach().cpu() for k, v in policy.actor.state_dict().items()}
    pc = {k: v.detach().cpu() for k, v in policy.critic.state_dict().items()}
    lp_ref, ent_ref = O.actor_evaluate(cfg, pa, t(obs), t(h), t(acts), t(masks), t(avail), t(active))
    v_ref, _ = O.critic_forward(cfg, pc, t(cent), t(h), t(masks))
    assert_close(logp.cpu().numpy(), lp_ref.numpy(), 1e-4, 1e-5, "log-probs")
    # (the perturbed value head has weights ~0.9: a value is a sum of 512 O(1) terms, compared relative to the values' scale)
    assert_close(values.cpu().numpy(), v_ref.numpy(), 1e-4, 2e-5 * float(np.abs(v_ref.numpy()).max()), "values")
    assert_close(float(ent), float(ent_ref), 1e-4, 1e-6, "entropy")


def test_big_net_tf32_rollout_values_and_logp(monkeypatch):
    """tcgen05 rollout forward (policy_step through the GEMM pipeline): values / log-probs of the SAMPLED actions against the
    oracle's forward on the same inputs and noise (integer actions may legitimately differ on near-ties in tf32: compare the
    log-prob the oracle assigns to the engine's action)."""
    monkeypatch.setenv("MAPPO_B200_GEMM", "tf32")
    cfg = O.PathConfig(episode_length=4, n_rollout_threads=4, num_agents=2, **CONFIGS["h512_relu"])
    torch.manual_seed(3)
    args, policy, trainer, buf = TP.build(cfg)
    _perturb(policy, 9)
    (cent, obs, h, _, _, _, _, masks, _, _, _, avail) = _sample(cfg, 300, 4)
    noise = np.random.RandomState(2).exponential(size=(300, 20)).astype(np.float32)
    v, a, lp, _, _ = policy._step(cent, obs, h, h, masks, avail, False, True, True, exp_noise=noise)
    t = torch.from_numpy
    pa = {k: x.detach().cpu() for k, x in policy.actor.state_dict().items()}
    pc = {k: x.detach().cpu() for k, x in policy.critic.state_dict().items()}
    lp_ref, _ = O.actor_evaluate(cfg, pa, t(obs), t(h), a.cpu().float(), t(masks), t(avail), None)
    v_ref, _ = O.critic_forward(cfg, pc, t(cent), t(h), t(masks))
    assert_close(lp.cpu().numpy(), lp_ref.numpy(), 5e-3, 5e-3, "log-prob of the sampled action")
    # (the perturbed value head has weights ~0.9: values are sums of 512 O(1) terms, compared relative to their scale)
    assert_close(v.cpu().numpy(), v_ref.numpy(), 5e-3, 5e-3 * float(np.abs(v_ref.numpy()).max()), "values")
    a_ref = O.actor_act(cfg, pa, t(obs), t(h), t(masks), t(avail), exp_noise=t(noise))[0]
    agree = float((a_ref.numpy().reshape(-1) == a.cpu().numpy().reshape(-1)).mean())
    assert agree > 0.97, f"only {agree:.3f} of the sampled actions agree with the fp32 oracle"
