"""The tcgen05 (kind::tf32, fp32 accumulate in TMEM) build of the fused MLP update against the reference's outputs.

TF32 keeps 10 mantissa bits of the GEMM inputs (round-to-nearest when the operand tiles are written); accumulation,
LayerNorm, activations, losses, the gradient reduction and Adam stay fp32.  Stated tolerances:
  first-update gradients   tanh nets: |err| <= 5e-3 * |ref| + 5e-3 * max|ref of that tensor| (tf32 rounding of a 64..9600-term dot;
                           measured worst case 1.7e-3 of the tensor's scale); ReLU nets (c5): relative L2 error of every tensor
                           <= 5e-2 (measured 1.6e-2) -- helpers.grad_agreement says why the element-wise maximum is not usable
  losses / ratio / entropy  rtol 2e-3 (policy_loss: + 2e-5 absolute, it is a difference of O(1) terms near zero)
  weights after a full train()   rtol 2e-2, atol 2.1 * (optimiser steps) * lr: Adam normalises the gradient, so a weight whose
                                   gradient is within tf32 noise of zero moves by lr per step in EITHER direction in each
                                   implementation -- 2 lr apart per step at worst (measured 1.7 steps * lr on the 72-row c5
                                   fixture, where most first-layer gradients are noise; 0.16 on c2)
  LayerNorm affine gradients     5e-2 of the tensor's scale (see _grad_check)
Cases: c1 (N = 8), c2 at the benchmark size (N = 128) and the c5 widths (hidden 512, layer_N 2: the TMA-fed GEMM pipeline).
The exact-fp32 build (MAPPO_B200_GEMM=fp32, tests/test_gpu_parity.py) keeps the tight tolerances.
"""
import numpy as np
import pytest
import torch

from helpers import Golden, INFO_KEYS, assert_close, grad_agreement
import test_gpu_parity as TP

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _tf32(monkeypatch):
    monkeypatch.setenv("MAPPO_B200_GEMM", "tf32")


def _grad_check(got, want, what, smooth=True, l2_tol=None):
    report = []
    ok, l2 = grad_agreement(got, want, what, smooth, True, report, l2_tol)
    assert ok, report[0]
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-12)) if smooth else l2


def test_tf32_path_is_active():
    g = Golden("c1_mlp_discrete")
    args, policy, trainer, buf = TP.build(g.cfg, g)
    from mappo_b200 import _lib
    assert trainer.gemm_mode == _lib.GEMM_TF32
    ws_a, ws_c = trainer._workspaces(600)
    assert ws_a.gemm_mode == _lib.GEMM_TF32 and ws_c.gemm_mode == _lib.GEMM_TF32
    assert ws_a.workspace.numel() > 1000          # folded tf32 weight image for the TMA bulk copy


TF32_CASES = ["c1_mlp_discrete", "c2_mlp_n128", "c5_h512_hanabi", "c3_gru_multidiscrete", "c4_gru_smac", "naive_rnn_ptl"]


def _collect_fp32(monkeypatch, cfg, policy, trainer, buf, feed, noise):
    """Rollout with fp32 forward passes so that the stored trajectory IS the reference's (a hidden >= 128 net would otherwise
    run its rollout GEMMs in tf32 as well and may flip a near-tie; tests/test_gpu_bignet.py covers that rollout), then
    back to the tcgen05 build for the update under test."""
    monkeypatch.setenv("MAPPO_B200_GEMM", "fp32")
    TP.collect_and_returns(cfg, policy, trainer, buf, feed, noise)
    monkeypatch.setenv("MAPPO_B200_GEMM", "tf32")


def _golden_rows(g, key, value):
    """(value restricted to the rows a compact fixture stores, stored tensor)."""
    value = np.asarray(value)
    if g.has(key + "@rows"):
        return value[g.get(key + "@rows")], g.get(key)
    return value, g.get(key)


@pytest.mark.parametrize("name", TF32_CASES)
def test_tf32_first_update_gradients(name, monkeypatch):
    g = Golden(name)
    cfg = g.cfg
    from oracle import mappo_oracle as O
    one = O.PathConfig(**{**cfg.to_dict(), "ppo_epoch": 1, "act_dims": tuple(cfg.act_dims)})
    args, policy, trainer, buf = TP.build(one, g)
    feed = g.feed(0)
    TP.warm(buf, feed)
    _collect_fp32(monkeypatch, cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    monkeypatch.setattr(torch, "randperm", TP.FakeRandperm([g.get("it0/perms")[0]]))
    if cfg.num_mini_batch == 1:
        info = trainer.train(buf)
    else:
        # several minibatches per epoch: drive ppo_update with the generator's first sample only (like tests/test_gpu_parity.py)
        info = None
        trainer.num_mini_batch = cfg.num_mini_batch
        st = buf._adv_stats.cpu().numpy()
        mean = st[0] / st[2]
        std = np.sqrt(max(st[1] / st[2] - mean * mean, 0.0))
        adv = (buf.advantages - float(mean)) / (float(std) + 1e-5)
        if cfg.use_recurrent_policy:
            gen = buf.recurrent_generator(adv, cfg.num_mini_batch, cfg.data_chunk_length)
        elif cfg.use_naive_recurrent_policy:
            gen = buf.naive_recurrent_generator(adv, cfg.num_mini_batch)
        else:
            gen = buf.feed_forward_generator(adv, cfg.num_mini_batch)
        trainer.ppo_update(next(gen))
    norms = g.get("it0/first_update/norms")
    worst = 0.0
    # ReLU on / off ties weigh 1 / sqrt(rows): the 120-row minibatch of the c4 fixture measures 7.3e-2 on the base MLP tensors of the
    # critic while every GRU / head tensor of the same update agrees to 2e-3 (scripts/diag_gru_tc.py, profiles/r2_summary.md)
    rows = cfg.episode_length * cfg.n_rollout_threads * cfg.num_agents // cfg.num_mini_batch
    l2_tol = 1e-1 if (cfg.use_ReLU and rows < 256) else None
    for net, nm, nrm in ((policy.actor, "actor", norms[0]), (policy.critic, "critic", norms[1])):
        coef = min(1.0, cfg.max_grad_norm / (nrm + 1e-6))
        for k, v in net.named_grads().items():
            got, want = _golden_rows(g, f"it0/first_update/{nm}/{k}", v.cpu().numpy() * coef)
            worst = max(worst, _grad_check(got, want, f"{nm} {k}", smooth=not cfg.use_ReLU, l2_tol=l2_tol))
    if info is None:
        print(f"\n[tf32] {name}: worst gradient error (first minibatch): {worst:.3e}")
        return
    losses = g.get("it0/first_update/losses")          # value_loss, policy_loss, dist_entropy, ratio
    assert_close(info["value_loss"], losses[0], 2e-3, 1e-6, "value_loss")
    assert_close(info["policy_loss"], losses[1], 2e-3, 2e-5, "policy_loss")
    assert_close(info["dist_entropy"], losses[2], 2e-3, 1e-6, "dist_entropy")
    assert_close(info["ratio"], losses[3], 2e-3, 1e-6, "ratio")
    assert_close([info["actor_grad_norm"], info["critic_grad_norm"]], norms, 5e-3, 1e-6, "grad norms")
    print(f"\n[tf32] {name}: worst gradient error relative to tensor scale: {worst:.3e}")


@pytest.mark.parametrize("name", TF32_CASES)
def test_tf32_full_iterations(name, monkeypatch):
    g = Golden(name)
    cfg = g.cfg
    args, policy, trainer, buf = TP.build(cfg, g)
    feed = g.feed(0)
    TP.warm(buf, feed)
    _collect_fp32(monkeypatch, cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), g.get("it0/buf/actions"))
    # (only the first iteration: iteration 2 samples actions from tf32-trained weights)
    monkeypatch.setattr(torch, "randperm", TP.FakeRandperm(g.get("it0/perms")))
    info = trainer.train(buf)
    buf.after_update()
    want = dict(zip(INFO_KEYS, g.get("it0/train_info")))
    for k in INFO_KEYS:
        assert_close(info[k], want[k], 2e-2, 2e-4, f"train_info[{k}]")
    worst = 0.0
    atol_w = max(2e-3, 2.1 * cfg.ppo_epoch * cfg.num_mini_batch * max(cfg.lr, cfg.critic_lr))
    for net, nm in ((policy.actor, "actor"), (policy.critic, "critic")):
        for k, v in net.state_dict().items():
            got, ref = _golden_rows(g, f"it0/{nm}/{k}", v.cpu().numpy())
            assert_close(got, ref, 2e-2, atol_w, f"{nm} {k} after it0")
            worst = max(worst, float(np.abs(got - ref).max()))
    print(f"\n[tf32] {name}: worst absolute weight deviation from the reference after one train() {worst:.3e}")


@pytest.mark.parametrize("relu", [False, True])
def test_tf32_ctas_with_several_tiles_match_the_fp32_build(relu, monkeypatch):
    """More 128-row tiles than SMs (256 threads x 3 agents x 25 steps = 19200 rows = 150 tiles on 148 CTAs): a CTA then walks
    several tiles, its weight-gradient accumulators stay in TMEM across them.  First-update gradients and losses of the
    tcgen05 build against the exact-fp32 build on identical rollouts."""
    from oracle import mappo_oracle as O
    cfg = O.PathConfig(episode_length=25, n_rollout_threads=256, num_agents=3, obs_dim=18, share_obs_dim=54,
                       act_dims=(5,), use_ReLU=relu, ppo_epoch=1, num_mini_batch=1, lr=7e-4, critic_lr=7e-4)
    feed = O.make_feed(cfg, seed=3)
    noise = np.random.RandomState(5).exponential(size=(cfg.episode_length, cfg.n_rollout_threads * cfg.num_agents, 5)) \
        .astype(np.float32)
    perm = np.random.RandomState(6).permutation(cfg.episode_length * cfg.n_rollout_threads * cfg.num_agents)
    res = {}
    for mode in ("fp32", "tf32"):
        monkeypatch.setenv("MAPPO_B200_GEMM", mode)
        torch.manual_seed(1)
        args, policy, trainer, buf = TP.build(cfg)
        TP.warm(buf, feed)
        TP.collect_and_returns(cfg, policy, trainer, buf, feed, noise)
        monkeypatch.setattr(torch, "randperm", TP.FakeRandperm([perm]))
        info = trainer.train(buf)
        res[mode] = (info, {("a", k): v.cpu().numpy().copy() for k, v in policy.actor.named_grads().items()} |
                     {("c", k): v.cpu().numpy().copy() for k, v in policy.critic.named_grads().items()})
    for key, want in res["fp32"][1].items():
        _grad_check(res["tf32"][1][key], want, f"{key}", smooth=not relu)
    for k in ("value_loss", "dist_entropy", "ratio", "actor_grad_norm", "critic_grad_norm"):
        assert_close(res["tf32"][0][k], res["fp32"][0][k], 1e-2, 1e-6, k)


def test_tf32_path_is_active_for_gru_nets():
    g = Golden("c3_gru_multidiscrete")
    args, policy, trainer, buf = TP.build(g.cfg, g)
    from mappo_b200 import _lib
    ws_a, ws_c = trainer._workspaces(200)
    assert ws_a.gemm_mode == _lib.GEMM_TF32 and ws_c.gemm_mode == _lib.GEMM_TF32
    assert ws_a.n_slots == 1 and ws_c.n_slots == 1          # the tcgen05 GRU pipeline leaves the flat gradient in slot 0


@pytest.mark.parametrize("ctas,relu", [("0", False), ("5", False), ("5", True)])
def test_tf32_gru_many_tiles_match_the_fp32_build(ctas, relu, monkeypatch):
    """GRU policy, 256 threads x 3 agents x 30 steps in chunks of 10: 2304 chunks = 18 sequence tiles, 23040 positions = 180
    position tiles (> 148 SMs, so CTAs walk several tiles and keep their weight-gradient accumulators in TMEM across them; with
    MAPPO_B200_GRU_CTAS=5 the sequence kernels walk several tiles per CTA as well).  First-update gradients and losses of the
    tcgen05 build (update_gru_tc.cu) against the exact-fp32 build (update_gru.cu) on identical rollouts."""
    from oracle import mappo_oracle as O
    if ctas != "0":
        monkeypatch.setenv("MAPPO_B200_GRU_CTAS", ctas)
    cfg = O.PathConfig(episode_length=30, n_rollout_threads=256, num_agents=3, obs_dim=30, share_obs_dim=48,
                       act_dims=(9,), use_ReLU=relu, use_recurrent_policy=True, data_chunk_length=10, ppo_epoch=1,
                       num_mini_batch=1, lr=5e-4, critic_lr=5e-4)
    feed = O.make_feed(cfg, seed=3, kind="smac")
    noise = np.random.RandomState(5).exponential(size=(cfg.episode_length, cfg.n_rollout_threads * cfg.num_agents, 9)) \
        .astype(np.float32)
    n_chunks = (cfg.episode_length // cfg.data_chunk_length) * cfg.n_rollout_threads * cfg.num_agents
    perm = np.random.RandomState(6).permutation(n_chunks)
    res = {}
    for mode in ("fp32", "tf32"):
        monkeypatch.setenv("MAPPO_B200_GEMM", mode)
        torch.manual_seed(1)
        args, policy, trainer, buf = TP.build(cfg)
        TP.warm(buf, feed)
        TP.collect_and_returns(cfg, policy, trainer, buf, feed, noise)
        monkeypatch.setattr(torch, "randperm", TP.FakeRandperm([perm]))
        info = trainer.train(buf)
        res[mode] = (info, {("a", k): v.cpu().numpy().copy() for k, v in policy.actor.named_grads().items()} |
                     {("c", k): v.cpu().numpy().copy() for k, v in policy.critic.named_grads().items()})
    worst = 0.0
    for key, want in res["fp32"][1].items():
        worst = max(worst, _grad_check(res["tf32"][1][key], want, f"{key}", smooth=not relu))
    for k in ("value_loss", "dist_entropy", "ratio", "actor_grad_norm", "critic_grad_norm"):
        assert_close(res["tf32"][0][k], res["fp32"][0][k], 1e-2, 1e-6, k)
    print(f"\n[tf32 gru] worst gradient error relative to tensor scale: {worst:.3e}")


@pytest.mark.parametrize("name", ["c1_mlp_discrete", "c2_mlp_n128"])
def test_fused_optimiser_tail_matches_the_separate_launches(name, monkeypatch):
    """mappo_update_tail (one 8-CTA cluster launch: slot sum -> unfold -> clip + Adam -> next weight image) against the four
    launches it replaces, over a full train() (several epochs, so the re-used weight image is exercised): every weight, both
    Adam moments, the step counts and the train_info sums agree to the run-to-run noise of the unfused path itself."""
    g = Golden(name)
    cfg = g.cfg
    res = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("MAPPO_B200_FUSED_TAIL", fused)
        args, policy, trainer, buf = TP.build(cfg, g)
        feed = g.feed(0)
        TP.warm(buf, feed)
        _collect_fp32(monkeypatch, cfg, policy, trainer, buf, feed, g.get("it0/noise"))
        monkeypatch.setattr(torch, "randperm", TP.FakeRandperm(g.get("it0/perms")))
        info = trainer.train(buf)
        ws_a, ws_c = next(iter(trainer._ws.values()))
        assert ws_a.fused_tail == (fused == "1") and ws_c.fused_tail == (fused == "1")
        state = {}
        for nm, net, opt in (("actor", policy.actor, policy.actor_optimizer), ("critic", policy.critic, policy.critic_optimizer)):
            state[nm + "/flat"] = net.flat.cpu().numpy().copy()
            state[nm + "/grad"] = net.grad.cpu().numpy().copy()
            state[nm + "/m"] = opt.exp_avg.cpu().numpy().copy()
            state[nm + "/v"] = opt.exp_avg_sq.cpu().numpy().copy()
            state[nm + "/step"] = opt.step_dev.cpu().numpy().copy()
            state[nm + "/beta_pow"] = opt.beta_pow.cpu().numpy().copy()
        res[fused] = (info, state)
    # (the advantage / return statistics are summed with fp64 atomics, so two runs of EITHER path differ in the last bits of a
    #  few dozen weights after ten epochs -- scripts/diag_tail.py; at one or two epochs the two paths are bit-identical.  A weight
    #  whose gradient is far below Adam's eps moves by lr * m / eps = 70 m per step, so last-bit noise of ~5e-9 in its gradient
    #  sums becomes ~4e-7 in the weight: hence the absolute tolerance)
    for k, v in res["0"][1].items():
        np.testing.assert_allclose(res["1"][1][k], v, rtol=1e-4, atol=5e-6, err_msg=k)
    for k, v in res["0"][0].items():
        assert abs(res["1"][0][k] - v) <= 1e-9 * max(1.0, abs(v)), k
    assert res["0"][1]["actor/step"][0] == cfg.ppo_epoch * cfg.num_mini_batch
