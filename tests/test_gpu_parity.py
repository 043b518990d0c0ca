"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI / drop-in classes against
(1) the committed golden fixtures = outputs of the reference itself, and (2) the CPU oracle on seeded inputs.

Tolerances (fp32 everywhere, the kernels only re-associate sums):
  integer action indices                      bit exact
  row indices of the gathers                  bit exact
  GAE returns / advantages                    rtol 1e-5, atol 1e-5
  one forward (log-probs, values, rnn states) rtol 1e-4, atol 1e-5
  first-update gradients                      rtol 2e-3, atol 2e-6
  weights / train_info after a full train()   rtol 2e-3, atol 2e-5
"""
import numpy as np
import pytest
import torch

from oracle import mappo_oracle as O
from helpers import Golden, GOLDEN_CASES, INFO_KEYS, assert_close
from argsutil import make_args, make_spaces

pytestmark = pytest.mark.gpu

MLP_CASES = [c for c in GOLDEN_CASES if not Golden(c).cfg.recurrent]


def build(cfg, g=None):
    from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy
    from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO
    from onpolicy.utils.shared_buffer import SharedReplayBuffer
    args = make_args(cfg)
    obs_s, share_s, act_s = make_spaces(cfg)
    dev = torch.device("cuda:0")
    seeded = g is not None and g.has("init_seed")
    if seeded:                     # compact fixtures: the initialiser is seed-identical to the reference's (test_host_logic.py)
        torch.set_num_threads(1)
        torch.manual_seed(int(g.get("init_seed")))
        np.random.seed(int(g.get("init_seed")))
    policy = R_MAPPOPolicy(args, obs_s, share_s, act_s, device=dev)
    if seeded:
        g.check_init("actor", policy.actor.state_dict())
        g.check_init("critic", policy.critic.state_dict())
    elif g is not None:
        policy.actor.load_state_dict(g.params("init/actor/"))
        policy.critic.load_state_dict(g.params("init/critic/"))
    trainer = R_MAPPO(args, policy, device=dev)
    buf = SharedReplayBuffer(args, cfg.num_agents, obs_s, share_s, act_s)
    return args, policy, trainer, buf


def collect_and_returns(cfg, policy, trainer, buf, feed, noise):
    T, N, M, H = cfg.episode_length, cfg.n_rollout_threads, cfg.num_agents, cfg.hidden_size
    E = N * M
    for t in range(T):
        avail = buf.available_actions[t].reshape(E, -1) if feed.available_actions is not None else None
        v, a, lp, ha, hc = policy._step(buf.share_obs[t].reshape(E, -1), buf.obs[t].reshape(E, -1),
                                        buf.rnn_states[t].reshape(E, 1, H), buf.rnn_states_critic[t].reshape(E, 1, H),
                                        buf.masks[t].reshape(E, 1), avail, False, True, True,
                                        exp_noise=None if noise is None else noise[t])
        d = torch.from_numpy(feed.dones[t]).to(buf.device)
        ha = ha.reshape(N, M, 1, H).clone()
        hc = hc.reshape(N, M, 1, H).clone()
        ha[d] = 0.0
        hc[d] = 0.0
        masks = torch.ones(N, M, 1, device=buf.device)
        masks[d] = 0.0
        buf.insert(feed.share_obs[t + 1], feed.obs[t + 1], ha, hc, a.reshape(N, M, -1).float(), lp.reshape(N, M, -1),
                   v.reshape(N, M, 1), feed.rewards[t], masks,
                   active_masks=None if feed.active_masks is None else feed.active_masks[t],
                   available_actions=None if feed.available_actions is None else feed.available_actions[t + 1])
    nv = policy.get_values(buf.share_obs[-1].reshape(E, -1), buf.rnn_states_critic[-1].reshape(E, 1, H),
                           buf.masks[-1].reshape(E, 1))
    buf.compute_returns(nv.reshape(N, M, 1), trainer.value_normalizer)


class FakeRandperm:
    def __init__(self, perms):
        self.perms = list(perms)
        self.i = 0

    def __call__(self, n, *a, **k):
        p = torch.from_numpy(np.asarray(self.perms[self.i]).astype(np.int64))
        assert p.numel() == n
        self.i += 1
        return p


def warm(buf, feed):
    buf.obs[0].copy_(torch.from_numpy(feed.obs[0]))
    buf.share_obs[0].copy_(torch.from_numpy(feed.share_obs[0]))
    if feed.available_actions is not None:
        buf.available_actions[0].copy_(torch.from_numpy(feed.available_actions[0]))


# --------------------------------------------------------------------------------------------
# golden: the reference's own outputs
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_rollout_and_returns_match_reference(name):
    g = Golden(name)
    cfg = g.cfg
    args, policy, trainer, buf = build(cfg, g)
    feed = g.feed(0)
    warm(buf, feed)
    collect_and_returns(cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    torch.cuda.synchronize()
    pre = "it0/buf/"
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), g.get(pre + "actions"))        # integer: bit exact
    assert_close(buf.action_log_probs.cpu().numpy(), g.get(pre + "action_log_probs"), 1e-4, 1e-5, "logp")
    assert_close(buf.value_preds.cpu().numpy(), g.get(pre + "value_preds"), 1e-4, 1e-5, "value_preds")
    if g.has(pre + "rnn_states"):
        assert_close(buf.rnn_states.cpu().numpy()[1:], g.get(pre + "rnn_states")[1:], 1e-4, 1e-5, "rnn_states")
        assert_close(buf.rnn_states_critic.cpu().numpy()[1:], g.get(pre + "rnn_states_critic")[1:], 1e-4, 1e-5, "rnn_c")
    assert_close(buf.returns.cpu().numpy()[:-1], g.get(pre + "returns")[:-1], 1e-4, 1e-4, "returns")
    # normalised advantages as R_MAPPO.train forms them
    st = buf._adv_stats.cpu().numpy()
    mean = st[0] / st[2]
    std = np.sqrt(max(st[1] / st[2] - mean * mean, 0.0))
    adv = (buf.advantages.cpu().numpy() - mean) / (std + 1e-5)
    assert_close(adv, g.get("it0/advantages"), 1e-3, 2e-4, "advantages")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_first_update_gradients_match_reference(name, monkeypatch):
    g = Golden(name)
    cfg = g.cfg
    one = O.PathConfig(**{**cfg.to_dict(), "ppo_epoch": 1, "act_dims": tuple(cfg.act_dims)})
    args, policy, trainer, buf = build(one, g)
    trainer.num_mini_batch = cfg.num_mini_batch
    feed = g.feed(0)
    warm(buf, feed)
    collect_and_returns(cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    # run exactly the first minibatch update: a 1-epoch train() with the recorded permutation, lr = 0 so the
    # parameters stay put and the gradient buffers hold the first update's (clipped) gradients
    perms = g.get("it0/perms")
    monkeypatch.setattr(torch, "randperm", FakeRandperm([perms[0]]))
    if cfg.num_mini_batch == 1 and not cfg.recurrent:
        trainer.train(buf)
        grads_a = {k: v.cpu().numpy() for k, v in policy.actor.named_grads().items()}
        grads_c = {k: v.cpu().numpy() for k, v in policy.critic.named_grads().items()}
    else:
        # first minibatch only: drive ppo_update with the generator's first sample
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
        sample = next(gen)
        trainer.ppo_update(sample)
        grads_a = {k: v.cpu().numpy() for k, v in policy.actor.named_grads().items()}
        grads_c = {k: v.cpu().numpy() for k, v in policy.critic.named_grads().items()}
    norms = g.get("it0/first_update/norms")
    coef_a = min(1.0, cfg.max_grad_norm / (norms[0] + 1e-6)) if cfg.use_max_grad_norm else 1.0
    coef_c = min(1.0, cfg.max_grad_norm / (norms[1] + 1e-6)) if cfg.use_max_grad_norm else 1.0
    # the engine keeps UNCLIPPED gradients in its buffer (clipping is applied inside the Adam kernel)
    for k, v in grads_a.items():
        g.cmp(f"it0/first_update/actor/{k}", v * coef_a, 2e-3, 2e-6, f"actor grad {k}")
    for k, v in grads_c.items():
        g.cmp(f"it0/first_update/critic/{k}", v * coef_c, 2e-3, 2e-6, f"critic grad {k}")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_full_iterations_match_reference(name, monkeypatch):
    g = Golden(name)
    cfg = g.cfg
    args, policy, trainer, buf = build(cfg, g)
    for it in range(g.iters):
        feed = g.feed(it)
        if it == 0:
            warm(buf, feed)
        collect_and_returns(cfg, policy, trainer, buf, feed, g.get(f"it{it}/noise"))
        np.testing.assert_array_equal(buf.actions.cpu().numpy(), g.get(f"it{it}/buf/actions"))
        monkeypatch.setattr(torch, "randperm", FakeRandperm(g.get(f"it{it}/perms")))
        info = trainer.train(buf)
        buf.after_update()
        want = dict(zip(INFO_KEYS, g.get(f"it{it}/train_info")))
        for k in INFO_KEYS:
            assert_close(info[k], want[k], 2e-3, 2e-5, f"{name} it{it} train_info[{k}]")
        for k, v in policy.actor.state_dict().items():
            g.cmp(f"it{it}/actor/{k}", v.cpu().numpy(), 2e-3, 2e-5, f"actor {k}")
        for k, v in policy.critic.state_dict().items():
            g.cmp(f"it{it}/critic/{k}", v.cpu().numpy(), 2e-3, 2e-5, f"critic {k}")
        assert_close(trainer.value_normalizer.state.cpu().numpy(), g.get(f"it{it}/valuenorm"), 1e-4, 1e-8, "valuenorm")


# --------------------------------------------------------------------------------------------
# oracle: per-kernel checks on seeded inputs
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_gae,ptl,use_vn", [(1, 0, 1), (1, 1, 1), (0, 0, 1), (0, 1, 1), (1, 0, 0)])
@pytest.mark.parametrize("T,N,M", [(25, 8, 3), (7, 1, 1), (400, 5, 3)])
def test_compute_returns_matches_oracle(T, N, M, use_gae, ptl, use_vn):
    cfg = O.PathConfig(episode_length=T, n_rollout_threads=N, num_agents=M, use_gae=bool(use_gae),
                       use_proper_time_limits=bool(ptl), use_valuenorm=bool(use_vn))
    rng = np.random.RandomState(T + N + M)
    store = O.RolloutStore(cfg)
    store.rewards[:] = rng.randn(*store.rewards.shape)
    store.value_preds[:] = rng.randn(*store.value_preds.shape)
    store.masks[:] = (rng.rand(*store.masks.shape) > 0.1)
    store.bad_masks[:] = (rng.rand(*store.bad_masks.shape) > 0.1)
    store.active_masks[:] = (rng.rand(*store.active_masks.shape) > 0.2)
    vn = O.ValueNormState()
    vn.load([0.3e-4, 1.7e-4, 1.2e-4])
    nv = rng.randn(N, M, 1).astype(np.float32)
    args, policy, trainer, buf = build(cfg)
    for nm in ("rewards", "value_preds", "masks", "bad_masks", "active_masks"):
        getattr(buf, nm).copy_(torch.from_numpy(getattr(store, nm)))
    if use_vn:
        trainer.value_normalizer.state.copy_(torch.from_numpy(vn.state()))
    buf.compute_returns(nv, trainer.value_normalizer)
    O.compute_returns(store, nv, vn if use_vn else None)
    n = T + 1 if not use_gae else T
    assert_close(buf.returns.cpu().numpy()[:n], store.returns[:n], 1e-5, 1e-5, "returns")
    v = store.value_preds[:-1]
    adv = store.returns[:-1] - (vn.denormalize(v) if use_vn else v)
    assert_close(buf.advantages.cpu().numpy(), adv, 1e-5, 1e-5, "raw advantages")
    act = store.active_masks[:-1] != 0
    st = buf._adv_stats.cpu().numpy()
    assert st[2] == act.sum()
    assert_close(st[0], adv[act].astype(np.float64).sum(), 1e-6, 1e-6, "sum adv")
    assert_close(st[1], (adv[act].astype(np.float64) ** 2).sum(), 1e-6, 1e-6, "sum adv^2")


@pytest.mark.parametrize("T,N,M,L,mb", [(25, 4, 2, 10, 1), (20, 4, 3, 10, 2), (7, 3, 1, 3, 2)])
def test_recurrent_generator_rows_bit_exact(T, N, M, L, mb, monkeypatch):
    cfg = O.PathConfig(episode_length=T, n_rollout_threads=N, num_agents=M, use_recurrent_policy=True,
                       data_chunk_length=L, num_mini_batch=mb)
    args, policy, trainer, buf = build(cfg)
    store = O.RolloutStore(cfg)
    rng = np.random.RandomState(0)
    for nm in ("share_obs", "obs", "rnn_states", "rnn_states_critic", "actions", "value_preds", "returns", "masks",
               "active_masks", "action_log_probs", "rewards", "available_actions"):
        a = getattr(store, nm)
        a[:] = rng.randn(*a.shape)
        getattr(buf, nm).copy_(torch.from_numpy(a))
    adv = rng.randn(T, N, M, 1).astype(np.float32)
    perm = np.random.RandomState(1).permutation(T * N * M // L)
    monkeypatch.setattr(torch, "randperm", FakeRandperm([perm]))
    got = list(buf.recurrent_generator(adv, mb, L))
    want = list(O.minibatches(store, adv, perm))
    assert len(got) == len(want) == mb
    for gs, ws in zip(got, want):
        for a, b in zip(gs, ws):
            np.testing.assert_array_equal(a.cpu().numpy().reshape(b.shape), b)     # pure data movement: bit exact


@pytest.mark.parametrize("mb", [1, 3])
def test_feed_forward_generator_bit_exact(mb, monkeypatch):
    cfg = O.PathConfig(episode_length=9, n_rollout_threads=5, num_agents=2, num_mini_batch=mb)
    args, policy, trainer, buf = build(cfg)
    store = O.RolloutStore(cfg)
    rng = np.random.RandomState(0)
    for nm in ("share_obs", "obs", "actions", "value_preds", "returns", "masks", "active_masks", "action_log_probs",
               "available_actions"):
        a = getattr(store, nm)
        a[:] = rng.randn(*a.shape)
        getattr(buf, nm).copy_(torch.from_numpy(a))
    adv = rng.randn(9, 5, 2, 1).astype(np.float32)
    perm = np.random.RandomState(1).permutation(90)
    monkeypatch.setattr(torch, "randperm", FakeRandperm([perm]))
    got = list(buf.feed_forward_generator(adv, mb))
    want = list(O.minibatches(store, adv, perm))
    assert len(got) == len(want) == mb
    for gs, ws in zip(got, want):
        for a, b in zip(gs, ws):
            np.testing.assert_array_equal(a.cpu().numpy().reshape(b.shape), b)


def test_device_randperm_is_a_permutation():
    import ctypes as C
    from mappo_b200 import _lib
    lib = _lib.load()
    for n in (1, 2, 640, 9600, 76800):
        out = torch.empty(n, dtype=torch.int32, device="cuda")
        ctr = torch.tensor([7], dtype=torch.int64, device="cuda")
        _lib.check(lib.mappo_randperm(n, 1234, ctr.data_ptr(), out.data_ptr(), None))
        o = np.sort(out.cpu().numpy())
        np.testing.assert_array_equal(o, np.arange(n))
    a = torch.empty(9600, dtype=torch.int32, device="cuda")
    b = torch.empty(9600, dtype=torch.int32, device="cuda")
    c0 = torch.tensor([0], dtype=torch.int64, device="cuda")
    c1 = torch.tensor([1], dtype=torch.int64, device="cuda")
    _lib.check(lib.mappo_randperm(9600, 1, c0.data_ptr(), a.data_ptr(), None))
    _lib.check(lib.mappo_randperm(9600, 1, c1.data_ptr(), b.data_ptr(), None))
    assert (a != b).float().mean().item() > 0.99
    assert abs(np.corrcoef(a.cpu().numpy(), np.arange(9600))[0, 1]) < 0.05


@pytest.mark.parametrize("hidden,share", [(64, 54), (512, 785)])      # 8 K parameters (scalar kernel) / 0.93 M, P % 4 == 3 (128-bit kernel + tail)
def test_clip_adam_matches_torch_adam(hidden, share):
    from mappo_b200.core import FusedAdam
    cfg = O.PathConfig(hidden_size=hidden, share_obs_dim=share, layer_N=2 if hidden > 64 else 1)
    args, policy, trainer, buf = build(cfg)
    net = policy.critic
    ref = net.flat.detach().cpu().clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=7e-4, eps=1e-5)
    mine = FusedAdam(net, lr=7e-4, eps=1e-5)
    rng = torch.Generator().manual_seed(0)
    norm_out = torch.zeros(1, dtype=torch.float64, device="cuda")
    for step in range(5):
        gcpu = torch.randn(net.n_params, generator=rng) * (30.0 if step % 2 else 0.01)
        ref.grad = gcpu.clone()
        total = torch.nn.utils.clip_grad_norm_([ref], 10.0)
        opt.step()
        net.grad.copy_(gcpu)
        norm_out.zero_()
        mine.apply(10.0, True, norm_out.data_ptr())
        # torch's fp32 CPU norm of 0.93 M elements is itself 1e-5 low against float64; ours is pairwise per 256 + fp64 on top
        assert_close(norm_out.item(), float(total), 3e-5, 1e-7, "grad norm")
        assert_close(norm_out.item(), float(gcpu.double().norm()), 2e-6, 1e-7, "grad norm against float64")
        assert_close(net.flat.cpu().numpy(), ref.detach().numpy(), 1e-4, 5e-7, f"params after step {step}")


@pytest.mark.parametrize("name", MLP_CASES)
def test_evaluate_actions_matches_oracle(name):
    g = Golden(name)
    cfg = g.cfg
    args, policy, trainer, buf = build(cfg, g)
    rng = np.random.RandomState(3)
    n = 150
    obs = rng.randn(n, cfg.obs_dim).astype(np.float32)
    cent = rng.randn(n, cfg.share_obs_dim).astype(np.float32)
    h = np.zeros((n, 1, cfg.hidden_size), np.float32)
    masks = np.ones((n, 1), np.float32)
    active = (rng.rand(n, 1) > 0.3).astype(np.float32)
    acts = np.stack([rng.randint(0, a, size=n) for a in cfg.act_dims], 1).astype(np.float32)
    avail = None
    if cfg.has_avail:
        avail = (rng.rand(n, cfg.act_dims[0]) > 0.3).astype(np.float32)
        avail[np.arange(n), acts[:, 0].astype(int)] = 1.0
    values, logp, ent = policy.evaluate_actions(cent, obs, h, h, acts, masks, avail, active)
    t = torch.from_numpy
    pa, pc = g.init_params("actor"), g.init_params("critic")
    lp_ref, ent_ref = O.actor_evaluate(cfg, pa, t(obs), t(h), t(acts), t(masks), None if avail is None else t(avail),
                                       t(active))
    v_ref, _ = O.critic_forward(cfg, pc, t(cent), t(h), t(masks))
    assert_close(logp.cpu().numpy(), lp_ref.numpy(), 1e-4, 1e-5, "log-probs")
    assert_close(values.cpu().numpy(), v_ref.numpy(), 1e-4, 1e-5, "values")
    assert_close(float(ent), float(ent_ref), 1e-4, 1e-6, "entropy")


def test_graft_smoke_runs():
    import __graft_entry__ as ge
    ge.smoke()


def test_engine_graph_replay_matches_eager_train():
    """The captured CUDA graph of one iteration reproduces the eager drop-in classes (same noise / permutations)."""
    from mappo_b200.engine import RolloutEngine
    g = Golden("c1_mlp_discrete")
    cfg = g.cfg
    feed = g.feed(0)
    # eager, through the public classes
    args, policy, trainer, buf = build(cfg, g)
    warm(buf, feed)
    collect_and_returns(cfg, policy, trainer, buf, feed, g.get("it0/noise"))
    import torch as _t
    real = _t.randperm
    _t.randperm = FakeRandperm(g.get("it0/perms"))
    try:
        info_eager = trainer.train(buf)
    finally:
        _t.randperm = real
    # graph, through the engine (host RNG mode so the very same noise / permutations are staged)
    args2, policy2, trainer2, buf2 = build(cfg, g)
    eng = RolloutEngine(args2, policy2, trainer2, buf2, rng="host", seed=1)
    eng.stage_feed(feed)
    eng.draw_host_rng = lambda: None
    eng.host["noise"].copy_(_t.from_numpy(g.get("it0/noise")))
    eng.host["perm"].copy_(_t.from_numpy(g.get("it0/perms").astype(np.int32)))
    eng.upload()
    _t.cuda.synchronize()
    # capture WITHOUT warm-up iterations mutating the weights: snapshot, capture, restore, replay once
    snap = [policy2.actor.flat.clone(), policy2.critic.flat.clone(), trainer2.value_normalizer.state.clone()]
    eng.capture(warmup=1)
    policy2.actor.flat.copy_(snap[0]); policy2.critic.flat.copy_(snap[1]); trainer2.value_normalizer.state.copy_(snap[2])
    for opt in (policy2.actor_optimizer, policy2.critic_optimizer):
        opt.exp_avg.zero_(); opt.exp_avg_sq.zero_(); opt.step_dev.zero_()
    for a in (buf2.rnn_states, buf2.rnn_states_critic):
        a.zero_()
    buf2.masks.fill_(1.0); buf2.active_masks.fill_(1.0)
    warm(buf2, feed)
    info_graph = eng.step_e2e()
    for k in INFO_KEYS:
        assert_close(info_graph[k], info_eager[k], 1e-5, 1e-7, f"graph vs eager train_info[{k}]")
    assert_close(policy2.actor.flat.cpu().numpy(), policy.actor.flat.cpu().numpy(), 1e-5, 1e-7, "actor weights")
    assert_close(policy2.critic.flat.cpu().numpy(), policy.critic.flat.cpu().numpy(), 1e-5, 1e-7, "critic weights")


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("persistent", [True, False])
def test_engine_rollout_matches_reference(name, persistent, monkeypatch):
    """The engine's collect phase -- one persistent launch for all T steps (or T x (policy_step + env_insert)) writing
    straight into the storage slots -- reproduces the reference's rollout: actions bit exact, values / log-probs /
    recurrent states / masks / returns within the forward tolerance."""
    from mappo_b200.engine import RolloutEngine
    monkeypatch.setenv("MAPPO_B200_PERSISTENT_ROLLOUT", "1" if persistent else "0")
    g = Golden(name)
    cfg = g.cfg
    args, policy, trainer, buf = build(cfg, g)
    feed = g.feed(0)
    eng = RolloutEngine(args, policy, trainer, buf, rng="host", seed=1)
    eng.stage_feed(feed)
    eng.draw_host_rng = lambda: None
    eng.host["noise"].copy_(torch.from_numpy(g.get("it0/noise")))
    eng.upload()
    lib = eng.lib
    import ctypes as C
    from mappo_b200._lib import check, ptr
    from mappo_b200.core import stream_ptr
    check(lib.mappo_pack_rollout_weights(C.byref(policy.actor.desc), ptr(policy.actor.flat), ptr(eng.img_actor), stream_ptr()))
    check(lib.mappo_pack_rollout_weights(C.byref(policy.critic.desc), ptr(policy.critic.flat), ptr(eng.img_critic), stream_ptr()))
    if persistent and eng.big:
        pytest.skip("hidden >= 128 nets run the per-step GEMM pipeline (no persistent rollout kernel)")
    if persistent:
        eng._rollout_persistent()
        eng._returns()
    else:
        for t in range(cfg.episode_length):
            eng._collect_and_insert(t)
        eng._compute()
    torch.cuda.synchronize()
    pre = "it0/buf/"
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), g.get(pre + "actions"))
    assert_close(buf.action_log_probs.cpu().numpy(), g.get(pre + "action_log_probs"), 1e-4, 1e-5, "logp")
    assert_close(buf.value_preds.cpu().numpy(), g.get(pre + "value_preds"), 1e-4, 1e-5, "value_preds")
    if g.has(pre + "rnn_states"):
        assert_close(buf.rnn_states.cpu().numpy()[1:], g.get(pre + "rnn_states")[1:], 1e-4, 1e-5, "rnn_states")
        assert_close(buf.rnn_states_critic.cpu().numpy()[1:], g.get(pre + "rnn_states_critic")[1:], 1e-4, 1e-5, "rnn_c")
    np.testing.assert_array_equal(buf.masks.cpu().numpy(), g.get(pre + "masks"))
    np.testing.assert_array_equal(buf.active_masks.cpu().numpy(), g.get(pre + "active_masks"))
    assert_close(buf.returns.cpu().numpy()[:-1], g.get(pre + "returns")[:-1], 1e-4, 1e-4, "returns")
    np.testing.assert_array_equal(buf.obs.cpu().numpy(), feed.obs)
    np.testing.assert_array_equal(buf.share_obs.cpu().numpy(), feed.share_obs)
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), feed.rewards)
