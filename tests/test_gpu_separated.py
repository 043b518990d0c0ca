"""Separated-policy path on the GPU against the reference's own outputs (tests/golden/sep_mlp_2agents.npz,
written by make_golden_separated.py): per-agent SeparatedReplayBuffer + R_MAPPOPolicy + R_MAPPO, the separated runner's
train loop (`train_agents`: randperm agent order, factor bookkeeping, train_infos appended in training order), the
13-tuple of the generators, plus the Hanabi insert variants and R_MAPPO.cal_value_loss against NumPy restatements."""
import numpy as np
import pytest
import torch

from oracle import mappo_oracle as O
from helpers import INFO_KEYS, assert_close
from argsutil import make_args, make_spaces
from test_oracle_golden import load_separated, SEPARATED_CASES
import test_gpu_parity as TP

pytestmark = pytest.mark.gpu


def _build_agents(z, M, cfgs, params):
    from onpolicy.utils.separated_buffer import SeparatedReplayBuffer
    if str(z["algo"]) == "happo":                          # what runner/separated/base_runner.py:69-71 selects
        from onpolicy.algorithms.happo.happo_trainer import HAPPO as R_MAPPO
        from onpolicy.algorithms.happo.policy import HAPPO_Policy as R_MAPPOPolicy
    else:
        from onpolicy.algorithms.r_mappo.algorithm.rMAPPOPolicy import R_MAPPOPolicy
        from onpolicy.algorithms.r_mappo.r_mappo import R_MAPPO
    dev = torch.device("cuda:0")
    pol, tr, buf = [], [], []
    for i, c in enumerate(cfgs):
        args = make_args(c)
        obs_s, share_s, act_s = make_spaces(c)
        p = R_MAPPOPolicy(args, obs_s, share_s, act_s, device=dev)
        p.actor.load_state_dict(params(f"agent{i}/init/actor/"))
        p.critic.load_state_dict(params(f"agent{i}/init/critic/"))
        pol.append(p)
        tr.append(R_MAPPO(args, p, device=dev))
        buf.append(SeparatedReplayBuffer(args, obs_s, share_s, act_s))
    return pol, tr, buf


@pytest.mark.parametrize("name", SEPARATED_CASES)
def test_separated_iteration_matches_reference(name, monkeypatch):
    """MAPPO: the factor is carried and ignored.  HAPPO (row f3): the factor and the joint (product over heads) importance weight
    enter the actor loss, ValueNorm is never updated, advantages use the raw value predictions -- all as in the reference."""
    from onpolicy.runner.separated.base_runner import train_agents
    z, M, cfgs, params, feed = load_separated(name)
    pol, tr, buf = _build_agents(z, M, cfgs, params)
    T, N = cfgs[0].episode_length, cfgs[0].n_rollout_threads
    feeds = [feed(i) for i in range(M)]
    for i in range(M):
        b, f = buf[i], feeds[i]
        b.obs[0].copy_(torch.from_numpy(f.obs[0][:, 0]))
        b.share_obs[0].copy_(torch.from_numpy(f.share_obs[0][:, 0]))
        if f.available_actions is not None:
            b.available_actions[0].copy_(torch.from_numpy(f.available_actions[0][:, 0]))
    for t in range(T):
        for i in range(M):                                        # per step, every agent in order (separated mpe_runner.py:100-131)
            b, f = buf[i], feeds[i]
            v, a, lp, ha, hc = pol[i]._step(b.share_obs[t], b.obs[t], b.rnn_states[t], b.rnn_states_critic[t], b.masks[t],
                                            None if b.available_actions is None else b.available_actions[t], False, True, True,
                                            exp_noise=z[f"agent{i}/noise"][t])
            d = torch.from_numpy(f.dones[t][:, 0]).to(b.device)
            masks = torch.ones(N, 1, device=b.device)
            masks[d] = 0.0
            b.insert(f.share_obs[t + 1][:, 0], f.obs[t + 1][:, 0], ha, hc, a.float(), lp, v, f.rewards[t][:, 0], masks,
                     active_masks=f.active_masks[t][:, 0],
                     available_actions=None if f.available_actions is None else f.available_actions[t + 1][:, 0])
    for i in range(M):
        nv = pol[i].get_values(buf[i].share_obs[-1], buf[i].rnn_states_critic[-1], buf[i].masks[-1])
        buf[i].compute_returns(nv, tr[i].value_normalizer)
        np.testing.assert_array_equal(buf[i].actions.cpu().numpy(), z[f"agent{i}/buf/actions"])
        assert_close(buf[i].action_log_probs.cpu().numpy(), z[f"agent{i}/buf/action_log_probs"], 1e-4, 1e-5, "logp")
        assert_close(buf[i].value_preds.cpu().numpy(), z[f"agent{i}/buf/value_preds"], 1e-4, 1e-5, "values")
        assert_close(buf[i].returns.cpu().numpy()[:-1], z[f"agent{i}/buf/returns"][:-1], 1e-4, 1e-4, "returns")
    order = [int(x) for x in z["agent_order"]]
    draws = [np.asarray(order)]
    for i in order:
        draws += list(z[f"agent{i}/perms"])
    monkeypatch.setattr(torch, "randperm", TP.FakeRandperm(draws))
    infos = train_agents(tr, buf, T, N)
    for pos in range(M):
        want = dict(zip(INFO_KEYS, z[f"train_info_pos{pos}"]))
        for k in INFO_KEYS:
            assert_close(float(infos[pos][k]), want[k], 2e-3, 2e-5, f"train_info[{k}] at training position {pos}")
    last = order[-1]
    assert_close(buf[last].factor.cpu().numpy(), z[f"agent{last}/factor_in"], 2e-3, 1e-5, "factor handed to the last agent")
    for i in range(M):
        for k, v in pol[i].actor.state_dict().items():
            assert_close(v.cpu().numpy(), z[f"agent{i}/final/actor/{k}"], 2e-3, 2e-5, f"agent {i} actor {k}")
        for k, v in pol[i].critic.state_dict().items():
            assert_close(v.cpu().numpy(), z[f"agent{i}/final/critic/{k}"], 2e-3, 2e-5, f"agent {i} critic {k}")
        assert_close(tr[i].value_normalizer.state.cpu().numpy(), z[f"agent{i}/valuenorm"], 1e-4, 1e-8, "valuenorm")


def test_separated_generator_yields_the_factor_as_13th_element(monkeypatch):
    z, M, cfgs, params, feed = load_separated()
    pol, tr, buf = _build_agents(z, M, cfgs, params)
    b, c = buf[0], cfgs[0]
    T, N = c.episode_length, c.n_rollout_threads
    rng = np.random.RandomState(0)
    for nm in ("share_obs", "obs", "actions", "value_preds", "returns", "masks", "active_masks", "action_log_probs"):
        a = getattr(b, nm)
        a.copy_(torch.from_numpy(rng.randn(*a.shape).astype(np.float32)))
    adv = rng.randn(T, N, 1).astype(np.float32)
    assert len(next(b.feed_forward_generator(adv, 2))) == 12                       # no factor yet: the shared 12-tuple
    factor = rng.rand(T, N, 1).astype(np.float32)
    b.update_factor(factor)
    perm = np.random.RandomState(1).permutation(T * N)
    monkeypatch.setattr(torch, "randperm", TP.FakeRandperm([perm]))
    got = list(b.feed_forward_generator(adv, 2))
    mb = T * N // 2
    for k, sample in enumerate(got):
        assert len(sample) == 13
        rows = perm[k * mb:(k + 1) * mb]
        np.testing.assert_array_equal(sample[12].cpu().numpy(), factor.reshape(-1, 1)[rows])
        np.testing.assert_array_equal(sample[1].cpu().numpy(), b.obs[:-1].reshape(T * N, -1).cpu().numpy()[rows])
        np.testing.assert_array_equal(sample[10].cpu().numpy(), adv.reshape(-1, 1)[rows])


def test_cal_value_loss_matches_reference_formula():
    """R_MAPPO.cal_value_loss (reference r_mappo.py:52-89) incl. the ValueNorm update, for the four switch combinations."""
    rng = np.random.RandomState(4)
    n = 500
    values, v_old = rng.randn(n, 1).astype(np.float32), rng.randn(n, 1).astype(np.float32)
    ret = (rng.randn(n, 1) * 3 + 1).astype(np.float32)
    active = (rng.rand(n, 1) > 0.3).astype(np.float32)
    for clipped in (True, False):
        for huber in (True, False):
            for vmask in (True, False):
                cfg = O.PathConfig(use_clipped_value_loss=clipped, use_huber_loss=huber, use_value_active_masks=vmask)
                args, policy, trainer, buf = TP.build(cfg)
                got = float(trainer.cal_value_loss(values, v_old, ret, active))
                vn = O.ValueNormState()
                vn.update(ret)
                mean, var = vn.mean_var()
                target = (ret - mean) / np.sqrt(var)
                v_clip = v_old + np.clip(values - v_old, -cfg.clip_param, cfg.clip_param)
                e_c, e_o = target - v_clip, target - values
                hub = lambda e: np.where(np.abs(e) <= cfg.huber_delta, e * e / 2, cfg.huber_delta * (np.abs(e) - cfg.huber_delta / 2))
                l_c, l_o = (hub(e_c), hub(e_o)) if huber else (e_c ** 2 / 2, e_o ** 2 / 2)
                vl = np.maximum(l_o, l_c) if clipped else l_o
                want = (vl * active).sum() / active.sum() if vmask else vl.mean()
                assert_close(got, want, 1e-5, 1e-7, f"value loss clipped={clipped} huber={huber} mask={vmask}")
                assert_close(trainer.value_normalizer.state.cpu().numpy(), vn.state(), 1e-5, 1e-9, "ValueNorm side effect")


def test_chooseinsert_and_chooseafter_update_match_reference_semantics():
    """Turn-based (Hanabi) variants, reference shared_buffer.py:125-158, 172-177: obs / share_obs / masks-family land in slot
    `step` (not step + 1), rnn states in step + 1; chooseafter_update copies only the rnn states and masks to slot 0."""
    cfg = O.PathConfig(episode_length=5, n_rollout_threads=3, num_agents=2, obs_dim=7, share_obs_dim=9, act_dims=(4,))
    args, policy, trainer, buf = TP.build(cfg)
    rng = np.random.RandomState(2)
    N, M, H = 3, 2, cfg.hidden_size
    ref = {k: getattr(buf, k).cpu().numpy().copy() for k in ("share_obs", "obs", "rnn_states", "rnn_states_critic", "actions",
                                                              "action_log_probs", "value_preds", "rewards", "masks", "bad_masks",
                                                              "active_masks", "available_actions")}
    for s in range(cfg.episode_length):
        x = dict(share_obs=rng.randn(N, M, 9), obs=rng.randn(N, M, 7), rnn_states=rng.randn(N, M, 1, H),
                 rnn_states_critic=rng.randn(N, M, 1, H), actions=rng.randint(0, 4, (N, M, 1)), action_log_probs=rng.randn(N, M, 1),
                 value_preds=rng.randn(N, M, 1), rewards=rng.randn(N, M, 1), masks=(rng.rand(N, M, 1) > 0.2),
                 bad_masks=(rng.rand(N, M, 1) > 0.1), active_masks=(rng.rand(N, M, 1) > 0.3), available_actions=(rng.rand(N, M, 4) > 0.5))
        x = {k: v.astype(np.float32) for k, v in x.items()}
        buf.chooseinsert(x["share_obs"], x["obs"], x["rnn_states"], x["rnn_states_critic"], x["actions"], x["action_log_probs"],
                         x["value_preds"], x["rewards"], x["masks"], x["bad_masks"], x["active_masks"], x["available_actions"])
        ref["share_obs"][s], ref["obs"][s] = x["share_obs"], x["obs"]
        ref["rnn_states"][s + 1], ref["rnn_states_critic"][s + 1] = x["rnn_states"], x["rnn_states_critic"]
        for k in ("actions", "action_log_probs", "value_preds", "rewards"):
            ref[k][s] = x[k]
        ref["masks"][s + 1] = x["masks"]
        ref["bad_masks"][s + 1], ref["active_masks"][s] = x["bad_masks"], x["active_masks"]
        ref["available_actions"][s] = x["available_actions"]
    assert buf.step == 0
    for k, v in ref.items():
        np.testing.assert_array_equal(getattr(buf, k).cpu().numpy(), v, err_msg=k)
    buf.chooseafter_update()
    for k in ("rnn_states", "rnn_states_critic", "masks", "bad_masks"):
        ref[k][0] = ref[k][-1]
    for k, v in ref.items():
        np.testing.assert_array_equal(getattr(buf, k).cpu().numpy(), v, err_msg="after chooseafter_update: " + k)
