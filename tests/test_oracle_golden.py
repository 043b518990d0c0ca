"""Pin the oracle (oracle/mappo_oracle.py) against outputs of the reference itself (tests/golden)."""
import numpy as np
import pytest

from oracle import mappo_oracle as O
from helpers import Golden, GOLDEN_CASES, INFO_KEYS, assert_close


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_iterations(name):
    g = Golden(name)
    cfg = g.cfg
    learner = O.Learner(cfg, g.init_params("actor"), g.init_params("critic"))
    store = O.RolloutStore(cfg)
    T, N, M = cfg.episode_length, cfg.n_rollout_threads, cfg.num_agents
    for it in range(g.iters):
        feed = g.feed(it)
        noise, perms = g.get(f"it{it}/noise"), g.get(f"it{it}/perms")
        # collect + returns (run_iteration minus train, to compare pre-train buffers)
        if it == 0:
            store.obs[0], store.share_obs[0] = feed.obs[0], feed.share_obs[0]
            if feed.available_actions is not None:
                store.available_actions[0] = feed.available_actions[0]
        info = O.run_iteration(cfg, learner, store, feed, noise=noise, perms=list(perms))
        pre = f"it{it}/"
        # integer action indices: bit exact
        np.testing.assert_array_equal(store.actions, g.get(pre + "buf/actions"))
        assert_close(store.action_log_probs, g.get(pre + "buf/action_log_probs"), 1e-5, 1e-6, "logp")
        assert_close(store.value_preds, g.get(pre + "buf/value_preds"), 1e-5, 1e-6, "value_preds")
        assert_close(store.returns[:-1], g.get(pre + "buf/returns")[:-1], 1e-5, 1e-5, "returns")
        # rnn states / masks: compare pre-after_update slots 1..T (slot 0 was overwritten by after_update)
        if g.has(pre + "buf/rnn_states"):
            assert_close(store.rnn_states[1:], g.get(pre + "buf/rnn_states")[1:], 1e-5, 1e-6, "rnn_states")
        want = dict(zip(INFO_KEYS, g.get(pre + "train_info")))
        for k in INFO_KEYS:
            assert_close(info[k], want[k], 2e-4, 1e-6, f"train_info[{k}] it{it}")
        for k, v in learner.actor.items():
            g.cmp(pre + f"actor/{k}", v.detach().numpy(), 1e-4, 2e-6, f"actor {k}")
        for k, v in learner.critic.items():
            g.cmp(pre + f"critic/{k}", v.detach().numpy(), 1e-4, 2e-6, f"critic {k}")
        if learner.vn is not None:
            assert_close(learner.vn.state(), g.get(pre + "valuenorm"), 1e-5, 1e-9, "valuenorm")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_first_update_gradients(name):
    g = Golden(name)
    cfg = g.cfg
    feed = g.feed(0)
    store2 = O.RolloutStore(cfg)
    learner2 = O.Learner(cfg, g.init_params("actor"), g.init_params("critic"))
    store2.obs[0], store2.share_obs[0] = feed.obs[0], feed.share_obs[0]
    if feed.available_actions is not None:
        store2.available_actions[0] = feed.available_actions[0]
    _collect_only(cfg, learner2, store2, feed, g.get("it0/noise"))
    adv = O.normalized_advantages(store2, learner2.vn)
    assert_close(adv, g.get("it0/advantages"), 1e-4, 1e-5, "advantages")
    sample = next(O.minibatches(store2, adv, g.get("it0/perms")[0]))
    out = learner2.ppo_update(sample, keep_grads=True)
    for k, v in out["actor_grads"].items():
        g.cmp(f"it0/first_update/actor/{k}", v.numpy(), 1e-3, 1e-7, f"actor grad {k}")
    for k, v in out["critic_grads"].items():
        g.cmp(f"it0/first_update/critic/{k}", v.numpy(), 1e-3, 1e-6, f"critic grad {k}")
    assert_close([out["actor_grad_norm"], out["critic_grad_norm"]], g.get("it0/first_update/norms"), 1e-4, 1e-7)


def _collect_only(cfg, learner, store, feed, noise):
    T, N, M = cfg.episode_length, cfg.n_rollout_threads, cfg.num_agents
    E = N * M
    cat = lambda a: a.reshape(E, *a.shape[2:])
    for t in range(T):
        avail = cat(store.available_actions[t]) if feed.available_actions is not None else None
        vals, acts, lps, h_a, h_c = learner.get_actions(cat(store.share_obs[t]), cat(store.obs[t]),
                                                        cat(store.rnn_states[t]), cat(store.rnn_states_critic[t]),
                                                        cat(store.masks[t]), avail, exp_noise=noise[t])
        un = lambda x: x.numpy().reshape(N, M, *x.shape[1:])
        h_a, h_c = un(h_a).copy(), un(h_c).copy()
        d = feed.dones[t]
        h_a[d] = 0.0
        h_c[d] = 0.0
        masks = np.ones((N, M, 1), np.float32)
        masks[d] = 0.0
        store.insert(feed.share_obs[t + 1], feed.obs[t + 1], h_a, h_c, un(acts).astype(np.float32), un(lps), un(vals),
                     feed.rewards[t], masks,
                     active_masks=None if feed.active_masks is None else feed.active_masks[t],
                     available_actions=None if feed.available_actions is None else feed.available_actions[t + 1])
    nv = learner.get_values(cat(store.share_obs[-1]), cat(store.rnn_states_critic[-1]), cat(store.masks[-1]))
    O.compute_returns(store, nv.numpy().reshape(N, M, 1), learner.vn)


# --------------------------------------------------------------------------------------------
# separated policies: one learner + store per agent, trained in the reference's randperm order
# --------------------------------------------------------------------------------------------
SEPARATED_CASES = ["sep_mlp_2agents", "sep_happo_2agents"]


def load_separated(name="sep_mlp_2agents"):
    import ast
    import os
    import torch
    from helpers import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    M = int(z["n_agents"])
    cfgs = []
    for i in range(M):
        d = ast.literal_eval(str(z[f"agent{i}/cfg_json"]))
        d["act_dims"] = tuple(d["act_dims"])
        cfgs.append(O.PathConfig(**d))
    params = lambda pre: {k[len(pre):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(pre)}
    feed = lambda i: O.SyntheticFeed(*[z[f"agent{i}/feed/{n}"].copy() if f"agent{i}/feed/{n}" in z.files else None for n in
                                       ("obs", "share_obs", "rewards", "dones", "active_masks", "available_actions")])
    return z, M, cfgs, params, feed


@pytest.mark.parametrize("name", SEPARATED_CASES)
def test_oracle_separated_matches_reference(name):
    import torch
    z, M, cfgs, params, feed = load_separated(name)
    happo = str(z["algo"]) == "happo"
    learners = [O.Learner(cfgs[i], params(f"agent{i}/init/actor/"), params(f"agent{i}/init/critic/"), happo=happo) for i in range(M)]
    stores = [O.RolloutStore(c) for c in cfgs]
    feeds = [feed(i) for i in range(M)]
    for i in range(M):
        stores[i].obs[0], stores[i].share_obs[0] = feeds[i].obs[0], feeds[i].share_obs[0]
        if feeds[i].available_actions is not None:
            stores[i].available_actions[0] = feeds[i].available_actions[0]
        _collect_only(cfgs[i], learners[i], stores[i], feeds[i], z[f"agent{i}/noise"])
        sq = lambda a: a.reshape(a.shape[0], a.shape[1], *a.shape[3:])
        np.testing.assert_array_equal(sq(stores[i].actions), z[f"agent{i}/buf/actions"])
        assert_close(sq(stores[i].action_log_probs), z[f"agent{i}/buf/action_log_probs"], 1e-5, 1e-6, "logp")
        assert_close(sq(stores[i].value_preds), z[f"agent{i}/buf/value_preds"], 1e-5, 1e-6, "values")
        assert_close(sq(stores[i].returns)[:-1], z[f"agent{i}/buf/returns"][:-1], 1e-5, 1e-5, "returns")
    T, N = cfgs[0].episode_length, cfgs[0].n_rollout_threads
    factor = np.ones((T, N, 1), np.float32)
    t = torch.from_numpy
    for pos, i in enumerate(z["agent_order"]):
        i = int(i)
        assert_close(factor, z[f"agent{i}/factor_in"], 1e-4, 1e-6, f"factor handed to agent {i}")
        s, c = stores[i], cfgs[i]
        flat = lambda a: t(a[:T].reshape(T * N, -1))
        ev = lambda: O.actor_evaluate(c, learners[i].actor, flat(s.obs), flat(s.rnn_states).reshape(T * N, 1, -1), flat(s.actions),
                                      flat(s.masks), None if s.available_actions is None else flat(s.available_actions),
                                      flat(s.active_masks))[0].detach()
        old = ev()
        info = learners[i].train(s, list(z[f"agent{i}/perms"]), factor=factor.reshape(T, N, 1, 1))
        new = ev()
        factor = factor * torch.prod(torch.exp(new - old), dim=-1).reshape(T, N, 1).numpy()
        s.after_update()
        want = dict(zip(INFO_KEYS, z[f"train_info_pos{pos}"]))
        for k in INFO_KEYS:
            assert_close(info[k], want[k], 2e-4, 1e-6, f"train_info[{k}] of the agent trained at position {pos}")
    assert_close(factor, z["factor_final"], 1e-4, 1e-6, "final factor")
    for i in range(M):
        for k, v in learners[i].actor.items():
            assert_close(v.detach().numpy(), z[f"agent{i}/final/actor/{k}"], 1e-4, 2e-6, f"agent {i} actor {k}")
        for k, v in learners[i].critic.items():
            assert_close(v.detach().numpy(), z[f"agent{i}/final/critic/{k}"], 1e-4, 2e-6, f"agent {i} critic {k}")
        assert_close(learners[i].vn.state(), z[f"agent{i}/valuenorm"], 1e-5, 1e-9, "valuenorm")
    if happo:                          # the reference's HAPPO never updates its ValueNorm (happo_trainer.py:42-84)
        assert np.all(z["agent0/valuenorm"] == 0)
