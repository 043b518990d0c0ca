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
