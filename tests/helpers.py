"""Shared test helpers: golden fixture loading (tests/golden/*.npz, written by make_golden.py)."""
import ast
import os

import numpy as np
import torch

from oracle import mappo_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["c1_mlp_discrete", "c3_gru_multidiscrete", "c4_gru_smac", "c5_mlp_switches", "naive_rnn_ptl"]
INFO_KEYS = ("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio")


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        d = ast.literal_eval(str(self.z["cfg_json"]))
        d["act_dims"] = tuple(d["act_dims"])
        self.cfg = O.PathConfig(**d)
        self.iters = 1 + max(int(k[2:k.index("/")]) for k in self.z.files if k.startswith("it"))

    def params(self, prefix):
        """{state_dict key: tensor} under e.g. 'init/actor/' or 'it0/critic/'."""
        return {k[len(prefix):]: torch.from_numpy(self.z[k].copy()) for k in self.z.files if k.startswith(prefix)}

    def feed(self, it):
        g = lambda n: self.z[f"it{it}/feed/{n}"].copy() if f"it{it}/feed/{n}" in self.z.files else None
        return O.SyntheticFeed(g("obs"), g("share_obs"), g("rewards"), g("dones"), g("active_masks"),
                               g("available_actions"))

    def get(self, key):
        return self.z[key]

    def has(self, key):
        return key in self.z.files


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} "
                             f"(|err|={err[i]:.3e}, tol={tol[i]:.3e}); max|err|={err.max():.3e}")
