"""Algebra of the layer-by-layer GEMM pipeline for hidden >= 128 nets (csrc/big_*.cu), checked on the CPU in float64.

The CUDA path never materialises LayerNorm outputs between the GEMMs.  It stores the raw activation a_l = act(z_l) plus
two row scalars (mu_l, rs_l) and folds the normalisation into the NEXT GEMM's epilogue; in the backward pass the two row
reductions LayerNorm-backward needs are obtained analytically from quantities the producing epilogue already holds.
This test restates exactly that algebra with NumPy and compares every parameter gradient with torch autograd of the
reference network (mlp.py:6-57 + a linear head), so that a GPU failure can be attributed to the kernels, not the maths.

    forward   z_l = rs_{l-1} (a_{l-1} W'_l^T - mu_{l-1} s_l) + b'_l          W'_l = W_l diag(g_{l-1}), b'_l = b_l + W_l be_{l-1}
    backward  P_{l+1} = dZ_{l+1} rs_l   (row scaled, what is stored)         s_l = rowsum(W'_l)
              acc = P_{l+1} W'_{l+1}                     (= rs_l dxhat_l)
              dA_l = acc - (m1 + xhat_l m2),   m1 = sum_o P_{l+1}[o] s_{l+1}[o] / H,  m2 = sum_o P_{l+1}[o] (z_{l+1}[o] - b'_{l+1}[o]) / H
              G_l = P_l^T [a_{l-1} | mu_{l-1} | 1/rs_{l-1}]  ->  dW'_l = G[:, :H] - G[:, H],  db'_l = G[:, H+1]
              dW_l = dW'_l diag(g_{l-1}) + db'_l be_{l-1}^T,  dg_{l-1} = colsum(dW'_l * W_l),  dbe_{l-1} = W_l^T db'_l
"""
import numpy as np
import pytest
import torch


def act_f(z, relu):
    return np.maximum(z, 0.0) if relu else np.tanh(z)


def act_d_from_out(a, relu):
    return (a > 0).astype(a.dtype) if relu else 1.0 - a * a


def act_inv(a, relu):
    return a if relu else np.arctanh(a)


def ln_stats(a, eps=1e-5):
    mu = a.mean(1, keepdims=True)
    var = ((a - mu) ** 2).mean(1, keepdims=True)
    return mu, 1.0 / np.sqrt(var + eps)


def deferred_forward_backward(p, x, dlogits, relu, n_layers):
    """p: dict of float64 arrays named like the reference state_dict.  Returns logits and a dict of gradients."""
    eps = 1e-5
    H = p["fc0.w"].shape[0]
    # explicit feature normalisation once (pre-affine), constant-1 column for the bias
    mu0, rs0 = ln_stats(x, eps)
    xh0 = (x - mu0) * rs0
    # folded weights
    W = [p[f"fc{l}.w"] for l in range(n_layers)] + [p["head.w"]]
    b = [p[f"fc{l}.b"] for l in range(n_layers)] + [p["head.b"]]
    gam = [p["fn.w"]] + [p[f"ln{l}.w"] for l in range(n_layers)]
    bet = [p["fn.b"]] + [p[f"ln{l}.b"] for l in range(n_layers)]
    Wf = [W[l] * gam[l][None, :] for l in range(n_layers + 1)]
    bf = [b[l] + W[l] @ bet[l] for l in range(n_layers + 1)]
    s = [w.sum(1) for w in Wf]
    # forward
    a, mu, rs, z_store = [None] * (n_layers + 1), [None] * (n_layers + 1), [None] * (n_layers + 1), [None] * (n_layers + 2)
    acc = xh0 @ Wf[0].T
    z = acc + bf[0]
    for l in range(1, n_layers + 1):
        a[l] = act_f(z, relu)
        mu[l], rs[l] = ln_stats(a[l], eps)
        acc = a[l] @ Wf[l].T
        z = rs[l] * (acc - mu[l] * s[l][None, :]) + bf[l]
    logits = z
    L = n_layers
    # backward: head
    P = dlogits * rs[L]
    m1 = (P * s[L][None, :]).sum(1, keepdims=True) / H
    m2 = (P * (logits - bf[L][None, :])).sum(1, keepdims=True) / H
    grads = {}

    def unfold(l, G, in_is_x0):
        if in_is_x0:
            dWf, dbf = G[:, :-1], G[:, -1]
        else:
            dWf, dbf = G[:, :H] - G[:, H:H + 1], G[:, H + 1]
        name = f"fc{l}" if l < n_layers else "head"
        grads[name + ".w"] = dWf * gam[l][None, :] + np.outer(dbf, bet[l])      # b' = b + W beta depends on W too
        grads[name + ".b"] = dbf
        gname = "fn" if l == 0 else f"ln{l - 1}"
        grads[gname + ".w"] = (dWf * W[l]).sum(0)
        grads[gname + ".b"] = W[l].T @ dbf

    ext = lambda l: np.concatenate([a[l], mu[l], 1.0 / rs[l]], axis=1)
    unfold(L, P.T @ ext(L), False)
    for l in range(L, 0, -1):            # layer l produced a[l]; its weights are W[l-1]
        acc = P @ Wf[l]                   # = rs_l * dxhat_l
        xh = (a[l] - mu[l]) * rs[l]
        dA = acc - (m1 + xh * m2)
        dZ = dA * act_d_from_out(a[l], relu)
        rs_prev = rs[l - 1] if l > 1 else 1.0
        P = dZ * rs_prev
        zl = act_inv(a[l], relu)
        m1 = (P * s[l - 1][None, :]).sum(1, keepdims=True) / H
        m2 = (P * (zl - bf[l - 1][None, :])).sum(1, keepdims=True) / H
        if l > 1:
            unfold(l - 1, P.T @ ext(l - 1), False)
        else:
            unfold(0, P.T @ np.concatenate([xh0, np.ones((x.shape[0], 1))], axis=1), True)
    return logits, grads


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("n_layers", [1, 3])
def test_deferred_layernorm_pipeline_matches_autograd(relu, n_layers):
    rng = np.random.RandomState(3 + n_layers)
    B, I, H, A = 37, 23, 32, 7
    p = {"fn.w": rng.uniform(0.5, 1.5, I), "fn.b": rng.normal(0, 0.3, I), "head.w": rng.normal(0, 0.3, (A, H)),
         "head.b": rng.normal(0, 0.1, A)}
    for l in range(n_layers):
        p[f"fc{l}.w"] = rng.normal(0, 0.4, (H, I if l == 0 else H))
        p[f"fc{l}.b"] = rng.normal(0, 0.2, H)
        p[f"ln{l}.w"] = rng.uniform(0.5, 1.5, H)
        p[f"ln{l}.b"] = rng.normal(0, 0.3, H)
    x = rng.normal(0.3, 1.2, (B, I))
    R = rng.normal(0, 1, (B, A))

    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    h = torch.nn.functional.layer_norm(torch.tensor(x), (I,), tp["fn.w"], tp["fn.b"], 1e-5)
    for l in range(n_layers):
        h = h @ tp[f"fc{l}.w"].T + tp[f"fc{l}.b"]
        h = torch.relu(h) if relu else torch.tanh(h)
        h = torch.nn.functional.layer_norm(h, (H,), tp[f"ln{l}.w"], tp[f"ln{l}.b"], 1e-5)
    logits_t = h @ tp["head.w"].T + tp["head.b"]
    (logits_t * torch.tensor(R)).sum().backward()

    logits, g = deferred_forward_backward(p, x, R, relu, n_layers)
    np.testing.assert_allclose(logits, logits_t.detach().numpy(), rtol=1e-9, atol=1e-9)
    for k in p:
        np.testing.assert_allclose(g[k], tp[k].grad.numpy(), rtol=1e-7, atol=1e-8, err_msg=k)
