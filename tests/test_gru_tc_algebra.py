"""Algebra of the tcgen05 GRU pipeline (csrc/update_gru_tc.cu + the TC_BASE_FWD / TC_HEAD / TC_BASE_BWD modes of update_mlp_tc.cu),
restated with NumPy in float64 and compared with torch autograd of the reference network (mlp.py:6-57, rnn.py:24-79, a linear head).

The CUDA path never runs autograd: it stages the step as six kernels around `[position][64]` planes and folds every LayerNorm affine
into the next GEMM.  This test follows exactly that staging so that a GPU failure can be attributed to the kernels, not the maths:

    1. base forward     X = xhat2 (pre-affine output of the last base LayerNorm)
    2. sequence forward hm_l = h_{l-1} m_l;  pre_rz = xaug W_ih'[rz]^T + hm W_hh[rz]^T  (b' carries b_ih + W_ih beta + b_hh[rz]),
                        gi_n = xaug W_ih'[n]^T,  gh_n = hm W_hh[n]^T + b_hn,  r, z = sigma(.),  n = tanh(gi_n + r gh_n),
                        h_l = (1 - z) n + z hm                                                     -> planes R, Z, N, GHN, H
    3. heads            xhat_h = LN(h) pre-affine, logits = xhat_h_aug Wh'^T, dxhat = dL Wh', LN backward       -> DHH, Gh, dbh
    4. BPTT             d = dh + DHH;  dn = d (1 - z)(1 - n^2),  dz = d (hm - n) z (1 - z),  dr = dn ghn r (1 - r)
                        dh_{l-1} = (d z + [dr | dz | dn r] W_hh) m_l                                              -> planes DR, DZ, DN
    5. gate gradients   DFEAT = [dr | dz | dn] W_ih';  Gih' += dgi^T [x | 1];  Ghh += [dr | dz | dn r]^T [hm | 1]
    6. base backward    from DFEAT = dL/dxhat2
    7. unfold           dW_ih = Gih'[:, :H] diag(g) + db' be^T,  db_ih = db',  dg = colsum(Gih' * W_ih),  dbe = W_ih^T db',
                        dW_hh = Ghh[:, :H],  db_hh = Ghh[:, H]   (b_hh[rz] enters b' additively, so its gradient equals db'[rz])
"""
import numpy as np
import pytest
import torch


def ln_stats(a, eps=1e-5):
    mu = a.mean(1, keepdims=True)
    var = ((a - mu) ** 2).mean(1, keepdims=True)
    return mu, 1.0 / np.sqrt(var + eps)


def ln_bwd(d, xh, rs):
    """dL/dxhat -> dL/d(input of the LayerNorm)"""
    return rs * (d - d.mean(1, keepdims=True) - xh * (d * xh).mean(1, keepdims=True))


def sigm(x):
    return 1.0 / (1.0 + np.exp(-x))


def staged_pipeline(p, x, h0, masks, dlogits_fn, L, Nc, relu):
    """x [P, in], h0 [Nc, H], masks [P] (position p = l * Nc + c); returns (logits, grads) of the staged computation."""
    H = h0.shape[1]
    P = L * Nc
    act = (lambda z: np.maximum(z, 0.0)) if relu else np.tanh
    dact = (lambda a: (a > 0).astype(a.dtype)) if relu else (lambda a: 1.0 - a * a)
    # ---- 1. base forward (fold: fn -> fc1, ln1 -> fc2; output xhat2 pre-affine) ----
    mu0, rs0 = ln_stats(x)
    xh0 = (x - mu0) * rs0
    W1f, b1f = p["fc1.w"] * p["fn.w"][None, :], p["fc1.b"] + p["fc1.w"] @ p["fn.b"]
    a1 = act(xh0 @ W1f.T + b1f)
    mu1, rs1 = ln_stats(a1)
    xh1 = (a1 - mu1) * rs1
    W2f, b2f = p["fc2.w"] * p["ln1.w"][None, :], p["fc2.b"] + p["fc2.w"] @ p["ln1.b"]
    a2 = act(xh1 @ W2f.T + b2f)
    mu2, rs2 = ln_stats(a2)
    X = (a2 - mu2) * rs2
    # ---- 2. sequence forward ----
    Wihf = p["gru.wih"] * p["ln2.w"][None, :]
    bx = p["gru.bih"] + p["gru.wih"] @ p["ln2.b"]
    bx[:2 * H] += p["gru.bhh"][:2 * H]
    Whh, bhn = p["gru.whh"], p["gru.bhh"][2 * H:]
    R, Z, N, GHN, Hs, HM = (np.zeros((P, H)) for _ in range(6))
    h = h0.copy()
    for l in range(L):
        sl = slice(l * Nc, (l + 1) * Nc)
        hm = h * masks[sl, None]
        gi = X[sl] @ Wihf.T + bx
        gh = hm @ Whh.T
        r = sigm(gi[:, :H] + gh[:, :H])
        z = sigm(gi[:, H:2 * H] + gh[:, H:2 * H])
        ghn = gh[:, 2 * H:] + bhn
        n = np.tanh(gi[:, 2 * H:] + r * ghn)
        h = (1.0 - z) * n + z * hm
        R[sl], Z[sl], N[sl], GHN[sl], Hs[sl], HM[sl] = r, z, n, ghn, h, hm
    # ---- 3. heads (fold: rnn.norm -> heads) ----
    muh, rsh = ln_stats(Hs)
    xhh = (Hs - muh) * rsh
    Whf, bhf = p["head.w"] * p["rln.w"][None, :], p["head.b"] + p["head.w"] @ p["rln.b"]
    logits = xhh @ Whf.T + bhf
    dL = dlogits_fn(logits)
    Gh = dL.T @ np.concatenate([xhh, np.ones((P, 1))], 1)                 # [A, H + 1]: dWh', dbh'
    DHH = ln_bwd(dL @ Whf, xhh, rsh)
    # ---- 4. BPTT ----
    DR, DZ, DN = (np.zeros((P, H)) for _ in range(3))
    dh = np.zeros((Nc, H))
    for l in range(L - 1, -1, -1):
        sl = slice(l * Nc, (l + 1) * Nc)
        d = dh + DHH[sl]
        dn = d * (1.0 - Z[sl]) * (1.0 - N[sl] ** 2)
        dz = d * (HM[sl] - N[sl]) * Z[sl] * (1.0 - Z[sl])
        dr = dn * GHN[sl] * R[sl] * (1.0 - R[sl])
        DR[sl], DZ[sl], DN[sl] = dr, dz, dn
        dgh = np.concatenate([dr, dz, dn * R[sl]], 1)
        dh = (d * Z[sl] + dgh @ Whh) * masks[sl, None]
    # ---- 5. gate gradients ----
    dgi = np.concatenate([DR, DZ, DN], 1)
    dgh = np.concatenate([DR, DZ, DN * R], 1)
    DFEAT = dgi @ Wihf
    Gih = dgi.T @ np.concatenate([X, np.ones((P, 1))], 1)                 # [3H, H + 1]
    Ghh = dgh.T @ np.concatenate([HM, np.ones((P, 1))], 1)
    # ---- 6. base backward ----
    dZ2 = ln_bwd(DFEAT, X, rs2) * dact(a2)
    G2 = dZ2.T @ np.concatenate([xh1, np.ones((P, 1))], 1)
    dZ1 = ln_bwd(dZ2 @ W2f, xh1, rs1) * dact(a1)
    G1 = dZ1.T @ np.concatenate([xh0, np.ones((P, 1))], 1)
    # ---- 7. unfold ----
    g = {}

    def unfold(G, W, gam, bet, wk, bk, gk, bek):
        dWf, dbf = G[:, :-1], G[:, -1]
        g[wk] = dWf * gam[None, :] + np.outer(dbf, bet)
        g[bk] = dbf
        g[gk] = (dWf * W).sum(0)
        g[bek] = W.T @ dbf

    unfold(G1, p["fc1.w"], p["fn.w"], p["fn.b"], "fc1.w", "fc1.b", "fn.w", "fn.b")
    unfold(G2, p["fc2.w"], p["ln1.w"], p["ln1.b"], "fc2.w", "fc2.b", "ln1.w", "ln1.b")
    unfold(Gih, p["gru.wih"], p["ln2.w"], p["ln2.b"], "gru.wih", "gru.bih", "ln2.w", "ln2.b")
    unfold(Gh, p["head.w"], p["rln.w"], p["rln.b"], "head.w", "head.b", "rln.w", "rln.b")
    g["gru.whh"] = Ghh[:, :-1]
    g["gru.bhh"] = Ghh[:, -1]
    return logits, g


def torch_reference(p, x, h0, masks, dlogits_fn, L, Nc, relu):
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    H = h0.shape[1]
    act = torch.relu if relu else torch.tanh
    F = torch.nn.functional
    xt = torch.tensor(x)
    y = F.layer_norm(xt, (x.shape[1],), t["fn.w"], t["fn.b"])
    y = F.layer_norm(act(F.linear(y, t["fc1.w"], t["fc1.b"])), (H,), t["ln1.w"], t["ln1.b"])
    y = F.layer_norm(act(F.linear(y, t["fc2.w"], t["fc2.b"])), (H,), t["ln2.w"], t["ln2.b"])
    h = torch.tensor(h0)
    m = torch.tensor(masks)
    outs = []
    for l in range(L):
        sl = slice(l * Nc, (l + 1) * Nc)
        hm = h * m[sl, None]
        gi = F.linear(y[sl], t["gru.wih"], t["gru.bih"])
        gh = F.linear(hm, t["gru.whh"], t["gru.bhh"])
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * hm
        outs.append(h)
    hs = F.layer_norm(torch.cat(outs, 0), (H,), t["rln.w"], t["rln.b"])
    logits = F.linear(hs, t["head.w"], t["head.b"])
    logits.backward(torch.tensor(dlogits_fn(logits.detach().numpy())))
    return logits.detach().numpy(), {k: v.grad.numpy() for k, v in t.items()}


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("L,Nc", [(10, 7), (4, 33)])
def test_staged_gru_pipeline_matches_autograd(relu, L, Nc):
    rng = np.random.RandomState(3 + L)
    H, ind, A = 64, 21, 15
    P = L * Nc

    def w(*shape, scale=0.3):
        return rng.randn(*shape) * scale

    p = {"fn.w": 1 + w(ind, scale=0.2), "fn.b": w(ind, scale=0.2), "fc1.w": w(H, ind), "fc1.b": w(H), "ln1.w": 1 + w(H, scale=0.2),
         "ln1.b": w(H, scale=0.2), "fc2.w": w(H, H, scale=0.15), "fc2.b": w(H), "ln2.w": 1 + w(H, scale=0.2), "ln2.b": w(H, scale=0.2),
         "gru.wih": w(3 * H, H, scale=0.15), "gru.whh": w(3 * H, H, scale=0.15), "gru.bih": w(3 * H), "gru.bhh": w(3 * H),
         "rln.w": 1 + w(H, scale=0.2), "rln.b": w(H, scale=0.2), "head.w": w(A, H), "head.b": w(A)}
    x = rng.randn(P, ind)
    h0 = rng.randn(Nc, H) * 0.5
    masks = (rng.rand(P) > 0.15).astype(np.float64)        # episode boundaries inside the chunks
    cot = rng.randn(P, A)

    def dlogits_fn(logits):                                 # any smooth loss: d/dlogits of sum(cot * tanh(logits))
        return cot * (1.0 - np.tanh(logits) ** 2)

    lg, g = staged_pipeline(p, x, h0, masks, dlogits_fn, L, Nc, relu)
    lg_ref, g_ref = torch_reference(p, x, h0, masks, dlogits_fn, L, Nc, relu)
    np.testing.assert_allclose(lg, lg_ref, rtol=1e-9, atol=1e-10)
    for k in p:
        np.testing.assert_allclose(g[k], g_ref[k], rtol=1e-7, atol=1e-9 * (1 + np.abs(g_ref[k]).max()), err_msg=k)
