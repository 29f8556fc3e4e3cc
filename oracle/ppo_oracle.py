"""numpy restatement of the reference's PPO arithmetic on the hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functions cite the reference file:line they follow.  Everything is float32 with the reference's operation order
(python scalars meeting fp32 tensors are rounded to fp32 first, exactly as torch does on CPU).
Pinned against the unmodified reference by tools/make_golden.py -> tests/golden/ppo_kats.npz.
"""
from __future__ import annotations

import numpy as np

F = np.float32
FEATURE_DIMS = [4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2, 2, 1, 4, 1, 1, 1, 1, 1, 1, 2]  # networks.py:8
FEATURE_AMOUNT, OBS_SPACE, EMBEDDING_DIM = 23, 65, 20                                  # networks.py:9-11


def get_gaes(ep_rew, ep_values, ep_dones, gamma=0.99, lam=0.95):
    """PPO.get_GAEs, PPO.py:193-203, one complete episode.  ep_values are fp32 (critic outputs)."""
    n = len(ep_rew)
    adv = np.zeros(n, np.float64)  # np.zeros_like(list of python floats)
    g = F(gamma)
    gl = F(np.float64(gamma) * np.float64(lam))  # python float product, rounded when it meets the tensor
    a = F(0.0)
    for t in reversed(range(n)):
        if t + 1 == n:
            delta = F(F(ep_rew[t]) - F(ep_values[t]))
        else:
            nd = F(1 - int(bool(ep_dones[t + 1])))
            delta = F(F(F(ep_rew[t]) + F(F(g * F(ep_values[t + 1])) * nd)) - F(ep_values[t]))
        a = F(delta + F(F(gl * F(1 - int(bool(ep_dones[t])))) * a))
        adv[t] = a
    return adv.astype(np.float32)


def gae_fixed_horizon(reward, value, done, v_boot, gamma=0.99, lam=0.95):
    """[T,E] buffers with episode boundaries at done[t,e]; complete episodes go through get_gaes verbatim, the open tail
    of each env is bootstrapped with v_boot[e] (extension documented in DESIGN.md): it is treated as an episode whose
    successor value is v_boot and whose last step is not terminal."""
    T, E = reward.shape
    adv = np.zeros((T, E), np.float32)
    g = F(gamma); gl = F(np.float64(gamma) * np.float64(lam))
    for e in range(E):
        start = 0
        for t in range(T):
            if done[t, e]:
                adv[start:t + 1, e] = get_gaes(reward[start:t + 1, e], value[start:t + 1, e], done[start:t + 1, e], gamma, lam)
                start = t + 1
        if start < T:  # open tail
            a = F(0.0); vn = F(v_boot[e])
            for t in reversed(range(start, T)):
                delta = F(F(F(reward[t, e]) + F(F(g * vn) * F(1.0))) - F(value[t, e]))
                a = F(delta + F(gl * a)) if t + 1 < T else delta
                adv[t, e] = a; vn = F(value[t, e])
    return adv


# ------------------------------------------------------------------------------------------------ networks.py
def _linear(x, w, b=None):
    y = x.astype(F) @ w.astype(F).T
    return y + b.astype(F) if b is not None else y


def actor_forward(sd: dict, obs: np.ndarray, faithful: bool = True):
    """Actor.forward, networks.py:31-41 with Projection (:58-65) and m_Attention (:75-82).
    faithful=True reproduces Projection.forward never advancing `index` (networks.py:59-63): every projection reads
    obs[:, 0:FEATURE_DIMS[i]].  Returns (move_logits [B,5], mark_logits [B,1])."""
    x = np.asarray(obs, F).reshape(-1, OBS_SPACE)
    toks = []
    index = 0
    for i, d in enumerate(FEATURE_DIMS):
        sl = x[:, 0:d] if faithful else x[:, index:index + d]
        toks.append(_linear(sl, sd[f"projection.layers.{i}.weight"], sd[f"projection.layers.{i}.bias"]))
        index += d
    tok = np.stack(toks, axis=1)  # [B,23,20]
    k = tok @ sd["attention.keys.weight"].astype(F).T
    q = tok @ sd["attention.querys.weight"].astype(F).T
    v = tok @ sd["attention.values.weight"].astype(F).T
    logits = np.einsum("bij,bkj->bik", q, k) / F(np.sqrt(10))
    logits = logits - logits.max(-1, keepdims=True)
    om = np.exp(logits); om = om / om.sum(-1, keepdims=True)
    ctx = np.einsum("bij,bjk->bik", om, v)
    h = (tok + ctx).reshape(-1, FEATURE_AMOUNT * EMBEDDING_DIM).astype(F)
    for l in range(3):
        h = np.maximum(_linear(h, sd[f"layers.{l}.weight"], sd[f"layers.{l}.bias"]), 0)
    return _linear(h, sd["move_head.weight"], sd["move_head.bias"]), _linear(h, sd["mark_head.weight"], sd["mark_head.bias"])


def critic_forward(sd: dict, obs: np.ndarray):
    """Critic.forward, networks.py:96-102: [E, 2*65] -> 64 -> 64 -> 1."""
    h = np.asarray(obs, F).reshape(-1, 2 * OBS_SPACE)
    h = np.maximum(_linear(h, sd["layers.0.weight"], sd["layers.0.bias"]), 0)
    h = np.maximum(_linear(h, sd["layers.1.weight"], sd["layers.1.bias"]), 0)
    return _linear(h, sd["layers.2.weight"], sd["layers.2.bias"])


def action_log_prob(move_logits, mark_logits, masks, moves, marks):
    """Joint log-prob of one agent's action, PPO.get_action / get_log_probs (PPO.py:154-186):
    log_softmax(masked move logits)[move] + log(p_mark if mark else 1 - p_mark), p_mark = sigmoid(logit) if mask[5] else 0."""
    ml = np.where(masks[:, :5].astype(bool), move_logits.astype(np.float64), -np.inf)
    ml = ml - ml.max(-1, keepdims=True)
    lse = np.log(np.exp(ml).sum(-1, keepdims=True))
    lp_move = np.take_along_axis(ml - lse, moves.astype(np.int64)[:, None], 1)[:, 0]
    p = np.where(masks[:, 5].astype(bool), 1.0 / (1.0 + np.exp(-mark_logits.reshape(-1).astype(np.float64))), 0.0)
    pm = np.where(marks.astype(bool), p, 1.0 - p)
    with np.errstate(divide="ignore"):
        return (lp_move + np.log(pm)).astype(np.float32)


def seeded_state_dicts(seed: int):
    """Deterministic, platform-independent weights with the reference's parameter names and shapes
    (actor 265 774 params, critic 12 609) -- so golden outputs can be reproduced without shipping a checkpoint."""
    rng = np.random.default_rng(seed)

    def lin(o, i, scale=None, bias=True):
        s = scale if scale is not None else 1.0 / np.sqrt(i)
        w = (rng.standard_normal((o, i)) * s).astype(F)
        return w, ((rng.standard_normal(o) * 0.1).astype(F) if bias else None)
    actor = {}
    for i, d in enumerate(FEATURE_DIMS):
        actor[f"projection.layers.{i}.weight"], actor[f"projection.layers.{i}.bias"] = lin(EMBEDDING_DIM, d)
    actor["attention.keys.weight"], _ = lin(10, EMBEDDING_DIM, bias=False)
    actor["attention.querys.weight"], _ = lin(10, EMBEDDING_DIM, bias=False)
    actor["attention.values.weight"], _ = lin(EMBEDDING_DIM, EMBEDDING_DIM, bias=False)
    dims = [FEATURE_AMOUNT * EMBEDDING_DIM, 264, 264, 264]
    for l in range(3):
        actor[f"layers.{l}.weight"], actor[f"layers.{l}.bias"] = lin(dims[l + 1], dims[l])
    actor["move_head.weight"], actor["move_head.bias"] = lin(5, 264, scale=0.05)
    actor["mark_head.weight"], actor["mark_head.bias"] = lin(1, 264, scale=0.05)
    critic = {}
    cd = [2 * OBS_SPACE, 64, 64, 1]
    for l in range(3):
        critic[f"layers.{l}.weight"], critic[f"layers.{l}.bias"] = lin(cd[l + 1], cd[l])
    return actor, critic


def actor_update(sd: dict, obs, masks, actions, old_logp, adv, clip=0.2, faithful: bool = True):
    """Actor half of PPO.train's inner loop (PPO.py:58-76) with its backward pass written out by hand, in float64:
        joint_e = sum_i get_log_probs(i)           (PPO.py:65-67, 154-168)
        ratio   = exp(joint - old_logp);  loss = -mean(min(ratio*A, clamp(ratio, 1-clip, 1+clip)*A))     (PPO.py:68-74)
    and d loss / d (every actor parameter) -- what actor_loss.backward() leaves in .grad (PPO.py:76).  The reference's backward is torch
    autograd; this restates the chain rule through the heads, the three Linear+ReLU layers, the residual single-head attention
    (networks.py:75-82) and the 23 projections (networks.py:58-65), with torch's sub-gradient conventions: minimum() splits a tie evenly,
    clamp() passes the gradient on its closed range, relu'(0) = 0.  obs [E,2,65], masks [E,2,6], actions [E,2,2] (move, mark).
    Returns (loss, joint [E], grads: dict name -> array)."""
    D = np.float64
    P = {k: np.asarray(v, D) for k, v in sd.items()}
    x = np.asarray(obs, D).reshape(-1, OBS_SPACE); B = x.shape[0]; E = B // 2
    mk = np.asarray(masks).reshape(B, 6).astype(bool); act = np.asarray(actions).reshape(B, 2).astype(np.int64)
    # ---- forward
    sl, index = [], 0
    for d in FEATURE_DIMS:
        sl.append((0, d) if faithful else (index, index + d)); index += d
    tok = np.stack([x[:, a:b] @ P[f"projection.layers.{i}.weight"].T + P[f"projection.layers.{i}.bias"] for i, (a, b) in enumerate(sl)], 1)
    Wk, Wq, Wv = P["attention.keys.weight"], P["attention.querys.weight"], P["attention.values.weight"]
    k, q, v = tok @ Wk.T, tok @ Wq.T, tok @ Wv.T
    S = np.einsum("bij,bkj->bik", q, k) / np.sqrt(10.0)
    Pm = np.exp(S - S.max(-1, keepdims=True)); Pm /= Pm.sum(-1, keepdims=True)
    hs = [(tok + np.einsum("bij,bjk->bik", Pm, v)).reshape(B, FEATURE_AMOUNT * EMBEDDING_DIM)]
    zs = []
    for l in range(3):
        zs.append(hs[-1] @ P[f"layers.{l}.weight"].T + P[f"layers.{l}.bias"]); hs.append(np.maximum(zs[-1], 0))
    ml = hs[-1] @ P["move_head.weight"].T + P["move_head.bias"]
    kl = (hs[-1] @ P["mark_head.weight"].T + P["mark_head.bias"]).reshape(B)
    mlm = np.where(mk[:, :5], ml, -np.inf)
    sm = np.exp(mlm - mlm.max(-1, keepdims=True)); sm /= sm.sum(-1, keepdims=True)
    p = np.where(mk[:, 5], 1.0 / (1.0 + np.exp(-kl)), 0.0)
    rows = np.arange(B)
    with np.errstate(divide="ignore"):
        lp = np.log(sm[rows, act[:, 0]]) + np.log(np.where(act[:, 1] == 1, p, 1.0 - p))
    joint = lp.reshape(E, 2).sum(1)
    A = np.asarray(adv, D); ratio = np.exp(joint - np.asarray(old_logp, D))
    s1, s2 = ratio * A, np.clip(ratio, 1 - clip, 1 + clip) * A
    loss = -np.minimum(s1, s2).mean()
    # ---- backward
    inside = (ratio >= 1 - clip) & (ratio <= 1 + clip)
    g1 = np.where(s1 < s2, 1.0, np.where(s1 == s2, 0.5, 0.0))     # share of min()'s gradient that goes to surrogate1
    dratio = -(g1 * A + (1.0 - g1) * A * inside) / E
    dlp = np.repeat(dratio * ratio, 2)                             # d loss / d (each agent row's log-prob)
    onehot = np.zeros((B, 5)); onehot[rows, act[:, 0]] = 1.0
    dml = np.where(mk[:, :5], dlp[:, None] * (onehot - sm), 0.0)
    dkl = np.where(mk[:, 5], dlp * (act[:, 1] - p), 0.0)
    G = {"move_head.weight": dml.T @ hs[-1], "move_head.bias": dml.sum(0), "mark_head.weight": (dkl[:, None] * hs[-1]).sum(0, keepdims=True),
         "mark_head.bias": dkl.sum(keepdims=True)}
    dh = dml @ P["move_head.weight"] + dkl[:, None] * P["mark_head.weight"]
    for l in (2, 1, 0):
        dz = dh * (zs[l] > 0)
        G[f"layers.{l}.weight"], G[f"layers.{l}.bias"] = dz.T @ hs[l], dz.sum(0)
        dh = dz @ P[f"layers.{l}.weight"]
    dctx = dh.reshape(B, FEATURE_AMOUNT, EMBEDDING_DIM)
    dPm = np.einsum("bik,bjk->bij", dctx, v); dv = np.einsum("bij,bik->bjk", Pm, dctx)
    dS = Pm * (dPm - (dPm * Pm).sum(-1, keepdims=True)) / np.sqrt(10.0)
    dq, dk = np.einsum("bij,bjk->bik", dS, k), np.einsum("bij,bik->bjk", dS, q)
    G["attention.querys.weight"] = np.einsum("bik,bij->kj", dq, tok); G["attention.keys.weight"] = np.einsum("bik,bij->kj", dk, tok)
    G["attention.values.weight"] = np.einsum("bik,bij->kj", dv, tok)
    dtok = dctx + dq @ Wq + dk @ Wk + dv @ Wv
    for i, (a, b) in enumerate(sl):
        G[f"projection.layers.{i}.weight"], G[f"projection.layers.{i}.bias"] = dtok[:, i].T @ x[:, a:b], dtok[:, i].sum(0)
    return loss, joint, G
