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
