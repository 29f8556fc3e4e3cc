"""oracle/ppo_oracle.py (numpy restatement of PPO.get_GAEs, Actor/Critic forward, action log-probs) and the product's
torch networks against known answers recorded from the reference's own PPO.py / networks.py (tests/golden/ppo_kats.npz)."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN
from oracle import ppo_oracle as po

Z = np.load(os.path.join(GOLDEN, "ppo_kats.npz"))


def test_gae_oracle_bit_exact_vs_reference():
    for k in range(int(Z["gae/n"])):
        rew, val, adv = Z[f"gae/{k}/rew"], Z[f"gae/{k}/val"], Z[f"gae/{k}/adv"]
        dones = [False] * (len(rew) - 1) + [True]
        got = po.get_gaes(rew, val, dones)
        assert np.array_equal(got.view(np.uint32), adv.view(np.uint32)), (k, np.abs(got - adv).max())


def test_gae_kat4():
    got = po.get_gaes([0, .5, 0, 1], np.full(4, 0.1, np.float32), [False, False, False, True])
    assert np.array_equal(got, Z["gae/kat4"])
    assert np.allclose(got, [1.12857449, 1.20103621, 0.74644995, 0.89999998], rtol=0, atol=1e-7)  # SURVEY 8c KAT(4)


def test_gae_fixed_horizon_equals_per_episode():
    rng = np.random.default_rng(0)
    T, E = 64, 9
    rew = rng.choice([0.0, 0.5, 1.0], size=(T, E)).astype(np.float32); val = rng.standard_normal((T, E)).astype(np.float32)
    done = (rng.random((T, E)) < 0.08); done[-1, ::2] = True
    vb = rng.standard_normal(E).astype(np.float32)
    adv = po.gae_fixed_horizon(rew, val, done, vb)
    e = 0
    cuts = [-1] + [t for t in range(T) if done[t, e]]
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert np.array_equal(adv[a + 1:b + 1, e], po.get_gaes(rew[a + 1:b + 1, e], val[a + 1:b + 1, e], done[a + 1:b + 1, e]))


@pytest.mark.parametrize("seed", [11, 12])
def test_network_oracle_and_torch_modules_vs_reference(seed):
    """tolerance: 1e-5 relative (north_star) on logits / values / log-probs; reductions are re-associated, nothing else differs."""
    from marl_maze_b200.networks import Actor, Critic
    asd, csd = po.seeded_state_dicts(seed)
    obs, masks, acts = Z["net/obs"], Z["net/masks"], Z["net/actions"]
    mv_ref, mk_ref, v_ref, lp_ref = (Z[f"net/{seed}/{k}"] for k in ("move_logits", "mark_logits", "values", "log_probs"))
    mv, mk = po.actor_forward(asd, obs.reshape(-1, 65)); v = po.critic_forward(csd, obs)
    tol = dict(rtol=1e-5, atol=2e-6)
    assert np.allclose(mv, mv_ref, **tol) and np.allclose(mk, mk_ref, **tol) and np.allclose(v, v_ref, **tol)
    lp = np.stack([po.action_log_prob(mv.reshape(-1, 2, 5)[:, i], mk.reshape(-1, 2)[:, i], masks[:, i], acts[:, i, 0], acts[:, i, 1]) for i in range(2)], 1)
    fin = np.isfinite(lp_ref)
    assert np.array_equal(np.isfinite(lp), fin) and np.allclose(lp[fin], lp_ref[fin], rtol=1e-5, atol=2e-6)
    actor = Actor([264, 264, 264]); critic = Critic(2, hidden_sizes=[64, 64])
    actor.load_state_dict({k: torch.from_numpy(x) for k, x in asd.items()}); critic.load_state_dict({k: torch.from_numpy(x) for k, x in csd.items()})
    with torch.no_grad():
        tmv, tmk = actor(torch.from_numpy(obs.reshape(-1, 65))); tv = critic(torch.from_numpy(obs))
    assert np.allclose(tmv.numpy(), mv_ref, **tol) and np.allclose(tmk.numpy(), mk_ref, **tol) and np.allclose(tv.numpy(), v_ref, **tol)
    # the faithful actor is a function of obs[:, 0:4] only (Projection never advances its index, networks.py:59-63)
    obs2 = obs.reshape(-1, 65).copy(); obs2[:, 4:] = 0.123
    with torch.no_grad():
        assert torch.allclose(actor(torch.from_numpy(obs2))[0], tmv, atol=1e-6)
    indexed = Actor([264, 264, 264], faithful_projection=False); indexed.load_state_dict(actor.state_dict())
    with torch.no_grad():
        assert not torch.allclose(indexed(torch.from_numpy(obs2))[0], indexed(torch.from_numpy(obs.reshape(-1, 65)))[0], atol=1e-4)


def test_reference_checkpoint_loads_and_matches_kat5():
    """SURVEY 8c KAT(5): PPO.pth actor on the 4 facing one-hots.  Only where the reference checkout is mounted."""
    path = "/root/reference/PPO.pth"
    if not os.path.exists(path):
        pytest.skip("reference checkpoint not present on this box")
    from marl_maze_b200.networks import Actor, Critic
    sd = torch.load(path, map_location="cpu")
    actor = Actor([264, 264, 264]); critic = Critic(2, hidden_sizes=[64, 64])
    actor.load_state_dict(sd["actor"]); critic.load_state_dict(sd["critic"])
    x = torch.zeros(4, 65); x[torch.arange(4), torch.arange(4)] = 1
    with torch.no_grad():
        mv, mk = actor(x)
    assert np.allclose(mv[0].numpy(), [5.7483, -0.3037, -7.4898, 10.1972, -9.9180], atol=2e-4)
    assert np.allclose(torch.sigmoid(mk).reshape(-1).numpy(), [.5609, .5664, .5459, .5300], atol=2e-4)


def test_kat5_fixture_matches_reference_outputs_on_cpu():
    """tests/golden/kat5_ppo_pth.npz (the reference's PPO.pth and what the reference's networks.py computes from it, recorded by
    tools/make_golden.py --kat5) through this repo's torch modules and the numpy oracle: the fixture the GPU test drives the kernels with."""
    from marl_maze_b200.networks import Actor, Critic
    Z = np.load(os.path.join(GOLDEN, "kat5_ppo_pth.npz"))
    asd = {k[6:]: Z[k] for k in Z.files if k.startswith("actor/")}; csd = {k[7:]: Z[k] for k in Z.files if k.startswith("critic/")}
    actor = Actor([264, 264, 264]); critic = Critic(2, hidden_sizes=[64, 64])
    actor.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()}); critic.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()})
    with torch.no_grad():
        mv4, mk4 = actor(torch.from_numpy(Z["kat5/obs"]))
        mv, mk = actor(torch.from_numpy(Z["trace/obs"].reshape(-1, 65)))
        val = critic(torch.from_numpy(Z["trace/obs"]))
    assert np.allclose(mv4.numpy(), Z["kat5/move_logits"], rtol=1e-6, atol=1e-6) and np.allclose(mk4.numpy().reshape(-1), Z["kat5/mark_logits"], rtol=1e-6, atol=1e-6)
    assert np.allclose(mv4[0].numpy(), [5.7483, -0.3037, -7.4898, 10.1972, -9.9180], atol=2e-4)
    assert np.allclose(mv.numpy(), Z["trace/move_logits"], rtol=1e-6, atol=2e-6) and np.allclose(val.numpy().reshape(-1), Z["trace/values"], rtol=1e-6, atol=1e-6)
    omv, omk = po.actor_forward(asd, Z["trace/obs"].reshape(-1, 65))
    assert np.allclose(omv, Z["trace/move_logits"], rtol=1e-5, atol=2e-6) and np.allclose(omk.reshape(-1), Z["trace/mark_logits"], rtol=1e-5, atol=2e-6)
    if os.path.exists("/root/reference/PPO.pth"):   # the fixture IS the checkpoint
        sd = torch.load("/root/reference/PPO.pth", map_location="cpu")
        assert all(np.array_equal(sd["actor"][k].numpy(), v) for k, v in asd.items()) and all(np.array_equal(sd["critic"][k].numpy(), v) for k, v in csd.items())


def test_actor_embedding_dedup_matches_per_row_evaluation():
    """Large faithful batches evaluate projection+attention once per distinct obs[:, 0:4] prefix; values and gradients must equal
    the plain per-row evaluation."""
    from marl_maze_b200.networks import Actor
    torch.manual_seed(0)
    actor = Actor([264, 264, 264])
    obs = torch.rand(6000, 65); obs[:, :4] = torch.eye(4)[torch.randint(0, 4, (6000,))]
    out = actor(obs)[0].sum() + actor(obs)[1].sum()
    g1 = torch.autograd.grad(out, list(actor.parameters()), allow_unused=True)
    small = [actor(obs[i:i + 1000]) for i in range(0, 6000, 1000)]  # below the dedup threshold: per-row path
    mv = torch.cat([s[0] for s in small]); mk = torch.cat([s[1] for s in small])
    assert torch.allclose(actor(obs)[0], mv, atol=1e-6) and torch.allclose(actor(obs)[1], mk, atol=1e-6)
    g2 = torch.autograd.grad(mv.sum() + mk.sum(), list(actor.parameters()), allow_unused=True)
    for a, b in zip(g1, g2):
        assert (a is None and b is None) or torch.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_token_maps_match_the_packed_policy_weights_and_the_oracle_embedding():
    """update.token_maps (the differentiable folding [I; Wk; Wq; Wv](P_a, b_a) the K5 token kernels are driven with) equals the maps
    policy.pack_weights builds for K4, and evaluating them reproduces the oracle's Projection + attention inputs (tokens, keys,
    queries, values) for both projection modes."""
    import torch
    from marl_maze_b200.networks import Actor, FEATURE_DIMS
    from marl_maze_b200.update import token_maps
    torch.manual_seed(5)
    for faithful in (True, False):
        actor = Actor([264, 264, 264], faithful_projection=faithful)
        tokm, tokb = token_maps(actor)
        assert tokm.shape == (60, 23, 4) and tokb.shape == (60, 23) and tokm.requires_grad
        x = torch.rand(7, 65)
        tok = actor.projection(x)                                   # [7,23,20]
        k, q, v = actor.attention.keys(tok), actor.attention.querys(tok), actor.attention.values(tok)
        want = torch.cat([tok, k, q, v], -1)                        # [7,23,60]
        cols = [0 if faithful else sum(FEATURE_DIMS[:i]) for i in range(23)]
        xin = torch.stack([torch.nn.functional.pad(x[:, c:c + d], (0, 4 - d)) for c, d in zip(cols, FEATURE_DIMS)], 1)   # [7,23,4]
        got = torch.einsum("jac,rac->raj", tokm, xin) + tokb.t().unsqueeze(0)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
        tokm.sum().backward()
        assert all(l.weight.grad is not None for l in actor.projection.layers) and actor.attention.keys.weight.grad is not None
