"""The actor-update oracle (oracle/ppo_oracle.actor_update: PPO.py:58-76 with its backward written out by hand, float64) pinned against
the UNMODIFIED reference: tests/golden/upd_kats.npz holds the loss, the joint log-probs and every actor parameter gradient that the
reference's own get_log_probs + actor_loss.backward() produced on torch CPU fp32 (tools/make_golden.py --upd)."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN
from oracle import ppo_oracle as po

BIG = ("layers.0.weight", "layers.1.weight", "layers.2.weight")


def golden_grads(Z, seed):
    pre = f"upd/{seed}/grad/"
    return {k[len(pre):]: Z[k] for k in Z.files if k.startswith(pre)}


def compare(name, got, ref, stride, tol):
    got = np.asarray(got, np.float64)
    if name in BIG:
        got = got.reshape(-1)[::stride]
    assert got.shape == ref.shape, name
    scale = np.abs(ref).max()
    assert scale > 0, name
    err = np.abs(got - ref).max() / scale
    assert err < tol, (name, err)
    return err


@pytest.mark.parametrize("seed", [11, 12])
def test_actor_update_oracle_matches_reference_backward(seed):
    Z = np.load(os.path.join(GOLDEN, "upd_kats.npz"))
    asd, _ = po.seeded_state_dicts(seed)
    loss, joint, G = po.actor_update(asd, Z["upd/obs"], Z["upd/masks"], Z["upd/actions"], Z[f"upd/{seed}/old"], Z[f"upd/{seed}/adv"], float(Z["upd/clip"]))
    assert 0.3 < float(Z[f"upd/{seed}/frac_clipped"]) < 0.7                      # the clip branch is really exercised
    assert abs(loss - float(Z[f"upd/{seed}/loss"])) < 2e-6 * max(1.0, abs(loss))
    assert np.allclose(joint, Z[f"upd/{seed}/joint"], rtol=1e-5, atol=1e-5)
    ref = golden_grads(Z, seed)
    assert set(ref) == set(G) and len(ref) == 59
    worst = max(compare(n, G[n], ref[n], int(Z["upd/stride"]), 5e-5) for n in ref)  # the reference is fp32 autograd: ~1e-6 expected
    print("worst", worst)


def test_actor_update_oracle_clip_subgradients():
    """Ratios exactly inside, below and above the clip range, both advantage signs: the gradient is A*ratio inside or when the
    unclipped surrogate is the smaller one, 0 otherwise (torch.minimum / torch.clamp conventions)."""
    Z = np.load(os.path.join(GOLDEN, "upd_kats.npz"))
    asd, _ = po.seeded_state_dicts(11)
    pick = np.flatnonzero((Z["upd/masks"][:, :, :5].sum(-1) >= 2).all(1))[:6]     # envs whose log-prob depends on the parameters at all
    obs, masks, acts = Z["upd/obs"][pick], Z["upd/masks"][pick], Z["upd/actions"][pick]
    _, joint, _ = po.actor_update(asd, obs, masks, acts, np.zeros(6), np.ones(6))
    old = joint - np.log(np.array([1.0, 0.5, 0.5, 2.0, 2.0, 1.1]))      # ratio = 1, .5, .5, 2, 2, 1.1
    adv = np.array([1.0, 1.0, -1.0, 1.0, -1.0, -1.0])
    base = po.actor_update(asd, obs, masks, acts, old, adv)[2]["move_head.bias"]
    live = []
    for e in range(6):   # drop env e's advantage to 0: its contribution to the gradient disappears iff it had one
        a2 = adv.copy(); a2[e] = 0.0
        live.append(not np.allclose(po.actor_update(asd, obs, masks, acts, old, a2)[2]["move_head.bias"], base, rtol=0, atol=1e-12))
    assert live == [True, True, False, False, True, True]
