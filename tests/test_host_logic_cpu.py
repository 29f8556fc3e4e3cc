"""Host-side logic of the reference-surface mirror that needs no GPU (maze-pool keys, checkpoint format)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Brain:
    maze = None


def _maze(**kw):
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent
    b = _Brain()
    return Maze(agents=(Agent("RED", b, None, None, 2), Agent("BLUE", b, None, None, 3)), max_timestep=1200, rand_sizes=True, rand_range=[12, 13],
                rand_start=True, **kw)


def test_pool_keys_never_repeat_across_refills():
    """K1 keys Philox on (64-bit seed, 32-bit maze id).  The id is the maze's global slot and the refill counter is
    folded into the seed, so refill 64 does not replay refill 0 (a 32-bit id of generation * 2^26 + slot wrapped there), and a high
    rank's refill g never meets a low rank's refill g + 1."""
    m = _maze(num_envs=1 << 20, pool_episodes=8, env_offset=7 << 20, seed=5)
    keys = set()
    for g in list(range(0, 130)) + [1 << 20, (1 << 32) + 1]:
        a = m._pool_args(g)
        assert a["id_base"] == (7 << 20) * 8 and a["id_mod"] == 1 << 20 and a["id_mul"] == 8      # ids: (global env) * K + episode
        assert 0 <= a["seed"] < 1 << 64
        keys.add(a["seed"])
    assert len(keys) == 132, "every refill must draw from its own seed"
    assert m._pool_args(0)["seed"] == 5                                                             # refill 0 keeps the user's seed
    # ids of different ranks are disjoint for the same refill (global env slots), so equal seeds there are fine
    lo, hi = _maze(num_envs=1024, pool_episodes=4, env_offset=0, seed=5), _maze(num_envs=1024, pool_episodes=4, env_offset=1024, seed=5)
    assert lo._pool_args(3)["seed"] == hi._pool_args(3)["seed"]
    assert lo._pool_args(3)["id_base"] + 1024 * 4 <= hi._pool_args(3)["id_base"]
    with pytest.raises(ValueError):
        _maze(num_envs=1 << 20, pool_episodes=8192, env_offset=0)._pool_args(0)


def test_checkpoint_written_here_loads_in_the_reference(tmp_path):
    """SURVEY 8(f).2 / VERDICT r1: a checkpoint written by marl_maze_b200.PPO.save_parameters is read by the REFERENCE's own
    PPO.load_parameters (PPO.py:229-238, bare torch.load) and gives identical actor logits and critic values, Adam state and decayed lr
    included.  Runs the unmodified reference in a subprocess whose CWD holds the checkpoint; only where the reference is present."""
    import json
    import subprocess
    import numpy as np
    from baseline import stage_reference
    ref = stage_reference.stage()
    if ref is None:
        pytest.skip("reference not present")
    from marl_maze_b200.PPO import PPO
    path = str(tmp_path / "PPO.pth")
    brain = PPO(agent_amount=2, lr=0.00014, device="cpu", model_path=path, verbose=False, seed=11)
    # one optimiser step each so that the Adam state is populated, and a decayed learning rate
    obs = torch.rand(6, 65)
    mv, mk = brain.actor(obs)
    (mv.sum() + mk.sum()).backward(); brain.actor_optim.step()
    brain.critic(torch.rand(3, 2, 65)).sum().backward(); brain.critic_optim.step()
    brain.decay_lr(); brain.decay_lr()
    brain.save_parameters()
    sd = torch.load(path)                      # bare torch.load, as the reference does
    assert all(not v.is_cuda for v in sd["actor"].values()) and sorted(sd) == ["actor", "actor_optim", "critic", "critic_optim"]
    probe = torch.zeros(4, 65); probe[torch.arange(4), torch.arange(4)] = 1.0; probe[:, 30:] = torch.linspace(0, 1, 35)
    cobs = torch.rand(5, 2, 65, generator=torch.Generator().manual_seed(3))
    np.save(str(tmp_path / "probe.npy"), probe.numpy()); np.save(str(tmp_path / "cobs.npy"), cobs.numpy())
    code = (
        "import sys, types, json, io, contextlib, numpy as np, torch\n"
        "stub = types.ModuleType('pygame'); stub.Color = lambda *a: a; sys.modules['pygame'] = stub\n"
        f"sys.path.insert(0, {ref!r})\n"
        "with contextlib.redirect_stdout(io.StringIO()) as out:\n"
        "    import PPO as ref\n"
        "    b = ref.PPO(agent_amount=2, lr=0.00014)\n"
        "loaded = 'successfuly loaded' in out.getvalue()\n"
        "with torch.no_grad():\n"
        "    mv, mk = b.actor(torch.from_numpy(np.load('probe.npy')))\n"
        "    v = b.critic(torch.from_numpy(np.load('cobs.npy')))\n"
        "st = b.actor_optim.state_dict()\n"
        "print(json.dumps(dict(loaded=loaded, mv=mv.tolist(), mk=mk.tolist(), v=v.tolist(), lr=st['param_groups'][0]['lr'],\n"
        "                      step=float(list(st['state'].values())[0]['step']), clr=b.critic_optim.state_dict()['param_groups'][0]['lr'])))\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["loaded"], "the reference did not pick the checkpoint up"
    with torch.no_grad():
        mv, mk = brain.actor(probe)
        v = brain.critic(cobs)
    assert np.array_equal(np.asarray(d["mv"], np.float32), mv.numpy()) and np.array_equal(np.asarray(d["mk"], np.float32), mk.numpy())
    assert np.array_equal(np.asarray(d["v"], np.float32), v.numpy())
    assert d["lr"] == pytest.approx(0.00014 * 0.997 ** 2, rel=1e-12) and d["clr"] == pytest.approx(0.00014 * 0.997 ** 2, rel=1e-12) and d["step"] == 1.0


def test_sample_action_is_the_reference_draw_bit_for_bit():
    """PPO._sample_action (the host-side draw of PPO.get_action on a GPU, and the whole of it on a CPU) against the reference's own lines
    (PPO.py:175-186: masked Categorical, sigmoid / Bernoulli mark, joint log-prob) evaluated with torch.distributions on the same logits and the same
    generator state: identical actions and bit-identical log-probs, for every mask pattern that leaves a legal move."""
    from marl_maze_b200.PPO import PPO
    g = torch.Generator().manual_seed(0)
    n = 0
    for trial in range(300):
        mv = torch.randn(1, 5, generator=g) * (0.1 + trial % 7)
        mk = torch.randn(1, 1, generator=g) * 3
        mask = [bool(v) for v in (torch.rand(6, generator=g) < 0.6).tolist()]
        if not any(mask[:5]):
            mask[trial % 5] = True
        torch.manual_seed(1000 + trial)
        m = torch.tensor(mask)
        dist = torch.distributions.Categorical(logits=torch.where(m[:5], mv, torch.tensor(-float("inf"))))
        move = dist.sample()
        p = torch.sigmoid(mk) if mask[5] else torch.zeros(1, 1)
        mark = torch.bernoulli(p)
        p = p if mark == 1 else 1 - p
        want = ([int(move.item()), float(mark.item())], dist.log_prob(move) + torch.log(p))
        torch.manual_seed(1000 + trial)
        got = PPO._sample_action(mv, mk, mask)
        assert got[0] == want[0] and mask[got[0][0]] and (mask[5] or got[0][1] == 0.0), (trial, got, want)
        assert torch.equal(got[1].reshape(-1), want[1].reshape(-1)), (trial, got[1], want[1])
        n += 1
    assert n == 300
