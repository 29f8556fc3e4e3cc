"""The reference's class surface end to end on the GPU: Maze / Agent / PPO wiring of main.py, list semantics at one env,
batched rollout + GAE + update, checkpoint round trip."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(num_envs, tmp_path, batch_size=600, **kw):
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent
    os.makedirs(tmp_path, exist_ok=True)
    brain = PPO(agent_amount=2, batch_size=batch_size, lr=0.00014, epochs=1, verbose=False, model_path=str(tmp_path / "PPO.pth"), **kw)
    agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
    maze = Maze(agents=agents, max_timestep=120, rand_sizes=True, rand_range=[12, 13], rand_start=True, difficulty=1, default_size=[4, 4],
                num_envs=num_envs, seed=5)  # main.py:17-20
    return brain, agents, maze


def test_single_env_list_semantics_match_oracle(tmp_path):
    """E=1: reset()/step() take and return python lists like the reference; values equal the oracle's on the same maze."""
    from oracle import OracleMaze
    brain, agents, maze = _make(1, tmp_path)
    assert brain.maze is maze and agents[0].maze is maze
    obs, masks = maze.reset()
    assert isinstance(obs, list) and len(obs) == 2 and len(obs[0]) == 65 and isinstance(masks[0][0], bool)
    o = OracleMaze(max_timestep=120)
    m = maze.engine.pool_maze(int(maze.engine.envs()[0, 3]))
    oo, om = o.reset_injected(m)
    assert np.array_equal(np.asarray(obs, np.float32), oo) and np.array_equal(np.asarray(masks, np.uint8), om)
    assert (agents[0].x, agents[0].y) == m["path0"] and (agents[1].x, agents[1].y) == m["path1"] and agents[0].direction == 2
    assert maze.shortest_path[0] == m["start"] and maze.shortest_path[-1] == m["end"] and len(maze.shortest_path) == maze.shortest_path_len
    rng = np.random.default_rng(0)
    for t in range(150):
        act = [[int(rng.choice([k for k in range(5) if masks[i][k]])), int(rng.integers(0, 2)) if masks[i][5] else 0] for i in range(2)]
        obs, masks, r, d = maze.step(act)
        oo, om, orr, od = o.step(np.asarray(act).reshape(4))
        assert np.array_equal(np.asarray(obs, np.float32), oo) and np.array_equal(np.asarray(masks, np.uint8), om) and r == orr and d == od
        assert maze.current_t == t % 120 + 1
        if d:
            obs, masks = maze.reset()
            oo, om = o.reset_injected(maze.engine.pool_maze(int(maze.engine.envs()[0, 3])))
            assert np.array_equal(np.asarray(obs, np.float32), oo)
    a, p = agents[0].get_action(obs[0], masks[0])
    assert masks[0][a[0]] and 0 < p <= 1
    assert "R" in maze.render_ascii() and "E" in maze.render_ascii()


def test_batched_rollout_gae_and_update(tmp_path):
    from oracle import ppo_oracle as po
    brain, agents, maze = _make(256, tmp_path, batch_size=10000, horizon=40)
    b_obs, b_act, b_logp, b_sp, ep_lens, b_masks, b_advs, b_vals = brain.get_batch()
    N = 256 * 40
    assert b_obs.shape == (N, 2, 65) and b_act.shape == (N, 2, 2) and b_logp.shape == (N,) and b_masks.dtype == torch.bool and b_advs.shape == (N,)
    # recorded actions are legal under the recorded masks, and the recorded joint log-prob is what the autograd path computes
    legal = torch.gather(b_masks[:, :, :5], 2, b_act[:, :, :1].long())
    assert legal.all() and (b_act[:, :, 1] <= b_masks[:, :, 5].float()).all()
    with torch.no_grad():
        lp = brain.get_log_probs(0, b_obs, b_act, b_masks) + brain.get_log_probs(1, b_obs, b_act, b_masks)
        lp_joint = brain.joint_log_probs(b_obs, b_act, b_masks)  # what the update evaluates: both agents in one actor pass
        assert torch.allclose(lp_joint, lp, rtol=1e-5, atol=2e-6)
        v = brain.get_state_values(b_obs)
    assert torch.allclose(lp, b_logp, rtol=1e-5, atol=2e-6) and torch.allclose(v, b_vals, rtol=1e-5, atol=2e-6)
    assert len(ep_lens) == len(b_sp) == brain.last_stats["episodes"]
    before = [p.detach().clone() for p in brain.actor.parameters()]
    stats = brain.update((b_obs, b_act, b_logp, b_sp, ep_lens, b_masks, b_advs, b_vals))
    assert stats["steps"] == 25 and np.isfinite(stats["actor_loss"]) and np.isfinite(stats["critic_loss"])
    assert any((a != b).any() for a, b in zip(before, brain.actor.parameters()))
    assert abs(brain.actor_optim.param_groups[0]["lr"] - 0.00014 * 0.997 ** 5) < 1e-12


def test_train_epoch_and_checkpoint_round_trip(tmp_path):
    brain, agents, maze = _make(64, tmp_path, batch_size=64 * 30 - 1, horizon=30)
    brain.epochs = 2
    brain.train()
    path = tmp_path / "PPO.pth"
    assert path.exists()
    sd = torch.load(path, map_location="cpu")
    assert set(sd) == {"actor", "critic", "actor_optim", "critic_optim"} and "projection.layers.22.weight" in sd["actor"] and "layers.2.bias" in sd["critic"]
    brain2, _, _ = _make(64, tmp_path, batch_size=64 * 30 - 1, horizon=30)  # loads the file in __init__ (PPO.py:31)
    for a, b in zip(brain.actor.parameters(), brain2.actor.parameters()):
        assert torch.equal(a, b)
    assert brain2.actor_optim.param_groups[0]["lr"] == brain.actor_optim.param_groups[0]["lr"]


def test_reference_checkpoint_drives_the_kernel_policy(tmp_path):
    """Config 1 flavour: the reference's shipped PPO.pth (its tensors travel as tests/golden/kat5_ppo_pth.npz) written as a checkpoint in
    the reference's format, picked up by PPO.__init__ (PPO.py:31), and driven through Agent.get_action + Maze.step on ONE maze -- and
    through the batched kernel policy (PolicyRunner) on the same observations: the kernel's logits are the checkpoint's."""
    from golden_util import GOLDEN
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.policy import PolicyRunner
    Z = np.load(os.path.join(GOLDEN, "kat5_ppo_pth.npz"))
    seed_brain = PPO(agent_amount=2, lr=0.00014, verbose=False, model_path=str(tmp_path / "PPO.pth"))
    seed_brain.actor.load_state_dict({k[6:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("actor/")})
    seed_brain.critic.load_state_dict({k[7:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("critic/")})
    seed_brain.save_parameters()
    brain, agents, maze = _make(1, tmp_path)           # loads tmp_path/PPO.pth
    assert torch.equal(brain.actor.move_head.weight.cpu(), torch.from_numpy(Z["actor/move_head.weight"]))
    obs, masks = maze.reset()
    run = PolicyRunner(brain.actor, brain.critic, 1, "cuda")
    seen = 0
    for _ in range(60):
        o_t = torch.tensor(obs, dtype=torch.float32, device="cuda").view(1, 2, 65)
        m_t = torch.tensor(masks, dtype=torch.uint8, device="cuda").view(1, 2, 6)
        logits = torch.zeros(1, 2, 6, device="cuda")
        k_act, _, k_val = run.forward(o_t, m_t, logits=logits)
        with torch.no_grad():
            mv, mk = brain.actor(o_t.view(2, 65))
            v = brain.critic(o_t)
        assert torch.allclose(logits[0, :, :5], mv, rtol=1e-5, atol=8e-6) and torch.allclose(logits[0, :, 5], mk.view(-1), rtol=1e-5, atol=8e-6)
        assert torch.allclose(k_val, v.view(-1), rtol=1e-5, atol=2e-6)
        assert all(masks[i][int(k_act[0, i, 0])] for i in range(2))      # the kernel's sampled moves are legal
        acts = [agents[i].get_action(obs[i], masks[i])[0] for i in range(2)]
        assert all(masks[i][acts[i][0]] for i in range(2))
        obs, masks, r, d = maze.step(acts)
        seen += 1
        if d:
            obs, masks = maze.reset()
    assert seen == 60


def test_reference_main_py_flow_single_env_train(tmp_path):
    """brain.train() exactly as wired in the reference's main.py, on ONE maze (num_envs=1): rollout of batch_size+1 steps, update, checkpoint."""
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent
    brain = PPO(agent_amount=2, batch_size=300, lr=0.00014, epochs=1, verbose=False, model_path=str(tmp_path / "PPO.pth"))
    agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
    maze = Maze(agents=agents, max_timestep=60, rand_sizes=True, rand_range=[4, 5], rand_start=True, difficulty=1, default_size=[4, 4])
    brain.train()
    assert brain.last_stats["env_steps"] == 301 and brain.last_stats["num_envs"] == 1 and brain.last_stats["episodes"] >= 4
    assert (tmp_path / "PPO.pth").exists() and brain.last_update["steps"] == 25


def test_cuda_graph_rollout_equals_eager_rollout(tmp_path):
    """From the second rollout on, get_batch replays one captured CUDA graph (T x 6 launches); it must produce exactly what the
    eager launches produce, rollout after rollout (the sampling stream advances through a device-side counter)."""
    outs = {}
    for mode in (False, True):
        brain, agents, maze = _make(128, tmp_path / f"g{int(mode)}", batch_size=128 * 24 - 2, horizon=24, use_cuda_graph=mode)
        res = []
        for _ in range(3):
            b = brain.get_batch()
            res.append([b[0].clone(), b[1].clone(), b[2].clone(), b[5].clone(), b[6].clone(), b[7].clone()])
        outs[mode] = res
        assert int(brain._policy().counter_dev.item()) == 3 * 24
        assert (brain._graph is not None) == mode
    for r_e, r_g in zip(outs[False], outs[True]):
        for x, y in zip(r_e, r_g):
            assert torch.equal(x, y)
    assert not torch.equal(outs[True][1][1], outs[True][2][1])  # different rollouts draw different actions


def test_prefetched_pool_equals_inline_generation(tmp_path):
    """get_batch starts carving the NEXT rollout's mazes on a side stream (Maze.prefetch_pool, in place: the consumed slots) and the next refill
    only waits for it; the rollouts must be exactly those of refilling the pool inline."""
    outs = {}
    for mode in (False, True):
        brain, agents, maze = _make(128, tmp_path / f"p{int(mode)}", batch_size=128 * 24 - 2, horizon=24, prefetch_pool=mode)
        res = []
        for _ in range(3):
            b = brain.get_batch()
            res.append([b[0].clone(), b[1].clone(), b[2].clone(), b[5].clone(), b[6].clone(), b[7].clone(), torch.as_tensor(b[3].astype("int64"))])
        outs[mode] = res
        assert (getattr(maze, "_staged", None) is not None) == mode
    for r_a, r_b in zip(outs[False], outs[True]):
        for x, y in zip(r_a, r_b):
            assert torch.equal(x, y)   # observations, actions, log-probs, masks, advantages, values, shortest paths of the finished episodes
    assert not torch.equal(outs[True][0][0], outs[True][1][0])  # every rollout runs on new mazes


@pytest.mark.parametrize("faithful", [True, False], ids=["column0_projection", "indexed_projection"])
def test_graphed_update_equals_eager_update(tmp_path, faithful):
    """the optimiser steps of PPO.update (PPO.py:51-85) replayed as CUDA graphs -- one graph per minibatch position, captured in the
    first update and replayed in every later epoch and update -- against the same steps launched eagerly with the same (capturable) Adam: identical
    parameters after two rollouts + updates, same losses, and the graphs were really used."""
    outs = {}
    for mode in (False, True):
        brain, agents, maze = _make(256, tmp_path / f"g{int(mode)}", batch_size=256 * 40 // 5 * 5, horizon=40, faithful_projection=faithful, graph_update=mode)
        stats = []
        for _ in range(2):
            stats.append(brain.update(brain.get_batch()))
        assert all(st["graphed"] == mode for st in stats) and stats[0]["steps"] == 25
        if mode:
            assert len(brain._ug["graphs"]) == 5      # five minibatch positions, each captured once and replayed 9 or 10 times
        outs[mode] = ([p.detach().clone() for p in list(brain.actor.parameters()) + list(brain.critic.parameters())], stats)
    for a, b in zip(outs[False][0], outs[True][0]):
        assert torch.equal(a, b)
    for sa, sb in zip(outs[False][1], outs[True][1]):
        assert sa["actor_loss"] == sb["actor_loss"] and sa["critic_loss"] == sb["critic_loss"]


def test_agent_reset_and_move_setters_match_oracle(tmp_path):
    """Agent.reset(x, y) / Agent.move(x, y, direction) (maze_agent.py:59-87) and Maze.build_maze() on the device state: after the same calls on the
    oracle, the next steps' observations, masks and agent state agree."""
    from oracle import OracleMaze
    brain, agents, maze = _make(1, tmp_path)
    maze.build_maze()                                  # == reset() without the return value
    m = maze._pool()
    o = OracleMaze(max_timestep=120)
    o.reset_injected(dict(layout=m["layout"], path0=m["path0"], path1=m["path1"], end=m["end"], key=m["key"], shortest_path_len=m["shortest_path_len"]))
    path = maze.shortest_path
    stop = [[4, 0], [4, 0]]
    for k in range(3):
        maze.step(stop); o.step([4, 0, 4, 0])
    (x, y), (x2, y2) = path[min(4, len(path) - 1)], path[min(6, len(path) - 1)]
    agents[0].move(x, y, 1); o.agent_move(0, x, y, 1)
    agents[1].reset(x2, y2); o.agent_reset(1, x2, y2)
    assert (agents[0].x, agents[0].y, agents[0].direction) == (x, y, 1) and (agents[1].x, agents[1].y, agents[1].direction) == (x2, y2, 2)
    for k in range(4):
        obs, masks, r, d = maze.step(stop)
        oo, om, orr, od = o.step([4, 0, 4, 0])
        assert np.array_equal(np.asarray(obs, np.float32).view(np.uint32), oo.view(np.uint32)), k
        assert np.array_equal(np.asarray(masks, np.uint8), om) and r == orr and d == od
        assert np.array_equal(maze.engine.agents()[0][:, :17], o.agents()[:, :17])   # (column 17, len(exit_route), is a stack in the reference: move() leaves it stale)


def test_incremental_refill_rebuilds_consumed_slots_only(tmp_path):
    """The reference builds one maze per reset (maze.py:57).  Between two rollouts the pool refill rebuilds exactly the slots (env e, episode k) an
    episode started on -- k < env_episode[e] -- with the new refill's seed; the other slots keep their (unseen) mazes."""
    brain, agents, maze = _make(64, tmp_path / "inc", batch_size=64 * 130 - 2, horizon=130)   # max_timestep 120 < horizon: every env consumes >= 2 slots
    brain.get_batch()
    eng = maze.engine
    E, K = eng.E, eng.P // eng.E
    hdr0 = eng.pool_hdr.view(torch.int32).view(K, E, 4).clone()
    grid0 = eng.pool_grid.view(K, E, -1).clone()
    used = eng.env_episode.view(torch.int32).clone()              # episodes started per env in that rollout
    assert int(used.min()) >= 2 and int(used.max()) <= K
    brain.get_batch()                                             # refills first
    hdr1 = eng.pool_hdr.view(torch.int32).view(K, E, 4)
    grid1 = eng.pool_grid.view(K, E, -1)
    consumed = torch.arange(K, device=used.device).view(K, 1) < used.view(1, E)
    same_grid = (grid0 == grid1).all(-1)
    assert bool(same_grid[~consumed].all()) and bool((hdr0 == hdr1).all(-1)[~consumed].all())      # untouched slots: identical mazes
    assert float(same_grid[consumed].float().mean()) < 0.05                                         # consumed slots: new mazes (same id, new seed)
    assert bool((hdr0[..., 3] == hdr1[..., 3]).all())                                               # the slot's maze id is its (env, episode) key


@pytest.mark.parametrize("faithful", [True, False], ids=["column0_projection", "indexed_projection"])
def test_fused_update_tracks_autograd_update(tmp_path, faithful):
    """PPO.update with the K5 kernels (fused_update=True, default) against the same schedule on PyTorch autograd: same rollout, same
    minibatch permutation, two optimiser epochs.  Losses agree to 1e-5 and the actors move together.  Gradient parity proper is
    tests/test_update_gpu.py; here Adam sits in between, and Adam's first steps are lr * sign(g): an element whose gradient is zero up to
    rounding (or carries a ReLU-gate flip, see that file) steps +-lr in either implementation -- so the check is the direction of the
    whole update and the share of far-apart elements, not their size."""
    import copy
    a, _, maze_a = _make(256, tmp_path / "a", batch_size=256 * 40 // 5 * 5, horizon=40, faithful_projection=faithful, use_cuda_graph=False)
    b, _, _ = _make(256, tmp_path / "b", batch_size=256 * 40 // 5 * 5, horizon=40, faithful_projection=faithful, use_cuda_graph=False, fused_update=False)
    b.actor.load_state_dict(a.actor.state_dict()); b.critic.load_state_dict(a.critic.state_dict())
    b.actor_optim = torch.optim.Adam(b.actor.parameters(), lr=0.00014); b.critic_optim = torch.optim.Adam(b.critic.parameters(), lr=0.00014)
    a.actor_optim = torch.optim.Adam(a.actor.parameters(), lr=0.00014); a.critic_optim = torch.optim.Adam(a.critic.parameters(), lr=0.00014)
    batch = a.get_batch()
    b._rollouts = a._rollouts
    a.updates_per_batch = b.updates_per_batch = 2
    before = copy.deepcopy(a.actor.state_dict())
    sa, sb = a.update(batch), b.update(batch)
    assert sa["steps"] == sb["steps"] == 10
    assert abs(sa["actor_loss"] - sb["actor_loss"]) < 1e-5 * max(1.0, abs(sb["actor_loss"])) + 1e-7
    assert abs(sa["critic_loss"] - sb["critic_loss"]) < 1e-4 * abs(sb["critic_loss"]) + 1e-7
    moved = moved_ref = far = total = 0
    da, db = [], []
    for (name, p), (_, q) in zip(a.actor.state_dict().items(), b.actor.state_dict().items()):
        moved += int((p != before[name]).sum()); moved_ref += int((q != before[name]).sum()); total += p.numel()
        far += int(((p - q).abs() > 2e-5).sum())
        da.append((p - before[name]).flatten().double()); db.append((q - before[name]).flatten().double())
        assert bool((p != before[name]).any()), name    # the fused path trained every block of the actor
    assert moved > 0.5 * total and abs(moved - moved_ref) < 0.01 * total, (moved, moved_ref, total)   # (dead ReLU units never move, in either path)
    da, db = torch.cat(da), torch.cat(db)
    cos = float((da @ db) / (da.norm() * db.norm()))
    assert cos > (0.999 if faithful else 0.98), cos
    assert far < (2e-3 if faithful else 0.15) * total, (far, total)


def test_get_action_kernel_path_equals_the_module_path(tmp_path):
    """PPO.get_action on a GPU asks the K4 kernels for the six logits of ONE observation (PPO._get_action_kernel) and draws on the host like the
    reference (PPO.py:170-186).  Same torch seed => the same action and the same log-prob as the autograd modules followed by the same draw; a changed
    actor (in-place optimiser-style step, load_state_dict) is picked up by the next call."""
    brain, agents, maze = _make(1, tmp_path)
    obs, masks = maze.reset()
    rng = np.random.default_rng(1)

    def module_path(o, m, seed):
        with torch.no_grad():
            mv, mk = brain.actor(torch.tensor(o, dtype=torch.float32, device="cuda"))
        torch.manual_seed(seed)
        return brain._sample_action(mv.cpu(), mk.cpu(), m), (mv.cpu(), mk.cpu())

    hits = 0

    def check(n, seed0):
        nonlocal obs, masks, hits
        for k in range(n):
            i = k % 2
            (a_ref, lp_ref), (mv, mk) = module_path(obs[i], masks[i], seed0 + k)
            torch.manual_seed(seed0 + k)
            a, lp = brain.get_action(obs[i], masks[i])
            lg = brain._act1["last_logits"]
            hits += brain._act1["cache"] is not None and obs[i] in brain._act1["cache"][0]
            assert torch.allclose(lg[:5], mv.view(-1), rtol=1e-5, atol=8e-6) and torch.allclose(lg[5:], mk.view(-1), rtol=1e-5, atol=8e-6)
            assert a == a_ref and masks[i][a[0]] and torch.allclose(lp, lp_ref, rtol=1e-5, atol=2e-6)
            act = [[int(rng.choice([j for j in range(5) if masks[q][j]])), int(rng.integers(0, 2)) if masks[q][5] else 0] for q in range(2)]
            obs, masks, r, d = maze.step(act)
            if d:
                obs, masks = maze.reset()

    check(30, 100)
    assert hits >= 25 and maze._prefetch_policy   # after the first call the logits arrive with the observation (PPO._prefetch_logits)
    with torch.no_grad():   # an in-place step on every parameter, as an optimiser does
        for p_ in brain.actor.parameters():
            p_.add_(0.01 * torch.randn_like(p_))
    check(10, 200)
    brain._weights_version = getattr(brain, "_weights_version", 0)  # attribute exists / is an int
    sd = {k: v + 0.01 for k, v in brain.actor.state_dict().items()}
    brain.actor.load_state_dict(sd)
    check(10, 300)


def test_full_size_config3_rollout_properties_and_sampled_parity(tmp_path):
    """BASELINE config[2] at its FULL size: 65 536 mazes x 2 agents, T = 128 (8.4 M env-steps), main.py's maze settings.  Size-independent properties of the
    whole rollout (every recorded action legal under its recorded mask, finite log-probs / values / advantages, episode bookkeeping consistent) and, on
    4 096 sampled (t, env) rows, the kernel policy against the autograd modules (joint log-prob and value, 1e-5 relative) plus the GAE of 64 sampled
    environments against the oracle's restatement of PPO.get_GAEs bootstrapped at the horizon."""
    from oracle import ppo_oracle as po
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent
    E, T = 65536, 128
    brain = PPO(agent_amount=2, batch_size=E * T - 5, lr=0.00014, epochs=1, verbose=False, model_path=None, horizon=T)
    agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
    Maze(agents=agents, max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True, num_envs=E, seed=1)
    for rollout in range(2):        # the second one is the captured CUDA graph
        b_obs, b_act, b_logp, b_sp, ep_lens, b_masks, b_advs, b_vals = brain.get_batch()
        N = E * T
        assert b_obs.shape == (N, 2, 65) and b_logp.shape == (N,)
        legal = torch.gather(b_masks[:, :, :5], 2, b_act[:, :, :1].long())
        assert bool(legal.all()) and bool((b_act[:, :, 1] <= b_masks[:, :, 5].float()).all())
        assert bool(torch.isfinite(b_logp).all()) and bool(torch.isfinite(b_vals).all()) and bool(torch.isfinite(b_advs).all()) and bool((b_logp <= 0).all())
        assert brain.last_stats["env_steps"] == N and len(ep_lens) == len(b_sp) == brain.last_stats["episodes"]
        g = torch.Generator(device="cuda").manual_seed(rollout)
        rows = torch.randint(0, N, (4096,), device="cuda", generator=g)
        with torch.no_grad():
            lp = brain.joint_log_probs(b_obs[rows], b_act[rows], b_masks[rows])
            v = brain.get_state_values(b_obs[rows])
        assert torch.allclose(lp, b_logp[rows], rtol=1e-5, atol=2e-6) and torch.allclose(v, b_vals[rows], rtol=1e-5, atol=2e-6)
        buf = brain._buf
        envs = torch.randint(0, E, (64,), generator=torch.Generator().manual_seed(rollout)).tolist()
        r, vals, d, adv = (buf[k].cpu().numpy() for k in ("reward", "values", "done", "adv"))
        want = po.gae_fixed_horizon(r[:, envs], vals[:T][:, envs], d[:, envs].astype(bool), vals[T][envs], brain.discount_rate, brain.lam)
        assert np.array_equal(adv[:, envs].view(np.uint32), want.view(np.uint32)), "GAE of the sampled envs differs from the oracle"
