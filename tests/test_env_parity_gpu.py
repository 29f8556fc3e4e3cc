"""CUDA step+obs kernel (K2) and reset against the golden reference traces and against the C oracle.

Everything here goes through the C ABI (marl_maze_b200.MazeEngine -> libmarl_maze_b200.so).  Bar: bit-exact --
observations (fp32 bit patterns), masks, rewards, dones and the full unpacked agent/env state.
"""
import numpy as np
import pytest
import torch

from golden_util import load_traces, load_vision_traces
from oracle import OracleBatch, OracleMaze

pytestmark = pytest.mark.gpu
TRACES = load_traces()
VTRACES = load_vision_traces()
ALL_TRACES = {**TRACES, **VTRACES}


def _engine(*a, **k):
    from marl_maze_b200 import MazeEngine
    return MazeEngine(*a, **k)


def _assert_obs_equal(got, want, ctx):
    got = np.asarray(got, np.float32); want = np.asarray(want, np.float32)
    if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
        bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
        i = tuple(bad[0])
        raise AssertionError(f"{ctx}: obs differ at {bad[:8].tolist()} got {got[i]} want {want[i]} ({len(bad)} entries)")


@pytest.mark.parametrize("name", sorted(ALL_TRACES))
def test_golden_trace_explicit_reset(name):
    """Maze.reset()/Maze.step() semantics, one env, terminal observation returned and reset called by the host.  The v* traces were recorded
    from the reference with Agent(..., vision_range != 4) per agent."""
    tr = ALL_TRACES[name]
    S = max(max(m["width"], m["height"]) for m in tr.mazes)
    eng = _engine(1, smax=S, max_timestep=tr.max_timestep, pool_size=len(tr.mazes), vision=tr.vision)
    eng.load_layouts(0, tr.mazes)
    n = tr.n
    dev = eng.device
    acts = torch.from_numpy(tr.actions.reshape(n, 1, 2, 2).copy()).to(dev)
    obs = torch.zeros(n, 1, 2, 65, device=dev); masks = torch.zeros(n, 1, 2, 6, dtype=torch.uint8, device=dev)
    rew = torch.zeros(n, 1, device=dev); done = torch.zeros(n, 1, dtype=torch.uint8, device=dev)
    robs = torch.zeros(len(tr.mazes), 1, 2, 65, device=dev); rmasks = torch.zeros(len(tr.mazes), 1, 2, 6, dtype=torch.uint8, device=dev)
    agents = []
    ep = 0
    eng.reset(obs=robs[0], masks=rmasks[0])
    ragents = [eng.agents()[0]]
    for i in range(n):
        eng.step(acts[i], auto_reset=False, obs=obs[i], masks=masks[i], reward=rew[i], done=done[i])
        agents.append(eng.agents()[0])
        if tr.done[i]:  # host-driven reset, like PPO.get_batch (PPO.py:127-130); golden `done` == device `done` is asserted below
            ep += 1
            eng.reset(obs=robs[ep], masks=rmasks[ep])
            ragents.append(eng.agents()[0])
    torch.cuda.synchronize()
    assert np.array_equal(done.cpu().numpy()[:, 0], tr.done), name
    assert np.array_equal(rew.cpu().numpy()[:, 0], tr.reward), name
    _assert_obs_equal(obs.cpu().numpy()[:, 0], tr.step_obs, name + " step")
    assert np.array_equal(masks.cpu().numpy()[:, 0], tr.step_masks), name
    _assert_obs_equal(robs.cpu().numpy()[:, 0], tr.reset_obs, name + " reset")
    assert np.array_equal(rmasks.cpu().numpy()[:, 0], tr.reset_masks), name
    agents = np.asarray(agents)
    bad = np.argwhere(agents != tr.agents_after)
    assert len(bad) == 0, (name, "agent state", bad[:5].tolist(), agents[tuple(bad[0][:2])], tr.agents_after[tuple(bad[0][:2])])
    nre = min(len(ragents), len(tr.reset_agents))
    assert np.array_equal(np.asarray(ragents)[:nre - 1], tr.reset_agents[:nre - 1]), name
    assert eng.envs()[0, 6] == 0  # no illegal-move flag


@pytest.mark.parametrize("name", ["guided_a", "tiny7", "mixed_sizes_d3", "s49_guided", "v33_guided", "v13_slow"])
def test_golden_trace_auto_reset(name):
    """Same traces with the in-launch auto-reset: at `done` the emitted obs/masks are those of the next episode."""
    tr = ALL_TRACES[name]
    S = max(max(m["width"], m["height"]) for m in tr.mazes)
    eng = _engine(1, smax=S, max_timestep=tr.max_timestep, pool_size=len(tr.mazes), vision=tr.vision)
    eng.load_layouts(0, tr.mazes)
    n = tr.n
    dev = eng.device
    acts = torch.from_numpy(tr.actions.reshape(n, 1, 2, 2).copy()).to(dev)
    obs = torch.zeros(n, 1, 2, 65, device=dev); masks = torch.zeros(n, 1, 2, 6, dtype=torch.uint8, device=dev)
    rew = torch.zeros(n, 1, device=dev); done = torch.zeros(n, 1, dtype=torch.uint8, device=dev)
    eng.reset()
    for i in range(n):
        eng.step(acts[i], auto_reset=True, obs=obs[i], masks=masks[i], reward=rew[i], done=done[i])
    torch.cuda.synchronize()
    want_obs = tr.step_obs.copy(); want_masks = tr.step_masks.copy()
    ep = 0
    for i in range(n):
        if tr.done[i]:
            ep += 1
            want_obs[i] = tr.reset_obs[ep]; want_masks[i] = tr.reset_masks[ep]
    assert np.array_equal(done.cpu().numpy()[:, 0], tr.done)
    assert np.array_equal(rew.cpu().numpy()[:, 0], tr.reward)
    _assert_obs_equal(obs.cpu().numpy()[:, 0], want_obs, name)
    assert np.array_equal(masks.cpu().numpy()[:, 0], want_masks)


def _oracle_pool(E, K, cfg, seed0=1000):
    gen = OracleMaze(max_timestep=10, **cfg)
    mazes = []
    for p in range(E * K):
        gen.seed(seed0 + p); gen.build(); mazes.append(gen.maze())
    return mazes


def _run_vs_oracle(E, K, T, max_t, cfg, p_follow, p_mark, auto_reset=True, check_state_every=97, fused_actions=False, vision=(4, 4)):
    mazes = _oracle_pool(E, K, cfg)
    S = max(max(m["width"], m["height"]) for m in mazes)
    ob = OracleBatch(E, E * K, max_timestep=max_t, threads=8, vision=vision)
    for p, m in enumerate(mazes):
        ob.set_pool_maze(p, m)
    eng = _engine(E, smax=S, max_timestep=max_t, pool_size=E * K, vision=vision)
    eng.load_layouts(0, mazes)
    o_obs, o_masks = ob.reset_all()
    g_obs, g_masks = eng.reset()
    _assert_obs_equal(g_obs.cpu().numpy(), o_obs, "reset")
    assert np.array_equal(g_masks.cpu().numpy(), o_masks)
    rng = (np.arange(E, dtype=np.uint64) + 1) * np.uint64(0x9E3779B97F4A7C15)
    tot_r = 0.0; tot_d = 0
    act_out = torch.zeros(E, 2, 2, dtype=torch.uint8, device=eng.device)
    for t in range(T):
        if fused_actions:
            g_obs, g_masks, g_r, g_d = eng.step(None, auto_reset=auto_reset, action_seed=7, actions_out=act_out)
            act = act_out.cpu().numpy()
            legal = np.take_along_axis(o_masks[:, :, :5], act[:, :, :1].astype(np.int64), axis=2)[..., 0]
            assert legal.all(), "kernel-sampled move is not mask-legal"
            assert (act[:, :, 1] <= o_masks[:, :, 5]).all(), "kernel-sampled mark is not mask-legal"
        else:
            act = ob.guided_actions(rng, p_follow=p_follow, p_mark=p_mark)
            g_obs, g_masks, g_r, g_d = eng.step(torch.from_numpy(act).to(eng.device), auto_reset=auto_reset)
        o_obs, o_masks, o_r, o_d = ob.step(act, auto_reset=auto_reset)
        assert np.array_equal(g_d.cpu().numpy(), o_d), f"done differs at step {t}"
        assert np.array_equal(g_r.cpu().numpy(), o_r), f"reward differs at step {t}"
        _assert_obs_equal(g_obs.cpu().numpy(), o_obs, f"step {t}")
        assert np.array_equal(g_masks.cpu().numpy(), o_masks), f"masks differ at step {t}"
        if not auto_reset and o_d.any():
            o_obs, o_masks = ob.reset_masked(o_d)
            g_obs, g_masks = eng.reset(torch.from_numpy(o_d.copy()))
            _assert_obs_equal(g_obs.cpu().numpy(), o_obs, f"masked reset {t}")
            assert np.array_equal(g_masks.cpu().numpy(), o_masks)
        tot_r += float(o_r.sum()); tot_d += int(o_d.sum())
        if t % check_state_every == 0 or t == T - 1:
            assert np.array_equal(eng.agents(), ob.agents()), f"agent state differs at step {t}"
            ge, oe = eng.envs(), ob.env_state()
            assert np.array_equal(ge[:, :4], oe), f"env state differs at step {t}"
            for e in (0, E // 2, E - 1):
                m = ob.env(e).maze()
                assert np.array_equal(eng.layout(e)[:m["height"], :m["width"]], m["layout"]), f"layout (marks) differs at step {t} env {e}"
    assert ob.errors() == 0 and int(eng.envs()[:, 6].sum()) == 0
    return tot_r, tot_d


MAIN = dict(rand_sizes=True, rand_range=(12, 13), rand_start=True, difficulty=1, default_size=(4, 4))  # main.py:20


def test_config2_4096_envs_bit_exact():
    """BASELINE config 2: 4096 mazes x 2 agents, S in {23,25}, >= 1200 steps so truncation, key pickup, marks, route sharing
    and >= 1 reset per env all occur; every field compared every step."""
    r, d = _run_vs_oracle(E=4096, K=12, T=1300, max_t=1200, cfg=MAIN, p_follow=0.8, p_mark=0.3)
    assert d >= 4096 and r > 4096  # every env finished at least once; key pickups and joint exits happened


def test_uniform_random_truncation():
    r, d = _run_vs_oracle(E=512, K=3, T=450, max_t=200, cfg=MAIN, p_follow=0.0, p_mark=0.5)
    assert d >= 2 * 512


def test_fused_action_sampler_is_legal_and_replays():
    r, d = _run_vs_oracle(E=1024, K=4, T=300, max_t=120, cfg=MAIN, p_follow=0, p_mark=0, fused_actions=True)
    assert d >= 2 * 1024


@pytest.mark.parametrize("vision", [(3, 3), (2, 4), (1, 2)], ids=lambda v: f"vision{v[0]}{v[1]}")
def test_vision_range_batch_bit_exact(vision):
    """per-agent vision_range != 4 over a batch, every field every step against the oracle (itself pinned to the reference's
    vision_range traces by tests/test_oracle_golden.py): truncations, key pickups, route sharing and resets included."""
    r, d = _run_vs_oracle(E=1024, K=6, T=500, max_t=200, cfg=MAIN, p_follow=0.8, p_mark=0.3, vision=vision)
    assert d >= 2 * 1024 and r > 100


def test_explicit_masked_reset_matches():
    _run_vs_oracle(E=300, K=8, T=400, max_t=150, cfg=MAIN, p_follow=0.9, p_mark=0.2, auto_reset=False)


def test_s49_and_ragged_env_count():
    """4x default area (S=49), E not a multiple of the 16 envs a warp owns."""
    cfg = dict(rand_sizes=True, rand_range=(25, 25), rand_start=True, difficulty=2, default_size=(4, 4))
    _run_vs_oracle(E=1000 + 7, K=3, T=500, max_t=400, cfg=cfg, p_follow=0.85, p_mark=0.4)


def test_tiny_fixed_start_and_single_env():
    cfg = dict(rand_sizes=False, rand_start=False, difficulty=1, default_size=(4, 4))
    _run_vs_oracle(E=1, K=40, T=600, max_t=60, cfg=cfg, p_follow=0.7, p_mark=0.3)
    _run_vs_oracle(E=33, K=30, T=400, max_t=60, cfg=cfg, p_follow=0.5, p_mark=0.6)


def test_world_size_independence():
    """Sharding envs over ranks changes nothing: envs [lo,hi) of a big engine == a small engine holding the same pool slice."""
    E, K = 256, 1
    mazes = _oracle_pool(E, K, MAIN)
    S = 25
    big = _engine(E, smax=S, max_timestep=100, pool_size=E); big.load_layouts(0, mazes); big.reset()
    lo, hi = 64, 192
    small = _engine(hi - lo, smax=S, max_timestep=100, pool_size=hi - lo, env_offset=lo); small.load_layouts(0, mazes[lo:hi]); small.reset()
    ao_b = torch.zeros(E, 2, 2, dtype=torch.uint8, device=big.device); ao_s = torch.zeros(hi - lo, 2, 2, dtype=torch.uint8, device=big.device)
    for t in range(150):
        ob_, mb, rb, db = big.step(None, action_seed=3, actions_out=ao_b)
        os_, ms, rs, ds = small.step(None, action_seed=3, actions_out=ao_s)
        assert torch.equal(ao_b[lo:hi], ao_s) and torch.equal(ob_[lo:hi], os_) and torch.equal(mb[lo:hi], ms) and torch.equal(db[lo:hi], ds)


def test_reciprocal_division_is_ieee_exact_exhaustively():
    """K2's ratio features use rcp + 2 FMA instead of div.rn; every quotient a/b the observation can form must match div.rn."""
    import ctypes as C
    from marl_maze_b200 import _abi
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    _abi.check(_abi.lib().mm_selftest_div(4096, 4096, C.c_void_p(cnt.data_ptr()), None), "mm_selftest_div")
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0


def test_rectangular_mazes_injected():
    """Maze(default_size=[w, h]) with w != h (maze.py:26-27): the step kernel keeps W and H apart; layouts come from the oracle."""
    for ds in [(4, 9), (12, 5)]:
        cfg = dict(rand_sizes=False, rand_start=True, difficulty=2, default_size=ds)
        _run_vs_oracle(E=96, K=10, T=260, max_t=90, cfg=cfg, p_follow=0.8, p_mark=0.4)


def test_illegal_actions_follow_the_stated_convention():
    """SURVEY H3 / VERDICT r1 weak #1a: what happens on input the reference's masks forbid, through the C ABI, against the oracle.
      * a move that would leave the grid (maze.py:141-145 prints, then indexes out of range): `stop` + the env's error flag;
      * a move code above 4 (the reference would IndexError in get_memory): `stop` + error flag;
      * an in-bounds move INTO A WALL: reproduced like the reference, which has no wall check (the agent stands in the wall, marks
        overwrite wall cells, observations follow) -- bit-exact against the oracle -- and the error flag is raised.
    Marks ordered with an illegal move are still applied (maze.py:132-134 runs before the move).  Envs fed legal actions keep err = 0."""
    E, max_t = 512, 400
    cfg = dict(difficulty=1, rand_start=True, rand_sizes=True, rand_range=(12, 13), default_size=(4, 4))
    mazes = _oracle_pool(E, 1, cfg, seed0=4000)
    S = max(max(m["width"], m["height"]) for m in mazes)
    ob = OracleBatch(E, E, max_timestep=max_t, threads=8)
    for p, m in enumerate(mazes):
        ob.set_pool_maze(p, m)
    eng = _engine(E, smax=S, max_timestep=max_t, pool_size=E)
    eng.load_layouts(0, mazes)
    g_obs, g_masks = eng.reset(); o_obs, o_masks = ob.reset_all()
    _assert_obs_equal(g_obs.cpu().numpy(), o_obs, "reset")
    rng = np.random.default_rng(17)
    rs = np.arange(1, E + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    DX, DY = [0, 1, 0, -1], [-1, 0, 1, 0]
    kind = np.zeros(E, np.int32)            # 0 legal, 1 out of bounds, 2 code > 4, 3 into a wall
    for t in range(6):                      # a few legal steps first so that agents have left their start cells
        act = ob.random_actions(rs)
        eng.step(torch.from_numpy(act).cuda(), auto_reset=False); ob.step(act, auto_reset=False)
    ag = ob.agents(); act = ob.random_actions(rs)
    for e in range(E):
        m = mazes[e]; lay = ob.layout(e); W, H = m["width"], m["height"]
        if ag[e, :, 3].any():               # somebody already knows the exit: keep this env legal (route bookkeeping is defined on open cells only)
            continue
        x, y, d = int(ag[e, 0, 0]), int(ag[e, 0, 1]), int(ag[e, 0, 2])
        choice = e % 4
        if choice == 1:                     # out of bounds, if agent 0 stands on the border
            for mv in range(4):
                nd = (mv + d) % 4; nx, ny = x + DX[nd], y + DY[nd]
                if not (0 <= nx < W and 0 <= ny < H):
                    act[e, 0] = [mv, 1]; kind[e] = 1; break
        elif choice == 2:
            act[e, 0] = [5 + int(rng.integers(0, 200)), int(rng.integers(0, 2))]; kind[e] = 2
        elif choice == 3:                   # in-bounds wall
            for mv in range(4):
                nd = (mv + d) % 4; nx, ny = x + DX[nd], y + DY[nd]
                if 0 <= nx < W and 0 <= ny < H and lay[ny, nx] == 1:
                    act[e, 0] = [mv, int(rng.integers(0, 2))]; kind[e] = 3; break
    assert (kind == 1).sum() > 5 and (kind == 2).sum() > 50 and (kind == 3).sum() > 50, np.bincount(kind)
    g = eng.step(torch.from_numpy(act).cuda(), auto_reset=False)
    o = ob.step(act, auto_reset=False)
    ga, oa = eng.agents(), ob.agents()
    # an agent may SIGHT the exit from inside a wall; from then on its route bookkeeping is outside R2's domain: compare the others
    keep = ~(oa[:, :, 3].any(1) & (kind == 3))
    assert keep.sum() > E - 40
    _assert_obs_equal(g[0].cpu().numpy()[keep], o[0][keep], "illegal step")
    assert np.array_equal(g[1].cpu().numpy()[keep], o[1][keep]) and np.array_equal(g[2].cpu().numpy(), o[2]) and np.array_equal(g[3].cpu().numpy(), o[3])
    assert np.array_equal(ga[keep], oa[keep]), np.argwhere(ga != oa)[:5]
    err = eng.envs()[:, 6]
    assert np.array_equal(err != 0, kind != 0), (np.bincount(kind), np.bincount(kind[err != 0], minlength=4))
    # stop semantics: position, facing and move memory of agent 0 unchanged by kinds 1 and 2; kind 3 really stands in the wall
    for e in np.flatnonzero((kind == 1) | (kind == 2)):
        assert tuple(ga[e, 0, :3]) == tuple(ag[e, 0, :3])
    for e in np.flatnonzero(kind == 3)[:50]:
        lay = mazes[e]["layout"]
        assert lay[ga[e, 0, 1], ga[e, 0, 0]] == 1
    # the wall-walkers keep following the reference afterwards (three more steps of whatever the masks now allow), flags stay raised
    for t in range(3):
        act = ob.random_actions(rs)
        g = eng.step(torch.from_numpy(act).cuda(), auto_reset=False); o = ob.step(act, auto_reset=False)
        keep = ~ob.agents()[:, :, 3].any(1)      # compare where nobody has learnt the exit meanwhile
        _assert_obs_equal(g[0].cpu().numpy()[keep], o[0][keep], f"after illegal step +{t}")
        assert np.array_equal(g[1].cpu().numpy()[keep], o[1][keep])
    assert np.array_equal(eng.envs()[:, 6] != 0, kind != 0)


def test_full_size_config4_sampled_envs_bit_exact_and_global_invariants():
    """BASELINE config[3] at its FULL size -- 1 Mi mazes of side 49 (4x the default area), two pool mazes per env built by K1, uniform mask-legal actions
    sampled in the kernel, auto-reset on -- checked two ways: (1) 1 024 environments spread over the whole batch (first / last warps, block boundaries,
    random ones) are replayed by the oracle on the SAME mazes (generated by the oracle itself from the maze keys) with the recorded actions: observations
    (bit patterns), masks, rewards, dones and the unpacked agent state must be identical at every one of 160 steps, resets included (max_timestep 64);
    (2) size-independent properties of the whole batch: no error flag, every sampled action legal under the previous masks, every observation finite with a
    one-hot facing and a one-hot agent id, rewards in {0, 0.5, 1}, a done either a joint exit (reward 1) or the truncation."""
    E, S, max_t, T, n_s = 1 << 20, 49, 64, 160, 1024
    seed, base = 4242, 17
    eng = _engine(E, smax=S, max_timestep=max_t, pool_size=2 * E)
    eng.generate(seed, side_range=(25, 25), rand_start=True, difficulty=1, id_base=base)
    rng = np.random.default_rng(0)
    pick = np.unique(np.concatenate([np.arange(0, 40), np.arange(E - 40, E), np.arange(63, 66), np.arange(4095, 4098), rng.integers(0, E, n_s)]))[:n_s]
    n_s = len(pick)
    ob = OracleBatch(n_s, 2 * n_s, max_timestep=max_t, threads=8)
    o = OracleMaze(max_timestep=10, difficulty=1, rand_start=True, rand_sizes=True, rand_range=(25, 25), default_size=(4, 4))
    for k in range(2):
        for i, e in enumerate(pick):
            o.seed_philox(seed, base + int(e) + k * E); o.build(); ob.set_pool_maze(i + k * n_s, o.maze())
    idx = torch.from_numpy(pick).to(eng.device)
    o_obs, o_masks = ob.reset_all()
    g_obs, g_masks = eng.reset()
    _assert_obs_equal(g_obs[idx].cpu().numpy(), o_obs, "reset")
    assert np.array_equal(g_masks[idx].cpu().numpy(), o_masks)
    act_out = torch.zeros(E, 2, 2, dtype=torch.uint8, device=eng.device)
    prev_masks = g_masks.clone()
    n_done = n_exit = 0
    for t in range(T):
        g_obs, g_masks, g_r, g_d = eng.step(None, auto_reset=True, action_seed=11, actions_out=act_out)
        # (2) whole-batch properties
        mv = act_out[:, :, 0].long()
        assert bool(torch.gather(prev_masks[:, :, :5], 2, mv.unsqueeze(-1)).all()) and bool((act_out[:, :, 1] <= prev_masks[:, :, 5]).all()), t
        assert bool(torch.isfinite(g_obs).all()) and bool((g_obs[:, :, 0:4].sum(-1) == 1).all()) and bool((g_obs[:, :, 63:65].sum(-1) == 1).all()), t
        assert bool(((g_r == 0) | (g_r == 0.5) | (g_r == 1)).all()), t
        # a finished env either solved the maze (reward 1) or ran into the truncation; the oracle comparison below pins which
        assert bool((g_d.bool() | (g_r != 1)).all()), t
        n_done += int(g_d.sum()); n_exit += int((g_r == 1).sum())
        prev_masks.copy_(g_masks)
        # (1) sampled envs against the oracle
        act = act_out[idx].cpu().numpy()
        o_obs, o_masks, o_r, o_d = ob.step(act, auto_reset=True)
        assert np.array_equal(g_d[idx].cpu().numpy(), o_d), f"done differs at step {t}"
        assert np.array_equal(g_r[idx].cpu().numpy(), o_r), f"reward differs at step {t}"
        _assert_obs_equal(g_obs[idx].cpu().numpy(), o_obs, f"step {t}")
        assert np.array_equal(g_masks[idx].cpu().numpy(), o_masks), f"masks differ at step {t}"
    assert np.array_equal(eng.agents()[pick], ob.agents())
    assert int(eng.envs()[:, 6].sum()) == 0 and ob.errors() == 0
    assert n_done >= 2 * E          # every env was truncated at least twice in 160 steps of max_timestep 64
