"""K1 (mm_generate) against the oracle's generator under the same Philox stream, plus structural properties."""
import numpy as np
import pytest

from oracle import OracleMaze

pytestmark = pytest.mark.gpu


def _check_tree_properties(m):
    lay = np.asarray(m["layout"]); H, W = lay.shape
    open_ = lay == 0
    n_open = int(open_.sum())
    edges = int((open_[:, :-1] & open_[:, 1:]).sum() + (open_[:-1, :] & open_[1:, :]).sum())
    assert edges == n_open - 1, "open cells must form a tree (perfect maze)"
    sx, sy = m["start"]; ex, ey = m["end"]; kx, ky = m["key"]
    assert sx % 2 == 0 and sy % 2 == 0 and open_[sy, sx]
    assert ex in (0, W - 1) and open_[ey, ex] and (ex, ey) != (sx, sy)
    assert open_[ky, kx] and (kx, ky) not in ((sx, sy), (ex, ey))
    p1 = m["path1"]; assert abs(p1[0] - sx) + abs(p1[1] - sy) == 1 and open_[p1[1], p1[0]]


@pytest.mark.parametrize("cfg", [dict(side=(13, 13), rand_start=True, difficulty=1), dict(side=(12, 13), rand_start=True, difficulty=1),
                                 dict(side=(25, 25), rand_start=True, difficulty=3), dict(side=(4, 9), rand_start=False, difficulty=2),
                                 dict(side=(27, 27), rand_start=True, difficulty=1)])
def test_generator_matches_oracle_philox(cfg):
    from marl_maze_b200 import MazeEngine
    n = 300
    smax = cfg["side"][1] * 2 - 1
    eng = MazeEngine(4, smax=smax, max_timestep=100, pool_size=n)
    seed, id_base = 0x1234_5678_9ABC_DEF0, 77
    eng.generate(seed, side_range=cfg["side"], rand_start=cfg["rand_start"], difficulty=cfg["difficulty"], id_base=id_base)
    o = OracleMaze(max_timestep=10, difficulty=cfg["difficulty"], rand_start=cfg["rand_start"], rand_sizes=True, rand_range=cfg["side"], default_size=(4, 4))
    for p in list(range(0, 40)) + [n - 1]:
        g = eng.pool_maze(p)
        o.seed_philox(seed, id_base + p); o.build(); m = o.maze()
        assert (g["width"], g["height"]) == (m["width"], m["height"])
        assert np.array_equal(g["layout"], m["layout"]), f"maze {p} layout"
        assert g["start"] == m["start"] and g["end"] == m["end"] and g["key"] == m["key"], (p, g["start"], m["start"], g["end"], m["end"], g["key"], m["key"])
        assert g["path1"] == m["path1"] and g["shortest_path_len"] == m["shortest_path_len"]
        # dir-to-exit field: following it from every cell of the oracle's path walks that path
        path = m["path"]
        for (x, y), (nx, ny) in zip(path[:-1], path[1:]):
            k = int(g["d2e"][y, x]); assert (x + (k == 1) - (k == 3), y + (k == 2) - (k == 0)) == (nx, ny)
        _check_tree_properties(g)


@pytest.mark.parametrize("wh", [(6, 11), (13, 5), (4, 27), (20, 9)], ids=lambda v: f"{2 * v[0] - 1}x{2 * v[1] - 1}")
def test_rectangular_mazes_match_oracle_and_step(wh):
    """Maze(default_size=[w, h]) with rand_sizes False (maze.py:26-27,170-178): rectangular mazes.  K1 against the oracle's generator under the same Philox
    stream (no size draw in this mode), then K2 stepping those mazes against the oracle, every field."""
    import torch
    from marl_maze_b200 import MazeEngine
    from oracle import OracleBatch
    w, h = wh
    E, K, max_t = 64, 2, 90
    smax = 2 * max(w, h) - 1
    eng = MazeEngine(E, smax=smax, max_timestep=max_t, pool_size=E * K)
    seed, id_base = 4242, 9
    eng.generate(seed, side_range=(w, w), rand_start=True, difficulty=2, id_base=id_base, height_cells=h)
    o = OracleMaze(max_timestep=10, difficulty=2, rand_start=True, rand_sizes=False, default_size=(w, h))
    ob = OracleBatch(E, E * K, max_timestep=max_t, threads=4)
    for p in range(E * K):
        g = eng.pool_maze(p)
        o.seed_philox(seed, id_base + p); o.build(); m = o.maze()
        assert (g["width"], g["height"]) == (m["width"], m["height"]) == (2 * w - 1, 2 * h - 1)
        assert np.array_equal(g["layout"], m["layout"]), f"maze {p} layout"
        assert (g["start"], g["path1"], g["end"], g["key"], g["shortest_path_len"]) == (m["start"], m["path1"], m["end"], m["key"], m["shortest_path_len"]), p
        _check_tree_properties(g)
        ob.set_pool_maze(p, m)
    oo, om = ob.reset_all(); go, gm = eng.reset()
    assert np.array_equal(go.cpu().numpy().view(np.uint32), oo.view(np.uint32)) and np.array_equal(gm.cpu().numpy(), om)
    rng = (np.arange(E, dtype=np.uint64) + 3) * np.uint64(0x9E3779B97F4A7C15)
    for t in range(200):
        act = ob.guided_actions(rng, p_follow=0.8, p_mark=0.3)
        go, gm, gr, gd = eng.step(torch.from_numpy(act).to(eng.device))
        oo, om, orr, od = ob.step(act)
        assert np.array_equal(go.cpu().numpy().view(np.uint32), oo.view(np.uint32)), t
        assert np.array_equal(gm.cpu().numpy(), om) and np.array_equal(gr.cpu().numpy(), orr) and np.array_equal(gd.cpu().numpy(), od), t
    assert np.array_equal(eng.agents(), ob.agents()) and ob.errors() == 0


def test_generated_pool_steps_bit_exact_vs_oracle():
    """End to end without any injected data: K1 fills the pool, K2 steps it; the oracle generates the same mazes itself."""
    import torch
    from marl_maze_b200 import MazeEngine
    from oracle import OracleBatch
    E, S, max_t = 512, 25, 150
    eng = MazeEngine(E, smax=S, max_timestep=max_t, pool_size=2 * E)
    eng.generate(99, side_range=(12, 13), difficulty=2, id_base=5)
    ob = OracleBatch(E, 2 * E, max_timestep=max_t, threads=8)
    o = OracleMaze(max_timestep=10, difficulty=2, rand_start=True, rand_sizes=True, rand_range=(12, 13), default_size=(4, 4))
    for p in range(2 * E):
        o.seed_philox(99, 5 + p); o.build(); ob.set_pool_maze(p, o.maze())
    oo, om = ob.reset_all(); go, gm = eng.reset()
    assert np.array_equal(go.cpu().numpy().view(np.uint32), oo.view(np.uint32)) and np.array_equal(gm.cpu().numpy(), om)
    rng = (np.arange(E, dtype=np.uint64) + 3) * np.uint64(0x9E3779B97F4A7C15)
    for t in range(320):
        act = ob.guided_actions(rng, p_follow=0.8, p_mark=0.3)
        go, gm, gr, gd = eng.step(torch.from_numpy(act).to(eng.device))
        oo, om, orr, od = ob.step(act)
        assert np.array_equal(go.cpu().numpy().view(np.uint32), oo.view(np.uint32)), t
        assert np.array_equal(gm.cpu().numpy(), om) and np.array_equal(gr.cpu().numpy(), orr) and np.array_equal(gd.cpu().numpy(), od), t
    assert np.array_equal(eng.agents(), ob.agents())


def test_generator_regression_zero_draw_maze():
    """Maze id whose first carve draw is 0 in 24 bits: a [0,1) mapping stopped the carve at one cell (found at 8 x 1 Mi mazes)."""
    from marl_maze_b200 import MazeEngine
    bad = 4 * (1 << 20) + 9 * 65536 + 25010
    eng = MazeEngine(4, smax=49, max_timestep=100, pool_size=64)
    eng.generate(2026, side_range=(25, 25), id_base=bad - 10)
    o = OracleMaze(max_timestep=10, difficulty=1, rand_start=True, rand_sizes=True, rand_range=(25, 25), default_size=(4, 4))
    for p in range(64):
        g = eng.pool_maze(p)
        o.seed_philox(2026, bad - 10 + p); o.build(); m = o.maze()
        assert np.array_equal(g["layout"], m["layout"]) and g["end"] == m["end"] and g["key"] == m["key"]
        _check_tree_properties(g)
        assert int((g["layout"] == 0).sum()) == 1249  # every room carved
    assert o.error() == 0


def test_generator_regression_partial_maze_without_eligible_exit():
    """A carve that is popped early at every frontier leaves a partial maze whose left edge has no open cell: the reference's
    set_end would spin forever; K1 and the oracle must take the same documented way out (and stay in RNG lock-step afterwards)."""
    from marl_maze_b200 import MazeEngine
    eng = MazeEngine(4, smax=25, max_timestep=100, pool_size=8)
    eng.generate(1000, side_range=(4, 13), difficulty=3, id_base=187600)
    o = OracleMaze(max_timestep=10, difficulty=3, rand_start=True, rand_sizes=True, rand_range=(4, 13), default_size=(4, 4))
    for p in range(8):
        g = eng.pool_maze(p)
        o.seed_philox(1000, 187600 + p); o.build(); m = o.maze()
        assert np.array_equal(g["layout"], m["layout"])
        assert (g["start"], g["path1"], g["end"], g["key"], g["shortest_path_len"]) == (m["start"], m["path1"], m["end"], m["key"], m["shortest_path_len"]), p
    g = eng.pool_maze(2)
    assert int((g["layout"] == 0).sum()) == 11 and g["layout"][g["end"][1], g["end"][0]] == 0


@pytest.mark.parametrize("cfg", [dict(side=(13, 13), difficulty=1, height_cells=0), dict(side=(25, 25), difficulty=4, height_cells=0),
                                 dict(side=(4, 9), difficulty=2, height_cells=0), dict(side=(6, 6), difficulty=3, height_cells=11),
                                 dict(side=(27, 27), difficulty=1, height_cells=0)])
def test_dir_to_exit_field_leads_every_open_cell_to_the_exit(cfg):
    """K1 builds the dir-to-exit field from the carve's own parent pointers (tree rooted at the start) by turning the start -> exit path around.  A perfect
    maze is a tree, so the field is fully determined: from EVERY open cell, following it must walk open cells only, never revisit one, and end on the exit;
    the exit's own entry and every wall's entry are 0."""
    from marl_maze_b200 import MazeEngine
    n = 48
    smax = max(cfg["side"][1], cfg["height_cells"]) * 2 - 1
    eng = MazeEngine(4, smax=smax, max_timestep=100, pool_size=n)
    eng.generate(99, side_range=cfg["side"], rand_start=True, difficulty=cfg["difficulty"], id_base=5, height_cells=cfg["height_cells"])
    for p in range(n):
        g = eng.pool_maze(p)
        lay, d2e = np.asarray(g["layout"]), np.asarray(g["d2e"])
        H, W = lay.shape
        ex, ey = g["end"]
        assert d2e[ey, ex] == 0 and (d2e[lay != 0] == 0).all()
        dist = -np.ones((H, W), np.int64); dist[ey, ex] = 0
        for y0 in range(H):
            for x0 in range(W):
                if lay[y0, x0] != 0 or dist[y0, x0] >= 0:
                    continue
                chain, x, y = [], x0, y0
                while dist[y, x] < 0:
                    assert lay[y, x] == 0 and (x, y) not in chain and len(chain) <= H * W, (p, x0, y0)
                    chain.append((x, y))
                    k = int(d2e[y, x]); x, y = x + (k == 1) - (k == 3), y + (k == 2) - (k == 0)
                    assert 0 <= x < W and 0 <= y < H, (p, x0, y0)
                for i, (cx, cy) in enumerate(reversed(chain)):
                    dist[cy, cx] = dist[y, x] + 1 + i
        sx, sy = g["start"]
        assert dist[sy, sx] + 1 == g["shortest_path_len"]


def test_masked_generation_ragged_counts_and_sparse_masks():
    """mm_generate_masked compacts the consumed slots warp-wide before carving: slot counts that are not multiples of 32, an all-zero mask, a single slot
    and a random mask must rebuild exactly the masked slots (the oracle's maze under the new seed) and leave every other slot untouched."""
    import torch
    from marl_maze_b200 import MazeEngine
    n, side = 1000 + 77, (7, 9)
    eng = MazeEngine(4, smax=side[1] * 2 - 1, max_timestep=100, pool_size=n)
    eng.generate(11, side_range=side, rand_start=True, difficulty=2, id_base=3)
    grid0, hdr0 = eng.pool_grid.clone(), eng.pool_hdr.clone()
    rng = np.random.default_rng(5)
    masks = [np.zeros(n, np.uint8), np.eye(1, n, n - 1, dtype=np.uint8)[0], (rng.random(n) < 0.17).astype(np.uint8), (rng.random(n) < 0.9).astype(np.uint8)]
    o = OracleMaze(max_timestep=10, difficulty=2, rand_start=True, rand_sizes=True, rand_range=side, default_size=(4, 4))
    for it, mk in enumerate(masks):
        seed = 100 + it
        eng.generate(seed, side_range=side, rand_start=True, difficulty=2, id_base=3, only=torch.from_numpy(mk).cuda())
        same = (eng.pool_grid.view(n, -1) == grid0.view(n, -1)).all(1).cpu().numpy() & (eng.pool_hdr.view(n, -1) == hdr0.view(n, -1)).all(1).cpu().numpy()
        assert same[mk == 0].all(), it                                       # untouched slots: bit-identical
        for p in list(np.flatnonzero(mk)[:12]) + list(np.flatnonzero(mk)[-3:]):
            g = eng.pool_maze(int(p))
            o.seed_philox(seed, 3 + int(p)); o.build(); m = o.maze()
            assert np.array_equal(g["layout"], m["layout"]) and g["start"] == m["start"] and g["end"] == m["end"] and g["key"] == m["key"], (it, p)
        grid0, hdr0 = eng.pool_grid.clone(), eng.pool_hdr.clone()
