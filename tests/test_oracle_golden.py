"""The C oracle (oracle/maze_oracle.c) against golden vectors recorded from the unmodified reference.

Two routes, both bit-exact:
  * seeded: the oracle regenerates every maze itself from random.seed(k) through its restatement of Python's
    Mersenne Twister + Maze.build_maze (maze.py:170-273) and replays the recorded actions;
  * injected: mazes are injected from the recording (the route the CUDA parity tests use).
"""
import hashlib

import numpy as np
import pytest

from golden_util import load_traces, load_vision_traces, kat2_sha256, load_gen_kats
from oracle import OracleMaze

TRACES = load_traces()
VTRACES = load_vision_traces()


def _replay(tr, injected):
    o = OracleMaze(vision=tr.vision, **tr.maze_kw)
    if injected:
        obs, masks = o.reset_injected(tr.mazes[0])
    else:
        o.seed(tr.maze_seed)
        obs, masks = o.reset()
    ep = 0
    emitted = [(obs, masks, 0.0, False)]
    for i in range(tr.n):
        if i == 0 or tr.done[i - 1]:
            assert np.array_equal(obs, tr.reset_obs[ep]), (tr.name, i, "reset obs")
            assert np.array_equal(masks, tr.reset_masks[ep]), (tr.name, i, "reset masks")
            assert np.array_equal(o.agents(), tr.reset_agents[ep]), (tr.name, i, "reset agents")
            m = o.maze()
            g = tr.mazes[ep]
            assert np.array_equal(m["layout"], g["layout"]) and m["end"] == g["end"] and m["key"] == g["key"]
            assert m["path0"] == g["path0"] and m["path1"] == g["path1"] and m["shortest_path_len"] == g["shortest_path_len"]
        obs, masks, r, d = o.step(tr.actions[i].reshape(4))
        assert np.array_equal(obs, tr.step_obs[i]), (tr.name, i, "obs", np.nonzero(obs != tr.step_obs[i]))
        assert np.array_equal(masks, tr.step_masks[i]), (tr.name, i, "masks")
        assert r == tr.reward[i] and d == bool(tr.done[i]), (tr.name, i, r, d)
        assert np.array_equal(o.agents(), tr.agents_after[i]), (tr.name, i, "agents")
        emitted.append((obs, masks, r, d))
        if d:
            ep += 1
            obs, masks = o.reset_injected(tr.mazes[ep]) if injected else o.reset()
            emitted.append((obs, masks, 0.0, False))
    assert o.error() == 0
    return emitted


@pytest.mark.parametrize("name", sorted(TRACES))
def test_oracle_seeded_replay(name):
    _replay(TRACES[name], injected=False)


@pytest.mark.parametrize("name", sorted(TRACES))
def test_oracle_injected_replay(name):
    _replay(TRACES[name], injected=True)


@pytest.mark.parametrize("injected", [False, True], ids=["seeded", "injected"])
@pytest.mark.parametrize("name", sorted(VTRACES))
def test_oracle_vision_range_replay(name, injected):
    """Agent(..., vision_range=r) with r != 4, per agent (maze_agent.py:16,148,165,218,264) -- observations (the ray features step in
    units of 1/r accumulated in float64), masks, rewards and agent state against reference traces recorded with vision (3,3), (2,2), (1,1), (4,2), (1,3)."""
    assert VTRACES[name].vision != (4, 4)
    _replay(VTRACES[name], injected=injected)


def test_oracle_kat2_sha256():
    """SURVEY.md 8c KAT(2): sha256 over every emitted (obs, masks, reward, done) of a 5000-step seeded run."""
    h = hashlib.sha256()
    for obs, masks, r, d in _replay(TRACES["kat2_uniform"], injected=False):
        h.update(np.asarray(obs, np.float32).tobytes()); h.update(np.asarray(masks, np.uint8).tobytes())
        h.update(np.float32(r).tobytes()); h.update(bytes([int(d)]))
    assert h.hexdigest() == kat2_sha256() == "4809ce85defd322829727dd95048c32b7d522acc9a7d29685b8b4d9d489325ef"


def test_oracle_kat1_first_reset():
    """SURVEY.md 8c KAT(1): random.seed(0), main.py configuration."""
    o = OracleMaze(max_timestep=1200, rand_sizes=True, rand_range=(12, 13), rand_start=True, difficulty=1, default_size=(4, 4))
    o.seed(0)
    obs, masks = o.reset()
    m = o.maze()
    assert (m["width"], m["height"], m["start"], m["end"], m["key"], m["shortest_path_len"]) == (25, 25, (24, 12), (0, 15), (6, 20), 152)
    assert obs[0][:8].tolist() == [0, 0, 1, 0, 1, 0, 1, 1] and obs[0][48:52].tolist() == [1, 0, 1, 0]
    assert obs[0][57] == np.float32(-0.025) and obs[0][61] == np.float32(0.025)
    assert masks.tolist() == [[0, 1, 0, 0, 0, 1]] * 2


def test_oracle_kat3_tiny_maze():
    """SURVEY.md 8c KAT(3): 7x7 fixed-start maze, random.seed(5)."""
    o = OracleMaze(default_size=(4, 4), rand_sizes=False, rand_start=False)
    o.seed(5)
    o.build()
    m = o.maze()
    assert (m["start"], m["end"], m["key"]) == ((2, 0), (0, 2), (2, 2))
    assert m["path"].tolist() == [[2, 0], [1, 0], [0, 0], [0, 1], [0, 2]]
    rows = ["".join(str(int(c)) for c in r) for r in m["layout"]]
    assert rows == ["0001000", "0111110", "0000000", "1111110", "0100000", "0101110", "0000010"]


@pytest.mark.parametrize("key", sorted(load_gen_kats()))
def test_oracle_generator_vs_reference_seeds(key):
    k = load_gen_kats()[key]
    o = OracleMaze(max_timestep=10, **k["kw"])
    o.seed(k["seed"])
    for j in range(3):
        o.build()
        m = o.maze()
        hdr = k["hdr"][j]
        W, H = int(hdr[0]), int(hdr[1])
        assert [m["width"], m["height"], *m["start"], *m["end"], *m["key"], m["shortest_path_len"], *m["path1"]] == hdr.tolist()
        assert np.array_equal(m["layout"], np.unpackbits(k["layout"][j], axis=-1)[:H, :W])
        assert hashlib.sha256(m["path"].astype(np.int32).tobytes()).digest() == k["path_sha"][32 * j:32 * j + 32]
