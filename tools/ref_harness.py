"""Drive the UNMODIFIED reference (rhuangr/MARL-Maze, mounted read-only at /root/reference) headless.

This file is test tooling: it exists only in the build container (the GPU box has no /root/reference).
It is used by tools/make_golden.py to record golden traces and by the `-m "not gpu"` tests that
cross-check the C oracle against the live reference when the reference happens to be present.

Nothing in the product path (marl_maze_b200/) imports this module.

How the reference is reached:
  * maze.py:1 imports pygame (not installed) -> a stub module exposing `Color` is injected into
    sys.modules; non-render code only evaluates pygame.Color(...) at import (maze.py:6-10).
  * Maze.__init__ only touches agent.maze / agent.brain.maze (maze.py:40-42) so a dummy brain is
    enough for environment-only traces.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys
import types

REFERENCE_DIR = os.environ.get("MARL_MAZE_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "maze.py"))


def _install_pygame_stub() -> None:
    if "pygame" in sys.modules:
        return
    stub = types.ModuleType("pygame")

    class Color:  # maze.py:6-10 / main.py:6-13 only construct colours
        def __init__(self, *a):
            self.args = a

    stub.Color = Color
    sys.modules["pygame"] = stub


_mods = None


def load_reference():
    """Returns (maze_module, maze_agent_module)."""
    global _mods
    if _mods is None:
        if not reference_available():
            raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
        _install_pygame_stub()
        sys.path.insert(0, REFERENCE_DIR)
        try:
            import maze as ref_maze  # noqa
            import maze_agent as ref_agent  # noqa
        finally:
            sys.path.remove(REFERENCE_DIR)
        _mods = (ref_maze, ref_agent)
    return _mods


class _DummyBrain:
    maze = None


MAIN_PY_KW = dict(max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True,
                  difficulty=1, default_size=[4, 4])  # main.py:20


def make_env(vision=(4, 4), **maze_kw):
    """Reference Maze with two Agents tagged 2 and 3 (main.py:18-20); vision = the agents' vision_range (maze_agent.py:16)."""
    ref_maze, ref_agent = load_reference()
    brain = _DummyBrain()
    agents = (ref_agent.Agent("RED", brain, None, None, 2, vision_range=vision[0]), ref_agent.Agent("BLUE", brain, None, None, 3, vision_range=vision[1]))
    kw = dict(MAIN_PY_KW)
    kw.update(maze_kw)
    return ref_maze.Maze(agents=agents, **kw)


def maze_snapshot(env) -> dict:
    """Everything a batched implementation needs to re-create the episode the reference just built."""
    return dict(
        width=env.width, height=env.height,
        layout=[list(r) for r in env.layout],
        start=tuple(env.start), end=tuple(env.end), key=tuple(env.key),
        path0=tuple(env.shortest_path[0]), path1=tuple(env.shortest_path[1]),
        shortest_path_len=env.shortest_path_len,
    )


def agent_snapshot(a) -> list:
    lm = a.last_mark_pos
    return [a.x, a.y, a.direction, int(a.knows_end), int(a.other_knows_end), int(a.has_key),
            int(a.team_has_key), a.exit_len, a.time_from_last_seen,
            a.other_last_seen[0], a.other_last_seen[1],
            -1 if lm is None else lm[0], -1 if lm is None else lm[1],
            a.min_x_visited, a.max_x_visited, a.min_y_visited, a.max_y_visited,
            -1 if a.exit_route is None else len(a.exit_route)]


AGENT_FIELDS = ["x", "y", "direction", "knows_end", "other_knows_end", "has_key", "team_has_key", "exit_len",
                "time_from_last_seen", "ols_x", "ols_y", "lm_x", "lm_y", "min_x", "max_x", "min_y", "max_y",
                "route_len"]


def legal_random_action(rng: random.Random, masks):
    """The action rule of SURVEY 8c KAT(2): per agent, agent 0 drawn first."""
    act = []
    for m in masks:
        move = rng.choice([k for k in range(5) if m[k]])
        mark = rng.randint(0, 1) if m[5] else 0
        act.append([move, mark])
    return act


def run_trace(maze_seed: int, action_seed: int, n_steps: int, maze_kw=None, policy="uniform", vision=(4, 4)):
    """Runs the reference for n_steps env-steps with reset-on-done, exactly like PPO.get_batch's
    loop (PPO.py:104-141) minus the networks.  Returns a dict of python lists:

      mazes[k]            : maze_snapshot of episode k (k-th reset)
      emit_obs/masks[i]   : the (obs, masks) the policy would see before step i  (i = 0 is the first reset)
      actions[i], reward[i], done[i] : step i
      step_obs/step_masks[i] : what Maze.step returned at step i (the terminal observation included)
      agents_after[i]     : agent_snapshot x2 after step i (before any reset)
      episode_of[i]       : episode index step i belongs to
    """
    with contextlib.redirect_stdout(io.StringIO()):
        env = make_env(vision=vision, **(maze_kw or {}))
        random.seed(maze_seed)
        rng = random.Random(action_seed)
        out = dict(mazes=[], emit_obs=[], emit_masks=[], actions=[], reward=[], done=[], step_obs=[],
                   step_masks=[], agents_after=[], agents_emit=[], episode_of=[], max_timestep=env.max_timestep)
        obs, masks = env.reset()
        out["mazes"].append(maze_snapshot(env))
        ep = 0
        for i in range(n_steps):
            out["emit_obs"].append([[float(v) for v in o] for o in obs])
            out["emit_masks"].append([[bool(v) for v in m] for m in masks])
            out["agents_emit"].append([agent_snapshot(a) for a in env.agents])
            if policy == "uniform":
                act = legal_random_action(rng, masks)
            elif policy == "nomark":  # fewer marks -> different mask/mark statistics
                act = legal_random_action(rng, masks)
                for a in act:
                    if rng.random() < 0.8:
                        a[1] = 0
            elif policy == "guided":
                act = guided_action(env, rng, masks)
            elif policy == "guided_slow":  # long wandering with occasional progress: drifting exit_len, marks
                act = guided_action(env, rng, masks, p_follow=0.45, p_mark=0.5)
            else:
                raise ValueError(policy)
            out["actions"].append(act)
            out["episode_of"].append(ep)
            obs, masks, reward, done = env.step([[a[0], float(a[1])] for a in act])
            out["reward"].append(float(reward))
            out["done"].append(bool(done))
            out["step_obs"].append([[float(v) for v in o] for o in obs])
            out["step_masks"].append([[bool(v) for v in m] for m in masks])
            out["agents_after"].append([agent_snapshot(a) for a in env.agents])
            if done:
                obs, masks = env.reset()
                out["mazes"].append(maze_snapshot(env))
                ep += 1
        out["final_obs"] = [[float(v) for v in o] for o in obs]
        out["final_masks"] = [[bool(v) for v in m] for m in masks]
        out["final_t"] = env.current_t
    return out


def _bfs_dirs(env, target):
    """abs direction (0=N,1=E,2=S,3=W; DELTAS of maze.py:19) of the first step from every open cell to target."""
    from collections import deque
    deltas = [(0, -1), (1, 0), (0, 1), (-1, 0)]
    W, H = env.width, env.height
    d = {tuple(target): -1}
    q = deque([tuple(target)])
    while q:
        x, y = q.popleft()
        for k, (dx, dy) in enumerate(deltas):
            nx, ny = x + dx, y + dy
            if 0 <= nx < W and 0 <= ny < H and env.layout[ny][nx] != 1 and (nx, ny) not in d:
                d[(nx, ny)] = (k + 2) % 4  # from (nx,ny) walk back towards (x,y)
                q.append((nx, ny))
    return d


def guided_action(env, rng: random.Random, masks, p_follow=0.85, p_mark=0.3):
    """Mask-legal actions biased towards key-then-exit so that traces contain key pickups, route sharing,
    the exit_ready mask override (maze.py:107-113) and reward-1 terminations."""
    act = []
    for i, a in enumerate(env.agents):
        m = masks[i]
        legal = [k for k in range(5) if m[k]]
        target = env.key if env.key != 0 else env.end
        move = None
        if rng.random() < p_follow:
            dirs = _bfs_dirs(env, target)
            ad = dirs.get((a.x, a.y), -1)
            if ad >= 0:
                rel = (ad - a.direction) % 4
                if m[rel]:
                    move = rel
        if move is None:
            move = rng.choice(legal)
        mark = int(rng.random() < p_mark) if m[5] else 0
        act.append([move, mark])
    return act
