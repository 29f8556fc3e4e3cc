"""Time the UNMODIFIED Python reference (staged under baseline/_ref/, see stage_reference.py) on the host cores.

    python baseline/ref_python_bench.py --mode env    [--seconds 8] [--side-half 13]   # Maze.step + obs, mask-legal uniform random actions
    python baseline/ref_python_bench.py --mode policy [--seconds 8]                    # display_policy.update_env loop with PPO.pth (maze.py:477-493)
    python baseline/ref_python_bench.py --mode batch  [--batch 400]                    # PPO.get_batch (PPO.py:89-152)
    python baseline/ref_python_bench.py --mode env --procs P                           # P independent processes, summed

Prints ONE JSON line.  BASELINE.md section 3: the reference is single-threaded Python, so "all cores" means P independent processes.
pygame (not installed) is replaced by a stub exposing `Color` -- the only pygame name non-render code touches (maze.py:6-10).
Nothing in the product path imports this file; bench.py runs it as a subprocess for its cpu_baseline.python_reference entry.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import random
import subprocess
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_dir():
    for d in (os.path.join(HERE, "_ref"), os.environ.get("MARL_MAZE_REFERENCE", "/root/reference")):
        if d and os.path.isfile(os.path.join(d, "maze.py")):
            return d
    return None


def _load(with_brain: bool):
    d = ref_dir()
    if d is None:
        raise SystemExit(json.dumps({"unavailable": "reference not staged (run python baseline/stage_reference.py in the build container)"}))
    stub = types.ModuleType("pygame")

    class Color:
        def __init__(self, *a):
            self.args = a
    stub.Color = Color
    sys.modules.setdefault("pygame", stub)
    sys.path.insert(0, d)
    os.chdir(d if not with_brain else _scratch_copy(d))  # PPO.py loads / saves "PPO.pth" relative to the CWD
    import maze as ref_maze
    import maze_agent as ref_agent
    return ref_maze, ref_agent


def _scratch_copy(d):
    import shutil
    import tempfile
    t = tempfile.mkdtemp(prefix="mmref_")
    shutil.copy2(os.path.join(d, "PPO.pth"), os.path.join(t, "PPO.pth"))
    return t


MAIN_KW = dict(max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True, difficulty=1)  # main.py:20


def bench_env(seconds: float, side_half: int, seed: int):
    ref_maze, ref_agent = _load(False)

    class Brain:
        maze = None
    brain = Brain()
    agents = (ref_agent.Agent("RED", brain, None, None, 2), ref_agent.Agent("BLUE", brain, None, None, 3))
    kw = dict(MAIN_KW)
    if side_half:
        kw["rand_range"] = [side_half, side_half]
    m = ref_maze.Maze(agents=agents, **kw)
    random.seed(seed)
    rng = random.Random(seed + 1)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        obs, masks = m.reset()
        steps = 0
        t0 = time.perf_counter()
        while True:
            act = []
            for a in range(2):
                legal = [k for k in range(5) if masks[a][k]]
                act.append([rng.choice(legal) if legal else 4, rng.randint(0, 1) if masks[a][5] else 0])
            obs, masks, r, d = m.step(act)
            steps += 1
            if d:
                obs, masks = m.reset()
            if (steps & 255) == 0 and time.perf_counter() - t0 >= seconds:
                break
        dt = time.perf_counter() - t0
    return dict(mode="env", env_steps=steps, seconds=dt, agent_steps_per_s=2 * steps / dt, side=m.width)


def bench_policy(seconds: float, seed: int):
    import torch
    torch.set_num_threads(1)
    ref_maze, ref_agent = _load(True)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        import PPO as ref_ppo
        brain = ref_ppo.PPO(agent_amount=2, batch_size=15000, lr=0.00014)
        agents = (ref_agent.Agent("RED", brain, None, None, 2), ref_agent.Agent("BLUE", brain, None, None, 3))
        m = ref_maze.Maze(agents=agents, **MAIN_KW)
        random.seed(seed)
        obs, masks = m.reset()
        steps = 0
        t0 = time.perf_counter()
        with torch.no_grad():
            while True:  # maze.py:477-493 update_env without the drawing
                actions = []
                for i, agent in enumerate(m.agents):
                    action, _ = agent.get_action(obs[i], masks[i])
                    actions.append(action)
                obs, masks, r, d = m.step(actions)
                steps += 1
                if d:
                    obs, masks = m.reset()
                if (steps & 31) == 0 and time.perf_counter() - t0 >= seconds:
                    break
        dt = time.perf_counter() - t0
    return dict(mode="policy", env_steps=steps, seconds=dt, agent_steps_per_s=2 * steps / dt, torch_threads=1)


def bench_batch(batch: int, seed: int):
    import torch
    torch.set_num_threads(1)
    ref_maze, ref_agent = _load(True)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        import warnings
        warnings.filterwarnings("ignore")
        import PPO as ref_ppo
        brain = ref_ppo.PPO(agent_amount=2, batch_size=batch, lr=0.00014)
        agents = (ref_agent.Agent("RED", brain, None, None, 2), ref_agent.Agent("BLUE", brain, None, None, 3))
        ref_maze.Maze(agents=agents, **MAIN_KW)
        random.seed(seed)
        t0 = time.perf_counter()
        out = brain.get_batch()
        dt = time.perf_counter() - t0
        n = int(out[0].shape[0])
    return dict(mode="batch", env_steps=n, seconds=dt, agent_steps_per_s=2 * n / dt, env_steps_per_s=n / dt, torch_threads=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="env", choices=["env", "policy", "batch"])
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--side-half", type=int, default=0, help="0 = main.py's rand_range [12,13]; 25 = side 49")
    ap.add_argument("--batch", type=int, default=400)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    if a.procs > 1:
        cmd = [sys.executable, os.path.abspath(__file__), "--mode", a.mode, "--seconds", str(a.seconds), "--side-half", str(a.side_half), "--batch", str(a.batch)]
        t0 = time.perf_counter()
        ps = [subprocess.Popen(cmd + ["--seed", str(a.seed + 1000 * i)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                               env=dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")) for i in range(a.procs)]
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
        wall = time.perf_counter() - t0
        print(json.dumps(dict(mode=a.mode, procs=a.procs, agent_steps_per_s=sum(o["agent_steps_per_s"] for o in outs), env_steps=sum(o["env_steps"] for o in outs),
                              seconds=max(o["seconds"] for o in outs), wall_seconds=wall, side=outs[0].get("side"))))
        return
    if a.mode == "env":
        r = bench_env(a.seconds, a.side_half, a.seed)
    elif a.mode == "policy":
        r = bench_policy(a.seconds, a.seed)
    else:
        r = bench_batch(a.batch, a.seed)
    print(json.dumps(r))


if __name__ == "__main__":
    main()
