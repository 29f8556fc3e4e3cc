"""Quick K2-only timing (device-resident, in-kernel action sampling) for A/B-testing kernel variants on the GPU box.
    MARL_MAZE_LIB=path/to/variant.so python tools/k2_bench.py [--envs N] [--side-half 25] [--steps 200]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200 import MazeEngine

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20); ap.add_argument("--side-half", type=int, default=25)
ap.add_argument("--steps", type=int, default=200); ap.add_argument("--warmup", type=int, default=40); ap.add_argument("--max-t", type=int, default=1200)
a = ap.parse_args()
S = 2 * a.side_half - 1
eng = MazeEngine(a.envs, smax=S, max_timestep=a.max_t, pool_size=a.envs)
eng.generate(2026, side_range=(a.side_half, a.side_half)); eng.reset()
ao = torch.zeros(a.envs, 2, 2, dtype=torch.uint8, device="cuda")
for _ in range(a.warmup): eng.step(None, action_seed=1, actions_out=ao)
torch.cuda.synchronize()
best = 1e9; tot = 0
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): eng.step(None, action_seed=1, actions_out=ao)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps; best = min(best, ms); tot += ms
bpe = 685 + (S * S + 3) // 4
print(json.dumps({"lib": os.environ.get("MARL_MAZE_LIB", "default"), "envs": a.envs, "side": S, "ms_per_step_best": best, "ms_per_step_mean": tot / 3,
                  "agent_steps_per_s": 2 * a.envs / (best * 1e-3), "alg_GBs": bpe * a.envs / (best * 1e-3) / 1e9, "err": int(eng.envs()[:, 6].sum())}))
