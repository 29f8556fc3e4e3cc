"""Time K1 (mm_generate_masked, Maze.build_maze maze.py:170-273): a full pool build and an incremental refill (a sixth of the slots, as after a config-3 rollout).
    python tools/k1_bench.py [--mazes 1048576] [--side-half 25]         (side = 2 * side_half - 1)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200.engine import MazeEngine
ap = argparse.ArgumentParser(); ap.add_argument("--mazes", type=int, default=1 << 20); ap.add_argument("--side-half", type=int, default=25)
ap.add_argument("--difficulty", type=int, default=1); a = ap.parse_args()
n, sh = a.mazes, a.side_half
eng = MazeEngine(min(n, 1024), smax=2 * sh - 1, max_timestep=1200, pool_size=n, device="cuda")


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


full = timed(lambda: eng.generate(seed=1, side_range=(sh, sh), rand_start=True, difficulty=a.difficulty))
hdr = eng.pool_hdr.view(torch.int32).view(-1, 4).clone()
only = (torch.rand(n, device="cuda") < 1 / 6).to(torch.uint8)
part = timed(lambda: eng.generate(seed=2, side_range=(sh, sh), rand_start=True, difficulty=a.difficulty, only=only))
print(json.dumps({"lib": os.environ.get("MARL_MAZE_LIB", "default"), "mazes": n, "side": 2 * sh - 1, "difficulty": a.difficulty, "full_build_ms": full,
                  "mazes_per_s": n / (full * 1e-3), "refill_one_sixth_ms": part, "refilled": int(only.sum()),
                  "hdr_checksum": int(hdr.to(torch.int64).sum().item())}))
