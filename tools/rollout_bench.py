"""BASELINE config[2]/[3]/[4] flavour: full PPO rollout (K4 policy + K2 step, T steps) + K3 GAE + the 5x5 update on this rank's envs.
    python tools/rollout_bench.py [--envs 65536] [--horizon 128] [--side-half 13] [--epochs 2]
    torchrun --nproc-per-node N tools/rollout_bench.py ...      # envs sharded, NCCL gradient all-reduce
Prints one JSON line (rank 0) with device-timed phases (max over ranks)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--horizon", type=int, default=128); ap.add_argument("--side-half", type=int, default=13)
ap.add_argument("--epochs", type=int, default=2); ap.add_argument("--max-t", type=int, default=1200); ap.add_argument("--no-update", action="store_true"); ap.add_argument("--update-tf32", action="store_true"); ap.add_argument("--simt", action="store_true"); ap.add_argument("--autograd-update", action="store_true", help="PyTorch autograd actor update instead of the K5 kernels"); ap.add_argument("--prefetch", action="store_true", help="PPO(prefetch_pool=True): build the next pool in the background")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from marl_maze_b200.PPO import PPO
from marl_maze_b200.maze import Maze
from marl_maze_b200.maze_agent import Agent

E, T = a.envs, a.horizon
brain = PPO(agent_amount=2, batch_size=E * T - 1 if (E * T) % 5 else E * T - 5, lr=0.00014, epochs=1, verbose=False, model_path=None, horizon=T, device=f"cuda:{local}", update_tf32=a.update_tf32, fused_update=not a.autograd_update, prefetch_pool=a.prefetch)
agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
maze = Maze(agents=agents, max_timestep=a.max_t, rand_sizes=True, rand_range=[a.side_half, a.side_half], rand_start=True, num_envs=E, device=f"cuda:{local}",
            seed=1, env_offset=rank * E)


def timed(fn):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return out, float(ms.item())


res = []
for ep in range(a.epochs):
    batch, ms_roll = timed(brain.get_batch)
    st = dict(brain.last_stats)
    ms_upd = None
    if not a.no_update:
        _, ms_upd = timed(lambda: brain.update(batch))
    res.append(dict(rollout_ms=ms_roll, update_ms=ms_upd, **st))
    del batch
if rank == 0:
    r = res[-1]
    steps = E * T * world
    print(json.dumps({"what": "PPO rollout (K4+K2 per step) + K3 GAE, then 5x5 minibatch update", "n_gpus": world, "envs_per_gpu": E, "horizon": T, "side": 2 * a.side_half - 1,
                      "rollout_ms": r["rollout_ms"], "update_ms": r["update_ms"], "rollout_agent_steps_per_s": 2 * steps / (r["rollout_ms"] * 1e-3),
                      "end_to_end_env_steps_per_s": steps / ((r["rollout_ms"] + (r["update_ms"] or 0)) * 1e-3), "episodes": r["episodes"], "solved": r["solved"], "keys": r["keys"],
                      "mean_reward_per_step": r["mean_reward_per_step"], "epochs": res, "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()
