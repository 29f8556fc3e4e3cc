"""Where the PPO update spends its time: torch.profiler kernel table over ONE actor + ONE critic minibatch at config-3 size.
    python tools/update_profile.py [--envs 65536] [--horizon 128] [--tf32] [--indexed]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from marl_maze_b200.PPO import PPO
from marl_maze_b200.maze import Maze
from marl_maze_b200.maze_agent import Agent

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--horizon", type=int, default=128)
ap.add_argument("--tf32", action="store_true"); ap.add_argument("--indexed", action="store_true")
ap.add_argument("--micro", type=int, default=1 << 17)
a = ap.parse_args()
E, T = a.envs, a.horizon
brain = PPO(agent_amount=2, batch_size=E * T // 5 * 5, lr=2e-4, epochs=1, verbose=False, model_path=None, horizon=T,
            faithful_projection=not a.indexed, update_tf32=a.tf32, micro_batch=a.micro)
agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
maze = Maze(agents=agents, max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True, num_envs=E, seed=3)
batch = brain.get_batch()
brain.updates_per_batch = 1
torch.cuda.synchronize(); t0 = time.time(); brain.update(batch); torch.cuda.synchronize()
print(f"one epoch (5 minibatches, actor+critic): {time.time() - t0:.3f} s   tf32={a.tf32} indexed={a.indexed} micro={a.micro}")
# profile a single minibatch by shrinking 'batch_size' to one minibatch worth of samples
brain.batch_size = brain.mbatch_size
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    brain.update(batch); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
