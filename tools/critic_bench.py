"""Time mm_critic_forward alone (Critic.forward, networks.py:96-102) at a given number of environments."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200.networks import Actor, Critic
from marl_maze_b200.policy import PolicyRunner
ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=65536); a = ap.parse_args()
E = a.envs
actor = Actor([264, 264, 264]).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
run = PolicyRunner(actor, critic, E, "cuda")
obs = torch.rand(E, 2, 65, device="cuda")
val = torch.empty(E, device="cuda")
for _ in range(5): run.values(obs, val)
torch.cuda.synchronize()
ref = critic(obs.reshape(E, 130).double().float()).reshape(-1)
err = ((val - ref).abs() / (ref.abs() + 1e-3)).max().item()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run.values(obs, val)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(json.dumps({"lib": os.environ.get("MARL_MAZE_LIB", "default"), "envs": E, "critic_ms": ms, "GFLOPs": 2 * 12480 * E / (ms * 1e-3) / 1e9, "max_rel_err_vs_torch": err}))
