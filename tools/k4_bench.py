"""Time the policy kernels in isolation (CUDA events): python tools/k4_bench.py [--envs 65536]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200.networks import Actor, Critic
from marl_maze_b200.policy import PolicyRunner
ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--simt", action="store_true"); ap.add_argument("--no-overlap", action="store_true"); ap.add_argument("--no-fused", action="store_true"); ap.add_argument("--no-value", action="store_true", help="actor only (no critic launch)"); a = ap.parse_args()
E = a.envs
actor = Actor([264, 264, 264]).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
run = PolicyRunner(actor, critic, E, "cuda", tensor_cores=not a.simt, overlap_critic=not a.no_overlap, fused_trunk=not a.no_fused)
obs = torch.rand(E, 2, 65, device="cuda"); masks = torch.ones(E, 2, 6, dtype=torch.uint8, device="cuda")
for _ in range(5): run.forward(obs, masks, want_value=not a.no_value)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run.forward(obs, masks, want_value=not a.no_value)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"lib": os.environ.get("MARL_MAZE_LIB", "default"), "envs": E, "fused_trunk": not a.no_fused, "critic": "none" if a.no_value else ("serial" if a.no_overlap else "side stream"), "policy_forward_ms": ms, "agent_rows_per_s": 2 * E / (ms * 1e-3), "trunk_TFLOPs_fp32_equiv": 2 * E * 0.525e6 / (ms * 1e-3) / 1e12}))
