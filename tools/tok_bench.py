"""Time the token kernel (projection + attention, K4 stage 1) alone: python tools/tok_bench.py [--rows 131072]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200 import _abi
from marl_maze_b200.networks import Actor, Critic
from marl_maze_b200.policy import pack_weights
ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=131072); a = ap.parse_args()
R = a.rows
actor = Actor([264, 264, 264]).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
w = pack_weights(actor, critic, "cuda")
obs = torch.rand(R, 65, device="cuda"); x0 = torch.empty(R, 460, device="cuda")
L = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
f = lambda: _abi.check(L.mm_tokens_forward(w.data_ptr(), obs.data_ptr(), R, x0.data_ptr(), st), "mm_tokens_forward")
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"lib": os.path.basename(os.environ.get("MARL_MAZE_LIB", "default")), "rows": R, "tokens_ms": ms, "checksum": float(x0.double().sum())}))
