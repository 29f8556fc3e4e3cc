"""Where the roles of k_trunk_fused wait (needs a -DMM_TF_PROFILE build: MARL_MAZE_LIB=variants/tf_prof.so python tools/trunk_profile.py)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200 import _abi
from marl_maze_b200.networks import Actor, Critic
from marl_maze_b200.policy import PolicyRunner
import ctypes as C
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
actor = Actor([264, 264, 264]).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
run = PolicyRunner(actor, critic, E, "cuda")
obs = torch.rand(E, 2, 65, device="cuda"); masks = torch.ones(E, 2, 6, dtype=torch.uint8, device="cuda")
prof = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
L = C.CDLL(os.environ.get("MARL_MAZE_LIB") or _abi._build.LIB)
L.mm_debug_trunk_profile_buffer(C.c_void_p(prof.data_ptr()))
for _ in range(3): run.forward(obs, masks)
torch.cuda.synchronize()
p = prof.cpu().double()
tiles = (2 * E + 127) // 128
names = ["prod:d_full", "prod:a_free", "prod:w_empty", "mma:d_free", "mma:h_ready", "mma:op_full", "mma:w_full(L0)", "mma:w_full(L1,2)",
         "epi:a_full", "epi:op_free", "epi:d_full(L0)", "epi:d_full(L1,2)", "epi:compute(L0,L1,L2)", "epi:compute(heads)", "-", "mma:total"]
per_tile = tiles / 148
print(json.dumps({"tiles": tiles, "kclk_per_tile": {n: round(float(p[:, i].mean()) / per_tile / 1e3, 2) for i, n in enumerate(names) if n != "-"}}))
