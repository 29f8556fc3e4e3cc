"""Time the token kernel (projection + attention, K4 stage 1) alone: python tools/tok_bench.py [--rows 131072]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200 import _abi
from marl_maze_b200.networks import Actor, Critic
from marl_maze_b200.policy import pack_weights
ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=131072); ap.add_argument("--maps-only", action="store_true", help="mm_tokens_forward (second-generation kernel, a function of the per-token maps alone) instead of mm_tokens_forward_full (what the rollout runs)"); a = ap.parse_args()
R = a.rows
actor = Actor([264, 264, 264]).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
w = pack_weights(actor, critic, "cuda")
obs = torch.rand(R, 65, device="cuda"); x0 = torch.empty(R, 460, device="cuda")
L = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
fwd = L.mm_tokens_forward if a.maps_only else L.mm_tokens_forward_full
f = lambda: _abi.check(fwd(w.data_ptr(), obs.data_ptr(), R, x0.data_ptr(), st), "mm_tokens_forward")
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
# accuracy against the modules in float64 (projection + attention, networks.py:58-65,75-82), both projection modes, on a sample of the rows
acc = {}
for faithful in (True, False):
    actor.projection.faithful = faithful
    w2 = pack_weights(actor, critic, "cuda")
    n = min(R, 8192)
    _abi.check(fwd(w2.data_ptr(), obs.data_ptr(), n, x0.data_ptr(), st), "mm_tokens_forward")
    import copy
    a64 = copy.deepcopy(actor).double()
    with torch.no_grad():
        ref = a64.attention(a64.projection(obs[:n].double()))
    err = (x0[:n].double() - ref).abs()
    acc["faithful" if faithful else "indexed"] = {"max_abs_err": float(err.max()), "max_rel_err": float((err / (ref.abs() + 1e-2)).max()), "ref_absmax": float(ref.abs().max())}
print(json.dumps({"lib": os.path.basename(os.environ.get("MARL_MAZE_LIB", "default")), "rows": R, "kernel": "maps only (gen 2)" if a.maps_only else "full (rollout)", "tokens_ms": ms, "checksum": float(x0.double().sum()), "accuracy_vs_fp64": acc}))
