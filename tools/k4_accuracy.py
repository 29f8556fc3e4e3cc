"""Accuracy of both policy paths against a float64 evaluation of the same network (random observations, indexed projection)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ppo_oracle as po
from marl_maze_b200.networks import Actor, Critic
from marl_maze_b200.policy import PolicyRunner
asd, csd = po.seeded_state_dicts(5)
rng = np.random.default_rng(3); E = 3000
obs = rng.random((E, 2, 65)).astype(np.float32); masks = np.ones((E, 2, 6), np.uint8)
asd64 = {k: v.astype(np.float64) for k, v in asd.items()}
po.F = np.float64
mv64, mk64 = po.actor_forward(asd64, obs.reshape(-1, 65).astype(np.float64), faithful=False)
po.F = np.float32
mv32, mk32 = po.actor_forward(asd, obs.reshape(-1, 65), faithful=False)
ref = np.concatenate([mv64, mk64], 1)
print("numpy fp32 oracle vs fp64: max abs", np.abs(np.concatenate([mv32, mk32], 1) - ref).max())
for tc in (False, True):
    actor = Actor([264, 264, 264], faithful_projection=False).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
    actor.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()}); critic.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()})
    run = PolicyRunner(actor, critic, E, "cuda", tensor_cores=tc)
    logits = torch.zeros(E, 2, 6, device="cuda")
    run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(masks).cuda(), logits=logits)
    got = logits.cpu().numpy().reshape(-1, 6).astype(np.float64)
    err = np.abs(got - ref)
    print("tcgen05 3xTF32" if tc else "SIMT fp32    ", "vs fp64: max abs", err.max(), "max rel(|ref|>0.05)", (err / np.maximum(np.abs(ref), 0.05)).max(), "mean abs", err.mean())
