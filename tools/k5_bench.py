"""Timing of the K5 update kernels and of one PPO update epoch, fused vs autograd.
    python tools/k5_bench.py [--rows 262144] [--envs 65536] [--horizon 128]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200 import update as U

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=262144); ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--horizon", type=int, default=128)
ap.add_argument("--skip-update", action="store_true"); ap.add_argument("--skip-kernels", action="store_true"); ap.add_argument("--micro", type=int, nargs="*", default=[1 << 17])
a = ap.parse_args()
R = a.rows
dev = "cuda"


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


def kernels():
    global R
    out = {}
    x0 = torch.randn(R, 460, device=dev); dz = torch.randn(R, 264, device=dev); h = torch.relu(torch.randn(R, 264, device=dev))
    w0 = torch.randn(264, 460, device=dev) / 20; w1 = torch.randn(264, 264, device=dev) / 16; b = torch.zeros(264, device=dev)
    s0, s1, s1t = U.tf32_split(w0), U.tf32_split(w1), U.tf32_split(w1.t())
    y = torch.empty(R, 264, device=dev)
    out["fwd_460_ms"] = timeit(lambda: U.linear_tc(x0, s0, U.MM_LINEAR_RELU, bias=b, out=y))
    out["fwd_264_ms"] = timeit(lambda: U.linear_tc(h, s1, U.MM_LINEAR_RELU, bias=b, out=y))
    from marl_maze_b200.policy import f16_split
    s0h, s1h = f16_split(w0, 480), f16_split(w1, 288)
    out["fwd16_460_ms"] = timeit(lambda: U.linear_f16(x0, s0h, b, out=y))
    out["fwd16_264_ms"] = timeit(lambda: U.linear_f16(h, s1h, b, out=y))
    _, bits = U.linear_tc(h, s1, U.MM_LINEAR_RELU, bias=b, want_bits=True)
    out["dgrad_264_ms"] = timeit(lambda: U.linear_tc(dz, s1t, U.MM_LINEAR_GATE, gate_bits=bits, out=y))
    out["plain_264_dz_ms"] = timeit(lambda: U.linear_tc(dz, s1t, U.MM_LINEAR_PLAIN, out=y))
    out["relu_264_dz_ms"] = timeit(lambda: U.linear_tc(dz, s1t, U.MM_LINEAR_RELU, bias=b, out=y))
    out["wgrad_264_ms"] = timeit(lambda: U.wgrad(dz, h))
    out["wgrad_460_ms"] = timeit(lambda: U.wgrad(dz, x0))
    out["torch_wgrad_264_fp32_ms"] = timeit(lambda: dz.t() @ h)
    out["torch_wgrad_460_fp32_ms"] = timeit(lambda: dz.t() @ x0)
    out["torch_fwd_460_fp32_ms"] = timeit(lambda: torch.relu(torch.addmm(b, x0, w0.t())))
    E = R // 2
    masks = torch.ones(R, 6, dtype=torch.uint8, device=dev); acts = torch.zeros(R, 2, dtype=torch.uint8, device=dev)
    old = torch.zeros(E, device=dev); adv = torch.randn(E, device=dev); wh = torch.randn(6, 264, device=dev) / 100; bh = torch.zeros(6, device=dev)
    out["heads_loss_ms"] = timeit(lambda: U.ppo_heads_loss(h, wh, bh, masks, acts, old, adv, 0.2, 1.0 / E))
    fl = lambda n, k: 2.0 * R * n * k / 1e9
    out["fwd_460_tflops_fp32eq"] = fl(264, 460) / out["fwd_460_ms"]; out["wgrad_264_tflops_fp32eq"] = fl(264, 264) / out["wgrad_264_ms"]
    out["wgrad_460_tflops_fp32eq"] = fl(264, 460) / out["wgrad_460_ms"]; out["dgrad_264_tflops_fp32eq"] = fl(264, 264) / out["dgrad_264_ms"]
    out["wgrad_264_GBps"] = R * (264 + 264) * 4 / 1e6 / out["wgrad_264_ms"]
    print(json.dumps({"rows": R, **{k: round(v, 4) for k, v in out.items()}}), flush=True)
    del x0, dz, h, y


if not a.skip_kernels:
    kernels()
if not a.skip_update:
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent
    E, T = a.envs, a.horizon
    res = {}
    for fused, micro in [(True, m) for m in a.micro] + [(False, 1 << 17)]:
        brain = PPO(agent_amount=2, batch_size=E * T // 5 * 5, lr=2e-4, epochs=1, verbose=False, model_path=None, horizon=T, fused_update=fused, micro_batch=micro)
        agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
        maze = Maze(agents=agents, max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True, num_envs=E, seed=3)
        batch = brain.get_batch()
        brain.updates_per_batch = 1
        brain.update(batch); torch.cuda.synchronize()
        t0 = time.time(); st = brain.update(batch); torch.cuda.synchronize()
        res[f"fused_micro{micro}" if fused else "autograd"] = dict(epoch_s=round(time.time() - t0, 3), actor_loss=st["actor_loss"] / st["steps"], critic_loss=st["critic_loss"] / st["steps"])
        del brain, maze, batch, agents
        torch.cuda.empty_cache()
    print(json.dumps({"update_epoch_5_minibatches": res, "envs": E, "horizon": T}))
