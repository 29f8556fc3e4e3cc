"""Short PPO training run on the batched environment (sanity: the whole loop -- K1 pool, K4 policy, K2 step, K3 GAE, update -- learns).
    python tools/train_demo.py [--envs 4096] [--horizon 128] [--iters 40] [--indexed] [--fp32-update] [--autograd-update]
    torchrun --nproc-per-node N tools/train_demo.py ...   # --envs mazes PER RANK, NCCL gradient all-reduce (BASELINE config 5)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200.PPO import PPO
from marl_maze_b200.maze import Maze
from marl_maze_b200.maze_agent import Agent

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096); ap.add_argument("--horizon", type=int, default=128); ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--indexed", action="store_true", help="fixed Projection slicing instead of the reference's column-0 behaviour")
ap.add_argument("--side-half", type=int, default=6); ap.add_argument("--max-t", type=int, default=300); ap.add_argument("--lr", type=float, default=3e-4)
ap.add_argument("--fp32-update", action="store_true", help="with --autograd-update: cuBLAS fp32 instead of TF32"); ap.add_argument("--autograd-update", action="store_true")
a = ap.parse_args()
E, T = a.envs, a.horizon
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
brain = PPO(agent_amount=2, batch_size=E * T - 5 if (E * T) % 5 == 0 else (E * T) // 5 * 5, lr=a.lr, epochs=1, verbose=False, model_path=None, horizon=T,
            faithful_projection=not a.indexed, update_tf32=not a.fp32_update, fused_update=not a.autograd_update, device=f"cuda:{local}")
agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
maze = Maze(agents=agents, max_timestep=a.max_t, rand_sizes=True, rand_range=[a.side_half, a.side_half], rand_start=True, num_envs=E, seed=3,
            device=f"cuda:{local}", env_offset=rank * E)
log = []
t0 = time.time()
for it in range(a.iters):
    batch = brain.get_batch()
    st = dict(brain.last_stats)
    up = brain.update(batch)
    row = dict(iter=it, solved=st["solved"], keys=st["keys"], episodes=st["episodes"], reward_per_kstep=1000 * st["mean_reward_per_step"],
               mean_ep_len=float(batch[4].mean()) if len(batch[4]) else None, actor_loss=up["actor_loss"] / up["steps"], critic_loss=up["critic_loss"] / up["steps"])
    log.append(row)
    if rank == 0 and (it % 5 == 0 or it == a.iters - 1):
        print(json.dumps(row), flush=True)
torch.cuda.synchronize()
if rank == 0:
  print(json.dumps({"summary": True, "n_gpus": world, "envs_per_gpu": E, "horizon": T, "iters": a.iters, "env_steps_total": world * E * T * a.iters, "indexed_projection": a.indexed, "first5_reward_per_kstep": sum(r["reward_per_kstep"] for r in log[:5]) / 5,
                  "last5_reward_per_kstep": sum(r["reward_per_kstep"] for r in log[-5:]) / 5, "seconds": time.time() - t0}))
if world > 1:
    dist.destroy_process_group()
