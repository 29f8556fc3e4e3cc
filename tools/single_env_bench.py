"""BASELINE config[0]: the reference's own use -- ONE maze, two agents, the policy asked for one action at a time through the
reference's interface (Maze.reset / Agent.get_action / Maze.step with python lists, the loop of maze.py:477-493) -- on this framework.
It is launch- and synchronisation-bound by construction (every step returns python lists); the batched path is the product.  Per step: two
get_action calls (each: one pinned copy in, tokens + fused trunk at one env, six logits back, the reference's draw on the host) and one Maze.step
(four action bytes in, the step kernel, observation / masks / reward / done back through pinned mirrors, one synchronisation).
    python tools/single_env_bench.py [--steps 3000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marl_maze_b200.PPO import PPO
from marl_maze_b200.maze import Maze
from marl_maze_b200.maze_agent import Agent

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=3000); a = ap.parse_args()
brain = PPO(agent_amount=2, batch_size=15000, lr=0.00014, verbose=False, model_path=None)
agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
maze = Maze(agents=agents, max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True, difficulty=1, default_size=[4, 4], num_envs=1, seed=0)
obs, masks = maze.reset()
for warm in (True, False):
    n = 200 if warm else a.steps
    torch.cuda.synchronize(); t0 = time.time()
    episodes = 0
    for _ in range(n):
        action = [agent.get_action(obs[i], masks[i])[0] for i, agent in enumerate(agents)]   # maze.py:484-486
        obs, masks, reward, done = maze.step(action)
        if done:
            obs, masks = maze.reset(); episodes += 1
    torch.cuda.synchronize(); dt = time.time() - t0
print(json.dumps({"what": "config[0]: 1 maze, reference list interface, get_action per agent + step", "steps": a.steps, "seconds": dt,
                  "env_steps_per_s": a.steps / dt, "agent_steps_per_s": 2 * a.steps / dt, "episodes": episodes}))
