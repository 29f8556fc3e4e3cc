"""The reference's main.py (main.py:1-23) with the imports switched to marl_maze_b200 -- headless (no pygame viewer).

    python examples/main.py                      # one maze, like the reference: loads ./PPO.pth if present, rolls the policy
    python examples/main.py --train --envs 4096  # brain.train() on a batch of mazes

Colours are opaque to this implementation (there is no renderer), so None is passed where the reference passes pygame Colors.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_maze_b200.maze import Maze
from marl_maze_b200.maze_agent import Agent
from marl_maze_b200.PPO import PPO

ap = argparse.ArgumentParser()
ap.add_argument("--train", action="store_true"); ap.add_argument("--envs", type=int, default=1); ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--steps", type=int, default=300)
a = ap.parse_args()

brain = PPO(agent_amount=2, batch_size=15000 if a.envs == 1 else a.envs * 128 - 5, lr=0.00014, epochs=a.epochs)
agents = (Agent('RED', brain, None, None, 2),
          Agent('BLUE', brain, None, None, 3))
maze = Maze(agents=agents, max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True, difficulty=1, default_size=[4, 4], num_envs=a.envs)

if a.train:
    brain.train()
else:  # the body of Maze.display_policy.update_env (maze.py:477-493) without the drawing calls
    assert a.envs == 1, "the viewer loop is the single-maze interface"
    obs, masks = maze.reset()
    total = 0.0
    for t in range(a.steps):
        action = [agents[i].get_action(obs[i], masks[i])[0] for i in range(2)]
        obs, masks, reward, done = maze.step(action)
        total += reward
        if done:
            print(f"episode finished at t={maze.current_t} reward so far {total}")
            obs, masks = maze.reset()
    print(maze.render_ascii())
    print(f"{a.steps} steps, total reward {total}")
