"""marl_maze_b200 -- B200-native hot path of MARL-Maze (batched env step + observation, rollout, GAE).

Host side is Python/PyTorch for buffers, streams and torch.distributed; all compute is hand-written
sm_100a CUDA reached through the C ABI in include/marl_maze_b200.h.  No CPU fallback.
"""
from .engine import MazeEngine, gae, OBS_DIM, MASK_DIM, AGENT_FIELDS  # noqa: F401

__all__ = ["MazeEngine", "gae", "OBS_DIM", "MASK_DIM", "AGENT_FIELDS"]
