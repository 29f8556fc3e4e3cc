"""CPU parity oracle for the MARL-Maze hot path.  TEST INFRASTRUCTURE ONLY.

`oracle/` is a scalar C restatement of the reference environment (maze_oracle.c) and a numpy restatement of
the reference's GAE / policy arithmetic (ppo_oracle.py), each function citing the reference file:line it
follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
it.  The product package (marl_maze_b200) never does, and fails loudly without its CUDA library.

Parity pin: the reference has no tests or golden vectors of its own; the oracle is pinned against the
reference itself, imported unmodified in the build container (tools/make_golden.py -> tests/golden/).
"""
from .build import build_oracle  # noqa: F401
from .env import OracleMaze, OracleBatch, load_lib  # noqa: F401
