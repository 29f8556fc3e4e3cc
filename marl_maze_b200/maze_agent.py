"""Agent: the reference's per-agent handle (maze_agent.py:15-57) over state that now lives packed in HBM.

The reference Agent owns python fields and computes its own observation; here observation, movement and marking for
ALL agents of ALL mazes happen inside the fused step kernel (csrc/mm_step_obs.cu), and an Agent is a thin view:
constructor signature, `maze` / `brain` wiring, `get_action`, `get_observations` and read-only state properties
(x, y, direction, knows_end, ...) that are fetched from the device on demand.  `env` selects which maze of a batched
Maze the properties look at (default 0).
"""
from __future__ import annotations

from math import exp

from .engine import AGENT_FIELDS

ACTIONS = ["forward", "right", "backward", "left"]
DIRECTIONS = ["north", "east", "south", "west"]


class Agent:
    def __init__(self, name, brain, color, mark_color, tag, vision_range=4):
        if not (isinstance(vision_range, int) and 1 <= vision_range <= 4):
            raise NotImplementedError("vision_range must be 1..4 (maze_agent.py:16 default 4): the packed grids keep a wall border of 5 = vision_range + 1 "
                                      "cells around every maze and an agent's window is the 11 rows / 11 bit columns around it")
        if tag not in (2, 3):
            raise ValueError("tags 2 and 3 are the two cell values the grid encodes for marks (main.py:18-19)")
        self.maze = None
        self.name, self.brain, self.color, self.mark_color, self.tag, self.vision_range = name, brain, color, mark_color, tag, vision_range
        self.env = 0
        self.average_exit = 5000

    # ---- the reference's two state setters (maze_agent.py:59-87), on the packed device state of env `self.env` ----------------------------
    def reset(self, x, y):
        """Agent.reset(x, y): position, facing south, flags / memory / route / bounding box cleared; time_from_last_seen kept (maze_agent.py:59-79)."""
        self.maze._ensure_engine().place_agent(self.env, self._index, x, y, 2, reset=True)

    def move(self, x, y, direction):
        """Agent.move(x, y, direction) (maze_agent.py:85-87)."""
        self.maze._ensure_engine().place_agent(self.env, self._index, x, y, direction, reset=False)

    # ---- behaviour -----------------------------------------------------------------------------------------------
    def get_action(self, obs, mask):
        """(action, probability of that action) -- maze_agent.py:81-83."""
        action, log_prob = self.brain.get_action(obs, mask)
        return action, exp(float(log_prob))

    def get_observations(self):
        """The observation / mask this agent was last given by Maze.reset/step (the kernel has already computed it;
        unlike the reference, asking again has no side effects)."""
        return self.maze._last_obs_of(self)

    # ---- state views ---------------------------------------------------------------------------------------------
    def _field(self, name):
        return int(self.maze._agent_state(self)[AGENT_FIELDS.index(name)])

    x = property(lambda s: s._field("x"))
    y = property(lambda s: s._field("y"))
    direction = property(lambda s: s._field("direction"))
    knows_end = property(lambda s: bool(s._field("knows_end")))
    other_knows_end = property(lambda s: bool(s._field("other_knows_end")))
    has_key = property(lambda s: bool(s._field("has_key")))
    team_has_key = property(lambda s: bool(s._field("team_has_key")))
    exit_len = property(lambda s: s._field("exit_len"))
    time_from_last_seen = property(lambda s: s._field("time_from_last_seen"))
    other_last_seen = property(lambda s: (s._field("ols_x"), s._field("ols_y")))
    min_x_visited = property(lambda s: s._field("min_x"))
    max_x_visited = property(lambda s: s._field("max_x"))
    min_y_visited = property(lambda s: s._field("min_y"))
    max_y_visited = property(lambda s: s._field("max_y"))

    @property
    def last_mark_pos(self):
        x = self._field("lm_x")
        return None if x < 0 else (x, self._field("lm_y"))

    @property
    def current_t(self):
        return self.maze.current_t

    def print_obs(self, obs):
        from .networks import FEATURE_DIMS
        names = ["Direction", "Dead Ends", "Own Mark Visible", "Others Mark Visible", "Agent Visible", "Others Direction", "Visible Key",
                 "Move t-4", "Move t-3", "Move t-2", "Move t-1", "Last Mark Pos", "Relative Position", "Other Agent Relative Position", "Sees End",
                 "Next Move to Exit", "Exit Path Length", "Visible Agent Knows End", "Has Key", "Team Has Key", "Time Last Agent Seen", "Timestep", "ID"]
        i = 0
        print(f"-------------- Agent {self.name} at {self.x},{self.y} --------------")
        for n, d in zip(names, FEATURE_DIMS):
            print(f"{n}: {list(obs[i:i + d])}")
            i += d
