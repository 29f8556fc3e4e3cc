"""Maze: the reference's environment surface (maze.py:21-163) over the batched CUDA engine.

    brain = PPO(agent_amount=2, ...); agents = (Agent('RED', brain, None, None, 2), Agent('BLUE', brain, None, None, 3))
    maze = Maze(agents=agents, max_timestep=1200, rand_sizes=True, rand_range=[12, 13], rand_start=True)   # main.py:17-20
    obs, masks = maze.reset(); obs, masks, reward, done = maze.step([[move, mark], [move, mark]])

With num_envs == 1 the calls take and return python lists exactly like the reference (done is reported, the caller
resets).  With num_envs > 1 they take/return device tensors ([E,2,2] u8 actions -> obs [E,2,65] f32, masks [E,2,6] u8,
reward [E] f32, done [E] u8) and finished environments restart inside the same kernel launch (auto_reset).
Generation (Maze.build_maze, maze.py:170-273) runs on the GPU (K1) into a maze pool; `seed` replaces Python's global
Mersenne Twister.  The pygame viewer (maze.py:276-522) is out of scope; `render_ascii()` is provided for debugging.
"""
from __future__ import annotations

from typing import Optional

import torch

from .engine import MazeEngine

DELTAS = [(0, -1), (1, 0), (0, 1), (-1, 0)]  # N, E, S, W
_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


class Maze:
    def __init__(self, agents, max_timestep=3500, difficulty=1, rand_start=False, rand_sizes=False, rand_range=[6, 12], default_size=[8, 8],
                 num_envs: int = 1, device="cuda", seed: int = 0, pool_episodes: Optional[int] = None, env_offset: int = 0):
        if len(agents) != 2:
            raise NotImplementedError("the reference (and the step kernel's lane pairing) is a two-agent game (README.md:34)")
        if sorted(a.tag for a in agents) != [2, 3] or agents[0].tag != 2:
            raise ValueError("agents must be tagged (2, 3) in that order (main.py:18-19)")
        self.agents = agents
        self.max_timestep, self.difficulty, self.rand_start, self.rand_sizes = max_timestep, difficulty, rand_start, rand_sizes
        self.rand_range, self.default_size = list(rand_range), list(default_size)
        self.num_envs, self.device, self.seed, self.env_offset = int(num_envs), torch.device(device), int(seed), int(env_offset)
        self.side_range = (rand_range[0], rand_range[1]) if rand_sizes else (default_size[0], default_size[0])
        # Maze(default_size=[w, h]) with rand_sizes False: every maze is 2w-1 by 2h-1 (maze.py:26-27); square default sizes keep the rand_sizes path
        self.height_cells = default_size[1] if (not rand_sizes and default_size[0] != default_size[1]) else 0
        self.pool_episodes = pool_episodes or (64 if self.num_envs == 1 else 4)
        for i, agent in enumerate(self.agents):  # maze.py:40-42
            agent.maze = self
            agent.brain.maze = self
            agent._index = i
        self.engine: Optional[MazeEngine] = None
        self._generation = 0
        self._resets_since_fill = 0
        self._obs = self._masks = None
        self.exit_found = False

    # ------------------------------------------------------------------ engine / pool
    def _ensure_engine(self):
        if self.engine is None:
            smax = max(self.side_range[1], self.height_cells) * 2 - 1
            self.engine = MazeEngine(self.num_envs, smax=smax, max_timestep=self.max_timestep, pool_size=self.num_envs * self.pool_episodes,
                                     device=self.device, env_offset=self.env_offset, vision=tuple(a.vision_range for a in self.agents))
            self.refill_pool()
        return self.engine

    def _pool_args(self, generation: Optional[int] = None):
        """K1 arguments of pool refill number `generation`.  A maze is keyed by (64-bit seed, 32-bit id): the id is the maze's GLOBAL slot
        (global env, episode within the pool) -- so results do not depend on how envs are sharded over ranks -- and the refill counter is
        folded into the SEED (splitmix64), so no (seed, id) pair ever repeats across refills: a 32-bit id alone wrapped after 64 refills."""
        g = self._generation if generation is None else generation
        if (self.env_offset + self.num_envs) * self.pool_episodes > 1 << 32:
            raise ValueError("(env_offset + num_envs) * pool_episodes must fit the 32-bit maze id")
        return dict(seed=(self.seed + _splitmix64(g)) & _M64 if g else self.seed & _M64, side_range=self.side_range, rand_start=self.rand_start,
                    difficulty=self.difficulty, id_base=self.env_offset * self.pool_episodes, id_mod=self.num_envs, id_mul=self.pool_episodes,
                    height_cells=self.height_cells)

    def _refill_mask(self):
        """The slots the next refill builds anew.  The reference builds ONE maze per reset (maze.py:57); a rollout draws its mazes from pool slots
        (env e, episode k), so between two refills only the slots an episode actually started on are consumed -- at config 3 that is 65 613 of
        393 216 -- and only those are rebuilt (with the new refill's seed); the others keep their unseen mazes.  Refill 0 builds everything."""
        return None if getattr(self, "_built_engine", None) != id(self.engine) else self.engine.consumed_slots()   # a new engine's pool is empty

    def prefetch_pool(self):
        """Start building the NEXT refill (the slots this rollout consumed) on a side stream, in place.  Only for callers that reset every env
        before they step again (PPO.get_batch does, PPO.py:104): the consumed slots include the mazes of the episodes still open."""
        if self.engine is None or getattr(self, "_staged", None) == (id(self.engine), self._generation) or getattr(self, "_built_engine", None) != id(self.engine):
            return
        self.engine.generate_background(self._refill_mask(), **self._pool_args())
        self._staged = (id(self.engine), self._generation)

    def refill_pool(self):
        """Refill the pool (K1): everything the first time, afterwards the consumed slots only.  Only between episodes of ALL envs: a live episode
        reads its maze's exit field from the pool."""
        eng = self.engine
        if getattr(self, "_staged", None) == (id(eng), self._generation):
            eng.wait_background()     # built ahead by prefetch_pool(): same seed, same slots, same mazes
        else:
            eng.generate(only=self._refill_mask(), **self._pool_args())
        self._staged = None
        self._built_engine = id(eng)
        eng.env_episode.zero_()
        self._generation += 1
        self._resets_since_fill = 0

    # ------------------------------------------------------------------ reference API
    def reset(self, mask: Optional[torch.Tensor] = None, obs: Optional[torch.Tensor] = None, masks: Optional[torch.Tensor] = None):
        eng = self._ensure_engine()
        if mask is None:
            self._resets_since_fill += 1
            if self._resets_since_fill > self.pool_episodes:
                self.refill_pool(); self._resets_since_fill = 1
        self._obs, self._masks = eng.reset(mask, obs=obs, masks=masks)
        self.exit_found = False
        return self._emit(self._obs, self._masks)

    def build_maze(self):
        """Maze.build_maze() (maze.py:170-218): every env gets its next maze.  The reference's grid and agents are separate python objects; here a
        maze's working copy and its agents are one device state, so taking a new maze also places the agents on shortest_path[0:2] -- i.e. this is
        reset() without the return value (the reference only ever calls build_maze from reset, maze.py:57).  Generation itself is K1
        (csrc/mm_generate.cu): get_neighbors / set_start / set_end / set_key / get_shortest_path (maze.py:220-273) are stages of that kernel."""
        self.reset()

    def step(self, action, auto_reset: Optional[bool] = None, **out):
        eng = self._ensure_engine()
        if self.num_envs == 1 and not torch.is_tensor(action):
            io = self._single_io()   # the reference's list interface: the four action bytes travel through one pinned buffer, no allocation per step
            h = io["h_act_np"]
            h[0], h[1], h[2], h[3] = int(action[0][0]), int(action[0][1]), int(action[1][0]), int(action[1][1])
            io["d_act"].copy_(io["h_act"], non_blocking=True)
            action = io["d_act"]
        elif torch.is_tensor(action) and (action.dtype != torch.uint8 or not action.is_contiguous() or action.device.type != self.device.type):
            action = action.to(device=self.device, dtype=torch.uint8).reshape(self.num_envs, 2, 2).contiguous()  # e.g. the float [E,2,2] of PPO.get_batch
        auto = (self.num_envs > 1) if auto_reset is None else auto_reset
        self._obs, self._masks, r, d = eng.step(action, auto_reset=auto, **out)
        if self.num_envs == 1:
            o, m = self._emit(self._obs, self._masks, r, d)
            io = self._io
            return o, m, float(io["h_rew_np"][0]), bool(io["h_done_np"][0])
        return self._obs, self._masks, r, d

    def _single_io(self):
        """Pinned host mirrors of ONE environment's step inputs / outputs (num_envs == 1: the reference's python-list interface).  A step is then one
        small host-to-device copy, the kernel, four small device-to-host copies and ONE stream synchronisation."""
        io = getattr(self, "_io", None)
        if io is None:
            pin = self.device.type == "cuda"
            mk = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, pin_memory=pin)
            io = dict(h_act=mk(1, 2, 2, dtype=torch.uint8), h_obs=mk(2, 65, dtype=torch.float32), h_masks=mk(2, 6, dtype=torch.uint8),
                      h_rew=mk(1, dtype=torch.float32), h_done=mk(1, dtype=torch.uint8))
            io["d_act"] = torch.zeros(1, 2, 2, dtype=torch.uint8, device=self.device)
            for k in ("act", "obs", "masks", "rew", "done"):
                io[f"h_{k}_np"] = io[f"h_{k}"].numpy()
            io["h_act_np"] = io["h_act_np"].reshape(-1)   # a view: the four action bytes
            self._io = io
        return io

    def _emit(self, obs, masks, reward=None, done=None):
        if self.num_envs == 1:
            io = self._single_io()
            io["h_obs"].copy_(obs[0], non_blocking=True); io["h_masks"].copy_(masks[0], non_blocking=True)
            if reward is not None:
                io["h_rew"].copy_(reward.reshape(1), non_blocking=True); io["h_done"].copy_(done.reshape(1), non_blocking=True)
            brain = self.agents[0].brain
            bdev = getattr(brain, "device", None)   # torch.device("cuda") and torch.device("cuda:0") do not compare equal
            pre = (bool(getattr(self, "_prefetch_policy", False)) and bdev is not None and bdev.type == obs.device.type == "cuda"
                   and bdev.index in (None, obs.device.index) and obs.is_contiguous())
            if pre:     # a get_action loop (maze.py:477-493): both agents' logits ride along with the observation (PPO._prefetch_logits)
                brain._prefetch_logits(obs)
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
            rows = io["h_obs_np"].tolist()
            if pre:
                brain._commit_logits(rows)
            return rows, [[bool(v) for v in row] for row in io["h_masks_np"].tolist()]
        return obs, masks

    def is_valid_cell(self, x, y):
        return 0 <= x < self.width and 0 <= y < self.height

    # ------------------------------------------------------------------ views used by Agent and by callers of the reference attrs
    def _agent_state(self, agent):
        return self.engine.agents()[agent.env, agent._index]

    def _last_obs_of(self, agent):
        o, m = self._obs[agent.env, agent._index], self._masks[agent.env, agent._index]
        return o.tolist(), [bool(v) for v in m.tolist()]

    def _env_row(self, e=0):
        return self.engine.envs()[e]

    def _pool(self, e=0):
        return self.engine.pool_maze(int(self._env_row(e)[3]))

    current_t = property(lambda s: int(s._env_row()[0]))
    width = property(lambda s: int(s._env_row()[4]) if s.engine is not None else s.side_range[1] * 2 - 1)
    height = property(lambda s: int(s._env_row()[5]) if s.engine is not None else s.side_range[1] * 2 - 1)
    start = property(lambda s: s._pool()["start"])
    end = property(lambda s: s._pool()["end"])
    shortest_path_len = property(lambda s: s._pool()["shortest_path_len"])

    @property
    def key(self):  # tuple, or 0 once picked up (maze.py:158)
        r = self._env_row()
        return 0 if r[1] < 0 else (int(r[1]), int(r[2]))

    @property
    def layout(self):
        return self.engine.layout(0)[:self.height, :self.width].tolist()

    @property
    def shortest_path(self):
        """start -> exit path, recovered by following the dir-to-exit field (get_shortest_path, maze.py:261-273)."""
        m = self._pool()
        (x, y), path = m["start"], [m["start"]]
        while (x, y) != m["end"] and len(path) < 4096:
            dx, dy = DELTAS[int(m["d2e"][y][x])]
            x, y = x + dx, y + dy
            path.append((x, y))
        return path

    @property
    def agent_positions(self):
        pos = {}
        for a in self.agents:
            pos.setdefault((a.x, a.y), []).append(a)
        return pos

    def render_ascii(self, e: int = 0) -> str:
        eng = self.engine
        row = eng.envs()[e]; W, H = int(row[4]), int(row[5])
        lay = eng.layout(e)[:H, :W]
        m = eng.pool_maze(int(row[3])); ag = eng.agents()[e]
        ch = {0: " ", 1: "#", 2: "r", 3: "b"}
        g = [[ch[int(v)] for v in r] for r in lay]
        g[m["end"][1]][m["end"][0]] = "E"
        if row[1] >= 0:
            g[int(row[2])][int(row[1])] = "K"
        for i, c in enumerate("RB"):
            g[int(ag[i][1])][int(ag[i][0])] = c
        return "\n".join("".join(r) for r in g)
