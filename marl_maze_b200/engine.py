"""MazeEngine: the batched environment state in HBM plus the kernel launches that advance it.

Torch is plumbing here: it owns the device buffers and the stream; every computation is a kernel of
libmarl_maze_b200.so reached through the C ABI (include/marl_maze_b200.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _abi

OBS_DIM = 65
MASK_DIM = 6
PAD = 5
MAX_SIDE = 54
AGENT_FIELDS = ["x", "y", "direction", "knows_end", "other_knows_end", "has_key", "team_has_key", "exit_len",
                "time_from_last_seen", "ols_x", "ols_y", "lm_x", "lm_y", "min_x", "max_x", "min_y", "max_y", "route_len"]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class MazeEngine:
    """E two-agent mazes of side <= smax stepped in lock-step on one GPU.

    Episodes draw their mazes from a pool of `pool_size` pre-built mazes (episode k of env e uses pool maze
    (e + k*E) mod P); the pool is filled either by `generate()` (K1) or by `load_layouts()` (parity injection).
    """

    def __init__(self, num_envs: int, smax: int = 25, max_timestep: int = 1200, pool_size: Optional[int] = None,
                 device: str | torch.device = "cuda", env_offset: int = 0, vision=(4, 4)):
        self.lib = _abi.lib()
        if not torch.cuda.is_available():
            raise _abi.MMError("MazeEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        if not (3 <= smax <= MAX_SIDE):
            raise ValueError(f"smax must be in [3, {MAX_SIDE}]")
        if len(vision) != 2 or not all(1 <= int(v) <= 4 for v in vision):
            raise ValueError("vision = (agent 0's, agent 1's) vision_range, each 1..4 (the stored grids keep a wall border of vision_range + 1 = 5 cells)")
        self.vision = (int(vision[0]), int(vision[1]))
        self.device = torch.device(device)
        self.E, self.smax, self.max_timestep = int(num_envs), int(smax), int(max_timestep)
        self.P = int(pool_size) if pool_size else self.E
        self.rows = self.smax + 2 * PAD
        L = self.lib
        u8 = lambda n: torch.zeros(int(n), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self.pool_grid = u8(L.mm_sizeof_pool_grid(self.P, self.smax))
            self.pool_d2e = u8(L.mm_sizeof_pool_d2e(self.P, self.smax))
            self.pool_hdr = u8(L.mm_sizeof_pool_hdr(self.P))
            self.env_grid = u8(L.mm_sizeof_env_grid(self.E, self.smax))
            self.env_hdr = u8(L.mm_sizeof_env_hdr(self.E))
            self.env_episode = u8(L.mm_sizeof_env_episode(self.E))
            self.agent_a = u8(L.mm_sizeof_agent_a(self.E))
            self.agent_b = u8(L.mm_sizeof_agent_b(self.E))
        self.st = _abi.MMState(self.pool_grid.data_ptr(), self.pool_d2e.data_ptr(), self.pool_hdr.data_ptr(),
                               self.env_grid.data_ptr(), self.env_hdr.data_ptr(), self.env_episode.data_ptr(),
                               self.agent_a.data_ptr(), self.agent_b.data_ptr(),
                               self.E, self.P, self.smax, self.max_timestep, int(env_offset), self.vision[0] | (self.vision[1] << 8))
        self.obs = torch.zeros(self.E, 2, OBS_DIM, dtype=torch.float32, device=self.device)
        self.masks = torch.zeros(self.E, 2, MASK_DIM, dtype=torch.uint8, device=self.device)
        self.reward = torch.zeros(self.E, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(self.E, dtype=torch.uint8, device=self.device)
        self._scratch = None
        self.launches = 0
        self._call = 0
        _abi.check(L.mm_init_state(C.byref(self.st), self._stream()), "mm_init_state")
        self.launches += 1

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _get_scratch(self, nbytes: int) -> torch.Tensor:
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._scratch

    def state_bytes(self) -> int:
        return sum(t.numel() for t in (self.pool_grid, self.pool_d2e, self.pool_hdr, self.env_grid, self.env_hdr,
                                       self.env_episode, self.agent_a, self.agent_b))

    # ------------------------------------------------------------------ pool
    def load_layouts(self, first: int, mazes: Sequence[dict], chunk: int = 4096):
        """Inject mazes recorded from the reference / the oracle (dicts with layout, path0, path1, end, key, shortest_path_len)."""
        S = self.smax
        for lo in range(0, len(mazes), chunk):
            part = mazes[lo:lo + chunk]
            n = len(part)
            lay = np.ones((n, S, S), np.uint8)
            hdr = np.zeros((n, 11), np.int32)
            for i, m in enumerate(part):
                a = np.asarray(m["layout"], np.uint8)
                H, W = a.shape
                if H > S or W > S:
                    raise ValueError(f"maze {W}x{H} does not fit smax={S}")
                lay[i, :H, :W] = (a == 1)
                hdr[i] = [W, H, m["path0"][0], m["path0"][1], m["path1"][0], m["path1"][1], m["end"][0], m["end"][1],
                          m["key"][0], m["key"][1], m["shortest_path_len"]]
            d_lay = torch.from_numpy(lay).to(self.device)
            d_hdr = torch.from_numpy(hdr).to(self.device)
            scratch = self._get_scratch(self.lib.mm_sizeof_finalize_scratch(n, S))
            _abi.check(self.lib.mm_load_layouts(C.byref(self.st), first + lo, n, _ptr(d_lay), _ptr(d_hdr), _ptr(scratch), self._stream()),
                       "mm_load_layouts")
            self.launches += 1
            torch.cuda.current_stream(self.device).synchronize()  # d_lay / d_hdr die here

    def generate(self, seed: int, side_range=(13, 13), rand_start: bool = True, difficulty: int = 1, first: int = 0,
                 count: Optional[int] = None, id_base: int = 0, id_mod: int = 0, id_mul: int = 0, only: Optional[torch.Tensor] = None,
                 max_blocks: int = 0, height_cells: int = 0):
        """K1: fill pool entries [first, first+count) with freshly generated mazes (maze.py:170-273).  only: u8 [count], build slot i only where
        only[i] != 0 (the incremental refill); max_blocks: cap on the thread blocks in flight (a background build on a side stream); height_cells > 0:
        rectangular mazes of (2 * side_range[0] - 1) x (2 * height_cells - 1) cells, Maze(default_size=[w, h]) with rand_sizes False."""
        n = self.P - first if count is None else count
        if only is not None:
            assert only.dtype == torch.uint8 and only.is_contiguous() and only.numel() == n and only.device == self.pool_hdr.device
        scratch = self._get_scratch(self.lib.mm_sizeof_generate_scratch(n, self.smax))
        _abi.check(self.lib.mm_generate_masked(C.byref(self.st), first, n, int(side_range[0]), int(side_range[1]), int(rand_start), int(difficulty),
                                               C.c_uint64(seed & (2**64 - 1)), C.c_uint32(id_base & 0xFFFFFFFF), int(id_mod), int(id_mul), _ptr(scratch),
                                               int(max_blocks), _ptr(only), int(height_cells), self._stream()), "mm_generate_masked")
        self.launches += 1

    def consumed_slots(self) -> torch.Tensor:
        """u8 [P]: 1 for the pool slots whose maze an episode has started on since the last refill (slot p = e + k*E is episode k of env e)."""
        K = self.P // self.E
        ep = self.env_episode.view(torch.int32)
        return (torch.arange(K, device=self.device, dtype=torch.int32).view(K, 1) < ep.view(1, self.E)).to(torch.uint8).reshape(-1).contiguous()

    def generate_background(self, only: Optional[torch.Tensor], **kw):
        """K1 on a side stream, in place (the slots in `only` are not read again before the next refill): the serial carve overlaps whatever the caller
        does next (the PPO update).  `wait_background()` orders the current stream behind it."""
        if getattr(self, "_bg", None) is None:
            self._bg = dict(stream=torch.cuda.Stream(self.device), event=torch.cuda.Event())
        bg = self._bg
        bg["stream"].wait_stream(torch.cuda.current_stream(self.device))   # the mask, and everything that still reads the consumed slots
        if only is not None:
            only.record_stream(bg["stream"])
        with torch.cuda.stream(bg["stream"]):
            # at most 3 blocks of 64 carving threads per SM: each holds its residency slot for milliseconds, and a full grid would keep the
            # kernels of the main stream waiting for slots
            self.generate(only=only, max_blocks=148 * 3, **kw)
            bg["event"].record(bg["stream"])

    def wait_background(self):
        torch.cuda.current_stream(self.device).wait_event(self._bg["event"])

    # ------------------------------------------------------------------ env API
    def reset(self, mask: Optional[torch.Tensor] = None, obs: Optional[torch.Tensor] = None, masks: Optional[torch.Tensor] = None):
        """Maze.reset() for the masked envs (all when mask is None).  Returns (obs [E,2,65] f32, masks [E,2,6] u8)."""
        obs = self.obs if obs is None else obs
        masks = self.masks if masks is None else masks
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _abi.check(self.lib.mm_reset(C.byref(self.st), _ptr(mask), _ptr(obs), _ptr(masks), self._stream()), "mm_reset")
        self.launches += 1
        return obs, masks

    def place_agent(self, env: int, agent: int, x: int, y: int, direction: int = 2, reset: bool = False):
        """Agent.reset(x, y) (reset=True) or Agent.move(x, y, direction) of one agent (maze_agent.py:59-87) on the packed device state."""
        _abi.check(self.lib.mm_agent_place(C.byref(self.st), int(env), int(agent), int(x), int(y), int(direction), int(reset), self._stream()), "mm_agent_place")
        self.launches += 1

    def step(self, actions: Optional[torch.Tensor], auto_reset: bool = True, obs: Optional[torch.Tensor] = None,
             masks: Optional[torch.Tensor] = None, reward: Optional[torch.Tensor] = None, done: Optional[torch.Tensor] = None,
             action_seed: int = 0, actions_out: Optional[torch.Tensor] = None):
        """Maze.step(action) for all envs.  actions [E,2,2] u8 (move, mark); None = kernel draws uniform legal actions."""
        obs = self.obs if obs is None else obs
        masks = self.masks if masks is None else masks
        reward = self.reward if reward is None else reward
        done = self.done if done is None else done
        if actions is not None:
            if actions.dtype != torch.uint8 or not actions.is_contiguous() or actions.numel() != self.E * 4 or actions.device != obs.device:
                raise ValueError("actions must be a contiguous uint8 device tensor of shape [E,2,2]")
        self._call += 1
        seed = (int(action_seed) * 0x9E3779B97F4A7C15 + self._call) & (2**64 - 1)
        _abi.check(self.lib.mm_step_obs(C.byref(self.st), _ptr(actions), _ptr(obs), _ptr(masks), _ptr(reward), _ptr(done),
                                        int(auto_reset), C.c_uint64(seed), _ptr(actions_out), self._stream()), "mm_step_obs")
        self.launches += 1
        return obs, masks, reward, done

    # ------------------------------------------------------------------ readback (tests / debugging)
    def agents(self) -> np.ndarray:
        out = torch.zeros(self.E, 2, len(AGENT_FIELDS), dtype=torch.int32, device=self.device)
        _abi.check(self.lib.mm_unpack_agents(C.byref(self.st), _ptr(out), self._stream()), "mm_unpack_agents")
        return out.cpu().numpy()

    def envs(self) -> np.ndarray:
        """[E,8] = t, key_x, key_y (-1 when taken), pool index, W, H, err, episode"""
        out = torch.zeros(self.E, 8, dtype=torch.int32, device=self.device)
        _abi.check(self.lib.mm_unpack_envs(C.byref(self.st), _ptr(out), self._stream()), "mm_unpack_envs")
        return out.cpu().numpy()

    def layout(self, e: int) -> np.ndarray:
        out = torch.zeros(self.smax, self.smax, dtype=torch.uint8, device=self.device)
        _abi.check(self.lib.mm_unpack_layout(C.byref(self.st), int(e), _ptr(out), self._stream()), "mm_unpack_layout")
        return out.cpu().numpy()

    def pool_maze(self, p: int) -> dict:
        lay = torch.zeros(self.smax, self.smax, dtype=torch.uint8, device=self.device)
        d2e = torch.zeros(self.smax, self.smax, dtype=torch.uint8, device=self.device)
        hdr = torch.zeros(11, dtype=torch.int32, device=self.device)
        _abi.check(self.lib.mm_unpack_pool(C.byref(self.st), int(p), _ptr(lay), _ptr(d2e), _ptr(hdr), self._stream()), "mm_unpack_pool")
        h = hdr.cpu().numpy()
        W, H = int(h[0]), int(h[1])
        return dict(width=W, height=H, layout=lay.cpu().numpy()[:H, :W], d2e=d2e.cpu().numpy()[:H, :W], path0=(int(h[2]), int(h[3])),
                    path1=(int(h[4]), int(h[5])), start=(int(h[2]), int(h[3])), end=(int(h[6]), int(h[7])), key=(int(h[8]), int(h[9])),
                    shortest_path_len=int(h[10]))


def gae(reward: torch.Tensor, value: torch.Tensor, done: torch.Tensor, v_boot: Optional[torch.Tensor], gamma: float = 0.99,
        lam: float = 0.95, with_rtg: bool = False, out: Optional[torch.Tensor] = None):
    """K3: PPO.get_GAEs over [T,E] buffers.  Returns adv (and rtg = adv + value when asked)."""
    T, E = reward.shape
    L = _abi.lib()
    adv = torch.empty_like(reward) if out is None else out
    rtg = torch.empty_like(reward) if with_rtg else None
    done = done.to(torch.uint8) if done.dtype != torch.uint8 else done
    for t in (reward, value, done):
        assert t.is_contiguous() and t.is_cuda
    _abi.check(L.mm_gae(_ptr(reward), _ptr(value), _ptr(done), _ptr(v_boot), _ptr(adv), _ptr(rtg), T, E, float(gamma), float(lam),
                        C.c_void_p(torch.cuda.current_stream(reward.device).cuda_stream)), "mm_gae")
    return (adv, rtg) if with_rtg else adv
