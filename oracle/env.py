"""ctypes face of oracle/maze_oracle.c.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import build_oracle, LIB

OBS_DIM = 65
N_AGENT_FIELDS = 18
AGENT_FIELDS = ["x", "y", "direction", "knows_end", "other_knows_end", "has_key", "team_has_key", "exit_len",
                "time_from_last_seen", "ols_x", "ols_y", "lm_x", "lm_y", "min_x", "max_x", "min_y", "max_y",
                "route_len"]

_lib = None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load_lib():
    global _lib
    if _lib is not None:
        return _lib
    path = LIB
    src_present = os.path.exists(os.path.join(os.path.dirname(path), "maze_oracle.c"))
    if src_present:
        try:
            path = build_oracle()
        except Exception:
            if not os.path.exists(path):
                raise
    lib = C.CDLL(path)
    vp, i32, u64, u32 = C.c_void_p, C.c_int, C.c_uint64, C.c_uint32
    pf, pu8, pi = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    sig = {
        "omaze_new": (vp, [i32] * 8),
        "omaze_free": (None, [vp]),
        "omaze_set_vision": (None, [vp, i32, i32]),
        "omaze_agent_reset": (None, [vp, i32, i32, i32]),
        "omaze_agent_move": (None, [vp, i32, i32, i32, i32]),
        "obatch_set_vision": (None, [vp, i32, i32]),
        "omaze_seed": (None, [vp, u64]),
        "omaze_seed_philox": (None, [vp, u64, u32]),
        "omaze_build": (None, [vp]),
        "omaze_reset": (None, [vp, pf, pu8]),
        "omaze_reset_injected": (None, [vp, i32, i32, pu8] + [i32] * 9 + [pf, pu8]),
        "omaze_step": (None, [vp, pi, pf, pu8, pf, pu8]),
        "omaze_get_maze": (None, [vp, pi, pu8]),
        "omaze_get_path": (i32, [vp, pi, i32]),
        "omaze_get_agents": (None, [vp, pi]),
        "omaze_error": (i32, [vp]),
        "obatch_new": (vp, [i32, i32, i32]),
        "obatch_free": (None, [vp]),
        "obatch_set_pool_maze": (None, [vp, i32, i32, i32, pu8] + [i32] * 9),
        "obatch_generate_pool_maze": (None, [vp, i32, i32, i32, i32, u64, u32]),
        "obatch_reset_all": (None, [vp, pf, pu8, i32]),
        "obatch_reset_masked": (None, [vp, pu8, pf, pu8]),
        "obatch_step": (None, [vp, pu8, pf, pu8, pf, pu8, i32, i32]),
        "obatch_get_agents": (None, [vp, pi]),
        "obatch_get_env": (None, [vp, pi]),
        "obatch_get_layout": (None, [vp, i32, pu8]),
        "obatch_env": (vp, [vp, i32]),
        "obatch_errors": (i32, [vp]),
        "obatch_random_actions": (None, [vp, pu8, pu8, C.POINTER(C.c_uint64)]),
        "obatch_run_random": (C.c_long, [vp, i32, i32, u64, pf, pu8, C.POINTER(C.c_double)]),
        "obatch_run_random_ex": (C.c_long, [vp, i32, i32, i32, u64, pf, pu8, C.POINTER(C.c_double)]),
        "obatch_guided_actions": (None, [vp, pu8, pu8, C.POINTER(C.c_uint64), i32, i32, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


class OracleMaze:
    """One literal reference environment: Maze(agents=(tag 2, tag 3), ...) of maze.py:22."""

    def __init__(self, max_timestep=3500, difficulty=1, rand_start=False, rand_sizes=False, rand_range=(6, 12),
                 default_size=(8, 8), _handle=None, vision=(4, 4)):
        self.lib = load_lib()
        self._own = _handle is None
        self.h = _handle or self.lib.omaze_new(max_timestep, difficulty, int(rand_start), int(rand_sizes),
                                               rand_range[0], rand_range[1], default_size[0], default_size[1])
        if _handle is None and tuple(vision) != (4, 4):   # Agent(..., vision_range=r), maze_agent.py:16
            self.lib.omaze_set_vision(self.h, int(vision[0]), int(vision[1]))
        self.obs = np.zeros((2, OBS_DIM), np.float32)
        self.masks = np.zeros((2, 6), np.uint8)

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            self.lib.omaze_free(self.h)
            self.h = None

    def agent_reset(self, a: int, x: int, y: int):   # Agent.reset(x, y), maze_agent.py:59-79
        self.lib.omaze_agent_reset(self.h, a, x, y)

    def agent_move(self, a: int, x: int, y: int, direction: int):   # Agent.move(x, y, direction), maze_agent.py:85-87
        self.lib.omaze_agent_move(self.h, a, x, y, direction)

    def seed(self, s: int):  # == random.seed(s) before Maze.reset()
        self.lib.omaze_seed(self.h, s)

    def seed_philox(self, s: int, maze_id: int):
        self.lib.omaze_seed_philox(self.h, s, maze_id)

    def build(self):
        self.lib.omaze_build(self.h)

    def reset(self):
        self.lib.omaze_reset(self.h, _p(self.obs, C.c_float), _p(self.masks, C.c_uint8))
        return self.obs.copy(), self.masks.copy()

    def reset_injected(self, maze: dict):
        lay = np.ascontiguousarray(np.asarray(maze["layout"], np.uint8))
        H, W = lay.shape
        self.lib.omaze_reset_injected(self.h, W, H, _p(lay, C.c_uint8), maze["path0"][0], maze["path0"][1],
                                      maze["path1"][0], maze["path1"][1], maze["end"][0], maze["end"][1],
                                      maze["key"][0], maze["key"][1], maze["shortest_path_len"],
                                      _p(self.obs, C.c_float), _p(self.masks, C.c_uint8))
        return self.obs.copy(), self.masks.copy()

    def step(self, action):
        act = np.asarray(action, np.int32).reshape(4).copy()
        r = C.c_float()
        d = C.c_uint8()
        self.lib.omaze_step(self.h, _p(act, C.c_int), _p(self.obs, C.c_float), _p(self.masks, C.c_uint8),
                            C.byref(r), C.byref(d))
        return self.obs.copy(), self.masks.copy(), float(r.value), bool(d.value)

    def maze(self) -> dict:
        hdr = np.zeros(12, np.int32)
        self.lib.omaze_get_maze(self.h, _p(hdr, C.c_int), None)
        W, H = int(hdr[0]), int(hdr[1])
        lay = np.zeros((H, W), np.uint8)
        self.lib.omaze_get_maze(self.h, _p(hdr, C.c_int), _p(lay, C.c_uint8))
        path = np.zeros((W * H, 2), np.int32)
        n = self.lib.omaze_get_path(self.h, _p(path, C.c_int), W * H)
        return dict(width=W, height=H, layout=lay, start=(int(hdr[2]), int(hdr[3])), end=(int(hdr[4]), int(hdr[5])),
                    key=(int(hdr[6]), int(hdr[7])), shortest_path_len=int(hdr[8]), path0=(int(hdr[2]), int(hdr[3])),
                    path1=(int(hdr[9]), int(hdr[10])), current_t=int(hdr[11]), path=path[:n].copy())

    def agents(self) -> np.ndarray:
        out = np.zeros((2, N_AGENT_FIELDS), np.int32)
        self.lib.omaze_get_agents(self.h, _p(out, C.c_int))
        return out

    def error(self) -> int:
        return self.lib.omaze_error(self.h)


class OracleBatch:
    """E literal environments behind the same host contract as the CUDA path (maze pool + auto-reset)."""

    def __init__(self, E: int, P: int, max_timestep: int = 1200, threads: int = 1, vision=(4, 4)):
        self.lib = load_lib()
        self.E, self.P, self.threads = E, P, threads
        self.h = self.lib.obatch_new(E, P, max_timestep)
        if tuple(vision) != (4, 4):
            self.lib.obatch_set_vision(self.h, int(vision[0]), int(vision[1]))
        self.obs = np.zeros((E, 2, OBS_DIM), np.float32)
        self.masks = np.zeros((E, 2, 6), np.uint8)
        self.reward = np.zeros(E, np.float32)
        self.done = np.zeros(E, np.uint8)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.obatch_free(self.h)
            self.h = None

    def set_pool_maze(self, p: int, maze: dict):
        lay = np.ascontiguousarray(np.asarray(maze["layout"], np.uint8))
        H, W = lay.shape
        self.lib.obatch_set_pool_maze(self.h, p, W, H, _p(lay, C.c_uint8), maze["path0"][0], maze["path0"][1],
                                      maze["path1"][0], maze["path1"][1], maze["end"][0], maze["end"][1],
                                      maze["key"][0], maze["key"][1], maze["shortest_path_len"])

    def generate_pool_maze(self, p: int, S: int, rand_start: bool, difficulty: int, seed: int, maze_id: int):
        self.lib.obatch_generate_pool_maze(self.h, p, S, int(rand_start), difficulty, seed, maze_id)

    def reset_all(self):
        self.lib.obatch_reset_all(self.h, _p(self.obs, C.c_float), _p(self.masks, C.c_uint8), self.threads)
        return self.obs, self.masks

    def reset_masked(self, which):
        w = np.ascontiguousarray(np.asarray(which, np.uint8))
        self.lib.obatch_reset_masked(self.h, _p(w, C.c_uint8), _p(self.obs, C.c_float), _p(self.masks, C.c_uint8))
        return self.obs, self.masks

    def step(self, actions, auto_reset=True):
        a = np.ascontiguousarray(np.asarray(actions, np.uint8).reshape(self.E, 2, 2))
        self.lib.obatch_step(self.h, _p(a, C.c_uint8), _p(self.obs, C.c_float), _p(self.masks, C.c_uint8),
                             _p(self.reward, C.c_float), _p(self.done, C.c_uint8), int(auto_reset), self.threads)
        return self.obs, self.masks, self.reward, self.done

    def agents(self) -> np.ndarray:
        out = np.zeros((self.E, 2, N_AGENT_FIELDS), np.int32)
        self.lib.obatch_get_agents(self.h, _p(out, C.c_int))
        return out

    def env_state(self) -> np.ndarray:
        out = np.zeros((self.E, 4), np.int32)
        self.lib.obatch_get_env(self.h, _p(out, C.c_int))
        return out

    def layout(self, e: int) -> np.ndarray:
        m = OracleMaze(_handle=self.lib.obatch_env(self.h, e)).maze()
        return m["layout"]

    def env(self, e: int) -> OracleMaze:
        return OracleMaze(_handle=self.lib.obatch_env(self.h, e))

    def errors(self) -> int:
        return self.lib.obatch_errors(self.h)

    def random_actions(self, rng_state: np.ndarray) -> np.ndarray:
        act = np.zeros((self.E, 2, 2), np.uint8)
        self.lib.obatch_random_actions(self.h, _p(self.masks, C.c_uint8), _p(act, C.c_uint8), _p(rng_state, C.c_uint64))
        return act

    def guided_actions(self, rng_state: np.ndarray, p_follow: float = 0.8, p_mark: float = 0.3) -> np.ndarray:
        """Test-driver helper (privileged BFS towards key then exit); see maze_oracle.c obatch_guided_actions."""
        act = np.zeros((self.E, 2, 2), np.uint8)
        self.lib.obatch_guided_actions(self.h, _p(self.masks, C.c_uint8), _p(act, C.c_uint8), _p(rng_state, C.c_uint64),
                                       int(p_follow * 1024), int(p_mark * 1024), self.threads)
        return act

    def run_random(self, steps: int, seed: int = 0, stagger: int = 0) -> int:
        """CPU-baseline driver: `steps` uniform-legal-random steps per env with auto-reset, env-major, OpenMP over envs.
        stagger > 0: env e runs (e * 2654435761 mod 2^32) mod stagger steps instead (spreads episode phases).  Returns env-steps run."""
        rs = C.c_double()
        return int(self.lib.obatch_run_random_ex(self.h, int(steps), int(stagger), self.threads, seed, _p(self.obs, C.c_float), _p(self.masks, C.c_uint8), C.byref(rs)))
