"""Readers for tests/golden/*.npz (recorded from the unmodified reference by tools/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAXS = 64


def _maze_from(hdr, packed):
    W, H = int(hdr[0]), int(hdr[1])
    walls = np.unpackbits(packed, axis=-1)[:H, :W].astype(np.uint8)
    return dict(width=W, height=H, layout=walls, path0=(int(hdr[2]), int(hdr[3])), path1=(int(hdr[4]), int(hdr[5])),
                end=(int(hdr[6]), int(hdr[7])), key=(int(hdr[8]), int(hdr[9])), shortest_path_len=int(hdr[10]),
                start=(int(hdr[2]), int(hdr[3])))


class Trace:
    def __init__(self, z, name):
        g = lambda k: z[f"{name}/{k}"]
        self.name = name
        self.mazes = [_maze_from(h, l) for h, l in zip(g("maze_hdr"), g("maze_layout"))]
        self.actions = g("actions")
        self.step_obs = g("step_obs")
        self.step_masks = g("step_masks")
        self.reward = g("reward")
        self.done = g("done")
        self.agents_after = g("agents_after")
        self.reset_obs = g("reset_obs")
        self.reset_masks = g("reset_masks")
        self.reset_agents = g("reset_agents")
        self.max_timestep = int(g("max_timestep"))
        cfg = g("cfg")
        self.maze_seed, self.action_seed, self.n = int(cfg[0]), int(cfg[1]), int(cfg[2])
        self.maze_kw = dict(max_timestep=int(cfg[3]), difficulty=int(cfg[4]), rand_start=bool(cfg[5]), rand_sizes=bool(cfg[6]),
                            rand_range=(int(cfg[7]), int(cfg[8])), default_size=(int(cfg[9]), int(cfg[10])))
        self.vision = tuple(int(v) for v in z[f"{name}/vision"]) if f"{name}/vision" in z.files else (4, 4)   # Agent(vision_range=...)


_cache = {}


def load_traces():
    if "t" not in _cache:
        z = np.load(os.path.join(GOLDEN, "env_traces.npz"))
        _cache["t"] = {str(n): Trace(z, str(n)) for n in z["names"]}
        _cache["kat2"] = bytes(z["kat2_sha256"]).hex()
    return _cache["t"]


def load_vision_traces():
    """Reference traces with vision_range != 4 (tools/make_golden.py --vision), SURVEY 8(f).4."""
    if "v" not in _cache:
        z = np.load(os.path.join(GOLDEN, "env_traces_vision.npz"))
        _cache["v"] = {str(n): Trace(z, str(n)) for n in z["names"]}
    return _cache["v"]


def kat2_sha256():
    load_traces()
    return _cache["kat2"]


def load_gen_kats():
    z = np.load(os.path.join(GOLDEN, "gen_kats.npz"))
    out = {}
    for key in z["keys"]:
        key = str(key)
        cfg = z[key + "/cfg"]
        out[key] = dict(seed=int(key.split("/")[1]), hdr=z[key + "/hdr"], layout=z[key + "/layout"], path_sha=bytes(z[key + "/path_sha"]),
                        kw=dict(difficulty=int(cfg[0]), rand_start=bool(cfg[1]), rand_sizes=bool(cfg[2]), rand_range=(int(cfg[3]), int(cfg[4])),
                                default_size=(int(cfg[5]), int(cfg[6]))))
    return out
