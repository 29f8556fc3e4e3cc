"""Oracle generator under the Philox stream (the stream K1 uses): structural properties, determinism."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import OracleMaze


def _props(m):
    lay = np.asarray(m["layout"]); H, W = lay.shape
    open_ = lay == 0
    edges = int((open_[:, :-1] & open_[:, 1:]).sum() + (open_[:-1, :] & open_[1:, :]).sum())
    assert edges == int(open_.sum()) - 1
    sx, sy = m["start"]; ex, ey = m["end"]; kx, ky = m["key"]
    assert sx % 2 == 0 and sy % 2 == 0 and ex in (0, W - 1) and open_[ey, ex] and (ex, ey) != (sx, sy)
    path = [tuple(p) for p in m["path"].tolist()]
    assert path[0] == (sx, sy) and path[-1] == (ex, ey) and len(path) == m["shortest_path_len"]
    assert all(abs(a[0] - b[0]) + abs(a[1] - b[1]) == 1 for a, b in zip(path, path[1:]))
    assert open_[ky, kx] and (kx, ky) not in path


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**63), mid=st.integers(0, 2**32 - 1), lo=st.integers(4, 27), span=st.integers(0, 3),
       rs=st.booleans(), diff=st.integers(1, 4))
def test_philox_generator_properties(seed, mid, lo, span, rs, diff):
    hi = min(27, lo + span)
    o = OracleMaze(max_timestep=10, difficulty=diff, rand_start=rs, rand_sizes=True, rand_range=(lo, hi), default_size=(4, 4))
    o.seed_philox(seed, mid); o.build(); m = o.maze()
    assert m["width"] == m["height"] and lo * 2 - 1 <= m["width"] <= hi * 2 - 1
    _props(m)
    o2 = OracleMaze(max_timestep=10, difficulty=diff, rand_start=rs, rand_sizes=True, rand_range=(lo, hi), default_size=(4, 4))
    o2.seed_philox(seed, mid); o2.build()
    assert np.array_equal(o2.maze()["layout"], m["layout"]) and o2.maze()["key"] == m["key"]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_mt_generator_properties(seed):
    o = OracleMaze(max_timestep=10, difficulty=1, rand_start=True, rand_sizes=True, rand_range=(12, 13), default_size=(4, 4))
    o.seed(seed)
    for _ in range(5):
        o.build(); _props(o.maze())
