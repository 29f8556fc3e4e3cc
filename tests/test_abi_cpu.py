"""CPU-side checks of the C-ABI library: it builds, loads, and exports every symbol include/marl_maze_b200.h declares.
No compute call is made here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "marl_maze_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from marl_maze_b200 import _abi, build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/marl_maze_b200.h but not exported"
    assert sorted(_abi.EXPORTS) == names, "marl_maze_b200/_abi.py EXPORTS out of sync with the header"


def test_size_helpers_and_error_strings():
    from marl_maze_b200 import _abi
    L = _abi.lib()
    assert L.mm_abi_version() == 1
    assert L.mm_sizeof_env_grid(4, 25) == 4 * 35 * 16 and L.mm_sizeof_pool_d2e(3, 49) == 3 * 49 * 16
    assert L.mm_sizeof_agent_a(10) == 320 and L.mm_sizeof_agent_b(10) == 80 and L.mm_sizeof_env_hdr(7) == 112
    assert L.mm_error_string(0) == b"ok" and L.mm_error_string(1) == b"bad argument"
    # bad arguments are rejected before any CUDA call
    assert L.mm_reset(None, None, None, None, None) == 1
    assert L.mm_gae(None, None, None, None, None, None, 1, 1, 0.99, 0.95, None) == 1


def test_loader_detects_a_stale_library():
    """a library compiled from other sources than the ones present must not be used silently -- the compiled-in content hash
    (mm_source_hash) is compared with the sources' hash on load."""
    from marl_maze_b200 import _abi, build
    build.build()
    assert build.built_hash() == build.source_hash() and not build.needs_build()
    assert _abi.lib().mm_source_hash().decode() == build.source_hash()
    real = build.source_hash
    try:
        build.source_hash = lambda: "0" * 64          # pretend a source changed
        assert build.needs_build()
    finally:
        build.source_hash = real


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "marl_maze_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), f"{f} imports oracle/"
                assert "maze_oracle" not in s or f.endswith((".cu", ".cuh")), f


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from marl_maze_b200 import MazeEngine, _abi
    with pytest.raises(_abi.MMError):
        MazeEngine(4)
