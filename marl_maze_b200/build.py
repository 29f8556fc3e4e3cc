"""Build libmarl_maze_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m marl_maze_b200.build        # or __graft_entry__.build()

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmarl_maze_b200.so")
SOURCES = ["mm_abi.cu", "mm_step_obs.cu", "mm_pool.cu", "mm_gae.cu", "mm_generate.cu", "mm_policy.cu", "mm_policy_tc.cu", "mm_linear16.cu", "mm_trunk_fused.cu", "mm_tokens_mma.cu", "mm_tokens_proj.cu", "mm_update_tc.cu", "mm_update.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: marl_maze_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def sources():
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(os.path.dirname(HERE), "include", "marl_maze_b200.h")]
    return srcs, hdrs


def source_hash() -> str:
    """SHA-256 over the names and contents of every source and header the library is built from (compiled in as MM_SRC_HASH)."""
    import hashlib
    h = hashlib.sha256()
    srcs, hdrs = sources()
    for f in sorted(srcs + hdrs):
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def built_hash(path: str | None = None) -> str | None:
    """The hash a built library reports (mm_source_hash), or None when it cannot be loaded / predates the export."""
    import ctypes
    try:
        lib = ctypes.CDLL(path or LIB)
        fn = lib.mm_source_hash
        fn.restype = ctypes.c_char_p
        return fn().decode()
    except Exception:  # noqa: BLE001
        return None


def needs_build() -> bool:
    """True when the library is missing or was compiled from other sources than the ones present (content hash, not mtime: the
    snapshot that travels to the GPU box does not keep modification times)."""
    if not os.path.exists(LIB):
        return True
    srcs, _ = sources()
    if not srcs:
        return False   # a binary-only install: nothing to compare against
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """extra_flags/out: build a kernel variant (e.g. -DMM_K2_MINBLOCKS=6) beside the default library for A/B runs."""
    LIB = out or globals()["LIB"]
    if not force and not extra_flags and not needs_build():
        return LIB
    srcs, _ = sources()
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [_nvcc()] + FLAGS + [f'-DMM_SRC_HASH="{source_hash()}"'] + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
