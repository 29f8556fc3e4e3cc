"""ctypes binding of include/marl_maze_b200.h.  There is NO CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

EXPORTS = [
    "mm_abi_version", "mm_source_hash", "mm_error_string", "mm_last_cuda_error",
    "mm_sizeof_pool_grid", "mm_sizeof_pool_d2e", "mm_sizeof_pool_hdr", "mm_sizeof_env_grid", "mm_sizeof_env_hdr",
    "mm_sizeof_env_episode", "mm_sizeof_agent_a", "mm_sizeof_agent_b", "mm_sizeof_finalize_scratch", "mm_sizeof_generate_scratch",
    "mm_init_state", "mm_load_layouts", "mm_generate", "mm_generate_ex", "mm_generate_masked", "mm_reset", "mm_agent_place", "mm_step_obs",
    "mm_unpack_agents", "mm_unpack_envs", "mm_unpack_layout", "mm_unpack_pool", "mm_gae",
    "mm_policy_offsets", "mm_sizeof_policy_scratch", "mm_policy_forward", "mm_critic_forward", "mm_selftest_div", "mm_counter_add",
    "mm_wgrad_geometry", "mm_wgrad_tf32x3", "mm_linear_tf32x3", "mm_linear_f16x3", "mm_ppo_loss_geometry", "mm_ppo_heads_loss",
    "mm_segment_sum_blocks", "mm_segment_sum", "mm_gather_rows", "mm_tokens_forward", "mm_tokens_forward_full", "mm_tokens_backward_blocks", "mm_tokens_backward", "mm_sizeof_tokens_backward_scratch",
]


class MMState(C.Structure):
    """struct mm_state of include/marl_maze_b200.h"""
    _fields_ = [("pool_grid", C.c_void_p), ("pool_d2e", C.c_void_p), ("pool_hdr", C.c_void_p),
                ("env_grid", C.c_void_p), ("env_hdr", C.c_void_p), ("env_episode", C.c_void_p),
                ("agent_a", C.c_void_p), ("agent_b", C.c_void_p),
                ("n_envs", C.c_int32), ("n_pool", C.c_int32), ("smax", C.c_int32), ("max_timestep", C.c_int32),
                ("env_offset", C.c_int32), ("vision", C.c_int32)]


class MMError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MARL_MAZE_LIB")  # override: A/B-testing kernel variants (used as it is)
    if path is None:
        path = _build.LIB
        # A checkout: (re)build in-tree when the library is missing OR older than any source / header -- a stale .so that still exports
        # every symbol would otherwise silently run the old kernels.  build() is a no-op when the library is up to date.
        if not os.path.exists(path) or _build.needs_build():
            try:
                path = _build.build()
            except Exception as e:  # noqa: BLE001
                if not os.path.exists(path):
                    raise MMError(f"libmarl_maze_b200.so is missing and could not be built ({e}); "
                                  "run `python -m marl_maze_b200.build` (needs nvcc). There is no CPU fallback.") from e
                raise MMError(f"libmarl_maze_b200.so is older than its sources and could not be rebuilt ({e}); rebuild it with "
                              "`python -m marl_maze_b200.build`, or point MARL_MAZE_LIB at the library you mean to run.") from e
    elif not os.path.exists(path):
        raise MMError(f"MARL_MAZE_LIB={path} does not exist")
    L = C.CDLL(path)
    vp, i32, u64, u32, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_size_t
    st = C.POINTER(MMState)
    sig = {
        "mm_abi_version": (i32, []),
        "mm_source_hash": (C.c_char_p, []),
        "mm_error_string": (C.c_char_p, [i32]),
        "mm_last_cuda_error": (C.c_char_p, []),
        "mm_sizeof_pool_grid": (sz, [i32, i32]), "mm_sizeof_pool_d2e": (sz, [i32, i32]), "mm_sizeof_pool_hdr": (sz, [i32]),
        "mm_sizeof_env_grid": (sz, [i32, i32]), "mm_sizeof_env_hdr": (sz, [i32]), "mm_sizeof_env_episode": (sz, [i32]),
        "mm_sizeof_agent_a": (sz, [i32]), "mm_sizeof_agent_b": (sz, [i32]),
        "mm_sizeof_finalize_scratch": (sz, [i32, i32]), "mm_sizeof_generate_scratch": (sz, [i32, i32]),
        "mm_init_state": (i32, [st, vp]),
        "mm_load_layouts": (i32, [st, i32, i32, vp, vp, vp, vp]),
        "mm_generate": (i32, [st, i32, i32, i32, i32, i32, i32, u64, u32, i32, i32, vp, vp]),
        "mm_generate_ex": (i32, [st, i32, i32, i32, i32, i32, i32, u64, u32, i32, i32, vp, i32, vp]),
        "mm_generate_masked": (i32, [st, i32, i32, i32, i32, i32, i32, u64, u32, i32, i32, vp, i32, vp, i32, vp]),
        "mm_reset": (i32, [st, vp, vp, vp, vp]),
        "mm_agent_place": (i32, [st, i32, i32, i32, i32, i32, i32, vp]),
        "mm_step_obs": (i32, [st, vp, vp, vp, vp, vp, i32, u64, vp, vp]),
        "mm_unpack_agents": (i32, [st, vp, vp]),
        "mm_unpack_envs": (i32, [st, vp, vp]),
        "mm_unpack_layout": (i32, [st, i32, vp, vp]),
        "mm_unpack_pool": (i32, [st, i32, vp, vp, vp, vp]),
        "mm_gae": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, C.c_double, C.c_double, vp]),
        "mm_policy_offsets": (i32, [C.POINTER(C.c_int32)]),
        "mm_sizeof_policy_scratch": (sz, [i32]),
        "mm_critic_forward": (i32, [vp, vp, i32, vp, vp]),
        "mm_selftest_div": (i32, [i32, i32, vp, vp]),
        "mm_policy_forward": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, u64, u64, i32, vp, vp]),
        "mm_counter_add": (i32, [vp, u64, vp]),
        "mm_wgrad_geometry": (i32, [i32, i32, i32] + [C.POINTER(C.c_int32)] * 4),
        "mm_wgrad_tf32x3": (i32, [vp, vp, i32, i32, i32, vp, vp]),
        "mm_linear_tf32x3": (i32, [vp, i32, i32, vp, vp, i32, vp, vp, vp, i32, i32, vp, vp]),
        "mm_linear_f16x3": (i32, [vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, i32, vp, vp]),
        "mm_ppo_loss_geometry": (i32, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "mm_tokens_forward": (i32, [vp, vp, i32, vp, vp]), "mm_tokens_forward_full": (i32, [vp, vp, i32, vp, vp]),
        "mm_tokens_backward_blocks": (i32, []),
        "mm_tokens_backward": (i32, [vp, vp, vp, i32, vp, vp, vp]),
        "mm_sizeof_tokens_backward_scratch": (sz, [i32]),
        "mm_segment_sum_blocks": (i32, [i32]),
        "mm_segment_sum": (i32, [vp, vp, i32, i32, i32, vp, vp]), "mm_gather_rows": (i32, [vp, vp, i32, i32, i32, vp, vp]),
        "mm_ppo_heads_loss": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, C.c_float, C.c_float, vp, vp, vp, vp]),
    }
    for name in EXPORTS:
        if not hasattr(L, name):
            raise MMError(f"{path} does not export {name}")
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(code: int, what: str):
    if code != 0:
        L = lib()
        raise MMError(f"{what}: {L.mm_error_string(code).decode()} {L.mm_last_cuda_error().decode()}")
