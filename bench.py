#!/usr/bin/env python
"""bench.py -- agent-steps/s of the fused env step + observation kernel (K2), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference arm: the CPU path on the host cores

Workload (config.workload): BASELINE config[3] -- 1,048,576 mazes x 2 agents PER GPU at 4x the default maze area
(side 49), max_timestep 1200, uniform mask-legal random actions, auto-reset on.  Weak scaling: envs are independent,
each rank owns its own shard (no data-path collective).

One "step" = one pass of the hot path over all envs of the rank (one K2 launch).
  value  : whole-job agent-steps/s with everything resident in HBM (actions are sampled inside the kernel).
  e2e    : the same through the host-buffer API: every step copies the step's actions from pinned host memory
           to the device and the step's full result (obs, masks, reward, done) back to pinned host memory.
  roofline: algorithmic bytes per launch (685 + ceil(S^2/4) B per env-step) / mean launch duration
           measured with CUDA events on the launching stream, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: the CPU oracle (a C port of the reference loop, oracle/) on the box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIDE_HALF = 25          # rand_range=[25,25] -> side 49 = "4x default area" 
SIDE = 2 * SIDE_HALF - 1
MAX_T = 1200            # main.py:20
ENVS_PER_GPU = 1 << 20
BYTES_PER_ENV_STEP = 685 + (SIDE * SIDE + 3) // 4   # SURVEY 8d -> 1286 B at S=49 (643 B per agent-step)
FALLBACK_HBM_GBS = 6650.0


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period=0.004):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_leg(n_envs: int, steps: int, warmup: int, threads: int):
    """The reference's CPU loop (C port in oracle/) on `threads` host threads; returns agent-steps/s and a description."""
    from oracle import OracleBatch
    b = OracleBatch(n_envs, n_envs, max_timestep=MAX_T, threads=threads)
    for p in range(n_envs):
        b.generate_pool_maze(p, SIDE, True, 1, 12345, p)
    b.reset_all()
    if warmup:
        b.run_random(warmup, seed=1)
    t0 = time.perf_counter()
    n = b.run_random(steps, seed=2)
    dt = time.perf_counter() - t0
    return 2.0 * n / dt, dt, f"{n_envs} mazes of side {SIDE} x {steps} steps, uniform legal random actions, auto-reset, env-major OpenMP"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_envs = args.ref_envs
    from oracle import OracleBatch
    b = OracleBatch(n_envs, n_envs, max_timestep=MAX_T, threads=threads)
    for p in range(n_envs):
        b.generate_pool_maze(p, SIDE, True, 1, 12345, p)
    b.reset_all()
    for _ in range(args.warmup):
        b.run_random(1, seed=1)
    t0 = time.perf_counter()
    n = 0
    for k in range(args.steps):  # one step = one pass of the path over the bounded sample
        n += b.run_random(1, seed=2 + k)
    dt = time.perf_counter() - t0
    val = 2.0 * n / dt
    sample = f"{n_envs} mazes of side {SIDE} per step (bounded sample of the {ENVS_PER_GPU}-maze workload), uniform legal random actions, auto-reset"
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec", "value": val, "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32/f64->f32", "data": "synthetic",
        "config": {"workload": f"config[3] sample: {sample}", "side": SIDE, "max_timestep": MAX_T, "envs_per_step": n_envs},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python and cannot travel to the GPU box; this arm times oracle/ (its C restatement, pinned bit-exactly to the "
                "reference by tests/golden) on all host threads",
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from marl_maze_b200 import MazeEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    E, K, Wm = args.envs, args.steps, args.warmup
    eng = MazeEngine(E, smax=SIDE, max_timestep=MAX_T, pool_size=E, env_offset=rank * E)
    eng.generate(seed=2026, side_range=(SIDE_HALF, SIDE_HALF), rand_start=True, difficulty=1, id_base=rank * E)
    eng.reset()
    torch.cuda.synchronize()
    actions_out = torch.zeros(E, 2, 2, dtype=torch.uint8, device=dev)

    # ---------------------------------------------------------------- leg 1: device-resident (value)
    # spread episode phases before timing so that resets are in steady state (about E/episode_len per step)
    for _ in range(max(Wm, 3)):
        eng.step(None, auto_reset=True, action_seed=1, actions_out=actions_out)
    barrier()
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(K):
            eng.step(None, auto_reset=True, action_seed=1, actions_out=actions_out)
        ev1.record()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launches - l0
    ms_per_step = ms_total / K
    value = 2.0 * E * world * K / (ms_total * 1e-3)
    errs = int(eng.envs()[:, 6].sum())

    # ---------------------------------------------------------------- leg 2: host buffers (e2e)
    Ke = max(8, min(K, args.e2e_steps))
    h_act = torch.zeros(Ke + 3, E, 2, 2, dtype=torch.uint8).pin_memory()
    h_obs = torch.zeros(E, 2, 65, dtype=torch.float32).pin_memory()
    h_masks = torch.zeros(E, 2, 6, dtype=torch.uint8).pin_memory()
    h_rew = torch.zeros(E, dtype=torch.float32).pin_memory()
    h_done = torch.zeros(E, dtype=torch.uint8).pin_memory()
    d_act = torch.zeros(E, 2, 2, dtype=torch.uint8, device=dev)
    # Host-side actions must be mask-legal for the state they meet: record Ke+3 steps of kernel-sampled actions, rewind the
    # environment state, and replay those very actions from pinned host memory (the trajectory is deterministic).
    state = [eng.env_grid, eng.env_hdr, eng.env_episode, eng.agent_a, eng.agent_b]
    snap = [t.clone() for t in state]
    d_rec = torch.zeros(Ke + 3, E, 2, 2, dtype=torch.uint8, device=dev)
    for k in range(Ke + 3):
        eng.step(None, auto_reset=True, action_seed=9, actions_out=d_rec[k])
    h_act.copy_(d_rec)
    for t, s_ in zip(state, snap):
        t.copy_(s_)
    del snap, d_rec
    torch.cuda.synchronize()
    step_i = [0]

    def e2e_step():
        d_act.copy_(h_act[step_i[0]], non_blocking=True)
        step_i[0] += 1
        o, m, r, d = eng.step(d_act, auto_reset=True)
        h_obs.copy_(o, non_blocking=True); h_masks.copy_(m, non_blocking=True)
        h_rew.copy_(r, non_blocking=True); h_done.copy_(d, non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(Ke):
        e2e_step()
    ev3.record()
    barrier()
    ms_e2e = max_over_ranks(ev2.elapsed_time(ev3))
    e2e_value = 2.0 * E * world * Ke / (ms_e2e * 1e-3)
    h2d = d_act.numel()
    d2h = h_obs.numel() * 4 + h_masks.numel() + h_rew.numel() * 4 + h_done.numel()

    # ---------------------------------------------------------------- roofline of K2 (dominant and only kernel of the step)
    peak, peak_src = _peaks()
    achieved = BYTES_PER_ENV_STEP * E / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU; every rank runs the same launch
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, sample = cpu_leg(args.cpu_envs, args.cpu_steps, 2, threads)
        cpu = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample, "seconds": round(dt, 2)}

    if rank == 0:
        line = {
            "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 bit-planes/int32 -> f32 obs",
            "data": "synthetic",
            "config": {"workload": f"config[3]: {E} mazes x 2 agents per GPU, side {SIDE} (4x default area), max_timestep {MAX_T}, uniform mask-legal "
                                   "random actions sampled in-kernel, auto-reset on, mazes from the K1 generator",
                       "envs_per_gpu": E, "side": SIDE, "max_timestep": MAX_T, "parallelism": f"env-shard x{world}",
                       "l2": f"per-step traffic ({BYTES_PER_ENV_STEP * E / 1e6:.0f} MB algorithmic) exceeds the 126 MB L2; no flush needed"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "ms_per_step": ms_e2e / Ke,
                    "what": "pinned host actions -> device, mm_step_obs, full result (obs, masks, reward, done) -> pinned host, every step"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "mm::k_step_obs<false>", "bytes_per_launch": BYTES_PER_ENV_STEP * E, "launch_ms": ms_per_step,
                         "peak_source": peak_src, "bytes_per_agent_step": BYTES_PER_ENV_STEP / 2},
            "cpu_baseline": cpu,
            "env_error_flags": errs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="mazes per GPU")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--cpu-envs", type=int, default=65536)
    ap.add_argument("--cpu-steps", type=int, default=1200)  # ~10 s of CPU work on a 16-thread host
    ap.add_argument("--ref-envs", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
