#!/usr/bin/env python
"""bench.py -- agent-steps/s of the fused env step + observation kernel (K2), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference arm: the CPU path on the host cores

Workload (config.workload): BASELINE config[3] -- 1,048,576 mazes x 2 agents PER GPU at 4x the default maze area
(side 49), max_timestep 1200, uniform mask-legal random actions, auto-reset on, episode phases spread uniformly over
[0, max_timestep) before timing (steady-state rate of truncations / resets).  Weak scaling: envs are independent, each
rank owns its own shard (no data-path collective).

One "step" = one pass of the hot path over all envs of the rank (one K2 launch).
  value   : whole-job agent-steps/s with everything resident in HBM (actions are sampled inside the kernel).  The timed
            region is `reps` back-to-back blocks of exactly K steps (reps chosen so that it lasts >= 400 ms: the clock sampler
            needs many NVML round trips); ms_per_step is the mean over all of them.
  e2e     : the same through the host-buffer API: every step copies the step's actions from pinned host memory to the device
            and the step's full result (obs, masks, reward, done) back to pinned host memory.
  roofline: algorithmic bytes per launch (685 + ceil(S^2/4) B per env-step) / mean launch duration measured with
            CUDA events on the launching stream, against MEASURED_PEAKS.json hbm_gbs (`frac`); `frac_dram` is the same with the
            DRAM bytes ncu measured for that launch (profiles/k2_traffic.json, `traffic_source`).
  cpu_baseline: the CPU oracle (a C port of the reference loop, oracle/) on the box's host cores, bounded sample -- and,
            under `python_reference`, the UNMODIFIED Python reference (staged under baseline/_ref/) timed on the same host.
Extra legs of our arm (extra keys; they do not change `value`):
  rollout   : BASELINE config[2] (65,536 mazes, T = 128, side 25): full PPO rollout (K4 policy + K2 step per step, K3 GAE) and the
              5 x 5 update, N = 1 only.
  train_iter: the same iteration per rank at N > 1 -- the one place a collective (NCCL gradient all-reduce) is on the path.
  strong    : K2 with 1,048,576 mazes in TOTAL (1 Mi / N per GPU): the strong-scaling point of config[3].
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
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
MIN_TIMED_MS = 400.0   # NVML clock queries take tens of ms each while the GPU is busy: a 60 ms region got ONE sample
STAGGER_HASH = 2654435761


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def workload_config(envs_per_gpu: int, world: int) -> dict:
    """The SAME dict in both arms: the driver compares the arms' configs."""
    return {"workload": f"config[3]: {envs_per_gpu} mazes x 2 agents per GPU, side {SIDE} (4x default area), max_timestep {MAX_T}, uniform mask-legal "
                        "random actions, auto-reset on, episode phases spread uniformly before timing",
            "envs_per_gpu": envs_per_gpu, "side": SIDE, "max_timestep": MAX_T, "parallelism": f"env-shard x{world}",
            "l2": f"per-step traffic ({BYTES_PER_ENV_STEP * envs_per_gpu / 1e6:.0f} MB algorithmic) exceeds the 126 MB L2; no flush needed"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period=0.002):
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


# ------------------------------------------------------------------------------------------------ CPU arms
class CpuPort:
    """The reference's CPU loop as its C restatement (oracle/), on `threads` host threads.  ONE code path for both CPU numbers the bench
    prints (the `cpu_baseline` leg of our arm and the whole `--impl reference` arm): env-major OpenMP over a bounded sample of the
    workload's mazes, every call one parallel region of `inner` consecutive steps per env, phases spread before timing."""

    def __init__(self, n_envs: int, threads: int):
        from oracle import OracleBatch
        self.n_envs, self.threads = n_envs, threads
        self.b = OracleBatch(n_envs, n_envs, max_timestep=MAX_T, threads=threads)
        for p in range(n_envs):
            self.b.generate_pool_maze(p, SIDE, True, 1, 12345, p)
        self.b.reset_all()
        self.b.run_random(0, seed=1, stagger=MAX_T)   # env e advances (e * 2654435761 mod 2^32) mod 1200 steps
        self.calls = 0

    def step(self, inner: int) -> int:
        self.calls += 1
        return self.b.run_random(inner, seed=100 + self.calls)

    def describe(self, inner: int) -> str:
        return (f"{self.n_envs} mazes of side {SIDE} (bounded sample of the workload), {inner} consecutive steps per env per call, uniform mask-legal "
                "random actions, auto-reset, phases spread, env-major OpenMP")


def python_reference_legs(seconds: float, threads: int) -> dict:
    """The UNMODIFIED Python reference on this host, staged under baseline/_ref/ by baseline/stage_reference.py."""
    script = os.path.join(ROOT, "baseline", "ref_python_bench.py")
    out = {"source": "baseline/_ref (unmodified rhuangr/MARL-Maze, staged by baseline/stage_reference.py)", "host_threads": threads}

    def run(*extra, timeout=600):
        try:
            r = subprocess.run([sys.executable, script, *extra], capture_output=True, text=True, timeout=timeout,
                               env=dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1"))
            return json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            return {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    first = run("--mode", "env", "--seconds", str(seconds), "--side-half", str(SIDE_HALF))
    if "unavailable" in first:
        out["unavailable"] = first["unavailable"]
        return out
    allp = run("--mode", "env", "--seconds", str(seconds), "--side-half", str(SIDE_HALF), "--procs", str(threads))
    out["env_step_obs"] = {"what": f"Maze.step + observations (maze.py:74-163), side {SIDE}, mask-legal uniform random actions, resets included",
                           "one_process_agent_steps_per_s": first.get("agent_steps_per_s"),
                           "all_cores_agent_steps_per_s": allp.get("agent_steps_per_s"), "processes": threads, "unit": "agent-steps/s"}
    pol = run("--mode", "policy", "--seconds", str(min(seconds, 5.0)))
    out["config1_policy_rollout"] = {"what": "display_policy.update_env loop with PPO.pth, 1 maze, main.py settings (maze.py:477-493), 1 torch thread",
                                     "agent_steps_per_s": pol.get("agent_steps_per_s"), "unit": "agent-steps/s"}
    bat = run("--mode", "batch", "--batch", "300")
    out["get_batch"] = {"what": "PPO.get_batch (PPO.py:89-152), batch_size 300, 1 torch thread", "env_steps_per_s": bat.get("env_steps_per_s"), "unit": "env-steps/s"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    inner = args.ref_inner
    cpu = CpuPort(args.ref_envs, threads)
    for _ in range(max(args.warmup, 1)):
        cpu.step(inner)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):  # one step = one call = one pass of the path over the bounded sample (inner consecutive steps per env)
        n += cpu.step(inner)
    dt = time.perf_counter() - t0
    val = 2.0 * n / dt
    sample = cpu.describe(inner)
    py = None if args.no_python_reference else python_reference_legs(args.py_seconds, threads)
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec", "value": val, "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32/f64->f32", "data": "synthetic",
        "config": workload_config(args.envs, world),
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample, "seconds": round(dt, 2),
                         "env_steps_per_step": args.ref_envs * inner, "python_reference": py},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "value = the reference's CPU loop as its C restatement (oracle/, pinned bit-exactly to the reference by tests/golden) on all host threads -- "
                "about 1000x the Python reference per core, i.e. a deliberately stringent CPU arm; the unmodified Python reference timed on this host is under "
                "cpu_baseline.python_reference",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def _bind_to_gpu_numa_node(local: int):
    """CPU affinity of this rank = the cores NVML reports as local to its GPU (no-op on single-node hosts / VMs that expose one domain):
    the rank's pinned host buffers are then first-touched on the GPU's own NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus and len(cpus) < (os.cpu_count() or 1):
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to GPU {local}"
        return "single affinity domain (nothing to bind)"
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__})"


def spread_phases(eng, torch, steps: int = MAX_T):
    """Bring the envs to uniformly spread episode times: `steps` K2 steps during which env e is reset once, at step
    (e * 2654435761 mod 2^32) mod steps -- afterwards about E / max_timestep envs hit the time limit (and reset in-launch) every step."""
    E = eng.E
    phase = ((torch.arange(E, device=eng.device, dtype=torch.int64) * STAGGER_HASH) % (1 << 32)) % steps
    for k in range(steps):
        eng.step(None, auto_reset=True, action_seed=3)
        eng.reset((phase == k).to(torch.uint8))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from marl_maze_b200 import MazeEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    affinity = _bind_to_gpu_numa_node(local)
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

    def timed_k2(eng, K, reps, sampler=None):
        """`reps` back-to-back blocks of exactly K device-resident steps; returns total ms (max over ranks)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(K * reps):
            eng.step(None, auto_reset=True, action_seed=1)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1))

    E, K, Wm = args.envs, args.steps, max(args.warmup, 3)
    eng = MazeEngine(E, smax=SIDE, max_timestep=MAX_T, pool_size=E, env_offset=rank * E)
    eng.generate(seed=2026, side_range=(SIDE_HALF, SIDE_HALF), rand_start=True, difficulty=1, id_base=rank * E)
    eng.reset()
    torch.cuda.synchronize()

    # ---------------------------------------------------------------- leg 1: device-resident (value)
    if not args.no_spread:
        spread_phases(eng, torch)
    for _ in range(Wm):
        eng.step(None, auto_reset=True, action_seed=1)
    barrier()
    # calibrate the number of K-step blocks so that the timed region lasts >= MIN_TIMED_MS on every rank
    probe = timed_k2(eng, K, 1)
    reps = max(1, int(math.ceil(MIN_TIMED_MS / max(probe, 1e-3))))
    l0 = eng.launches
    with ClockSampler(local) as clk:
        ms_total = timed_k2(eng, K, reps)
    launches = eng.launches - l0
    ms_per_step = ms_total / (K * reps)
    value = 2.0 * E * world * K * reps / (ms_total * 1e-3)
    envs_now = eng.envs()
    errs = int(envs_now[:, 6].sum())
    resets_per_step = None
    if not args.no_spread:  # evidence that the timed window sat in reset steady state: episodes started during it
        resets_per_step = float((envs_now[:, 0] < K * reps).sum()) / (K * reps) if K * reps < MAX_T else None

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
    del h_act, h_obs, h_masks, h_rew, h_done, d_act

    # ---------------------------------------------------------------- roofline of K2 (dominant and only kernel of the step)
    peak, peak_src = _peaks()
    achieved = BYTES_PER_ENV_STEP * E / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU; every rank runs the same launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), "cited, not measured by this run: " + tj.get("source", "profiles/k2_traffic.json")
    except Exception:
        pass
    frac_dram = (traffic * E / ENVS_PER_GPU / (ms_per_step * 1e-3) / 1e9 / peak) if traffic else None

    # ---------------------------------------------------------------- strong-scaling point: 1 Mi mazes in TOTAL
    strong = None
    if not args.no_extra_legs and world > 1 and ENVS_PER_GPU % world == 0:
        del eng
        torch.cuda.empty_cache()
        Es = ENVS_PER_GPU // world
        eng_s = MazeEngine(Es, smax=SIDE, max_timestep=MAX_T, pool_size=Es, env_offset=rank * Es)
        eng_s.generate(seed=2026, side_range=(SIDE_HALF, SIDE_HALF), rand_start=True, difficulty=1, id_base=rank * Es)
        eng_s.reset()
        if not args.no_spread:
            spread_phases(eng_s, torch)
        for _ in range(Wm):
            eng_s.step(None, auto_reset=True, action_seed=1)
        p_s = timed_k2(eng_s, K, 1)
        reps_s = max(1, int(math.ceil(MIN_TIMED_MS / max(p_s, 1e-3))))
        ms_s = timed_k2(eng_s, K, reps_s)
        strong = {"what": f"config[3] strong scaling: {ENVS_PER_GPU} mazes in total, {Es} per GPU", "n_gpus": world, "value": 2.0 * Es * world * K * reps_s / (ms_s * 1e-3),
                  "unit": "agent-steps/s", "ms_per_step": ms_s / (K * reps_s)}
        del eng_s
    else:
        del eng
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- PPO legs: config[2] rollout + update (N = 1), train iteration (N > 1)
    ppo_leg = None
    if not args.no_extra_legs:
        ppo_leg = ppo_iteration_leg(torch, dist, world, rank, local, args)

    single = None
    if rank == 0 and world == 1 and not args.no_extra_legs:
        single = single_env_leg(torch, local, seconds=2.0)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        port = CpuPort(args.cpu_envs, threads)
        port.step(args.ref_inner)  # warm-up call
        t0 = time.perf_counter()
        n = 0
        calls = max(1, args.cpu_steps // args.ref_inner)
        for _ in range(calls):
            n += port.step(args.ref_inner)
        dt = time.perf_counter() - t0
        cpu = {"value": 2.0 * n / dt, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": port.describe(args.ref_inner) + f", {calls} calls",
               "seconds": round(dt, 2), "python_reference": None if args.no_python_reference else python_reference_legs(args.py_seconds, threads)}

    if rank == 0:
        line = {
            "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 bit-planes / int32 -> f32 obs (bit-exact); policy legs: fp32-grade 3-term split on tcgen05, 1e-5 rel on log-probs / values",
            "data": "synthetic",
            "config": workload_config(E, world),
            "timed_region": {"reps": reps, "steps_total": K * reps, "ms_total": ms_total, "resets_per_step": resets_per_step,
                             "actions": "sampled in-kernel (Philox)", "mazes": "K1 generator"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "ms_per_step": ms_e2e / Ke, "d2h_GBs_aggregate": d2h * world / (ms_e2e / Ke * 1e-3) / 1e9, "cpu_affinity": affinity,
                    "what": "pinned host actions -> device, mm_step_obs, full result (obs, masks, reward, done) -> pinned host, every step"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "frac_dram": frac_dram,
                         "kernel": "mm::k_step_obs<false, 1> (ncu: void mm::k_step_obs<0, 1>(mm::StepParams))", "bytes_per_launch": BYTES_PER_ENV_STEP * E,
                         "launch_ms": ms_per_step, "peak_source": peak_src, "bytes_per_agent_step": BYTES_PER_ENV_STEP / 2},
            "cpu_baseline": cpu,
            "env_error_flags": errs,
        }
        if strong is not None:
            line["strong"] = strong
        if ppo_leg is not None:
            line["rollout" if world == 1 else "train_iter"] = ppo_leg
        if single is not None:
            line["single_env"] = single
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def ppo_iteration_leg(torch, dist, world, rank, local, args):
    """BASELINE config[2] per rank: 65,536 mazes x 128 steps of side 25 -- rollout (K4 + K2 per step, K3 at the end) and the 5 x 5 update.
    Device-timed, max over ranks.  At N > 1 the update's gradient all-reduce (NCCL) is on the path and is timed separately."""
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent

    E, T = args.ppo_envs, args.ppo_horizon
    dev = f"cuda:{local}"
    brain = PPO(agent_amount=2, batch_size=E * T - 5, lr=0.00014, epochs=1, verbose=False, model_path=None, horizon=T, device=dev)
    agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
    Maze(agents=agents, max_timestep=MAX_T, rand_sizes=True, rand_range=[13, 13], rand_start=True, num_envs=E, device=dev, seed=1, env_offset=rank * E)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms.item())

    roll_ms = upd_ms = None
    stats = {}
    for it in range(3):  # iteration 0 runs eagerly and warms up, 1 captures the rollout graph, 2 is the one reported
        batch, roll_ms = timed(brain.get_batch)
        stats = dict(brain.last_stats)
        upd, upd_ms = timed(lambda: brain.update(batch))
        del batch
    ar_ms = None
    if world > 1:   # the collective on its own: both gradient buckets (actor 1.06 MB, critic 50 KB), as one optimiser step issues them, 25 steps per update
        def ar_step():
            h_a = brain._allreduce_start(brain.actor); h_c = brain._allreduce_start(brain.critic)
            brain._allreduce_finish(h_a); brain._allreduce_finish(h_c)
        for _ in range(5):
            ar_step()
        _, ms25 = timed(lambda: [ar_step() for _ in range(25)])
        ar_ms = ms25
    steps = E * T * world
    out = {"what": f"config[2]: {E} mazes x 2 agents per GPU, T = {T}, side 25: PPO rollout (K4 policy forward + fused sampling, K2 step + obs, per step; K3 GAE) "
                   "then the 5 x 5 minibatch update (K5)",
           "n_gpus": world, "envs_per_gpu": E, "horizon": T, "rollout_ms": roll_ms, "update_ms": upd_ms,
           "rollout_agent_steps_per_s": 2.0 * steps / (roll_ms * 1e-3), "iteration_env_steps_per_s": steps / ((roll_ms + upd_ms) * 1e-3),
           "allreduce_ms": ar_ms, "allreduce_what": None if ar_ms is None else "25 x (actor + critic gradient bucket, NCCL AVG) timed on their own; inside the update the "
           "actor's bucket is in flight while the critic's forward / backward runs", "optimizer_steps": upd.get("steps") if isinstance(upd, dict) else None,
           "episodes": stats.get("episodes"), "mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    try:
        with open(os.path.join(ROOT, "profiles", "k4_tensor_pipe.json")) as f:
            out["policy_gemm_tensor_pipe"] = json.load(f)
    except Exception:
        pass
    return out


def single_env_leg(torch, local, seconds=2.0):
    """BASELINE config[0] on this framework: ONE maze, the reference's python-list interface (Maze.reset / Agent.get_action per agent / Maze.step, the
    loop of maze.py:477-493), main.py's settings, randomly initialised networks of the reference architecture.  Host-timed (every call returns python
    objects); cpu_baseline.python_reference.config1_policy_rollout is the unmodified reference in the same loop on this box's CPU."""
    from marl_maze_b200.PPO import PPO
    from marl_maze_b200.maze import Maze
    from marl_maze_b200.maze_agent import Agent
    dev = f"cuda:{local}"
    brain = PPO(agent_amount=2, batch_size=15000, lr=0.00014, verbose=False, model_path=None, device=dev)
    agents = (Agent("RED", brain, None, None, 2), Agent("BLUE", brain, None, None, 3))
    maze = Maze(agents=agents, max_timestep=MAX_T, rand_sizes=True, rand_range=[12, 13], rand_start=True, difficulty=1, default_size=[4, 4], num_envs=1, seed=0, device=dev)
    obs, masks = maze.reset()
    n = 0
    for budget in (0.3, seconds):   # warm-up, then the timed loop
        torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < budget:
            action = [agent.get_action(obs[i], masks[i])[0] for i, agent in enumerate(agents)]
            obs, masks, reward, done = maze.step(action)
            n += 1
            if done:
                obs, masks = maze.reset()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return {"what": "config[0]: 1 maze (side 23-25), 2 agents, Agent.get_action per agent + Maze.step through python lists; one stream synchronisation per env step "
                    "(both agents' logits are fetched behind the step kernel)", "env_steps": n, "seconds": round(dt, 3),
            "agent_steps_per_s": 2.0 * n / dt, "unit": "agent-steps/s", "timed": "host clock"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="mazes per GPU")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--cpu-envs", type=int, default=65536)
    ap.add_argument("--cpu-steps", type=int, default=1280)  # ~10 s of CPU work on a 16-thread host
    ap.add_argument("--ref-envs", type=int, default=65536)
    ap.add_argument("--ref-inner", type=int, default=64, help="consecutive steps per env inside one CPU call (one OpenMP region)")
    ap.add_argument("--py-seconds", type=float, default=6.0, help="seconds per timed leg of the unmodified Python reference")
    ap.add_argument("--ppo-envs", type=int, default=65536)
    ap.add_argument("--ppo-horizon", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the rollout / train_iter / strong legs")
    ap.add_argument("--no-spread", action="store_true", help="skip the phase spreader (all envs start at t = 0)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
