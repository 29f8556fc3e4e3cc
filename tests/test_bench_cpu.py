"""bench.py contract on CPU: the reference arm prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--ref-envs", "256", "--no-python-reference"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_sec" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the driver compares the two arms' configs: both come from ONE function
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(bench.ENVS_PER_GPU, 1)
    assert d["cpu_baseline"]["env_steps_per_step"] == 256 * 64 and "phases spread" in d["cpu_baseline"]["sample"]


def test_cpu_arms_share_one_loop_and_spread_phases():
    """both CPU numbers come from the same batched, env-major loop (bench.CpuPort), and the envs' episode times are spread
    over [0, max_timestep) before timing so that the timed window contains the steady-state share of truncations and resets."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    port = bench.CpuPort(512, 2)
    t = port.b.env_state()[:, 0]
    want = (np.arange(512, dtype=np.uint64) * bench.STAGGER_HASH % (1 << 32)) % bench.MAX_T
    # env e advanced `want[e]` steps from t = 0; an episode that ended on the way restarted its clock, so t <= want, equal unless it finished
    assert (t <= want).all() and (t == want).mean() > 0.9
    assert len(np.unique(t // 100)) == 12           # all twelve 100-step bands of the 1200-step episode are populated
    n = port.step(8)
    assert n == 512 * 8


def test_python_reference_leg_runs_the_staged_reference():
    """cpu_baseline.python_reference: the unmodified reference (baseline/_ref, staged from /root/reference in the build container)."""
    sys.path.insert(0, ROOT)
    from baseline import stage_reference
    if stage_reference.stage() is None:
        import pytest
        pytest.skip("no reference available to stage")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_python_bench.py"), "--mode", "env", "--seconds", "0.5", "--side-half", "25"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["side"] == 49 and d["agent_steps_per_s"] > 1000


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
