"""Long randomized bit-exactness soak of K2 (+K1 pools) against the C oracle, beyond what the regular gpu tests run.
    python tools/soak_parity.py [--envs 16384] [--steps 2500] [--cases 4]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import OracleBatch, OracleMaze
from marl_maze_b200 import MazeEngine

ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=16384); ap.add_argument("--steps", type=int, default=2500); ap.add_argument("--cases", type=int, default=4)
ap.add_argument("--vision", type=int, nargs=2, default=None, help="per-agent vision_range for every case (default: rotate through (4,4), (3,3), (2,4), (1,2), (4,1))")
a = ap.parse_args()
VISIONS = [(4, 4), (3, 3), (2, 4), (1, 2), (4, 1)]
CASES = [dict(side=(4, 13), diff=3, max_t=250, pf=0.85, pm=0.3), dict(side=(12, 13), diff=1, max_t=1200, pf=0.0, pm=0.5),
         dict(side=(25, 25), diff=2, max_t=600, pf=0.9, pm=0.2), dict(side=(4, 6), diff=4, max_t=80, pf=0.6, pm=0.7),
         dict(side=(27, 27), diff=1, max_t=400, pf=0.95, pm=0.05)][:a.cases]
tot = 0; t0 = time.time()
for ci, c in enumerate(CASES):
    E, K = a.envs, 16
    S = c["side"][1] * 2 - 1
    vis = tuple(a.vision) if a.vision else VISIONS[ci % len(VISIONS)]
    eng = MazeEngine(E, smax=S, max_timestep=c["max_t"], pool_size=E * K, vision=vis)
    eng.generate(1000 + ci, side_range=c["side"], difficulty=c["diff"], id_base=ci * 10_000_000)
    ob = OracleBatch(E, E * K, max_timestep=c["max_t"], threads=os.cpu_count() or 8, vision=vis)
    g = OracleMaze(max_timestep=10, difficulty=c["diff"], rand_start=True, rand_sizes=True, rand_range=c["side"], default_size=(4, 4))
    for p in range(E * K):
        g.seed_philox(1000 + ci, ci * 10_000_000 + p); g.build(); ob.set_pool_maze(p, g.maze())
    oo, om = ob.reset_all(); go, gm = eng.reset()
    assert np.array_equal(go.cpu().numpy().view(np.uint32), oo.view(np.uint32)) and np.array_equal(gm.cpu().numpy(), om)
    rng = (np.arange(E, dtype=np.uint64) + 11 + ci) * np.uint64(0x9E3779B97F4A7C15)
    dones = 0; rsum = 0.0
    for t in range(a.steps):
        act = ob.guided_actions(rng, p_follow=c["pf"], p_mark=c["pm"])
        go, gm, gr, gd = eng.step(torch.from_numpy(act).cuda())
        oo, om, orr, od = ob.step(act)
        ok = np.array_equal(go.cpu().numpy().view(np.uint32), oo.view(np.uint32)) and np.array_equal(gm.cpu().numpy(), om) and \
             np.array_equal(gr.cpu().numpy(), orr) and np.array_equal(gd.cpu().numpy(), od)
        if not ok:
            print(json.dumps({"case": ci, "step": t, "MISMATCH": True})); sys.exit(1)
        dones += int(od.sum()); rsum += float(orr.sum())
        if t % 500 == 499:
            assert np.array_equal(eng.agents(), ob.agents()) and np.array_equal(eng.envs()[:, :4], ob.env_state())
    tot += E * a.steps
    print(json.dumps({"case": ci, **{k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()}, "vision": list(vis), "env_steps": E * a.steps, "episodes": dones, "reward": rsum,
                      "errors": int(eng.envs()[:, 6].sum()) + ob.errors()}), flush=True)
print(json.dumps({"soak": "bit-exact", "total_env_steps": tot, "seconds": round(time.time() - t0, 1)}))
