"""Record golden vectors from the UNMODIFIED reference (run in the build container only).

    python tools/make_golden.py            # rewrites tests/golden/*.npz

Outputs (all small, committed):
  tests/golden/env_traces.npz  -- per-step traces of Maze.reset/step + Agent.get_observations under several
                                  mask-legal action policies and maze configurations (ref_harness.run_trace)
  tests/golden/gen_kats.npz    -- mazes produced by Maze.build_maze for given random.seed values
The GPU box has no /root/reference: tests only ever read these files.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, maze_seed, action_seed, steps, policy, maze kwargs overriding main.py:20)
TRACES = [
    ("kat2_uniform", 0, 7, 5000, "uniform", {}),  # SURVEY 8c KAT(2); sha256 recorded below
    ("guided_a", 100, 0, 2500, "guided", {}),
    ("guided_b", 101, 1, 2500, "guided", {}),
    ("slow_a", 200, 0, 2500, "guided_slow", {}),
    ("slow_b", 203, 3, 2500, "guided_slow", {}),
    ("nomark", 303, 3, 1500, "nomark", {}),
    ("tiny7", 5, 1, 1500, "guided", dict(default_size=[4, 4], rand_sizes=False, rand_start=False, max_timestep=60)),
    ("mixed_sizes_d3", 6, 1, 2000, "guided", dict(rand_range=[3, 9], difficulty=3, max_timestep=200)),
    ("s49_d2", 7, 2, 2500, "guided_slow", dict(rand_range=[25, 25], difficulty=2, max_timestep=800)),
    ("s49_guided", 8, 3, 2000, "guided", dict(rand_range=[25, 25], max_timestep=1200)),
]

# Agent(..., vision_range=r) per agent (maze_agent.py:16,148,165,218,264).  (name, maze_seed, action_seed, steps, policy, kwargs, vision)
VISION_TRACES = [
    ("v33_guided", 400, 0, 2000, "guided", {}, (3, 3)),
    ("v22_uniform", 401, 1, 1500, "uniform", dict(max_timestep=300), (2, 2)),
    ("v11_slow", 402, 2, 1500, "guided_slow", dict(rand_range=[5, 8], max_timestep=300), (1, 1)),
    ("v42_guided", 403, 3, 2000, "guided", {}, (4, 2)),
    ("v13_slow", 404, 4, 2000, "guided_slow", dict(rand_range=[6, 13], difficulty=2, max_timestep=400), (1, 3)),
]

MAXS = 64


def make_vision_traces():
    """tests/golden/env_traces_vision.npz: the reference with vision_range != 4 (per agent)."""
    blob, names = {}, []
    for name, ms, as_, n, pol, kw, vis in VISION_TRACES:
        tr = rh.run_trace(ms, as_, n, maze_kw=kw, policy=pol, vision=vis)
        for k, v in pack_trace(tr).items():
            blob[f"{name}/{k}"] = v
        cfg = dict(rh.MAIN_PY_KW); cfg.update(kw)
        blob[f"{name}/cfg"] = np.asarray([ms, as_, n, cfg["max_timestep"], cfg["difficulty"], int(cfg["rand_start"]), int(cfg["rand_sizes"]),
                                          cfg["rand_range"][0], cfg["rand_range"][1], cfg["default_size"][0], cfg["default_size"][1]], np.int64)
        blob[f"{name}/vision"] = np.asarray(vis, np.int32)
        names.append(name)
        print(f"{name}: vision={vis} steps={n} episodes={len(tr['mazes'])} dones={int(np.sum(tr['done']))} reward_sum={sum(tr['reward'])}")
    blob["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(OUT, "env_traces_vision.npz"), **blob)


def pack_trace(tr):
    n = len(tr["actions"])
    mz = tr["mazes"]
    K = len(mz)
    lay = np.ones((K, MAXS, MAXS), np.uint8)
    hdr = np.zeros((K, 11), np.int32)
    for k, m in enumerate(mz):
        a = np.asarray(m["layout"], np.uint8)
        lay[k, :a.shape[0], :a.shape[1]] = a
        hdr[k] = [m["width"], m["height"], m["path0"][0], m["path0"][1], m["path1"][0], m["path1"][1],
                  m["end"][0], m["end"][1], m["key"][0], m["key"][1], m["shortest_path_len"]]
    # obs the policy sees after each reset (episode k), i.e. emit_obs at the first step of episode k
    first = [0] + [i + 1 for i in range(n) if tr["done"][i]]
    reset_obs = np.zeros((K, 2, 65), np.float32)
    reset_masks = np.zeros((K, 2, 6), np.uint8)
    reset_agents = np.zeros((K, 2, 18), np.int32)
    for k, i in enumerate(first):
        if i < n:
            reset_obs[k] = tr["emit_obs"][i]; reset_masks[k] = tr["emit_masks"][i]; reset_agents[k] = tr["agents_emit"][i]
        else:
            reset_obs[k] = tr["final_obs"]; reset_masks[k] = tr["final_masks"]
    return dict(
        maze_layout=np.packbits(lay == 1, axis=-1),  # walls only; the reference's fresh layout is 0/1
        maze_hdr=hdr,
        actions=np.asarray(tr["actions"], np.uint8),
        step_obs=np.asarray(tr["step_obs"], np.float32),
        step_masks=np.asarray(tr["step_masks"], np.uint8),
        reward=np.asarray(tr["reward"], np.float32),
        done=np.asarray(tr["done"], np.uint8),
        agents_after=np.asarray(tr["agents_after"], np.int32),
        reset_obs=reset_obs, reset_masks=reset_masks, reset_agents=reset_agents,
        max_timestep=np.int32(tr["max_timestep"]), final_t=np.int32(tr["final_t"]),
    )


def kat2_hash(tr):
    """SHA-256 of SURVEY 8c KAT(2), computed from the recorded trace."""
    h = hashlib.sha256()
    n = len(tr["actions"])

    def emit(o, m, r, d):
        h.update(np.asarray(o, np.float32).tobytes()); h.update(np.asarray(m, np.uint8).tobytes())
        h.update(np.float32(r).tobytes()); h.update(bytes([int(d)]))
    emit(tr["emit_obs"][0], tr["emit_masks"][0], 0, False)
    for i in range(n):
        emit(tr["step_obs"][i], tr["step_masks"][i], tr["reward"][i], tr["done"][i])
        if tr["done"][i]:
            emit(tr["emit_obs"][i + 1] if i + 1 < n else tr["final_obs"], tr["emit_masks"][i + 1] if i + 1 < n else tr["final_masks"], 0, False)
    return h.hexdigest()


def main():
    import random
    os.makedirs(OUT, exist_ok=True)
    blob = {}
    names = []
    for name, ms, as_, n, pol, kw in TRACES:
        tr = rh.run_trace(ms, as_, n, maze_kw=kw, policy=pol)
        for k, v in pack_trace(tr).items():
            blob[f"{name}/{k}"] = v
        cfg = dict(rh.MAIN_PY_KW); cfg.update(kw)
        blob[f"{name}/cfg"] = np.asarray([ms, as_, n, cfg["max_timestep"], cfg["difficulty"], int(cfg["rand_start"]), int(cfg["rand_sizes"]),
                                          cfg["rand_range"][0], cfg["rand_range"][1], cfg["default_size"][0], cfg["default_size"][1]], np.int64)
        names.append(name)
        nd = int(np.sum(tr["done"])); print(f"{name}: steps={n} episodes={len(tr['mazes'])} dones={nd} reward_sum={sum(tr['reward'])}")
        if name == "kat2_uniform":
            hx = kat2_hash(tr); print("kat2 sha256", hx)
            assert hx == "4809ce85defd322829727dd95048c32b7d522acc9a7d29685b8b4d9d489325ef"
            blob["kat2_sha256"] = np.frombuffer(bytes.fromhex(hx), np.uint8)
    blob["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(OUT, "env_traces.npz"), **blob)

    # generator KATs: random.seed(s); Maze(**cfg).build_maze() x3 consecutive mazes per seed
    import contextlib, io
    g = {}
    cfgs = [("main", {}), ("tiny7", dict(default_size=[4, 4], rand_sizes=False, rand_start=False)),
            ("d3", dict(rand_range=[3, 9], difficulty=3)), ("s49", dict(rand_range=[25, 25], difficulty=2)),
            ("fixed17", dict(default_size=[9, 9], rand_sizes=False, rand_start=True))]
    for cname, kw in cfgs:
        for seed in (0, 1, 5, 1234567, 2**40 + 17):
            with contextlib.redirect_stdout(io.StringIO()):
                env = rh.make_env(**kw)
                random.seed(seed)
                hdrs, lays, paths = [], [], []
                for _ in range(3):
                    env.build_maze()
                    m = rh.maze_snapshot(env)
                    lay = np.ones((MAXS, MAXS), np.uint8); a = np.asarray(m["layout"], np.uint8); lay[:a.shape[0], :a.shape[1]] = a
                    lays.append(np.packbits(lay == 1, axis=-1))
                    hdrs.append([m["width"], m["height"], m["start"][0], m["start"][1], m["end"][0], m["end"][1], m["key"][0], m["key"][1],
                                 m["shortest_path_len"], m["path1"][0], m["path1"][1]])
                    paths.append(hashlib.sha256(np.asarray(env.shortest_path, np.int32).tobytes()).digest())
            cfg = dict(rh.MAIN_PY_KW); cfg.update(kw)
            key = f"{cname}/{seed}"
            g[key + "/hdr"] = np.asarray(hdrs, np.int32); g[key + "/layout"] = np.asarray(lays)
            g[key + "/path_sha"] = np.frombuffer(b"".join(paths), np.uint8)
            g[key + "/cfg"] = np.asarray([cfg["difficulty"], int(cfg["rand_start"]), int(cfg["rand_sizes"]), cfg["rand_range"][0], cfg["rand_range"][1],
                                          cfg["default_size"][0], cfg["default_size"][1]], np.int64)
    g["keys"] = np.asarray(sorted({k.rsplit("/", 1)[0] for k in g}))
    np.savez_compressed(os.path.join(OUT, "gen_kats.npz"), **g)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__" and "--ppo" not in sys.argv and "--upd" not in sys.argv and "--kat5" not in sys.argv and "--vision" not in sys.argv:
    main()
if __name__ == "__main__" and "--vision" in sys.argv:
    make_vision_traces()


def make_ppo_kats():
    """GAE and network known answers from the reference's own PPO.py / networks.py (torch CPU)."""
    import contextlib, io, random
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import ppo_oracle as po
    rh.load_reference()
    sys.path.insert(0, rh.REFERENCE_DIR)
    cwd = os.getcwd(); os.chdir("/tmp")  # PPO.__init__ would load ./PPO.pth (PPO.py:229-238): stay away from it
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import PPO as ref_ppo, networks as ref_net
    finally:
        os.chdir(cwd); sys.path.remove(rh.REFERENCE_DIR)
    out = {}

    class Dummy:
        discount_rate = 0.99; lam = 0.95
    rng = np.random.default_rng(42)
    lens = [1, 2, 3, 4, 7, 33, 200, 1200]
    for k, L in enumerate(lens):
        rew = [0.0] * L
        for _ in range(min(2, L)):
            rew[int(rng.integers(0, L))] = 0.5
        rew[-1] = 1 if (k % 2 == 0 and L > 1) else 0.0  # maze.py:118 sets the int 1; an all-int list (1-step success) is unreachable
        vals = rng.standard_normal(L).astype(np.float32)
        dones = [False] * (L - 1) + [True]
        with contextlib.redirect_stdout(io.StringIO()):
            adv = ref_ppo.PPO.get_GAEs(Dummy(), list(rew), [torch.tensor([[v]], dtype=torch.float32) for v in vals], dones)
        out[f"gae/{k}/rew"] = np.asarray(rew, np.float32); out[f"gae/{k}/val"] = vals; out[f"gae/{k}/adv"] = np.asarray(adv, np.float32)
    out["gae/n"] = np.int32(len(lens))
    # SURVEY 8c KAT(4)
    adv = ref_ppo.PPO.get_GAEs(Dummy(), [0, .5, 0, 1], [torch.tensor([[0.1]])] * 4, [False, False, False, True])
    out["gae/kat4"] = np.asarray(adv, np.float32)

    # networks: seeded weights with the reference's names/shapes, real observations from a recorded trace
    z = np.load(os.path.join(OUT, "env_traces.npz"))
    obs = z["guided_a/step_obs"][:192]            # [192,2,65]
    masks = z["guided_a/step_masks"][:192].astype(bool)
    acts = z["guided_a/actions"][1:193]           # the actions taken FROM those observations
    for seed in (11, 12):
        asd, csd = po.seeded_state_dicts(seed)
        actor = ref_net.Actor([264, 264, 264]); critic = ref_net.Critic(2, hidden_sizes=[64, 64])
        actor.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()}); critic.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()})
        with torch.no_grad():
            mv, mk = actor(torch.from_numpy(obs.reshape(-1, 65)))
            val = critic(torch.from_numpy(obs))
            holder = Dummy(); holder.actor = actor
            lps = [ref_ppo.PPO.get_log_probs(holder, i, torch.from_numpy(obs), torch.from_numpy(acts.astype(np.float32)), torch.from_numpy(masks)).numpy() for i in range(2)]
        out[f"net/{seed}/move_logits"] = mv.numpy(); out[f"net/{seed}/mark_logits"] = mk.numpy(); out[f"net/{seed}/values"] = val.numpy()
        out[f"net/{seed}/log_probs"] = np.stack(lps, 1)
    out["net/obs"] = obs; out["net/masks"] = masks.astype(np.uint8); out["net/actions"] = acts
    np.savez_compressed(os.path.join(OUT, "ppo_kats.npz"), **out)
    print("ppo_kats.npz", os.path.getsize(os.path.join(OUT, "ppo_kats.npz")))


def make_upd_kats():
    """Actor-update known answers from the reference's own PPO.py:58-76 (get_log_probs for both agents, ratio, clipped surrogate,
    actor_loss.backward()) on torch CPU fp32: loss, joint log-probs and the gradient of every actor parameter (the three trunk
    weight matrices as every 5th element of the flattened gradient, to keep the fixture small)."""
    import contextlib, io
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import ppo_oracle as po
    rh.load_reference()
    sys.path.insert(0, rh.REFERENCE_DIR)
    cwd = os.getcwd(); os.chdir("/tmp")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import PPO as ref_ppo, networks as ref_net
    finally:
        os.chdir(cwd); sys.path.remove(rh.REFERENCE_DIR)
    z = np.load(os.path.join(OUT, "env_traces.npz"))
    n = 384
    obs = z["guided_a/step_obs"][:n]; masks = z["guided_a/step_masks"][:n].astype(bool)
    rng = np.random.default_rng(77)
    acts = np.zeros((n, 2, 2), np.uint8)     # mask-legal actions (an action the mask forbids has log-prob -inf and makes the ratio NaN)
    for e in range(n):
        for a in range(2):
            acts[e, a, 0] = rng.choice(np.flatnonzero(masks[e, a, :5]))
            acts[e, a, 1] = rng.integers(0, 2) if masks[e, a, 5] else 0
    out = {"upd/obs": obs, "upd/masks": masks.astype(np.uint8), "upd/actions": acts, "upd/clip": np.float32(0.2), "upd/stride": np.int32(5)}

    class Holder:
        pass
    for seed in (11, 12):
        asd, _ = po.seeded_state_dicts(seed)
        actor = ref_net.Actor([264, 264, 264]); actor.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()})
        h = Holder(); h.actor = actor
        t_obs, t_act, t_masks = torch.from_numpy(obs), torch.from_numpy(acts.astype(np.float32)), torch.from_numpy(masks)
        with torch.no_grad():
            base = sum(ref_ppo.PPO.get_log_probs(h, i, t_obs, t_act, t_masks) for i in range(2))
        old = (base.numpy() + 0.3 * rng.standard_normal(n)).astype(np.float32)      # ratios on both sides of the clip range
        adv = rng.standard_normal(n).astype(np.float32)
        m_log_probs, m_advantage, clip = torch.from_numpy(old), torch.from_numpy(adv), 0.2
        current_log_prob = 0
        for i in range(2):                                                            # PPO.py:65-67
            current_log_prob += ref_ppo.PPO.get_log_probs(h, i, t_obs, t_act, t_masks)
        prob_ratios = torch.exp(current_log_prob - m_log_probs)                       # PPO.py:68
        surrogate1 = prob_ratios * m_advantage
        surrogate2 = torch.clamp(prob_ratios, 1 - clip, 1 + clip) * m_advantage
        actor_loss = -torch.mean(torch.min(surrogate1, surrogate2))                   # PPO.py:74
        actor.zero_grad(); actor_loss.backward()
        out[f"upd/{seed}/old"] = old; out[f"upd/{seed}/adv"] = adv
        out[f"upd/{seed}/loss"] = np.float32(actor_loss.item()); out[f"upd/{seed}/joint"] = current_log_prob.detach().numpy()
        out[f"upd/{seed}/frac_clipped"] = np.float32(((prob_ratios < 0.8) | (prob_ratios > 1.2)).float().mean().item())
        for name, p in actor.named_parameters():
            g = p.grad.numpy()
            out[f"upd/{seed}/grad/{name}"] = g.reshape(-1)[::5].copy() if name in ("layers.0.weight", "layers.1.weight", "layers.2.weight") else g
    np.savez_compressed(os.path.join(OUT, "upd_kats.npz"), **out)
    print("upd_kats.npz", os.path.getsize(os.path.join(OUT, "upd_kats.npz")))


if __name__ == "__main__" and "--ppo" in sys.argv:
    make_ppo_kats()
if __name__ == "__main__" and "--upd" in sys.argv:
    make_upd_kats()


def make_kat5():
    """SURVEY 8c KAT(5) as a fixture that travels: the reference's shipped checkpoint (PPO.pth: actor + critic tensors, fp32) and what the
    REFERENCE's own networks.py computes from it -- on the four facing one-hots (the survey's KAT) and on 192 recorded observations --
    plus the log-probs of the recorded actions through the reference's PPO.get_log_probs.  Lets the `-m gpu` tests push the real
    checkpoint through mm_policy_forward on a box that has neither /root/reference nor PPO.pth."""
    import contextlib, io
    import torch
    rh.load_reference()
    sys.path.insert(0, rh.REFERENCE_DIR)
    cwd = os.getcwd(); os.chdir("/tmp")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import PPO as ref_ppo, networks as ref_net
    finally:
        os.chdir(cwd); sys.path.remove(rh.REFERENCE_DIR)
    sd = torch.load(os.path.join(rh.REFERENCE_DIR, "PPO.pth"), map_location="cpu")
    actor = ref_net.Actor([264, 264, 264]); critic = ref_net.Critic(2, hidden_sizes=[64, 64])
    actor.load_state_dict(sd["actor"]); critic.load_state_dict(sd["critic"])
    out = {f"actor/{k}": v.numpy().astype(np.float32) for k, v in sd["actor"].items()}
    out.update({f"critic/{k}": v.numpy().astype(np.float32) for k, v in sd["critic"].items()})
    x = torch.zeros(4, 65); x[torch.arange(4), torch.arange(4)] = 1
    z = np.load(os.path.join(OUT, "env_traces.npz"))
    obs = z["guided_a/step_obs"][:192]; masks = z["guided_a/step_masks"][:192].astype(bool); acts = z["guided_a/actions"][1:193]

    class Holder:
        pass
    h = Holder(); h.actor = actor
    with torch.no_grad():
        mv4, mk4 = actor(x)
        mv, mk = actor(torch.from_numpy(obs.reshape(-1, 65)))
        val = critic(torch.from_numpy(obs))
        lps = [ref_ppo.PPO.get_log_probs(h, i, torch.from_numpy(obs), torch.from_numpy(acts.astype(np.float32)), torch.from_numpy(masks)).numpy() for i in range(2)]
    out["kat5/obs"] = x.numpy(); out["kat5/move_logits"] = mv4.numpy(); out["kat5/mark_logits"] = mk4.numpy().reshape(-1)
    out["trace/obs"] = obs; out["trace/masks"] = masks.astype(np.uint8); out["trace/actions"] = acts
    out["trace/move_logits"] = mv.numpy(); out["trace/mark_logits"] = mk.numpy().reshape(-1); out["trace/values"] = val.numpy().reshape(-1)
    out["trace/log_probs"] = np.stack(lps, 1)
    np.savez_compressed(os.path.join(OUT, "kat5_ppo_pth.npz"), **out)
    print("kat5_ppo_pth.npz", os.path.getsize(os.path.join(OUT, "kat5_ppo_pth.npz")), mv4.numpy()[0], torch.sigmoid(mk4).reshape(-1).numpy())


if __name__ == "__main__" and "--kat5" in sys.argv:
    make_kat5()
