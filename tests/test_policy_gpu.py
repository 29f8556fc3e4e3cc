"""K4 (fused actor/critic forward + sampling) against the numpy oracle restatement of networks.py / PPO.get_action and the
recorded reference outputs.  Tolerance: 1e-5 relative (+2e-6 absolute) on logits, values and log-probs -- fp32 with re-associated sums."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN
from oracle import ppo_oracle as po

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-5, atol=2e-6)
# Raw LOGITS of the tensor-core path: tcgen05 accumulates its K = 460/264-long sums with truncation inside the tensor core, which
# leaves ~1e-6 relative (max 4e-6 absolute, tools/k4_accuracy.py) on the logits -- 10x the fp32 SIMT path, still far inside the
# north-star bar on the quantities it names (log-probs and values: 1e-5 relative, asserted with TOL for BOTH paths below).
TOL_LOGITS = {"simt_fp32": TOL, "tcgen05_3xtf32": dict(rtol=1e-5, atol=8e-6), "tcgen05_3xfp16": dict(rtol=1e-5, atol=8e-6),
              "tcgen05_fused_trunk": dict(rtol=1e-5, atol=8e-6)}
PATHS = ["simt_fp32", "tcgen05_3xtf32", "tcgen05_3xfp16", "tcgen05_fused_trunk"]


def _runner(actor, critic, E, path, **kw):
    """The four trunk implementations behind mm_policy_forward: fp32 SIMT tiles, 3xTF32 tcgen05 (mm_policy_tc.cu), 3xFP16 tcgen05 with two
    CTAs per SM (mm_linear16.cu), and the three layers + heads as one persistent 3xFP16 kernel (mm_trunk_fused.cu, the default)."""
    from marl_maze_b200.policy import PolicyRunner
    return PolicyRunner(actor, critic, E, "cuda", tensor_cores=path != "simt_fp32", fp16_split=path in ("tcgen05_3xfp16", "tcgen05_fused_trunk"),
                        fused_trunk=path == "tcgen05_fused_trunk", **kw)


def _nets(seed, faithful=True):
    from marl_maze_b200.networks import Actor, Critic
    asd, csd = po.seeded_state_dicts(seed)
    actor = Actor([264, 264, 264], faithful_projection=faithful).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
    actor.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()}); critic.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()})
    return actor, critic, asd, csd


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("seed", [11, 12])
def test_policy_kernel_vs_reference_golden(seed, tc):
    Z = np.load(os.path.join(GOLDEN, "ppo_kats.npz"))
    obs, masks, acts = Z["net/obs"], Z["net/masks"], Z["net/actions"]
    actor, critic, _, _ = _nets(seed)
    E = obs.shape[0]
    run = _runner(actor, critic, E, tc)
    logits = torch.zeros(E, 2, 6, device="cuda")
    _, logp, val = run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(masks).cuda(), actions_in=torch.from_numpy(acts).cuda(), logits=logits)
    lg = logits.cpu().numpy()
    assert np.allclose(lg[:, :, :5].reshape(-1, 5), Z[f"net/{seed}/move_logits"], **TOL_LOGITS[tc])
    assert np.allclose(lg[:, :, 5].reshape(-1, 1), Z[f"net/{seed}/mark_logits"], **TOL_LOGITS[tc])
    assert np.allclose(val.cpu().numpy(), Z[f"net/{seed}/values"].reshape(-1), **TOL)
    want = Z[f"net/{seed}/log_probs"].sum(1)  # joint log-prob, PPO.py:118,121
    fin = np.isfinite(want)
    got = logp.cpu().numpy()
    assert np.array_equal(np.isfinite(got), fin) and np.allclose(got[fin], want[fin], **TOL)


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("faithful", [True, False])
def test_policy_kernel_vs_oracle_random_obs_and_sampling(faithful, tc):
    actor, critic, asd, csd = _nets(5, faithful)
    rng = np.random.default_rng(3)
    E = 3000
    obs = rng.random((E, 2, 65)).astype(np.float32)
    masks = (rng.random((E, 2, 6)) < 0.6).astype(np.uint8); masks[:, :, 0] |= (masks[:, :, :5].sum(-1) == 0).astype(np.uint8)
    run = _runner(actor, critic, E, tc, seed=9)
    logits = torch.zeros(E, 2, 6, device="cuda")
    act, logp, val = run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(masks).cuda(), logits=logits)
    act = act.cpu().numpy()
    mv, mk = po.actor_forward(asd, obs.reshape(-1, 65), faithful=faithful)
    assert np.allclose(logits.cpu().numpy()[:, :, :5].reshape(-1, 5), mv, **TOL_LOGITS[tc]) and np.allclose(logits.cpu().numpy()[:, :, 5].reshape(-1, 1), mk, **TOL_LOGITS[tc])
    assert np.allclose(val.cpu().numpy(), po.critic_forward(csd, obs).reshape(-1), **TOL)
    # sampled actions are mask-legal and their log-prob is the oracle's
    assert np.take_along_axis(masks[:, :, :5], act[:, :, :1].astype(np.int64), 2).all() and (act[:, :, 1] <= masks[:, :, 5]).all()
    lp = sum(po.action_log_prob(mv.reshape(E, 2, 5)[:, i], mk.reshape(E, 2)[:, i], masks[:, i], act[:, i, 0], act[:, i, 1]) for i in range(2))
    assert np.allclose(logp.cpu().numpy(), lp, **TOL)
    # a second call draws different actions (counter advances); evaluating recorded actions reproduces the log-prob
    act2, _, _ = run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(masks).cuda())
    assert (act2.cpu().numpy() != act).any()
    _, lp_eval, _ = run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(masks).cuda(), actions_in=torch.from_numpy(act).cuda())
    assert torch.equal(lp_eval, logp)


def test_sampling_distribution_chi_square():
    """Fused sampler follows the masked softmax / Bernoulli(sigmoid) distribution (chi-square, fixed seed)."""
    from marl_maze_b200.policy import PolicyRunner
    actor, critic, asd, _ = _nets(21)
    E = 1 << 17
    obs = np.zeros((E, 2, 65), np.float32); obs[:, :, 1] = 1.0   # every row identical: facing east
    masks = np.tile(np.array([1, 1, 0, 1, 0, 1], np.uint8), (E, 2, 1))
    run = PolicyRunner(actor, critic, E, "cuda", seed=1)
    act, _, _ = run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(masks).cuda())
    act = act.cpu().numpy().reshape(-1, 2)
    mv, mk = po.actor_forward(asd, obs[:1, 0])
    l = np.where(masks[0, 0, :5] == 1, mv[0].astype(np.float64), -np.inf); p = np.exp(l - l.max()); p /= p.sum()
    n = len(act)
    cnt = np.bincount(act[:, 0], minlength=5)
    assert cnt[2] == 0 and cnt[4] == 0
    chi = sum((cnt[j] - n * p[j]) ** 2 / (n * p[j]) for j in (0, 1, 3))
    assert chi < 20.0, (chi, cnt, p)          # 2 dof; p(chi2 > 20) ~ 5e-5
    pm = 1 / (1 + np.exp(-float(mk[0, 0])))
    z = (act[:, 1].mean() - pm) / np.sqrt(pm * (1 - pm) / n)
    assert abs(z) < 4.5, (z, pm)


def _ppo_pth_nets():
    """The reference's shipped checkpoint (tests/golden/kat5_ppo_pth.npz, recorded by tools/make_golden.py --kat5)."""
    from marl_maze_b200.networks import Actor, Critic
    Z = np.load(os.path.join(GOLDEN, "kat5_ppo_pth.npz"))
    actor = Actor([264, 264, 264]).cuda(); critic = Critic(2, hidden_sizes=[64, 64]).cuda()
    actor.load_state_dict({k[6:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("actor/")})
    critic.load_state_dict({k[7:]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("critic/")})
    return Z, actor, critic


@pytest.mark.parametrize("tc", PATHS)
def test_kat5_reference_checkpoint_through_mm_policy_forward(tc):
    """SURVEY 8c KAT(5) through the C ABI: the reference's own PPO.pth, packed by policy.pack_weights, evaluated by mm_policy_forward on
    the four facing one-hots -- against the logits the REFERENCE's networks.py computes from the same file (and the survey's printed
    values) -- and on 192 recorded observations with the recorded actions: logits, values and joint log-probs (PPO.get_log_probs)."""
    from marl_maze_b200.policy import PolicyRunner
    Z, actor, critic = _ppo_pth_nets()
    # --- the four facings.  Rows are agents; the kernel takes [E,2,65], so the 4 probes are laid out as 2 envs x 2 agents
    x = torch.from_numpy(Z["kat5/obs"]).cuda().view(2, 2, 65).contiguous()
    masks = torch.ones(2, 2, 6, dtype=torch.uint8, device="cuda")
    run = _runner(actor, critic, 2, tc)
    logits = torch.zeros(2, 2, 6, device="cuda")
    run.forward(x, masks, logits=logits)
    lg = logits.cpu().numpy().reshape(4, 6)
    assert np.allclose(lg[:, :5], Z["kat5/move_logits"], **TOL_LOGITS[tc]) and np.allclose(lg[:, 5], Z["kat5/mark_logits"], **TOL_LOGITS[tc])
    assert np.allclose(lg[0, :5], [5.7483, -0.3037, -7.4898, 10.1972, -9.9180], atol=2e-4)                 # SURVEY 8c KAT(5), as printed there
    assert np.allclose(1 / (1 + np.exp(-lg[:, 5])), [.5609, .5664, .5459, .5300], atol=2e-4)
    # --- recorded observations / actions
    obs, mk, acts = Z["trace/obs"], Z["trace/masks"], Z["trace/actions"]
    E = obs.shape[0]
    run = _runner(actor, critic, E, tc)
    logits = torch.zeros(E, 2, 6, device="cuda")
    _, logp, val = run.forward(torch.from_numpy(obs).cuda(), torch.from_numpy(mk).cuda(), actions_in=torch.from_numpy(acts).cuda(), logits=logits)
    lg = logits.cpu().numpy()
    assert np.allclose(lg[:, :, :5].reshape(-1, 5), Z["trace/move_logits"], **TOL_LOGITS[tc])
    assert np.allclose(lg[:, :, 5].reshape(-1), Z["trace/mark_logits"], **TOL_LOGITS[tc])
    assert np.allclose(val.cpu().numpy(), Z["trace/values"], **TOL)
    want = Z["trace/log_probs"].sum(1)
    fin = np.isfinite(want)
    got = logp.cpu().numpy()
    assert np.array_equal(np.isfinite(got), fin) and np.allclose(got[fin], want[fin], **TOL)


@pytest.mark.parametrize("faithful", [True, False])
def test_token_kernels_full_and_maps_only_agree_with_float64(faithful):
    """mm_tokens_forward (second generation: everything from the folded per-token maps -- the forward the update differentiates) and
    mm_tokens_forward_full (third generation, the rollout's: keys / queries / values as tensor-path products with att_q / att_k / att_v) are the same
    function of a consistently packed weight buffer: both within 1e-6 (of the largest element) of the float64 modules (Projection + m_Attention, networks.py:58-65,75-82),
    ragged row counts included."""
    import copy
    from marl_maze_b200 import _abi
    from marl_maze_b200.policy import pack_weights
    actor, critic, _, _ = _nets(7, faithful)
    w = pack_weights(actor, critic, "cuda")
    L = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
    a64 = copy.deepcopy(actor).double()
    rng = np.random.default_rng(11)
    for R in (1, 33, 2051):
        obs = torch.from_numpy((rng.random((R, 65)) * 2 - 0.5).astype(np.float32)).cuda()
        with torch.no_grad():
            ref = a64.attention(a64.projection(obs.double()))
        for fn in (L.mm_tokens_forward, L.mm_tokens_forward_full):
            x0 = torch.full((R, 460), float("nan"), device="cuda")
            _abi.check(fn(w.data_ptr(), obs.data_ptr(), R, x0.data_ptr(), st), "tokens")
            err, scale = float((x0.double() - ref).abs().max()), max(1.0, float(ref.abs().max()))
            assert err < 1e-6 * scale, (R, fn.__name__, err, scale)


def test_critic_kernel_ragged_env_counts_against_float64():
    """mm_critic_forward (3xFP16 mma.sync, 32 envs per warp) at env counts that do not fill a warp tile, against the float64 module."""
    import copy
    from marl_maze_b200.policy import PolicyRunner
    actor, critic, _, _ = _nets(9)
    W = [l.weight.detach().double() for l in critic.layers]; B = [l.bias.detach().double() for l in critic.layers]
    c64 = lambda x: torch.relu(torch.relu(x.reshape(-1, 130) @ W[0].t() + B[0]) @ W[1].t() + B[1]) @ W[2].t() + B[2]   # networks.py:96-102 in float64
    rng = np.random.default_rng(2)
    for E in (1, 15, 33, 4099):
        run = PolicyRunner(actor, critic, E, "cuda")
        obs = torch.from_numpy((rng.random((E, 2, 65)) * 2 - 0.5).astype(np.float32)).cuda()
        val = run.values(obs)
        with torch.no_grad():
            ref = c64(obs.double()).reshape(-1)
        assert torch.allclose(val.double(), ref, rtol=1e-5, atol=2e-6), (E, float((val.double() - ref).abs().max()))
