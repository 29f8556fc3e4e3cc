"""K5 (PPO actor update kernels) against torch fp64 autograd / matmul on the same inputs.  The kernels compute in 3xTF32 with fp32
accumulation, i.e. fp32-grade results with a different summation order: tolerances are relative to the magnitude of the result."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("R,k_in", [(32, 264), (1000, 264), (4096 + 17, 460), (70000, 264), (70000, 460)])
def test_wgrad_matches_fp64(R, k_in):
    from marl_maze_b200.update import wgrad
    g = torch.Generator(device="cuda"); g.manual_seed(R + k_in)
    dz = torch.randn(R, 264, device="cuda", generator=g) * (torch.rand(R, 264, device="cuda", generator=g) < 0.5)   # relu-gated gradient: half zeros
    h = torch.relu(torch.randn(R, k_in, device="cuda", generator=g)) + 0.25
    dW, db = wgrad(dz, h)
    ref_W = dz.double().t() @ h.double(); ref_b = dz.double().sum(0)
    assert dW.shape == (264, k_in) and db.shape == (264,)
    # the tensor core truncates when it accumulates: a slab of n rows sums ~n/3 MMA results into fp32, each rounded toward zero
    tol = 2e-6 + 3e-10 * R
    assert _rel(dW, ref_W) < tol, _rel(dW, ref_W)
    assert _rel(db, ref_b) < tol, _rel(db, ref_b)
    # a single-pass TF32 product would sit at ~3e-4: make sure the compensation terms are really there
    fp32 = dz.t() @ h
    assert _rel(dW, ref_W) < 4 * _rel(fp32, ref_W) + tol


def test_linear_modes_match_fp64():
    """mm_linear_tf32x3: forward (bias + ReLU), gated data gradient, and the 460-wide plain data gradient made of two column blocks."""
    from marl_maze_b200.update import linear_tc, tf32_split, MM_LINEAR_RELU, MM_LINEAR_GATE, MM_LINEAR_PLAIN
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    R = 1000
    x = torch.randn(R, 460, device="cuda", generator=g)
    w0 = torch.randn(264, 460, device="cuda", generator=g) / 20; b0 = torch.randn(264, device="cuda", generator=g)
    y = linear_tc(x, tf32_split(w0), MM_LINEAR_RELU, bias=b0)
    ref = torch.relu(x.double() @ w0.double().t() + b0.double())
    assert _rel(y, ref) < 5e-6   # K = 460 products summed by the tensor core with truncation (cf. TOL_LOGITS in test_policy_gpu.py)
    dz = torch.randn(R, 264, device="cuda", generator=g)
    w1 = torch.randn(264, 264, device="cuda", generator=g) / 16
    gate, bits = linear_tc(torch.randn(R, 264, device="cuda", generator=g), tf32_split(w1), MM_LINEAR_RELU, bias=b0, want_bits=True)
    unpacked = ((bits.unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(R, 288)[:, :264].bool()
    assert bool((unpacked == (gate > 0)).all()) and 0.3 < float(unpacked.float().mean()) < 0.7
    dh = linear_tc(dz, tf32_split(w1.t()), MM_LINEAR_GATE, gate_bits=bits)
    ref = (dz.double() @ w1.double()) * (gate > 0)
    assert _rel(dh, ref) < 5e-6
    assert bool(((gate > 0) | (dh == 0)).all())
    dx = torch.full((R, 460), float("nan"), device="cuda")
    w0t = w0.t().contiguous()
    for c0 in (0, 264):
        linear_tc(dz, tf32_split(w0t[c0:c0 + 264]), MM_LINEAR_PLAIN, out=dx, col0=c0)
    assert _rel(dx, dz.double() @ w0.double()) < 5e-6


def _loss_inputs(E, g, wide_ratio=True):
    masks = torch.rand(2 * E, 6, device="cuda", generator=g) < 0.6
    masks[torch.arange(2 * E, device="cuda"), torch.randint(0, 5, (2 * E,), device="cuda", generator=g)] = True      # at least one legal move
    moves = torch.multinomial(masks[:, :5].float(), 1, generator=g).squeeze(1)                                      # a legal move
    marks = (torch.rand(2 * E, device="cuda", generator=g) < 0.5) & masks[:, 5]
    actions = torch.stack([moves, marks.long()], 1).to(torch.uint8)
    adv = torch.randn(E, device="cuda", generator=g)
    return masks, actions, adv


@pytest.mark.parametrize("rows,n,k", [(1000, 264, 460), (4133, 264, 264), (129, 64, 132), (5000, 64, 64), (300, 196, 264), (128, 128, 36)])
def test_linear_f16x3_matches_fp64(rows, n, k):
    """mm_linear_f16x3 (3xFP16 on tcgen05, N-split tiles, two CTAs per SM) against fp64: the trunk shapes, the critic's, ragged row counts,
    a single-half width (n <= 128) and a second half narrower than 144; the ReLU bit words; weights spanning four decades."""
    from marl_maze_b200.policy import f16_split
    from marl_maze_b200.update import linear_f16
    g = torch.Generator(device="cuda"); g.manual_seed(rows + n + k)
    x = torch.randn(rows, k, device="cuda", generator=g) * (torch.rand(rows, 1, device="cuda", generator=g) * 4)
    w = torch.randn(n, k, device="cuda", generator=g) * 10 ** (torch.rand(n, 1, device="cuda", generator=g) * 4 - 3) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g) * 0.1
    y, bits = linear_f16(x, f16_split(w, (k + 31) // 32 * 32), b, want_bits=True)
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    scale = float((x.double().abs() @ w.double().abs().t()).max())        # the size of the sums that produced the outputs
    err = float((y.double() - ref).abs().max())
    assert err < 3e-6 * scale, (err, scale)
    # bit c*32+b of the row = y[row][32c+b] > 0, zero beyond column n
    want = torch.zeros(rows, 9 * 32, dtype=torch.bool, device="cuda"); want[:, :n] = y > 0
    got = ((bits.view(rows, 9, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).bool().view(rows, 288)
    assert torch.equal(got[:, :((n + 31) // 32) * 32], want[:, :((n + 31) // 32) * 32])


def test_ppo_heads_loss_matches_autograd():
    """Loss, joint log-probs and every gradient of mm_ppo_heads_loss against the reference formulas under torch fp64 autograd, with
    ratios on both sides of (and exactly inside) the clip range."""
    from marl_maze_b200.update import ppo_heads_loss
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    E = 5000
    z2 = torch.randn(2 * E, 264, device="cuda", generator=g)
    wh = torch.randn(6, 264, device="cuda", generator=g) / 8; bh = torch.randn(6, device="cuda", generator=g) / 4
    masks, actions, adv = _loss_inputs(E, g)

    def reference(z2, wh, bh, old=None):
        h2 = torch.relu(z2)
        logits = h2 @ wh.t() + bh
        mv = logits[:, :5].masked_fill(~masks[:, :5], float("-inf"))
        lp = torch.log_softmax(mv, -1).gather(1, actions[:, 0:1].long()).squeeze(1)
        p = torch.sigmoid(logits[:, 5].masked_fill(~masks[:, 5], float("-inf")))
        lp = lp + torch.log(torch.where(actions[:, 1].bool(), p, 1 - p))
        return lp.view(E, 2).sum(1)

    with torch.no_grad():
        joint0 = reference(z2.double(), wh.double(), bh.double())
    old = (joint0 + 0.35 * torch.randn(E, device="cuda", generator=g).double()).float()   # ratios from ~0.4 to ~2.5
    old[:100] = joint0[:100].float()                                                      # ratio == 1 up to rounding: inside the range
    clip, scale = 0.2, 1.0 / E
    zd, wd, bd = z2.double().requires_grad_(), wh.double().requires_grad_(), bh.double().requires_grad_()
    joint = reference(zd, wd, bd)
    ratio = torch.exp(joint - old.double())
    ref_loss = -(torch.min(ratio * adv.double(), torch.clamp(ratio, 1 - clip, 1 + clip) * adv.double())).sum() * scale
    ref_loss.backward()
    frac_clipped = float(((ratio < 1 - clip) | (ratio > 1 + clip)).float().mean())
    assert 0.2 < frac_clipped < 0.9
    loss, logp, dz2, dwh, dbh = ppo_heads_loss(torch.relu(z2), wh, bh, masks.view(torch.uint8), actions, old, adv, clip, scale)
    assert torch.allclose(logp.double(), joint.detach(), rtol=1e-5, atol=1e-5)
    assert abs(float(loss) - float(ref_loss.detach())) < 1e-5 * max(1.0, abs(float(ref_loss.detach())))
    # envs whose fp32 ratio lands on the other side of a clip boundary than the fp64 one would flip a whole gradient row: none expected
    assert _rel(dz2, zd.grad) < 2e-5, _rel(dz2, zd.grad)
    assert _rel(dwh, wd.grad) < 2e-5, _rel(dwh, wd.grad)
    assert _rel(dbh, bd.grad) < 2e-5, _rel(dbh, bd.grad)


@pytest.mark.parametrize("faithful", [True, False], ids=["column0_projection", "indexed_projection"])
def test_fused_actor_loss_gradients_match_autograd(faithful):
    """Every actor parameter gradient of the fused K5 path against the reference formulas under torch autograd in fp64.  The yardstick
    is the fp32 autograd path it replaces (PPO.joint_log_probs): log(1 - sigmoid(l)) in fp32 makes that path itself ~1e-4 off on
    saturated mark logits, so the fused gradients must be within 1e-5 of fp64 OR as close to it as fp32 autograd is (x4)."""
    import copy
    from marl_maze_b200.networks import Actor
    from marl_maze_b200.update import actor_loss, fused_available
    torch.manual_seed(21)
    actor = Actor([264, 264, 264], faithful_projection=faithful).cuda()
    with torch.no_grad():
        actor.move_head.weight.mul_(10); actor.mark_head.weight.mul_(10)   # un-do most of the 0.01 head scaling: non-trivial distributions
    assert fused_available(actor)
    g = torch.Generator(device="cuda"); g.manual_seed(22)
    E = 6000
    obs = torch.rand(2 * E, 65, device="cuda", generator=g)
    obs[:, :4] = torch.nn.functional.one_hot(torch.randint(0, 4, (2 * E,), device="cuda", generator=g), 4).float()
    masks, _, adv = _loss_inputs(E, g)
    ref_actor, a32 = copy.deepcopy(actor).double(), copy.deepcopy(actor)
    with torch.no_grad():   # actions SAMPLED from the policy, as in a rollout (a mark the policy gives probability ~0 would make log(1 - p) ill-conditioned in fp32)
        hh = ref_actor.attention(ref_actor.projection(obs.double()))
        for lin in ref_actor.layers:
            hh = torch.relu(lin(hh))
        pmove = torch.softmax(ref_actor.move_head(hh).masked_fill(~masks[:, :5], float("-inf")), -1)
        pmark = torch.sigmoid(ref_actor.mark_head(hh).reshape(-1)) * masks[:, 5]
        actions = torch.stack([torch.multinomial(pmove.float(), 1, generator=g).squeeze(1),
                               (torch.rand(2 * E, device="cuda", generator=g) < pmark.float()).long()], 1).to(torch.uint8)

    kept = {}
    # ReLU'(0): an fp32 pre-activation within rounding of zero may land on the other side than the fp64 one, and that single (row, unit)
    # then carries a gradient in one computation and none in the other -- a legitimate difference between ANY two float evaluations, as
    # large as the gradient element itself.  Take it out of the comparison: every reference below uses the fused forward's own gates.
    from marl_maze_b200.update import fwd_relu
    with torch.no_grad():
        from marl_maze_b200.update import token_embed
        hf, gates = (actor.embed(obs) if faithful else token_embed(actor, obs)).contiguous(), []   # the embedding path actor_loss takes
        for lin in actor.layers:
            hf = fwd_relu(hf, lin.weight, lin.bias)
            gates.append(hf > 0)

    def loss_ref(a, dt, old):   # Actor.forward + PPO.get_log_probs + the clipped surrogate in dtype dt (Actor.trunk itself casts to fp32)
        h = a.attention(a.projection(obs.to(dt)))
        if old is not None:
            h.retain_grad(); kept[dt] = h
        for lin, gate in zip(a.layers, gates):
            h = lin(h) * gate
        ml, kl = a.move_head(h), a.mark_head(h)
        lp = torch.log_softmax(ml.masked_fill(~masks[:, :5], float("-inf")), -1).gather(1, actions[:, 0:1].long()).squeeze(1)
        p = torch.sigmoid(kl.reshape(-1).masked_fill(~masks[:, 5], float("-inf")))
        joint = (lp + torch.log(torch.where(actions[:, 1].bool(), p, 1 - p))).view(E, 2).sum(1)
        if old is None:
            return joint
        ratio = torch.exp(joint - old.to(dt))
        return -(torch.min(ratio * adv.to(dt), torch.clamp(ratio, 0.8, 1.2) * adv.to(dt))).sum() / E

    with torch.no_grad():
        old = (loss_ref(ref_actor, torch.float64, None) + 0.3 * torch.randn(E, device="cuda", generator=g).double()).float()
    ref_loss = loss_ref(ref_actor, torch.float64, old); ref_loss.backward()
    loss_ref(a32, torch.float32, old).backward()
    parts = actor.embed_parts

    def parts_kept(x):
        emb, inv = parts(x)
        emb.retain_grad(); kept["fused"] = (emb, inv)
        return emb, inv

    actor.embed_parts = parts_kept
    loss, logp = actor_loss(actor, obs, masks, actions, old, adv, 0.2, 1.0 / E)
    loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < 2e-5 * max(1.0, abs(float(ref_loss.detach())))
    # the gradient handed back to autograd at the embedding output (element-wise, against the largest element): per agent row, or -- when the
    # embedding was evaluated once per distinct observation prefix -- summed over the rows that share an embedding row
    r_dx0 = r32_dx0 = float("nan")
    if faithful:   # (the indexed mode embeds every row with the token kernels: their gradients are checked below, parameter by parameter)
        emb, inv = kept["fused"]
        assert inv is not None
        want = {dt: torch.zeros(emb.shape[0], 460, device="cuda", dtype=dt).index_add_(0, inv, kept[dt].grad) for dt in (torch.float64, torch.float32)}
        r_dx0, r32_dx0 = _rel(emb.grad, want[torch.float64]), _rel(want[torch.float32], want[torch.float64])
        assert r_dx0 < max(2e-5, 4 * r32_dx0), (r_dx0, r32_dx0)
    else:
        assert "fused" not in kept
    worst = 0.0
    for (name, p), (_, q), (_, t) in zip(actor.named_parameters(), ref_actor.named_parameters(), a32.named_parameters()):
        if q.grad is None or float(q.grad.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-9, name
            continue
        assert p.grad is not None, name
        r, r32 = _rel(p.grad, q.grad), _rel(t.grad, q.grad)
        worst = max(worst, r)
        assert r < max(2e-5, 4 * r32), (name, r, r32)
    print("worst relative gradient error", worst, "dX0", r_dx0, r32_dx0)


@pytest.mark.parametrize("seed", [11, 12])
def test_fused_actor_loss_matches_reference_golden(seed):
    """The K5 path on the committed reference fixture (tests/golden/upd_kats.npz: the reference's own get_log_probs + actor_loss.backward(),
    torch CPU fp32, 384 envs of a recorded trace): loss, joint log-probs, every actor parameter gradient."""
    import os
    import numpy as np
    from golden_util import GOLDEN
    from oracle import ppo_oracle as po
    from marl_maze_b200.networks import Actor
    from marl_maze_b200.update import actor_loss
    Z = np.load(os.path.join(GOLDEN, "upd_kats.npz"))
    asd, _ = po.seeded_state_dicts(seed)
    actor = Actor([264, 264, 264]).cuda()
    actor.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()})
    obs = torch.from_numpy(Z["upd/obs"]).cuda().reshape(-1, 65); masks = torch.from_numpy(Z["upd/masks"]).cuda().reshape(-1, 6)
    acts = torch.from_numpy(Z["upd/actions"]).cuda().reshape(-1, 2)
    old = torch.from_numpy(Z[f"upd/{seed}/old"]).cuda(); adv = torch.from_numpy(Z[f"upd/{seed}/adv"]).cuda()
    E = old.shape[0]
    loss, logp = actor_loss(actor, obs, masks, acts, old, adv, float(Z["upd/clip"]), 1.0 / E)
    loss.backward()
    assert abs(float(loss.detach()) - float(Z[f"upd/{seed}/loss"])) < 1e-5 * max(1.0, abs(float(Z[f"upd/{seed}/loss"])))
    assert np.allclose(logp.cpu().numpy(), Z[f"upd/{seed}/joint"], rtol=1e-5, atol=1e-5)
    stride, worst = int(Z["upd/stride"]), 0.0
    for name, p in actor.named_parameters():
        ref = Z[f"upd/{seed}/grad/{name}"]
        got = p.grad.cpu().numpy()
        if name in ("layers.0.weight", "layers.1.weight", "layers.2.weight"):
            got = got.reshape(-1)[::stride]
        err = float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max())
        worst = max(worst, err)
        assert err < 2e-5, (name, err)
    print("worst relative gradient error vs the reference", worst)


@pytest.mark.parametrize("rows,n_seg", [(1, 1), (1000, 4), (200001, 8)])
def test_gather_rows_backward_is_segment_sum(rows, n_seg):
    from marl_maze_b200.update import GatherRows
    g = torch.Generator(device="cuda"); g.manual_seed(rows)
    src = torch.randn(n_seg, 460, device="cuda", generator=g, requires_grad=True)
    inv = torch.randint(0, n_seg, (rows,), device="cuda", generator=g)
    up = torch.randn(rows, 460, device="cuda", generator=g)
    out = GatherRows.apply(src, inv)
    assert torch.equal(out, src.detach()[inv])
    out.backward(up)
    ref = torch.zeros(n_seg, 460, device="cuda", dtype=torch.float64).index_add_(0, inv, up.double())
    assert _rel(src.grad, ref) < 2e-6


@pytest.mark.parametrize("faithful", [True, False], ids=["column0_projection", "indexed_projection"])
def test_token_embed_forward_and_backward_match_autograd(faithful):
    """update.token_embed (k_tokens / k_tokens_bwd + the differentiable folding of the parameters into per-token maps) against
    Projection + m_Attention under torch autograd in fp64: values and every embedding parameter gradient."""
    import copy
    from marl_maze_b200.networks import Actor
    from marl_maze_b200.update import token_embed
    torch.manual_seed(31)
    actor = Actor([264, 264, 264], faithful_projection=faithful).cuda()
    ref = copy.deepcopy(actor).double()
    g = torch.Generator(device="cuda"); g.manual_seed(32)
    B = 3001
    obs = torch.rand(B, 65, device="cuda", generator=g) * 2 - 0.5
    up = torch.randn(B, 460, device="cuda", generator=g)
    x0 = token_embed(actor, obs)
    want = ref.attention(ref.projection(obs.double()))
    assert _rel(x0, want) < 2e-6
    x0.backward(up); want.backward(up.double())
    worst = 0.0
    for (name, p), (_, q) in zip(actor.named_parameters(), ref.named_parameters()):
        if not name.startswith(("projection", "attention")):
            assert p.grad is None, name
            continue
        r = _rel(p.grad, q.grad)
        worst = max(worst, r)
        assert r < 2e-5, (name, r)
    print("worst embedding gradient error", worst)


def test_fused_critic_loss_matches_autograd():
    """update.critic_loss (K5 GEMM kernels for the critic's hidden layers) against Critic + MSE under torch autograd in fp64."""
    import copy
    from marl_maze_b200.networks import Critic
    from marl_maze_b200.update import critic_loss, critic_fused_available, pad_critic_obs, fwd_relu
    torch.manual_seed(41)
    critic = Critic(2, hidden_sizes=[64, 64]).cuda()
    assert critic_fused_available(critic)
    ref = copy.deepcopy(critic).double()
    g = torch.Generator(device="cuda"); g.manual_seed(42)
    n = 5003
    obs = torch.rand(n, 2, 65, device="cuda", generator=g)
    rtg = torch.randn(n, device="cuda", generator=g)
    xpad = pad_critic_obs(obs)
    loss = critic_loss(critic, xpad, rtg, 1.0 / n)
    loss.backward()
    # same ReLU gates in the reference as in the fused forward (see test_fused_actor_loss_gradients_match_autograd)
    with torch.no_grad():
        w0p = torch.nn.functional.pad(critic.layers[0].weight, (0, 2))
        h0 = fwd_relu(xpad, w0p, critic.layers[0].bias)
        h1 = fwd_relu(h0, critic.layers[1].weight, critic.layers[1].bias)
    x = obs.double().reshape(n, 130)
    r0 = ref.layers[0](x) * (h0 > 0)
    r1 = ref.layers[1](r0) * (h1 > 0)
    ref_loss = ((ref.layers[2](r1).squeeze(-1) - rtg.double()) ** 2).sum() / n
    ref_loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < 1e-5 * float(ref_loss.detach())
    for (name, p), (_, q) in zip(critic.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, q.grad) < 2e-5, (name, _rel(p.grad, q.grad))


def test_gather_rows_equals_index_select():
    """mm_gather_rows (the forward of the embedding gather whose adjoint is mm_segment_sum): a pure copy, bit-identical to index_select, for ragged row
    counts and every legal number of source rows."""
    from marl_maze_b200.update import gather_rows
    g = torch.Generator(device="cuda").manual_seed(3)
    for U, rows, cols in ((1, 1, 460), (4, 33, 460), (8, 100_003, 460), (5, 4097, 264), (3, 70_001, 64)):
        src = torch.randn(U, cols, device="cuda", generator=g)
        inv = torch.randint(0, U, (rows,), device="cuda", generator=g)
        assert torch.equal(gather_rows(src, inv), src.index_select(0, inv)), (U, rows, cols)
