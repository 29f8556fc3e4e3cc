"""Host side of the PPO update kernels (K5; PPO.py:58-85 of the reference: the clipped-surrogate actor loss and the critic's MSE, each
with its loss.backward()).  Thin wrappers over the C ABI -- torch owns every buffer, the library borrows pointers for the duration of a
call -- and the autograd nodes built from them:
    _ActorTrunkLoss  trunk + heads + masked log-probs + ratio + clipped surrogate, forward and backward (actor_loss)
    TokenEmbed       the 23-token projection + attention in front of the trunk, forward (K4's kernel) and backward (token_embed)
    GatherRows       embedding rows shared by many agents: gather forward, segment sum backward
    _CriticLoss      the centralised critic + MSE, forward and backward (critic_loss)
There is no CPU path: every entry point raises _abi.MMError when the library or a GPU is missing."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


WGRAD_MAX_ROWS = 1 << 18  # rows per mm_wgrad_tf32x3 launch: bounds how many MMA results one TMEM accumulator sums (the tensor core truncates)


def wgrad(dz: torch.Tensor, h: torch.Tensor):
    """(dW [n_out,k_in], db [n_out]) = (dz^T h, column sums of dz) for one Linear layer, 3xTF32 on tcgen05 (mm_wgrad_tf32x3)."""
    assert dz.is_cuda and h.is_cuda and dz.dtype == h.dtype == torch.float32 and dz.is_contiguous() and h.is_contiguous()
    R, n_out = dz.shape
    k_in = h.shape[1]
    assert h.shape[0] == R
    L = _abi.lib()
    tot = tr = None
    for r0 in range(0, R, WGRAD_MAX_ROWS):
        r = min(WGRAD_MAX_ROWS, R - r0)
        slabs, ld, out_rows, transposed = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _abi.check(L.mm_wgrad_geometry(r, n_out, k_in, C.byref(slabs), C.byref(ld), C.byref(out_rows), C.byref(transposed)), "mm_wgrad_geometry")
        part = torch.empty(slabs.value, out_rows.value, ld.value, device=dz.device, dtype=torch.float32)
        _abi.check(L.mm_wgrad_tf32x3(_ptr(dz[r0:r0 + r]), _ptr(h[r0:r0 + r]), r, n_out, k_in, _ptr(part), _stream(dz)), "mm_wgrad_tf32x3")
        t = part.sum(0)
        tot, tr = (t if tot is None else tot + t), transposed.value  # the orientation depends on (n_out, k_in) only
    if tr:  # [k_in + 1][ld]: rows = input columns (+ the bias row), columns = output units
        return tot[:k_in, :n_out].t().contiguous(), tot[k_in, :n_out]
    return tot[:, :k_in], tot[:, k_in]


def segment_sum(x: torch.Tensor, seg: torch.Tensor, n_seg: int) -> torch.Tensor:
    """out[s] = sum of the rows r of x with seg[r] == s (n_seg <= 8; mm_segment_sum)."""
    x = x.contiguous()
    L = _abi.lib()
    part = torch.empty(L.mm_segment_sum_blocks(x.shape[0]), n_seg, x.shape[1], device=x.device, dtype=torch.float32)
    _abi.check(L.mm_segment_sum(_ptr(x), _ptr(seg), x.shape[0], x.shape[1], n_seg, _ptr(part), _stream(x)), "mm_segment_sum")
    return part.sum(0)


def gather_rows(src: torch.Tensor, inv: torch.Tensor) -> torch.Tensor:
    """src[inv] for src [U <= 8, cols] f32 and inv [rows] int64 (mm_gather_rows: the source rows from shared memory, streaming stores at HBM write speed);
    anything else falls to index_select."""
    if not (src.is_cuda and src.dtype == torch.float32 and inv.dtype == torch.int64 and src.dim() == 2 and 1 <= src.shape[0] <= 8 and src.shape[1] % 4 == 0
            and src.shape[0] * src.shape[1] * 4 <= 48 * 1024 and inv.dim() == 1 and inv.numel() > 0):
        return src.index_select(0, inv)
    src, inv = src.contiguous(), inv.contiguous()
    out = torch.empty(inv.numel(), src.shape[1], device=src.device, dtype=torch.float32)
    _abi.check(_abi.lib().mm_gather_rows(_ptr(src), _ptr(inv), inv.numel(), src.shape[1], src.shape[0], _ptr(out), _stream(src)), "mm_gather_rows")
    return out


class GatherRows(torch.autograd.Function):
    """src[inv] for a handful (<= 8) of distinct source rows; the backward is mm_segment_sum instead of a scatter-add into those rows."""

    @staticmethod
    def forward(ctx, src, inv):
        ctx.save_for_backward(inv)
        ctx.n = src.shape[0]
        return gather_rows(src.detach(), inv)

    @staticmethod
    def backward(ctx, g):
        (inv,) = ctx.saved_tensors
        return segment_sum(g, inv, ctx.n), None


def token_maps(actor):
    """The per-token affine maps the token kernels evaluate, as a DIFFERENTIABLE function of the embedding parameters:
    [token; key; query; value]_a = [I; Wk; Wq; Wv] (P_a x_a + b_a)  ->  tokm [60,23,4], tokb [60,23]   (policy.pack_weights, csrc/mm_policy.cu)."""
    import torch.nn.functional as F
    from .networks import EMBEDDING_DIM
    layers = actor.projection.layers
    pw = torch.stack([F.pad(l.weight, (0, 4 - l.weight.shape[1])) for l in layers], 0)                     # [23,20,4]
    pb = torch.stack([l.bias for l in layers], 0)                                                          # [23,20]
    att = actor.attention
    stack = torch.cat([torch.eye(EMBEDDING_DIM, device=pw.device, dtype=pw.dtype), att.keys.weight, att.querys.weight, att.values.weight], 0)  # [60,20]
    return torch.einsum("jd,adc->jac", stack, pw).contiguous(), (stack @ pb.t()).contiguous()


_TOKEN_LAYOUT = {}


def token_layout(device, faithful):
    """(first observation column, width) of the 23 feature tokens as device tensors -- host -> device once per (device, projection mode): nothing in
    the update may copy from the host during a CUDA-graph capture (PPO._update calls this before it captures)."""
    from .networks import FEATURE_DIMS
    ck = (str(torch.device(device)), bool(faithful))
    if ck not in _TOKEN_LAYOUT:
        cols = [0 if faithful else sum(FEATURE_DIMS[:i]) for i in range(len(FEATURE_DIMS))]
        _TOKEN_LAYOUT[ck] = (torch.tensor(cols, dtype=torch.float32, device=device), torch.tensor(FEATURE_DIMS, dtype=torch.float32, device=device))
    return _TOKEN_LAYOUT[ck]


class TokenEmbed(torch.autograd.Function):
    """Projection + attention of every row by the token kernels: forward = K4's k_tokens, backward = k_tokens_bwd (gradients at the
    per-token maps; autograd carries them on to the parameters through token_maps)."""

    @staticmethod
    def forward(ctx, obs, tokm, tokb, faithful):
        from .networks import FEATURE_DIMS
        from .policy import offsets
        o = offsets()
        obs = obs.contiguous()
        buf = torch.zeros(o["total"], dtype=torch.float32, device=obs.device)   # K4 buffer layout; the token kernels read these four blocks only
        buf[o["tokm"]:o["tokm"] + tokm.numel()] = tokm.detach().reshape(-1)
        buf[o["tokb"]:o["tokb"] + tokb.numel()] = tokb.detach().reshape(-1)
        d_cols, d_dims = token_layout(obs.device, faithful)
        buf[o["proj_col"]:o["proj_col"] + d_cols.numel()] = d_cols
        buf[o["proj_dim"]:o["proj_dim"] + d_dims.numel()] = d_dims
        x0 = torch.empty(obs.shape[0], 460, device=obs.device, dtype=torch.float32)
        _abi.check(_abi.lib().mm_tokens_forward(_ptr(buf), _ptr(obs), obs.shape[0], _ptr(x0), _stream(obs)), "mm_tokens_forward")
        ctx.save_for_backward(obs, buf)
        return x0

    @staticmethod
    def backward(ctx, g):
        obs, buf = ctx.saved_tensors
        L = _abi.lib()
        g = g.contiguous()
        tot = None
        for r0 in range(0, obs.shape[0], TOKENS_BWD_MAX_ROWS):   # bounds the per-row scratch (3680 B per row)
            r = min(TOKENS_BWD_MAX_ROWS, obs.shape[0] - r0)
            scratch = torch.empty(int(L.mm_sizeof_tokens_backward_scratch(r)), dtype=torch.uint8, device=g.device)
            part = torch.empty(L.mm_tokens_backward_blocks(), 60, 23, 5, device=g.device, dtype=torch.float32)
            _abi.check(L.mm_tokens_backward(_ptr(buf), _ptr(obs[r0:r0 + r]), _ptr(g[r0:r0 + r]), r, _ptr(scratch), _ptr(part), _stream(g)), "mm_tokens_backward")
            t = part.sum(0)
            tot = t if tot is None else tot + t
        return None, tot[..., :4].contiguous(), tot[..., 4].contiguous(), None


TOKENS_BWD_MAX_ROWS = 1 << 20


def token_embed(actor, obs2):
    """[B,460] embedding of obs2 [B,65] through the token kernels (any projection mode, no de-duplication)."""
    tokm, tokb = token_maps(actor)
    return TokenEmbed.apply(obs2, tokm, tokb, bool(actor.projection.faithful))


MM_LINEAR_RELU, MM_LINEAR_GATE, MM_LINEAR_PLAIN = 0, 2, 3


def tf32_split(w: torch.Tensor):
    """w = hi + lo with hi = tf32(w), lo = tf32(w - hi), round-to-nearest (the operand form mm_linear_tf32x3 takes)."""
    w = w.detach().contiguous()
    hi = ((w.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    r = w - hi
    return hi, ((r.view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def linear_tc(x, w_split, mode, bias=None, gate_bits=None, out=None, col0=0, want_bits=False):
    """out[:, col0:col0+n] = epi(x @ W^T) for W [n<=264, k] given as tf32_split(W); out defaults to a fresh [rows, n] tensor.
    MM_LINEAR_RELU with want_bits=True also returns the ReLU pattern as bit words [rows, 9] int32 -- the `gate_bits` a later
    MM_LINEAR_GATE call (the data gradient through that ReLU) takes."""
    w_hi, w_lo = w_split
    rows, k = x.shape
    n = w_hi.shape[0]
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and w_hi.shape == w_lo.shape == (n, k) and w_hi.is_contiguous() and w_lo.is_contiguous()
    if out is None:
        out = torch.empty(rows, n, device=x.device, dtype=torch.float32)
    assert out.is_contiguous() and out.shape[0] == rows and col0 + n <= out.shape[1]
    bits = torch.empty(rows, 9, device=x.device, dtype=torch.int32) if want_bits else None
    assert gate_bits is None or (gate_bits.shape == (rows, 9) and gate_bits.dtype == torch.int32 and gate_bits.is_contiguous())
    y = C.c_void_p(out.data_ptr() + 4 * col0)
    _abi.check(_abi.lib().mm_linear_tf32x3(_ptr(x), rows, k, _ptr(w_hi), _ptr(w_lo), n, _ptr(bias), _ptr(gate_bits), y, out.shape[1], mode, _ptr(bits),
                                           _stream(x)), "mm_linear_tf32x3")
    return (out, bits) if want_bits else out


def linear_f16(x, w_split16, bias, out=None, want_bits=False):
    """relu(x @ W^T + bias) through the 3xFP16 kernel (mm_linear_f16x3): W [n<=264, k] given as policy.f16_split(W, kpad) = (hi16, lo16,
    acc_scale) with kpad = k rounded up to 32.  want_bits=True also returns the ReLU pattern [rows, 9] int32 (see linear_tc)."""
    w_hi, w_lo, asc = w_split16
    rows, k = x.shape
    n, kpad = w_hi.shape
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and w_hi.dtype == w_lo.dtype == torch.float16 and w_lo.shape == w_hi.shape
    assert kpad == (k + 31) // 32 * 32 and w_hi.is_contiguous() and w_lo.is_contiguous() and asc.dtype == torch.float32 and asc.is_cuda
    if out is None:
        out = torch.empty(rows, n, device=x.device, dtype=torch.float32)
    assert out.is_contiguous() and out.shape == (rows, n)
    bits = torch.empty(rows, 9, device=x.device, dtype=torch.int32) if want_bits else None
    _abi.check(_abi.lib().mm_linear_f16x3(_ptr(x), rows, k, _ptr(w_hi), _ptr(w_lo), n, kpad, _ptr(asc), _ptr(bias.detach().contiguous()), _ptr(out), n, _ptr(bits),
                                          _stream(x)), "mm_linear_f16x3")
    return (out, bits) if want_bits else out


FWD_FP16 = __import__("os").environ.get("MARL_MAZE_TF32_TRUNK", "0") != "1"   # forward GEMMs of the update on the 3xFP16 kernel (default)


def fwd_relu(x, w, bias, want_bits=False):
    """relu(x @ w^T + bias) through whichever forward GEMM the fused update uses (3xFP16 by default, 3xTF32 with MARL_MAZE_TF32_TRUNK=1):
    what tests call to reproduce the fused forward's own ReLU gates."""
    if FWD_FP16:
        from .policy import f16_split
        return linear_f16(x, f16_split(w, (w.shape[1] + 31) // 32 * 32), bias, want_bits=want_bits)
    return linear_tc(x, tf32_split(w), MM_LINEAR_RELU, bias=bias.detach().contiguous(), want_bits=want_bits)


def ppo_heads_loss(h2, head_w, head_b, masks, actions, old_logp, adv, clip, scale):
    """-> (loss [] , joint log-prob [E], dz2 [2E,264], dhead_w [6,264], dhead_b [6]); see mm_ppo_heads_loss in the header."""
    E = old_logp.shape[0]
    assert h2.shape == (2 * E, 264) and h2.is_contiguous() and masks.dtype == torch.uint8 and actions.dtype == torch.uint8
    assert masks.is_contiguous() and actions.is_contiguous() and masks.numel() == 12 * E and actions.numel() == 4 * E
    L = _abi.lib()
    blocks, ld = C.c_int32(), C.c_int32()
    _abi.check(L.mm_ppo_loss_geometry(C.byref(blocks), C.byref(ld)), "mm_ppo_loss_geometry")
    part = torch.empty(blocks.value, ld.value, device=h2.device, dtype=torch.float32)
    dz2 = torch.empty_like(h2)
    logp = torch.empty(E, device=h2.device, dtype=torch.float32)
    head_w, head_b = head_w.detach().contiguous(), head_b.detach().contiguous()
    _abi.check(L.mm_ppo_heads_loss(_ptr(h2), _ptr(head_w), _ptr(head_b), _ptr(masks), _ptr(actions), _ptr(old_logp.contiguous()), _ptr(adv.contiguous()), E,
                                   float(clip), float(scale), _ptr(dz2), _ptr(logp), _ptr(part), _stream(h2)), "mm_ppo_heads_loss")
    tot = part.sum(0)
    return tot[6 * 264 + 6], logp, dz2, tot[:6 * 264].view(6, 264), tot[6 * 264:6 * 264 + 6]


def trunk_splits(actor):
    """TF32 splits of the three trunk weights, plain (forward) and transposed (data gradient; W0^T as two blocks of <= 264 rows), cached
    on the module until an optimizer step changes a weight (tensor._version)."""
    ws = [l.weight for l in actor.layers]
    key = tuple((w.data_ptr(), w._version) for w in ws)
    cached = getattr(actor, "_k5_splits", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            w0t = ws[0].t().contiguous()
            from .policy import f16_split
            data = dict(fwd=[tf32_split(w) for w in ws], t1=tf32_split(ws[1].t()), t2=tf32_split(ws[2].t()),
                        t0=[tf32_split(w0t[c0:c0 + 264]) for c0 in range(0, w0t.shape[0], 264)],
                        fwd16=[f16_split(w, (w.shape[1] + 31) // 32 * 32) for w in ws])
        cached = (key, data)
        actor._k5_splits = cached
    return cached[1]


class _ActorTrunkLoss(torch.autograd.Function):
    """PPO actor loss as ONE autograd node from the embedding output x0 [2E,460] down: three Linear+ReLU layers, the two heads, masked
    Categorical / Bernoulli log-probs of the recorded actions, ratio, clipped surrogate (Actor.forward networks.py:36-41,
    PPO.get_log_probs PPO.py:154-168, PPO.train PPO.py:66-72).  The loss is a scalar, so the whole backward is evaluated inside
    forward() (no activation outlives the call) and backward() only scales the stored gradients."""

    @staticmethod
    def forward(ctx, x0, inv, w0, b0, w1, b1, w2, b2, wh, bh, sp, masks, actions, old_logp, adv, clip, scale):
        """x0 [2E,460], inv None -- or x0 = the few DISTINCT embedding rows [U<=8,460] and inv [2E] the row of each agent (Actor.embed_parts)."""
        src = x0.detach().contiguous()
        x0 = src if inv is None else gather_rows(src, inv)
        if FWD_FP16:   # forward GEMMs: 3xFP16, two CTAs per SM (the data / weight gradients below need TF32's exponent range)
            h0, bits0 = linear_f16(x0, sp["fwd16"][0], b0, want_bits=True)
            h1, bits1 = linear_f16(h0, sp["fwd16"][1], b1, want_bits=True)
            h2 = linear_f16(h1, sp["fwd16"][2], b2)
        else:
            h0, bits0 = linear_tc(x0, sp["fwd"][0], MM_LINEAR_RELU, bias=b0.detach().contiguous(), want_bits=True)
            h1, bits1 = linear_tc(h0, sp["fwd"][1], MM_LINEAR_RELU, bias=b1.detach().contiguous(), want_bits=True)
            h2 = linear_tc(h1, sp["fwd"][2], MM_LINEAR_RELU, bias=b2.detach().contiguous())
        loss, logp, dz2, dwh, dbh = ppo_heads_loss(h2, wh, bh, masks, actions, old_logp, adv, clip, scale)
        del h2
        dw2, db2 = wgrad(dz2, h1)
        dz1 = linear_tc(dz2, sp["t2"], MM_LINEAR_GATE, gate_bits=bits1)
        del dz2, h1
        dw1, db1 = wgrad(dz1, h0)
        dz0 = linear_tc(dz1, sp["t1"], MM_LINEAR_GATE, gate_bits=bits0)
        del dz1, h0
        dx0 = None
        if inv is not None:
            # Layer 0's input is a GATHER of a few source rows, X0 = G src.  Both gradients that involve X0 are adjoints of that gather,
            # and the gather's adjoint is a segment sum over agent rows that commutes with the products by W0 / src:
            #     d loss / d src = G^T (dZ0 W0) = (G^T dZ0) W0,      dW0 = dZ0^T (G src) = (G^T dZ0)^T src,      db0 = sum_u (G^T dZ0)_u
            # -- one streaming pass over [2E,264] (mm_segment_sum) and two [U x 264 x 460] products instead of the 460-wide data-gradient
            # GEMM, a pass over its [2E,460] result, and the 460-wide weight-gradient GEMM
            seg = segment_sum(dz0, inv, src.shape[0])                     # [U,264] = G^T dZ0
            dw0, db0 = seg.t() @ src, seg.sum(0)
            if ctx.needs_input_grad[0]:
                dx0 = seg @ w0.detach()
        else:
            dw0, db0 = wgrad(dz0, x0)
            if ctx.needs_input_grad[0]:  # the 460 output columns of dX0 = dZ0 W0 are produced as two blocks of <= 264
                dx0 = torch.empty_like(x0)
                for i, blk in enumerate(sp["t0"]):
                    linear_tc(dz0, blk, MM_LINEAR_PLAIN, out=dx0, col0=264 * i)
        ctx.grads = (dx0, None, dw0, db0, dw1, db1, dw2, db2, dwh, dbh)
        ctx.mark_non_differentiable(logp)
        return loss, logp

    @staticmethod
    def backward(ctx, g, _g_logp):
        grads, ctx.grads = ctx.grads, None
        # g is 1 for loss.backward(); reading it would cost a device synchronisation per micro-batch, so scale instead -- in place for a
        # full [2E,460] dX0 (one pass, no second buffer), out of place for the small tensors
        grads = tuple(None if t is None else (t.mul_(g) if t.numel() > (1 << 22) else g * t) for t in grads)
        return grads + (None,) * 7


def fused_available(actor) -> bool:
    """The fused update is built for the reference architecture (460 -> 264 -> 264 -> 264 -> 5+1, ReLU) on a CUDA device."""
    import torch.nn as nn
    ls = actor.layers
    return (len(ls) == 3 and tuple(l.out_features for l in ls) == (264, 264, 264) and ls[0].in_features == 460 and actor.activation is nn.ReLU
            and ls[0].weight.is_cuda)


def actor_loss(actor, obs2, masks2, actions2, old_logp, adv, clip, scale):
    """scale * sum_e -min(ratio_e A_e, clip(ratio_e) A_e) over E envs, and the new joint log-probs [E] (no grad).
    obs2 [2E,65] f32 (agent rows 2e, 2e+1), masks2 [2E,6] bool/u8, actions2 [2E,2] (move, mark) any integer/float dtype."""
    x0, inv = (actor.embed_parts(obs2) if actor.projection.faithful and obs2.shape[0] >= 4096 else (None, None))
    if inv is None or x0.shape[0] > 8:   # no (or too little) duplication among the rows: every row through the token kernels
        x0, inv = token_embed(actor, obs2), None
    ls = actor.layers
    wh = torch.cat([actor.move_head.weight, actor.mark_head.weight], 0)
    bh = torch.cat([actor.move_head.bias, actor.mark_head.bias], 0)
    masks = masks2.contiguous().view(torch.uint8) if masks2.dtype == torch.bool else masks2.to(torch.uint8).contiguous()
    actions = actions2.to(torch.uint8).contiguous()
    return _ActorTrunkLoss.apply(x0, inv, ls[0].weight, ls[0].bias, ls[1].weight, ls[1].bias, ls[2].weight, ls[2].bias, wh, bh, trunk_splits(actor), masks,
                                 actions, old_logp, adv, float(clip), float(scale))


# ------------------------------------------------------------------------------------------------ critic (PPO.py:79-84)
CRITIC_IN = 130   # both agents' observations (networks.py:96-102); padded to 132 columns: TMA rows must be multiples of 16 bytes


def pad_critic_obs(obs: torch.Tensor) -> torch.Tensor:
    """[n,2,65] -> [n,132] (two zero columns), the input layout critic_loss takes; PPO.update pads the whole shuffled rollout once."""
    return torch.nn.functional.pad(obs.reshape(obs.shape[0], CRITIC_IN), (0, 2))


class _CriticLoss(torch.autograd.Function):
    """scale * sum_e (V(s_e) - rtg_e)^2 through the centralised critic 130 -> 64 -> 64 -> 1 (Critic.forward networks.py:96-102, the MSE of
    PPO.py:79) as one autograd node evaluated forward AND backward in forward(), like _ActorTrunkLoss: the two hidden layers, their data
    gradient and their weight gradients are the K5 GEMM kernels (64 of the 272-column tile used); the 64 -> 1 output layer and the
    element-wise steps around it are torch."""

    @staticmethod
    def forward(ctx, xpad, w0, b0, w1, b1, w2, b2, rtg, scale):
        w0p = torch.nn.functional.pad(w0.detach(), (0, xpad.shape[1] - w0.shape[1]))
        if FWD_FP16:
            from .policy import f16_split
            h0, bits0 = linear_f16(xpad, f16_split(w0p, (w0p.shape[1] + 31) // 32 * 32), b0, want_bits=True)
            h1 = linear_f16(h0, f16_split(w1, (w1.shape[1] + 31) // 32 * 32), b1)
        else:
            h0, bits0 = linear_tc(xpad, tf32_split(w0p), MM_LINEAR_RELU, bias=b0.detach().contiguous(), want_bits=True)
            h1 = linear_tc(h0, tf32_split(w1), MM_LINEAR_RELU, bias=b1.detach().contiguous())
        v = torch.addmv(b2.detach(), h1, w2.detach()[0])
        diff = v - rtg
        loss = (diff * diff).sum() * scale
        dv = diff * (2.0 * scale)
        dw2, db2 = (dv @ h1).unsqueeze(0), dv.sum().reshape(1)
        dz1 = (dv.unsqueeze(1) * w2.detach()) * (h1 > 0)
        dw1, db1 = wgrad(dz1, h0)
        dz0 = linear_tc(dz1, tf32_split(w1.detach().t()), MM_LINEAR_GATE, gate_bits=bits0)
        dw0p, db0 = wgrad(dz0, xpad)
        ctx.grads = (dw0p[:, :w0.shape[1]].contiguous(), db0, dw1, db1, dw2, db2)
        return loss

    @staticmethod
    def backward(ctx, g):
        grads, ctx.grads = ctx.grads, None
        return (None,) + tuple(g * t for t in grads) + (None, None)


def critic_fused_available(critic) -> bool:
    ls = critic.layers
    import torch.nn as nn
    return (len(ls) == 3 and ls[0].in_features == CRITIC_IN and ls[0].out_features == 64 and ls[1].out_features == 64 and ls[2].out_features == 1
            and critic.activation is nn.ReLU and ls[0].weight.is_cuda)


def critic_loss(critic, xpad, rtg, scale):
    """scale * sum (critic(obs) - rtg)^2 for xpad = pad_critic_obs(obs) [n,132], rtg [n]."""
    ls = critic.layers
    return _CriticLoss.apply(xpad.contiguous(), ls[0].weight, ls[0].bias, ls[1].weight, ls[1].bias, ls[2].weight, ls[2].bias, rtg.contiguous(), float(scale))
