"""PolicyRunner: batched actor/critic forward with fused action sampling (K4) for the rollout.

Packs the reference-shaped state_dicts (networks.Actor / networks.Critic) into the flat fp32 buffer the kernels read and
launches mm_policy_forward.  Used by PPO.get_action / PPO.get_batch; autograd never sees this path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _abi
from .networks import FEATURE_DIMS, EMBEDDING_DIM

_NAMES = ["proj_w", "proj_b", "proj_col", "proj_dim", "att_k", "att_q", "att_v", "l0_w", "l0_b", "l1_w", "l1_b", "l2_w", "l2_b",
          "head_w", "head_b", "c0_w", "c0_b", "c1_w", "c1_b", "c2_w", "c2_b", "total", "l0_whi", "l0_wlo", "l1_whi", "l1_wlo", "l2_whi", "l2_wlo", "c0_wt", "c1_wt", "tokm", "tokb",
          "l0_h16", "l0_l16", "l1_h16", "l1_l16", "l2_h16", "l2_l16", "l0_asc", "l1_asc", "l2_asc", "lh_h16", "lh_l16", "lh_asc"]


def offsets() -> dict:
    out = (C.c_int32 * len(_NAMES))()   # MM_POLICY_N_OFFSETS
    _abi.check(_abi.lib().mm_policy_offsets(out), "mm_policy_offsets")
    return {n: int(out[i]) for i, n in enumerate(_NAMES)}


def _tf32_rn(x: torch.Tensor) -> torch.Tensor:
    """Nearest TF32 value (10 explicit mantissa bits, ties away from zero) -- same rounding as cvt.rna.tf32.f32 on the device."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def f16_split(w: torch.Tensor, kpad: int | None = None):
    """The operand form of mm_linear_f16x3 / MM_POLICY_FP16_SPLIT: (hi, lo, acc_scale) with hi = fp16(2^e w), lo = fp16(2^e w - hi) (both round
    to nearest; fp16 [n][kpad], zero padded columns) and acc_scale = 2^-e as a one-element fp32 DEVICE tensor.  e puts the largest |w| in
    [256, 512): exact scaling, every hi / lo a normal fp16 number unless |w| < 2^-23 max|w|.  No host synchronisation."""
    w = w.detach().to(torch.float32)
    n, k = w.shape
    kpad = k if kpad is None else kpad
    m = w.abs().max().clamp_min(1e-30)
    e = torch.floor(torch.log2(512.0 / m)).clamp(-14.0, 24.0)
    ws = w * torch.exp2(e)
    hi = ws.to(torch.float16)
    lo = (ws - hi.to(torch.float32)).to(torch.float16)
    if kpad != k:
        hi = torch.nn.functional.pad(hi, (0, kpad - k)); lo = torch.nn.functional.pad(lo, (0, kpad - k))
    return hi.contiguous(), lo.contiguous(), torch.exp2(-e).reshape(1).to(torch.float32)


def pack_weights(actor, critic, device=None) -> torch.Tensor:
    """Flatten actor+critic parameters into the layout of mm_policy_offsets (include/marl_maze_b200.h)."""
    if tuple(l.out_features for l in actor.layers) != (264, 264, 264) or actor.layers[0].in_features != 460:
        raise ValueError("the fused policy kernel is built for the reference architecture Actor([264,264,264])")
    if tuple(l.out_features for l in critic.layers) != (64, 64, 1) or critic.layers[0].in_features != 130:
        raise ValueError("the fused policy kernel is built for the reference architecture Critic(2, [64,64])")
    o = offsets()
    dev = device or actor.move_head.weight.device
    buf = torch.zeros(o["total"], dtype=torch.float32, device=dev)
    with torch.no_grad():
        pw = torch.zeros(len(FEATURE_DIMS), EMBEDDING_DIM, 4, device=dev)
        col = torch.zeros(len(FEATURE_DIMS), device=dev); dim = torch.zeros(len(FEATURE_DIMS), device=dev)
        c = 0
        for i, (lin, d) in enumerate(zip(actor.projection.layers, FEATURE_DIMS)):
            pw[i, :, :d] = lin.weight
            col[i] = 0 if actor.projection.faithful else c
            dim[i] = d
            c += d

        def put(name, t):
            t = t.detach().to(dev, torch.float32).reshape(-1)
            buf[o[name]:o[name] + t.numel()] = t
        put("proj_w", pw); put("proj_b", torch.cat([l.bias for l in actor.projection.layers])); put("proj_col", col); put("proj_dim", dim)
        put("att_k", actor.attention.keys.weight); put("att_q", actor.attention.querys.weight); put("att_v", actor.attention.values.weight)
        for i in range(3):
            w = actor.layers[i].weight.detach().to(dev, torch.float32).contiguous()
            put(f"l{i}_w", w); put(f"l{i}_b", actor.layers[i].bias)
            hi = _tf32_rn(w)
            put(f"l{i}_whi", hi); put(f"l{i}_wlo", _tf32_rn(w - hi))
            kpad = (w.shape[1] + 31) // 32 * 32
            h16, l16, asc = f16_split(w, kpad)               # fp16 [264][kpad]: two halves per float slot of the flat buffer
            n16 = h16.numel() // 2
            buf[o[f"l{i}_h16"]:o[f"l{i}_h16"] + n16] = h16.reshape(-1).view(torch.float32)
            buf[o[f"l{i}_l16"]:o[f"l{i}_l16"] + n16] = l16.reshape(-1).view(torch.float32)
            buf[o[f"l{i}_asc"]:o[f"l{i}_asc"] + 1] = asc
        head_w = torch.cat([actor.move_head.weight, actor.mark_head.weight], 0).detach().to(dev, torch.float32)
        put("head_w", head_w); put("head_b", torch.cat([actor.move_head.bias, actor.mark_head.bias]))
        h16, l16, asc = f16_split(torch.nn.functional.pad(head_w, (0, 0, 0, 10)), 288)   # the heads as a fourth layer of the fused trunk kernel: [16][288]
        buf[o["lh_h16"]:o["lh_h16"] + h16.numel() // 2] = h16.reshape(-1).view(torch.float32)
        buf[o["lh_l16"]:o["lh_l16"] + l16.numel() // 2] = l16.reshape(-1).view(torch.float32)
        buf[o["lh_asc"]:o["lh_asc"] + 1] = asc
        for i in range(3):
            put(f"c{i}_w", critic.layers[i].weight); put(f"c{i}_b", critic.layers[i].bias)
        put("c0_wt", critic.layers[0].weight.t().contiguous()); put("c1_wt", critic.layers[1].weight.t().contiguous())
        # per-token affine maps: [token; key; query; value] = [I; Wk; Wq; Wv] (P_a x + b_a) -> rows 0-19, 20-29, 30-39, 40-59
        att = actor.attention
        stack = torch.cat([torch.eye(EMBEDDING_DIM, device=dev), att.keys.weight.detach().to(dev, torch.float32),
                           att.querys.weight.detach().to(dev, torch.float32), att.values.weight.detach().to(dev, torch.float32)], 0).double()  # [60,20]
        tokm = torch.einsum("jd,adc->jac", stack, pw.double()).float()                       # [60,23,4]
        tokb = (stack @ torch.stack([l.bias.detach().to(dev) for l in actor.projection.layers], 0).double().t()).float()  # [60,23]
        put("tokm", tokm.contiguous()); put("tokb", tokb.contiguous())
    return buf


class PolicyRunner:
    def __init__(self, actor, critic, num_envs: int, device, env_offset: int = 0, seed: int = 0, tensor_cores: bool = True, overlap_critic: bool = True,
                 fp16_split: bool | None = None, fused_trunk: bool | None = None):
        self.lib = _abi.lib()
        self.E, self.device, self.env_offset, self.seed = int(num_envs), torch.device(device), int(env_offset), int(seed) & (2**64 - 1)
        self.actor, self.critic = actor, critic
        self.weights = pack_weights(actor, critic, self.device)
        self.scratch = torch.empty(int(self.lib.mm_sizeof_policy_scratch(self.E)), dtype=torch.uint8, device=self.device)
        self.counter = 0
        self.counter_dev = torch.zeros(1, dtype=torch.int64, device=self.device)  # added to `counter` on the device (CUDA-graph replays)
        self.launches = 0
        if fp16_split is None:   # default: the 3xFP16 two-CTA-per-SM kernel; MARL_MAZE_TF32_TRUNK=1 keeps the 3xTF32 one (A/B runs)
            fp16_split = os.environ.get("MARL_MAZE_TF32_TRUNK", "0") != "1"
        if fused_trunk is None:  # default: the three trunk layers + heads as one persistent kernel; MARL_MAZE_FUSED_TRUNK=0 keeps one kernel per layer
            fused_trunk = os.environ.get("MARL_MAZE_FUSED_TRUNK", "1") != "0"
        fused_trunk = bool(fused_trunk and tensor_cores and fp16_split)
        # MM_POLICY_TCGEN05 | MM_POLICY_OVERLAP_CRITIC | MM_POLICY_FP16_SPLIT | MM_POLICY_FUSED_TRUNK
        self.flags = (1 if tensor_cores else 0) | (2 if overlap_critic else 0) | (4 if (tensor_cores and fp16_split) else 0) | (8 if fused_trunk else 0)

    def refresh(self):
        """Re-pack after an optimiser step -- in place, so that captured CUDA graphs keep pointing at live weights."""
        self.weights.copy_(pack_weights(self.actor, self.critic, self.device))

    def bump(self, n: int):
        """Advance the device-side counter by n (captured at the end of a rollout graph)."""
        _abi.check(self.lib.mm_counter_add(C.c_void_p(self.counter_dev.data_ptr()), C.c_uint64(n),
                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "mm_counter_add")

    def values(self, obs: torch.Tensor, value: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Critic only (bootstrap value of the state after the last rollout step)."""
        E = obs.shape[0]
        value = torch.empty(E, dtype=torch.float32, device=self.device) if value is None else value
        _abi.check(self.lib.mm_critic_forward(C.c_void_p(self.weights.data_ptr()), C.c_void_p(obs.data_ptr()), E, C.c_void_p(value.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "mm_critic_forward")
        self.launches += 1
        return value

    def forward(self, obs: torch.Tensor, masks: torch.Tensor, actions_in: Optional[torch.Tensor] = None, actions_out: Optional[torch.Tensor] = None,
                logp: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None, want_value: bool = True,
                counter: Optional[int] = None):
        """Sample (actions_in None) or evaluate actions for all envs.  Returns (actions [E,2,2] u8, joint logp [E], value [E] | None)."""
        E = obs.shape[0]
        assert E <= self.E and obs.is_contiguous() and masks.is_contiguous() and obs.dtype == torch.float32 and masks.dtype == torch.uint8
        if actions_in is None and actions_out is None:
            actions_out = torch.empty(E, 2, 2, dtype=torch.uint8, device=self.device)
        logp = torch.empty(E, dtype=torch.float32, device=self.device) if logp is None else logp
        if want_value and value is None:
            value = torch.empty(E, dtype=torch.float32, device=self.device)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        if counter is None:
            self.counter += 1
            counter = self.counter
        _abi.check(self.lib.mm_policy_forward(p(self.weights), p(obs), p(masks), E, p(self.scratch), p(actions_in), p(actions_out), p(logp),
                                              p(value if want_value else None), p(logits), self.env_offset, C.c_uint64(self.seed), C.c_uint64(counter),
                                              self.flags, p(self.counter_dev), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "mm_policy_forward")
        # tokens + (fused trunk | 3 trunk layers (+ heads on the SIMT path, + k_heads_finish on the 3xFP16 path)) + critic
        self.launches += (2 if (self.flags & 13) == 13 else 5 if (self.flags & 5) == 5 else 4 if self.flags & 1 else 5) + (1 if want_value else 0)
        return (actions_in if actions_in is not None else actions_out), logp, (value if want_value else None)
