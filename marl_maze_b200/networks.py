"""Actor / Critic with the reference's parameter names and shapes (so `PPO.pth` loads), evaluated batched on the GPU.

Mirrors networks.py of the reference (Actor :13-47, Projection :50-65, m_Attention :67-82, Critic :84-106) in
*interface* -- class names, constructor arguments, state_dict keys -- but not in evaluation strategy:
  * the 23 per-feature projections are applied as ONE [B,65] x [65,460] matmul against a block matrix assembled from
    the 23 small weights (the reference loops over 23 tiny Linear layers);
  * `faithful_projection=True` (default) reproduces the reference's Projection.forward, which never advances its
    column index (networks.py:59-63): projection i reads obs[:, 0:FEATURE_DIMS[i]].  The shipped checkpoint was
    trained with that behaviour, so log-prob parity requires it.  False gives the evidently intended slicing.
The rollout does not call these modules: it goes through the fused forward+sampling kernel (policy.py).  They are
what autograd differentiates in PPO.train().
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

FEATURE_DIMS = [4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2, 2, 1, 4, 1, 1, 1, 1, 1, 1, 2]
FEATURE_AMOUNT = len(FEATURE_DIMS)
OBS_SPACE = sum(FEATURE_DIMS)
EMBEDDING_DIM = 20
_OFFSETS = [sum(FEATURE_DIMS[:i]) for i in range(FEATURE_AMOUNT)]


class Projection(nn.Module):
    """23 Linear(d_i -> 20) stored exactly like the reference (`layers.{i}.weight/bias`), applied as one block matmul."""

    def __init__(self, faithful: bool = True):
        super().__init__()
        self.faithful = faithful
        self.layers = nn.ModuleList(nn.Linear(d, EMBEDDING_DIM) for d in FEATURE_DIMS)

    def block_matrix(self):
        """[460, 65] weight and [460] bias equivalent to the 23 projections (columns chosen by `faithful`)."""
        W = self.layers[0].weight.new_zeros(FEATURE_AMOUNT * EMBEDDING_DIM, OBS_SPACE)
        rows = []
        for i, (lin, d) in enumerate(zip(self.layers, FEATURE_DIMS)):
            c0 = 0 if self.faithful else _OFFSETS[i]
            blk = torch.zeros(EMBEDDING_DIM, OBS_SPACE, dtype=W.dtype, device=W.device)
            blk[:, c0:c0 + d] = lin.weight
            rows.append(blk)
        return torch.cat(rows, 0), torch.cat([lin.bias for lin in self.layers], 0)

    def forward(self, x):
        W, b = self.block_matrix()
        return torch.addmm(b, x, W.t()).view(-1, FEATURE_AMOUNT, EMBEDDING_DIM)


class m_Attention(nn.Module):
    """Single-head self-attention over the 23 feature tokens with a residual connection (networks.py:67-82)."""

    def __init__(self, kq_dim: int = 10):
        super().__init__()
        self.kq_dim = kq_dim
        self.keys = nn.Linear(EMBEDDING_DIM, kq_dim, bias=False)
        self.querys = nn.Linear(EMBEDDING_DIM, kq_dim, bias=False)
        self.values = nn.Linear(EMBEDDING_DIM, EMBEDDING_DIM, bias=False)

    def forward(self, tok):
        scores = torch.bmm(self.querys(tok), self.keys(tok).transpose(1, 2)) / math.sqrt(self.kq_dim)
        return (tok + torch.bmm(torch.softmax(scores, -1), self.values(tok))).reshape(-1, FEATURE_AMOUNT * EMBEDDING_DIM)


_STATIC = {"on": False, "overflow": None}


class static_row_grouping:
    """Inside this context _few_distinct_rows runs a fixed four rounds with no host synchronisation (CUDA-graph capture of the update): a batch
    with more than four distinct rows -- impossible for environment observations, whose obs[:, 0:4] is the facing one-hot -- raises a device-side flag
    (row_grouping_overflow) instead of falling back to torch.unique."""

    def __enter__(self):
        self.prev = _STATIC["on"]; _STATIC["on"] = True

    def __exit__(self, *a):
        _STATIC["on"] = self.prev


def row_grouping_overflow() -> bool:
    f = _STATIC["overflow"]
    if f is None:
        return False
    v = bool(f.item()); f.zero_()
    return v


def _few_distinct_rows(p, max_rows: int = 8):
    """(rows [U, d], inverse [B]) with rows[inverse] == p, like torch.unique(p, dim=0, return_inverse=True) but without its
    lexicographic sort when there are at most `max_rows` distinct rows (environment observations have 4 distinct facing prefixes):
    rounds of "take the first row not yet matched, mark every row equal to it" -- 4 (or 8) streaming passes over [B, d] and one (or two)
    host synchronisations.  Unused slots repeat row 0 (harmless: nothing maps to them); more distinct rows fall back to torch.unique."""
    p = p.contiguous()
    inv = torch.full((p.shape[0],), -1, dtype=torch.int64, device=p.device)
    reps = []
    if _STATIC["on"]:
        for u in range(4):
            rem = inv < 0
            rep = p.index_select(0, torch.argmax(rem.to(torch.uint8)).reshape(1))[0]   # (p[tensor_index] would read the index back on the host)
            inv = torch.where(rem & (p == rep).all(1), u, inv)
            reps.append(rep)
        if _STATIC["overflow"] is None or _STATIC["overflow"].device != p.device:
            _STATIC["overflow"] = torch.zeros((), dtype=torch.bool, device=p.device)
        _STATIC["overflow"].logical_or_((inv < 0).any())
        return torch.stack(reps), inv.clamp_min(0)
    for u in range(max_rows):
        rem = inv < 0
        rep = p[torch.argmax(rem.to(torch.uint8))]          # first unmatched row (row 0 once everything is matched)
        inv = torch.where(rem & (p == rep).all(1), u, inv)
        reps.append(rep)
        if u % 4 == 3 and not bool((inv < 0).any()):        # the one host synchronisation, after 4 (the facings) or 8 rounds
            return torch.stack(reps), inv
    return torch.unique(p, dim=0, return_inverse=True)


class Actor(nn.Module):
    """obs [B,65] -> (move logits [B,5], mark logit [B,1]); hidden_sizes excludes the 460-wide input (networks.py:15)."""

    def __init__(self, hidden_sizes=(164, 164, 164, 164, 164), activation=nn.ReLU, faithful_projection: bool = True):
        super().__init__()
        self.projection = Projection(faithful_projection)
        self.attention = m_Attention()
        widths = [FEATURE_AMOUNT * EMBEDDING_DIM, *hidden_sizes]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        self.activation = activation
        self.move_head = nn.Linear(widths[-1], 5)
        self.mark_head = nn.Linear(widths[-1], 1)
        self.initialize_weights()

    def embed_parts(self, x):
        """Projection + attention as (rows [U, 460], inverse [B] or None): the embedding of observation b is rows[inverse[b]] (rows itself
        when inverse is None).  With the faithful projection every token reads obs[:, 0:4] only, so rows that agree on those four columns
        have identical embeddings: large batches are evaluated once per DISTINCT prefix (same values, same gradients, at a fraction of the
        batched-attention cost)."""
        if self.projection.faithful and x.shape[0] >= 4096:
            uniq, inv = _few_distinct_rows(x[:, :max(FEATURE_DIMS)])
            if uniq.shape[0] * 8 <= x.shape[0]:
                xin = x.new_zeros(uniq.shape[0], OBS_SPACE)
                xin[:, :uniq.shape[1]] = uniq
                return self.attention(self.projection(xin)), inv
        return self.attention(self.projection(x)), None

    def embed(self, x):
        """Projection + attention -> [B, 460] (embed_parts, gathered)."""
        emb, inv = self.embed_parts(x)
        if inv is None:
            return emb
        if emb.shape[0] <= 8 and x.is_cuda and x.dtype == torch.float32:  # backward = one streaming segment-sum kernel (update.GatherRows)
            from .update import GatherRows
            return GatherRows.apply(emb, inv)
        if emb.shape[0] <= 64:  # gather as a one-hot matmul: its backward is a dense [U,B]x[B,460] GEMM instead of a scatter-add
            return torch.nn.functional.one_hot(inv, emb.shape[0]).to(emb.dtype) @ emb  # into a handful of rows (atomics contention)
        return emb.index_select(0, inv)

    def trunk(self, x):
        dev = self.move_head.weight.device
        h = self.embed(torch.as_tensor(x, dtype=torch.float32, device=dev).reshape(-1, OBS_SPACE))
        act = self.activation()
        for lin in self.layers:
            h = act(lin(h))
        return h

    def forward(self, x):
        h = self.trunk(x)
        return [self.move_head(h), self.mark_head(h)]

    def initialize_weights(self):  # networks.py:43-48: orthogonal trunk, heads scaled by 0.01
        for lin in self.layers:
            nn.init.orthogonal_(lin.weight)
        with torch.no_grad():
            self.move_head.weight.mul_(0.01)
            self.mark_head.weight.mul_(0.01)


class Critic(nn.Module):
    """Centralised critic: both agents' observations concatenated, one value per environment (networks.py:84-106)."""

    def __init__(self, agent_amount, hidden_sizes=(128, 128), activation=nn.ReLU):
        super().__init__()
        self.agent_amount = agent_amount
        widths = [agent_amount * OBS_SPACE, *hidden_sizes, 1]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        self.activation = activation
        self.initialize_weights()

    def forward(self, x):
        h = torch.as_tensor(x, dtype=torch.float32, device=self.layers[0].weight.device).reshape(-1, self.agent_amount * OBS_SPACE)
        act = self.activation()
        for lin in self.layers[:-1]:
            h = act(lin(h))
        return self.layers[-1](h)

    def initialize_weights(self):
        for lin in self.layers:
            nn.init.orthogonal_(lin.weight)
