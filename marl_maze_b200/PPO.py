"""PPO: the reference's learner surface (PPO.py:11-238) driving the batched CUDA environment.

Same constructor, same methods (train, get_batch, get_action, get_log_probs, get_state_values, get_GAEs, decay_lr,
save_parameters, load_parameters), same checkpoint format -- a reference `PPO.pth` loads and a checkpoint written
here loads in the reference.  What changes is how a batch is produced:

  reference get_batch (PPO.py:89-152)  one maze, python loop, whole episodes until > batch_size steps
  here                                 E mazes in lock-step for a fixed horizon T (T*E > batch_size), every step is
                                       K4 (actor/critic forward + fused sampling) then K2 (fused env step + obs, with
                                       in-launch auto-reset), writing straight into the [T+1,E,...] rollout buffers;
                                       advantages by K3 (reverse scan; episodes still open at T are bootstrapped)

The update (PPO.py:46-85) keeps the reference's schedule -- clipped surrogate on the JOINT ratio of both agents, MSE
critic, grad-norm clip 0.5, two Adams, lr x0.997 per update -- with gradients averaged over ranks (NCCL) when
torch.distributed is initialised and advantage statistics taken over all ranks.  The actor's trunk, heads and loss run
forward AND backward in hand-written kernels (K5, update.py: tcgen05 3xTF32 GEMMs for the forward, data-gradient and
weight-gradient passes), so do the critic's hidden layers and -- when rows do not share their embedding -- the 23-token
embedding (`fused_update=False` puts the whole update back on autograd).  The shuffled rollout is gathered once per update and
the minibatches are views of it (the reference, too, shuffles once and reuses the order in every epoch).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np
import torch

from . import update as _upd
from .engine import gae as _gae
from .networks import Actor, Critic
from .policy import PolicyRunner

MODEL_PATH = "PPO.pth"


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _static_row_grouping():
    from . import networks
    return networks.static_row_grouping()


def _row_grouping_overflow() -> bool:
    from . import networks
    return networks.row_grouping_overflow()


def finished_episodes(done: torch.Tensor):
    """From done [T,E] (uint8 / bool): (length, index of the episode within its env, env) of every finished episode, in (time, env)
    order.  Finished episodes are sparse in [T,E]: one nonzero() (sorted by env, then time), then everything on the short list."""
    T, E = done.shape
    dev = done.device
    e_s, t_s = done.t().nonzero(as_tuple=True)
    n_ep = e_s.numel()
    pos = torch.arange(n_ep, device=dev)
    same = torch.zeros(n_ep, dtype=torch.bool, device=dev)
    same[1:] = e_s[1:] == e_s[:-1]                                    # the previous entry is an earlier episode of the same env
    lens_s = t_s + 1 - torch.where(same, torch.roll(t_s, 1) + 1, torch.zeros_like(t_s))
    k_s = pos - torch.cummax(torch.where(same, torch.zeros_like(pos), pos), 0).values if n_ep else pos
    order = torch.argsort(t_s * E + e_s)                              # report in (time, env) order
    return lens_s[order].to(torch.int32), k_s[order], e_s[order]


class PPO:
    def __init__(self, agent_amount, epochs=500, batch_size=15000, lr=0.0002, discount_rate=0.99, lam=0.95, updates_per_batch=5, clip=0.2, max_grad=0.5,
                 *, device=None, horizon: Optional[int] = None, seed: int = 3234, model_path: Optional[str] = MODEL_PATH, faithful_projection: bool = True,
                 verbose: bool = True, micro_batch: int = 1 << 20, update_tf32: bool = False, use_cuda_graph: bool = True,
                 fused_update: bool = True, prefetch_pool: bool = False, graph_update: Optional[bool] = None):
        if agent_amount != 2:
            raise NotImplementedError("two agents (README.md:34)")
        self.maze = None  # injected by Maze.__init__ (maze.py:40-42)
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        torch.manual_seed(seed)  # PPO.py:7
        self.actor = Actor([264, 264, 264], faithful_projection=faithful_projection).to(self.device)
        self.critic = Critic(agent_amount, hidden_sizes=[64, 64]).to(self.device)
        # torch.optim.Adam exactly as in the reference (PPO.py:20-21).  (fused=True was tried for its launch count and changed the training
        # trajectory from the first update on -- the plain implementation is the one that tracks the autograd reference path.)
        # On a GPU with use_cuda_graph the optimisers are built capturable (step counts and the learning rate live on the device) so that a whole
        # optimiser step of the update can be replayed as a CUDA graph (_update, small minibatches); the update rule is the same Adam, the bias
        # corrections are evaluated in fp32 on the device instead of in float64 on the host.  Checkpoints keep the reference's format (see
        # save_parameters / load_parameters).
        cap = bool(use_cuda_graph and fused_update and self.device.type == "cuda")
        mk = (lambda ps: torch.optim.Adam(ps, lr=torch.tensor(float(lr), device=self.device), capturable=True)) if cap else (lambda ps: torch.optim.Adam(ps, lr=lr))
        self.actor_optim = mk(self.actor.parameters())
        self.critic_optim = mk(self.critic.parameters())
        self._ug = None   # captured update graphs (see _update)
        self.graph_update = graph_update   # None: follow use_cuda_graph; False: eager update steps (tests: same optimiser, no graph)
        self.epochs, self.batch_size, self.lr, self.discount_rate, self.lam = epochs, batch_size, lr, discount_rate, lam
        self.updates_per_batch, self.mbatch_size, self.clip, self.max_grad = updates_per_batch, batch_size // 5, clip, max_grad
        self.horizon, self.seed, self.model_path, self.verbose, self.micro_batch = horizon, seed, model_path, verbose, micro_batch
        self.update_tf32 = update_tf32  # let cuBLAS use TF32 tensor cores in the autograd update (the reference is fp32; off by default)
        self.fused_update = fused_update  # actor trunk + heads + clipped surrogate, fwd and bwd, as hand-written tcgen05 3xTF32 kernels (update.py)
        self.prefetch_pool = prefetch_pool  # build the next rollout's maze pool in the background during the update (off by default: it only moves
                                            # the ~13 ms carve from the rollout to the update, and hurts when rollouts follow each other directly)
        self.batched_values = os.environ.get("MARL_MAZE_BATCHED_VALUES", "1") != "0"   # critic once per rollout over the whole buffer (see get_batch)
        self.use_cuda_graph = use_cuda_graph  # replay the T-step rollout (6 launches per step) as one captured CUDA graph from the 2nd rollout on
        self._buf = None
        self._graph = None
        self._runner: Optional[PolicyRunner] = None
        self._rollouts = 0
        self.last_stats: dict = {}
        d = _dist()
        if d is not None and d.get_world_size() > 1:  # identical initial weights on every rank
            for p in list(self.actor.parameters()) + list(self.critic.parameters()):
                d.broadcast(p.data, 0)
        self.load_parameters()

    # ------------------------------------------------------------------ rollout
    def _policy(self) -> PolicyRunner:
        E = self.maze.num_envs
        if self._runner is None or self._runner.E != E:
            self._runner = PolicyRunner(self.actor, self.critic, E, self.device, env_offset=self.maze.env_offset, seed=self.seed)
        return self._runner

    def get_batch(self):
        """Fixed-horizon vectorised rollout.  Returns the reference's tuple (PPO.py:151-152) flattened t-major:
        b_obs [N,2,65], b_actions [N,2,2] f32, b_log_probs [N], b_shortest_paths, episode_lens, b_masks [N,2,6] bool, b_advs [N], b_vals [N]."""
        maze = self.maze
        E = maze.num_envs
        T = self.horizon or (self.batch_size // E + 1)  # total_timesteps > batch_size (PPO.py:140)
        need = max(4, math.ceil(T / 24))                # pool mazes per env so that an env never meets a maze twice in one rollout
        if maze.engine is not None and maze.pool_episodes < need:
            maze.engine = None
        maze.pool_episodes = max(maze.pool_episodes, need)
        eng = maze._ensure_engine()
        if self._rollouts > 0 or maze._resets_since_fill > 0:
            maze.refill_pool()  # the reference builds a new maze for every episode (maze.py:57)
        dev = self.device
        key = (T, E, id(eng))
        if self._buf is None or self._buf["key"] != key:  # rollout buffers live across rollouts (a captured graph holds their addresses)
            self._buf = dict(key=key,
                             obs=torch.empty(T + 1, E, 2, 65, dtype=torch.float32, device=dev), masks=torch.empty(T + 1, E, 2, 6, dtype=torch.uint8, device=dev),
                             actions=torch.empty(T, E, 2, 2, dtype=torch.uint8, device=dev), logp=torch.empty(T, E, dtype=torch.float32, device=dev),
                             values=torch.empty(T + 1, E, dtype=torch.float32, device=dev), reward=torch.empty(T, E, dtype=torch.float32, device=dev),
                             done=torch.empty(T, E, dtype=torch.uint8, device=dev), adv=torch.empty(T, E, dtype=torch.float32, device=dev))
            self._graph = None
        b = self._buf
        obs, masks, actions, logp, values, reward, done, adv = (b[k] for k in ("obs", "masks", "actions", "logp", "values", "reward", "done", "adv"))
        pol = self._policy()
        pol.refresh()
        maze.reset(obs=obs[0], masks=masks[0])  # PPO.py:104

        def body():  # launches only: no allocation, no host sync -> capturable
            for t in range(T):  # PPO.py:108-141, one iteration = one step of every env
                pol.forward(obs[t], masks[t], actions_out=actions[t], logp=logp[t], value=None if self.batched_values else values[t],
                            want_value=not self.batched_values, counter=t + 1)
                eng.step(actions[t], auto_reset=True, obs=obs[t + 1], masks=masks[t + 1], reward=reward[t], done=done[t])
            if self.batched_values:
                # Nothing in the loop reads V(s_t) (PPO.py:116 only records it for the advantages): the critic runs ONCE over all (T + 1) E observation
                # pairs of the rollout buffer -- V(s_T) included, which bootstraps the episodes still open at the horizon -- instead of as a forked
                # side-stream launch in every step (128 block prologues and 128 fork / join edges of the captured graph less)
                rows_per_call = max(1, (1 << 28) // E)          # the C ABI counts environments in an int
                for t0 in range(0, T + 1, rows_per_call):
                    t1 = min(T + 1, t0 + rows_per_call)
                    pol.values(obs[t0:t1].view((t1 - t0) * E, 2, 65), values[t0:t1].view((t1 - t0) * E))
            else:
                pol.values(obs[T], values[T])  # V(s_T) bootstraps the episodes still open at the horizon
            _gae(reward, values[:T], done, values[T], self.discount_rate, self.lam, out=adv)
            pol.bump(T)                    # next rollout / replay continues the sampling stream

        if self.use_cuda_graph and T <= 2048 and self._rollouts >= 1:
            if self._graph is None:  # the first rollout ran eagerly (lazy kernel attributes, tensor-map cache are warm): capture now
                torch.cuda.synchronize(dev)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    body()
            self._graph.replay()
        else:
            body()
        self._rollouts += 1
        maze._obs, maze._masks = obs[T], masks[T]

        # episode statistics for the progress prints (PPO.py:36-43): lengths of finished episodes and their shortest paths
        episode_lens, k_ep, e_ep = finished_episodes(done)
        n_ep = episode_lens.numel()
        spl_all = (eng.pool_hdr.view(torch.int32).view(-1, 4)[:, 2] >> 16) & 0xFFFF
        b_shortest = spl_all[(e_ep + k_ep * E) % eng.P]
        solved, keys, rsum = torch.stack([(reward == 1).sum(), (reward == 0.5).sum(), reward.sum()]).tolist()
        self.last_stats = dict(env_steps=T * E, episodes=n_ep, solved=int(solved), keys=int(keys), mean_reward_per_step=rsum / (T * E), horizon=T, num_envs=E)
        if self.prefetch_pool:  # the mazes this rollout consumed are rebuilt on a side stream, in place, while the update runs (after the statistics
            maze.prefetch_pool()   # above: they read the consumed slots' headers)
        N = T * E
        return (obs[:T].reshape(N, 2, 65), actions.reshape(N, 2, 2).float(), logp.reshape(N), b_shortest.cpu().numpy(), episode_lens.cpu().numpy(),
                masks[:T].reshape(N, 2, 6).bool(), adv.reshape(N), values[:T].reshape(N))

    def get_action(self, obs, action_mask):
        """One agent's action from ONE observation (PPO.py:170-186) -- the reference's single-env interface, kept for
        Agent.get_action / viewers.  The batched rollout does not come through here (it uses the fused kernel)."""
        if self.device.type == "cuda":
            return self._get_action_kernel(obs, action_mask)
        with torch.no_grad():
            move_logits, mark_logits = self.actor(obs)
            return self._sample_action(move_logits, mark_logits, action_mask)

    @staticmethod
    def _sample_action(move_logits, mark_logits, action_mask):
        """The reference's draw (PPO.py:175-186) from the six logits of one observation: masked categorical move, Bernoulli mark (probability 0 when the
        mark action is masked), joint log-prob.  Logits [1,5] / [1,1] on any device; the generator of that device decides.  Written with the calls
        torch.distributions.Categorical makes underneath (normalised logits, softmax, multinomial): same random draws, bit-identical log-probs, half the
        host time (70 vs 138 us)."""
        dev = move_logits.device
        m = [bool(v) for v in (action_mask.reshape(-1).tolist() if torch.is_tensor(action_mask) else list(np.asarray(action_mask).reshape(-1)))]
        ml = move_logits.masked_fill(~torch.tensor([m[:5]], dtype=torch.bool, device=dev), float("-inf"))
        nl = ml - ml.logsumexp(-1, keepdim=True)
        move = torch.multinomial(torch.softmax(nl, -1), 1, True)
        p = torch.sigmoid(mark_logits) if m[5] else torch.zeros(1, 1, device=dev)
        mark = torch.bernoulli(p)
        p = p if mark == 1 else 1 - p
        log_prob = nl.gather(-1, move).view(1) + torch.log(p)
        return [int(move.item()), float(mark.item())], log_prob

    def _act1_state(self):
        a1 = getattr(self, "_act1", None)
        if a1 is None:
            run = PolicyRunner(self.actor, self.critic, 1, self.device, seed=self.seed)
            a1 = self._act1 = dict(run=run, key=None, cache=None, h_obs=torch.zeros(1, 2, 65, pin_memory=True), d_obs=torch.zeros(1, 2, 65, device=self.device),
                                   d_masks=torch.ones(1, 2, 6, dtype=torch.uint8, device=self.device), d_logits=torch.zeros(1, 2, 6, device=self.device),
                                   h_logits=torch.zeros(1, 2, 6, pin_memory=True), d_act=torch.zeros(1, 2, 2, dtype=torch.uint8, device=self.device),
                                   d_logp=torch.zeros(1, device=self.device))
            a1["h_obs_np"] = a1["h_obs"].numpy()
        key = (getattr(self, "_weights_version", 0),) + tuple(p._version for p in self.actor.parameters())
        if a1["key"] != key:     # optimiser steps (eager: tensor versions; graph replays: _weights_version), load_state_dict, load_parameters
            a1["run"].refresh(); a1["key"] = key; a1["cache"] = None
        return a1

    def _prefetch_logits(self, d_obs):
        """Maze.step / Maze.reset at ONE env, once get_action has been seen in the loop: the logits of both agents' NEW observations are computed right
        behind the step kernel and ride to the host under the step's own synchronisation; the two get_action calls that follow find them
        (_commit_logits) and touch neither the GPU nor the stream.  One synchronisation per environment step instead of three."""
        a1 = self._act1_state()
        a1["run"].forward(d_obs, a1["d_masks"], actions_out=a1["d_act"], logp=a1["d_logp"], logits=a1["d_logits"], want_value=False)
        a1["h_logits"].copy_(a1["d_logits"], non_blocking=True)

    def _commit_logits(self, obs_rows):
        a1 = self._act1
        a1["cache"] = (obs_rows, a1["h_logits"][0].clone())   # after the caller's stream synchronisation

    def _get_action_kernel(self, obs, action_mask):
        """get_action on the GPU without the ~60 small launches of the autograd modules: the six logits of the one observation come from the K4 kernels
        (mm_policy_forward at one env, this agent's row; weights re-packed only when the actor changed), travel to a pinned host buffer, and the
        draw is the reference's own sequence of torch calls (PPO.py:175-184) on the host -- as in the reference, which runs on the CPU, the
        global torch CPU generator decides the action.  In the single-env loop (maze.py:477-493) the logits were already fetched by the step that
        produced `obs` (_prefetch_logits)."""
        a1 = self._act1_state()
        lg = None
        c = a1["cache"]
        if c is not None and isinstance(obs, list):
            for i in (0, 1):
                if obs == c[0][i]:
                    lg = c[1][i]
                    break
        if lg is None:
            a1["h_obs_np"][0, 0, :] = obs.detach().cpu().numpy().reshape(-1) if torch.is_tensor(obs) else obs
            a1["d_obs"].copy_(a1["h_obs"], non_blocking=True)
            a1["run"].forward(a1["d_obs"], a1["d_masks"], actions_out=a1["d_act"], logp=a1["d_logp"], logits=a1["d_logits"], want_value=False)
            a1["h_logits"].copy_(a1["d_logits"], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            lg = a1["h_logits"][0, 0].clone()
            if self.maze is not None and self.maze.num_envs == 1:
                self.maze._prefetch_policy = True   # a policy loop at one env: from now on the step fetches the logits with the observation
        a1["last_logits"] = lg
        with torch.no_grad():
            return self._sample_action(lg[:5].view(1, 5), lg[5:6].view(1, 1), action_mask)

    # ------------------------------------------------------------------ update-side helpers (autograd)
    def get_log_probs(self, i, batch_obs, batch_actions, batch_masks):
        """Log-prob of agent i's recorded actions under the current actor (PPO.py:154-168)."""
        moves, marks = batch_actions[:, i, 0], batch_actions[:, i, 1]
        move_logits, mark_logits = self.actor(batch_obs[:, i, :])
        move_logits = move_logits.masked_fill(~batch_masks[:, i, 0:5], float("-inf"))
        lp_move = torch.distributions.Categorical(logits=move_logits).log_prob(moves)
        p = torch.sigmoid(mark_logits.reshape(-1).masked_fill(~batch_masks[:, i, 5], float("-inf")))
        p = torch.where(marks.bool(), p, 1 - p)
        return lp_move + torch.log(p)

    def joint_log_probs(self, batch_obs, batch_actions, batch_masks):
        """sum_i get_log_probs(i, ...) (PPO.py:66-68) with BOTH agents pushed through the shared actor in one batch of 2B rows."""
        B = batch_obs.shape[0]
        move_logits, mark_logits = self.actor(batch_obs.reshape(2 * B, -1))
        masks = batch_masks.reshape(2 * B, 6)
        acts = batch_actions.reshape(2 * B, 2)
        lp_move = torch.log_softmax(move_logits.masked_fill(~masks[:, 0:5], float("-inf")), dim=-1).gather(1, acts[:, 0:1].long()).squeeze(1)
        p = torch.sigmoid(mark_logits.reshape(-1).masked_fill(~masks[:, 5], float("-inf")))
        p = torch.where(acts[:, 1].bool(), p, 1 - p)
        return (lp_move + torch.log(p)).view(B, 2).sum(1)

    def get_state_values(self, batch_obs):
        return self.critic(batch_obs).squeeze(-1)

    def get_GAEs(self, ep_rew, ep_values, ep_dones):
        """One complete episode (PPO.py:193-203) through the reverse-scan kernel."""
        L = len(ep_rew)
        r = torch.as_tensor(np.asarray(ep_rew, np.float32), device=self.device).view(L, 1).contiguous()
        v = torch.as_tensor(np.asarray([float(x) for x in ep_values], np.float32), device=self.device).view(L, 1).contiguous()
        d = torch.as_tensor(np.asarray(ep_dones, np.uint8), device=self.device).view(L, 1).contiguous()
        return _gae(r, v, d, None, self.discount_rate, self.lam).view(L).cpu().numpy().astype(np.float64)

    def decay_lr(self):  # PPO.py:216-220
        for opt in (self.actor_optim, self.critic_optim):
            for g in opt.param_groups:
                g["lr"] *= 0.997

    # ------------------------------------------------------------------ training
    def _bucket(self, module):
        """The module's persistent flat gradient bucket: one fp32 buffer over all its parameters plus a view per parameter."""
        key = id(module)
        if not hasattr(self, "_buckets"):
            self._buckets = {}
        bk = self._buckets.get(key)
        params = list(module.parameters())
        if bk is None or bk["n"] != sum(p.numel() for p in params) or bk["flat"].device != params[0].device:
            flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
            views, o = [], 0
            for p in params:
                views.append(flat[o:o + p.numel()].view_as(p)); o += p.numel()
            bk = self._buckets[key] = dict(flat=flat, views=views, n=flat.numel())
        return bk, params

    def _pack(self, module):
        """Copy the module's gradients into its persistent flat bucket with ONE multi-tensor copy; returns the (parameter, bucket view) pairs."""
        bk, params = self._bucket(module)
        live = [(p, v) for p, v in zip(params, bk["views"]) if p.grad is not None]
        torch._foreach_copy_([v for _, v in live], [p.grad for p, _ in live])
        return dict(live=live, flat=bk["flat"])

    def _reduce_async(self, h):
        """Mean all-reduce of a packed bucket as ONE asynchronous collective (NCCL: on its own stream, averaging inside the collective)."""
        d = _dist()
        avg = d.get_backend() == "nccl"   # gloo has no AVG
        h.update(work=d.all_reduce(h["flat"], op=d.ReduceOp.AVG if avg else d.ReduceOp.SUM, async_op=True), avg=avg, world=d.get_world_size())
        return h

    def _allreduce_start(self, module):
        """Start the mean all-reduce of the module's gradients (PPO.py:76-78,82-84 see single-process gradients: the mean over ranks of per-rank minibatch
        means): pack + one asynchronous collective -- whatever the caller launches next (the critic's forward / backward after the actor's bucket)
        overlaps it.  `_allreduce_finish` makes the current stream wait and points every p.grad at its slice of the bucket: no copy back."""
        d = _dist()
        if d is None or d.get_world_size() == 1:
            return None
        return self._reduce_async(self._pack(module))

    @staticmethod
    def _allreduce_finish(h):
        if h is None:
            return
        h["work"].wait()
        if not h["avg"]:
            h["flat"] /= h["world"]
        for p, v in h["live"]:
            p.grad = v

    def _allreduce_grads(self, module):
        """Blocking form (tests, callers outside update)."""
        self._allreduce_finish(self._allreduce_start(module))

    def _normalise(self, adv):
        """(adv - mean) / (unbiased std + 1e-10), PPO.py:47, statistics over every rank's samples."""
        d = _dist()
        if d is None or d.get_world_size() == 1:
            return (adv - adv.mean()) / (adv.std() + 1e-10)
        s = torch.stack([torch.tensor(float(adv.numel()), device=adv.device, dtype=torch.float64), adv.double().sum(), (adv.double() ** 2).sum()])
        d.all_reduce(s)
        n, mean = s[0], s[1] / s[0]
        std = torch.sqrt(torch.clamp(s[2] - n * mean * mean, min=0) / (n - 1))
        return ((adv - mean.float()) / (std.float() + 1e-10))

    def update(self, batch):
        """The 5 x 5 minibatch schedule of PPO.train (PPO.py:46-85) on one rollout."""
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.update_tf32)
        try:
            return self._update(batch)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32
            self._weights_version = getattr(self, "_weights_version", 0) + 1   # graph replays change the parameters without touching tensor versions

    # minibatches of at most this many env-steps replay their optimiser step as a CUDA graph (launch-bound regime: at the reference's own
    # configuration an optimiser step is ~150 small launches, 7.5 ms eager against < 1 ms of GPU work); larger ones are compute-bound
    GRAPH_UPDATE_MAX_MINIBATCH = 1 << 18

    def _update(self, batch):
        b_obs, b_actions, b_log_probs, _, _, b_masks, b_advs, b_vals = batch
        N = b_obs.shape[0]
        b_rtgs = b_advs + b_vals.detach()
        b_advs = self._normalise(b_advs)
        g = torch.Generator(device=self.device); g.manual_seed(self.seed + 7919 * self._rollouts)
        index_list = torch.randperm(N, device=self.device, generator=g)
        used = min(self.batch_size, N)
        mb = max(1, self.mbatch_size if self.batch_size <= N else N // 5)
        fused = self.fused_update and _upd.fused_available(self.actor)
        fused_critic = self.fused_update and _upd.critic_fused_available(self.critic)
        graphed = bool(self.use_cuda_graph and self.graph_update is not False and fused and fused_critic and self.device.type == "cuda" and mb <= self.GRAPH_UPDATE_MAX_MINIBATCH
                       and all(gr.get("capturable", False) for opt in (self.actor_optim, self.critic_optim) for gr in opt.param_groups))
        # The reference shuffles ONCE per train iteration and walks the same minibatches in every update epoch (PPO.py:48-55): gather the
        # used part of the rollout into that order once, and every minibatch / micro-batch below is a contiguous view.
        sel = index_list[:used]
        srcs = (b_obs, b_actions, b_masks, b_log_probs, b_advs, b_rtgs)
        if graphed:  # static buffers: the captured graphs hold their addresses across update() calls
            key = (used, mb, self.micro_batch, tuple(t.dtype for t in srcs), id(self.actor_optim), id(self.critic_optim))
            if self._ug is None or self._ug["key"] != key:
                self._ug = dict(key=key, bufs=[torch.empty((used,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device) for t in srcs],
                                xpad=torch.empty(used, _upd.CRITIC_IN + 2, dtype=torch.float32, device=self.device),
                                a_sum=torch.zeros((), device=self.device), c_sum=torch.zeros((), device=self.device), graphs={}, pool=None, warm=False)
            ug = self._ug
            _upd.token_layout(self.device, self.actor.projection.faithful)   # host -> device copies happen here, never inside a capture
            for t, dst in zip(srcs, ug["bufs"]):
                torch.index_select(t, 0, sel, out=dst)
            p_obs, p_act, p_masks, p_logp, p_adv, p_rtg = ug["bufs"]
            ug["xpad"].copy_(_upd.pad_critic_obs(p_obs))
            p_xpad, a_sum, c_sum = ug["xpad"], ug["a_sum"], ug["c_sum"]
            a_sum.zero_(); c_sum.zero_()
        else:
            p_obs, p_act, p_masks, p_logp, p_adv, p_rtg = (t.index_select(0, sel) for t in srcs)
            a_sum = torch.zeros((), device=self.device); c_sum = torch.zeros((), device=self.device)   # summed on the device: no sync per micro-batch
            p_xpad = _upd.pad_critic_obs(p_obs) if fused_critic else None
        stats = dict(actor_loss=0.0, critic_loss=0.0, steps=0, graphed=graphed)

        def seg_actor(start, n):
            """Actor forward / backward on minibatch [start, start + n): launches only (no host synchronisation), so it can be captured."""
            self.actor_optim.zero_grad(set_to_none=True)
            for s0 in range(start, start + n, self.micro_batch):  # gradient accumulation bounds activation memory; the sum equals the minibatch mean
                s1 = min(s0 + self.micro_batch, start + n)
                m_obs, m_act, m_masks, adv = p_obs[s0:s1], p_act[s0:s1], p_masks[s0:s1], p_adv[s0:s1]
                if fused:  # K5: trunk + heads + clipped surrogate, forward and backward, in hand-written kernels (update.py)
                    loss, _ = _upd.actor_loss(self.actor, m_obs.reshape(-1, m_obs.shape[-1]), m_masks.reshape(-1, 6), m_act.reshape(-1, 2),
                                              p_logp[s0:s1], adv, self.clip, 1.0 / n)
                else:
                    cur = self.joint_log_probs(m_obs, m_act, m_masks)
                    ratio = torch.exp(cur - p_logp[s0:s1])
                    loss = -(torch.min(ratio * adv, torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * adv)).sum() / n
                loss.backward()
                a_sum.add_(loss.detach())

        def seg_critic(start, n):
            self.critic_optim.zero_grad(set_to_none=True)
            for s0 in range(start, start + n, self.micro_batch):
                s1 = min(s0 + self.micro_batch, start + n)
                if fused_critic:  # K5 GEMM kernels for the two hidden layers, forward and backward (update._CriticLoss)
                    loss = _upd.critic_loss(self.critic, p_xpad[s0:s1], p_rtg[s0:s1], 1.0 / n)
                else:
                    loss = ((self.get_state_values(p_obs[s0:s1]) - p_rtg[s0:s1]) ** 2).sum() / n
                loss.backward()
                c_sum.add_(loss.detach())

        def seg_opt():
            torch.nn.utils.clip_grad_norm_(self.actor.parameters(), self.max_grad)
            self.actor_optim.step()
            torch.nn.utils.clip_grad_norm_(self.critic.parameters(), self.max_grad)
            self.critic_optim.step()

        multi = _dist() is not None and _dist().get_world_size() > 1

        def step(start, n):
            """One optimiser step, eagerly.  The two networks share nothing: the critic's forward / backward runs while the actor's gradient bucket is
            being reduced, and both optimisers step once both buckets are back (same result as the reference's actor-then-critic order)."""
            seg_actor(start, n)
            h_a = self._allreduce_start(self.actor)
            seg_critic(start, n)
            h_c = self._allreduce_start(self.critic)
            self._allreduce_finish(h_a); self._allreduce_finish(h_c)
            seg_opt()

        def capture(fn):
            torch.cuda.synchronize(self.device)
            gr = torch.cuda.CUDAGraph()
            # thread_local: NCCL's watchdog thread polls its events while we capture; only this thread's calls must be capture-safe
            with _static_row_grouping(), torch.cuda.graph(gr, pool=ug["pool"], capture_error_mode="thread_local" if multi else "global"):
                out = fn()
            if ug["pool"] is None:
                ug["pool"] = gr.pool()
            return gr, out

        def graph_step(start, n):
            """The same step as CUDA graphs, one set per minibatch position (the same positions in every epoch and every update).  One rank: a single
            graph.  Several ranks: the collectives stay OUTSIDE the graphs (graph: actor fwd / bwd + bucket pack | NCCL | graph: critic + pack | NCCL |
            graph: clip + Adam on the bucket views) -- capturing the NCCL calls themselves hung the 2-GPU run."""
            gk = (start, n)
            if not multi:
                if gk not in ug["graphs"]:
                    ug["graphs"][gk] = capture(lambda: step(start, n))[0]
                ug["graphs"][gk].replay()
                return
            if gk not in ug["graphs"]:
                g_a, h_a = capture(lambda: (seg_actor(start, n), self._pack(self.actor))[1])
                g_a.replay(); self._reduce_async(h_a)
                g_c, h_c = capture(lambda: (seg_critic(start, n), self._pack(self.critic))[1])
                g_c.replay(); self._reduce_async(h_c)
                self._allreduce_finish(h_a); self._allreduce_finish(h_c)   # p.grad -> bucket views: what the optimiser graph reads
                g_o, _ = capture(seg_opt)
                g_o.replay()
                ug["graphs"][gk] = (g_a, h_a, g_c, h_c, g_o)
                return
            g_a, h_a, g_c, h_c, g_o = ug["graphs"][gk]
            g_a.replay(); self._reduce_async(h_a)
            g_c.replay(); self._reduce_async(h_c)
            h_a["work"].wait(); h_c["work"].wait()
            if not h_a["avg"]:
                h_a["flat"] /= h_a["world"]; h_c["flat"] /= h_c["world"]
            g_o.replay()

        for _ in range(self.updates_per_batch):
            self.decay_lr()
            for start in range(0, used, mb):
                n = min(mb, used - start)
                if not graphed:
                    step(start, n)
                elif not ug["warm"]:   # the very first step runs eagerly: lazy kernel attributes, Adam's state tensors, NCCL's communicator
                    with _static_row_grouping():
                        step(start, n)
                    ug["warm"] = True
                else:
                    graph_step(start, n)
                stats["steps"] += 1
        stats["actor_loss"], stats["critic_loss"] = float(a_sum), float(c_sum)
        if graphed and _row_grouping_overflow():
            raise RuntimeError("update graph: an observation batch had more than 4 distinct obs[:, 0:4] prefixes (not environment observations); "
                               "run with use_cuda_graph=False")
        return stats

    def train(self):
        d = _dist()
        rank0 = d is None or d.get_rank() == 0
        for epoch in range(self.epochs):
            batch = self.get_batch()
            episode_lens, b_shortest = batch[4], batch[3]
            if self.verbose and rank0:
                print(f"-------------------- Epoch #{epoch} --------------------")
                print(f"Mazes solved in current epoch: {self.last_stats['solved']} of {len(episode_lens)} finished episodes "
                      f"({self.last_stats['env_steps']} env-steps on {self.last_stats['num_envs']} mazes)")
                if len(episode_lens):
                    print(f"Average Exit Time: {np.mean(episode_lens):.1f}  Best: {np.min(episode_lens)}  Worst: {np.max(episode_lens)}")
                    print(f"Average Length of Shortest Path: {np.mean(b_shortest):.1f}")
                print("--------------------------------------------------", flush=True)
            self.last_update = self.update(batch)
            if rank0:
                self.save_parameters()

    # ------------------------------------------------------------------ checkpoint I/O (PPO.py:222-238, same format)
    def save_parameters(self):
        """Same four keys as the reference (PPO.py:222-227).  Every tensor is saved as a CPU tensor: the reference loads with a bare
        torch.load(MODEL_PATH) (PPO.py:231), which on the CPU-only machine it normally runs on cannot deserialise CUDA storages."""
        if not self.model_path:
            return

        def cpu(o):
            if torch.is_tensor(o):
                return o.detach().cpu()
            if isinstance(o, dict):
                return {k: cpu(v) for k, v in o.items()}
            if isinstance(o, (list, tuple)):
                return type(o)(cpu(v) for v in o)
            return o
        def plain(opt):  # the reference's optimisers are plain torch.optim.Adam (PPO.py:20-21): float learning rate, not capturable, step counts on the CPU
            sd = cpu(opt.state_dict())
            for gr in sd["param_groups"]:
                gr["lr"] = float(gr["lr"]); gr["capturable"] = False
            return sd
        torch.save({"actor": cpu(self.actor.state_dict()), "critic": cpu(self.critic.state_dict()),
                    "actor_optim": plain(self.actor_optim), "critic_optim": plain(self.critic_optim)}, self.model_path)

    def load_parameters(self):
        if self.model_path and os.path.exists(self.model_path):
            sd = torch.load(self.model_path, map_location=self.device)
            self.actor.load_state_dict(sd["actor"]); self.critic.load_state_dict(sd["critic"])
            for opt, osd in ((self.actor_optim, sd["actor_optim"]), (self.critic_optim, sd["critic_optim"])):
                cap = bool(opt.param_groups[0].get("capturable", False))   # ours may be capturable (device-side step counts and learning rate)
                for gr in osd["param_groups"]:
                    gr["capturable"] = cap
                opt.load_state_dict(osd)
                if cap:
                    for gr in opt.param_groups:
                        gr["lr"] = torch.as_tensor(float(gr["lr"]), dtype=torch.float32, device=self.device)
            if self.verbose:
                print("successfuly loaded existing parameters")
            return True
        return False
