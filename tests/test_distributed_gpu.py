"""Two ranks, two GPUs, NCCL: the one place a collective is on the path (SURVEY 8e -- the PPO gradient all-reduce).

SURVEY section 4 item 7: after the all-reduce, every rank's gradients equal the 1-GPU gradients of the concatenated batch.  Each rank pushes ITS half of
a fixed synthetic minibatch through the fused K5 actor / critic losses (scaled by 1 / n_local, as PPO.update does), PPO._allreduce_start / _finish
average the buckets over NCCL, and the result is compared with what one process computes on the whole minibatch -- the mean of two half-batch means is
the full-batch mean; what differs is the order of the sums (the weight-gradient kernel splits its rows into slabs by the row count, and each tcgen05
accumulation carries ~1e-6 of its own: tests/test_update_gpu.py allows 2e-6 + 3e-10 R against fp64).  Measured on 2 x B200: 1.5e-6 of the largest
element of a gradient tensor; the bar is 5e-6.  Skipped on a box with fewer than two GPUs.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(E, device):
    g = torch.Generator(device=device); g.manual_seed(1234)
    obs = torch.rand(E, 2, 65, device=device, generator=g)
    obs[:, :, :4] = torch.nn.functional.one_hot(torch.randint(0, 4, (E, 2), device=device, generator=g), 4).float()
    masks = torch.rand(E, 2, 6, device=device, generator=g) < 0.6
    masks[..., 0] |= ~masks[..., :5].any(-1)                                                     # at least one legal move
    moves = torch.multinomial(masks[..., :5].reshape(-1, 5).float(), 1, generator=g).view(E, 2)  # a legal move
    marks = (torch.rand(E, 2, device=device, generator=g) < 0.5) & masks[..., 5]
    actions = torch.stack([moves, marks.long()], -1).to(torch.uint8)
    adv = torch.randn(E, device=device, generator=g)
    old = -2.0 + 0.3 * torch.randn(E, device=device, generator=g)                                # ratios on both sides of the clip range
    rtg = torch.randn(E, device=device, generator=g)
    return obs, masks, actions, adv, old, rtg


def _grads(brain, obs, masks, actions, adv, old, rtg):
    """The per-rank part of one optimiser step of PPO.update: fused actor loss + fused critic loss on this rank's rows, mean over them."""
    from marl_maze_b200 import update as U
    n = obs.shape[0]
    brain.actor_optim.zero_grad(set_to_none=True); brain.critic_optim.zero_grad(set_to_none=True)
    loss, _ = U.actor_loss(brain.actor, obs.reshape(-1, 65), masks.reshape(-1, 6), actions.reshape(-1, 2), old, adv, brain.clip, 1.0 / n)
    loss.backward()
    U.critic_loss(brain.critic, U.pad_critic_obs(obs), rtg, 1.0 / n).backward()


def _worker(rank, world, port, tmp, faithful):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from marl_maze_b200.PPO import PPO
        # different initial weights per rank (seed): the constructor must broadcast rank 0's
        brain = PPO(agent_amount=2, batch_size=100, device=str(dev), model_path=None, verbose=False, faithful_projection=faithful, seed=100 + rank)
        E = 8192
        full = _batch(E, dev)
        # ---- sharded: each rank its half, then the bucketed asynchronous all-reduce (actor bucket in flight while the critic's is packed)
        mine = [t[rank * (E // world):(rank + 1) * (E // world)].contiguous() for t in full]
        _grads(brain, *mine)
        h_a = brain._allreduce_start(brain.actor)
        h_c = brain._allreduce_start(brain.critic)
        brain._allreduce_finish(h_a); brain._allreduce_finish(h_c)
        sharded = [p.grad.detach().clone() for p in list(brain.actor.parameters()) + list(brain.critic.parameters())]
        # ---- one process on the concatenated batch (every rank computes it: the weights are identical after the broadcast)
        _grads(brain, *full)
        whole = [p.grad.detach().clone() for p in list(brain.actor.parameters()) + list(brain.critic.parameters())]
        worst = 0.0
        for a, b in zip(sharded, whole):
            worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)))
        # every rank holds the same reduced gradients
        flat = torch.cat([g.reshape(-1) for g in sharded])
        other = flat.clone(); dist.broadcast(other, 0)
        assert torch.equal(flat, other), "ranks disagree after the all-reduce"
        assert worst < 5e-6, worst
        torch.save(torch.tensor(worst), os.path.join(tmp, f"ok{rank}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("faithful", [True, False], ids=["column0_projection", "indexed_projection"])
def test_two_rank_nccl_gradients_equal_single_process(tmp_path, faithful):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), faithful), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
