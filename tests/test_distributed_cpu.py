"""world_size-2 gloo tests (CPU) of the multi-rank host logic: gradient averaging and global advantage statistics of
PPO.update, and the env-shard bookkeeping (env_offset -> global maze ids / RNG keys) that keeps results independent of
the number of ranks.  The data path itself (K1-K4) has no collective; it is covered per-rank by the gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from marl_maze_b200.PPO import PPO
        brain = PPO(agent_amount=2, batch_size=100, device="cpu", model_path=None, verbose=False, seed=100 + rank)
        # (1) construction broadcasts rank 0's weights
        ref = [p.detach().clone() for p in brain.actor.parameters()]
        for p in ref:
            q = p.clone(); dist.broadcast(q, 0)
            assert torch.equal(p, q)
        # (2) gradient all-reduce = mean over ranks, identical on both ranks afterwards
        for i, p in enumerate(brain.actor.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        brain._allreduce_grads(brain.actor)
        for i, p in enumerate(brain.actor.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))
        # (3) advantage normalisation uses global mean / unbiased std (PPO.py:47) over both ranks' samples
        g = torch.Generator().manual_seed(5)
        full = torch.randn(2000, generator=g)
        mine = full[rank * 1000:(rank + 1) * 1000].clone()
        want = ((full - full.mean()) / (full.std() + 1e-10))[rank * 1000:(rank + 1) * 1000]
        assert torch.allclose(brain._normalise(mine), want, atol=1e-5)
        torch.save(torch.tensor(1), os.path.join(tmp, f"ok{rank}"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_update_plumbing(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_global_maze_ids_do_not_depend_on_sharding():
    """Maze.refill_pool keys pool slot (env e, episode k) with the GLOBAL id (env_offset+e)*K + k (mm_generate id_mod/id_mul)."""
    K = 4

    def ids(env_offset, E):
        base = env_offset * K
        return {(env_offset + (i % E), i // E): base + (i % E) * K + i // E for i in range(E * K)}
    one = ids(0, 8)
    two = {**ids(0, 4), **ids(4, 4)}
    assert one == two and len(set(one.values())) == 32


def test_finished_episode_statistics_and_row_grouping_cpu():
    """Host logic of the rollout / update that needs no GPU: PPO.finished_episodes against the dense [T,E] formulation, and
    networks._few_distinct_rows against torch.unique."""
    import torch
    from marl_maze_b200.PPO import finished_episodes
    from marl_maze_b200.networks import _few_distinct_rows
    torch.manual_seed(0)
    T, E = 37, 53
    for trial in range(10):
        done = (torch.rand(T, E) < (0.0 if trial == 0 else 0.07)).to(torch.uint8)
        t_idx = torch.arange(1, T + 1, dtype=torch.int32).view(T, 1).expand(T, E)
        d = done.bool()
        last = torch.where(d, t_idx, torch.zeros_like(t_idx))
        prev_end = torch.cat([torch.zeros(1, E, dtype=torch.int32), torch.cummax(last, 0).values[:-1]], 0)
        lens, k, e = finished_episodes(done)
        assert torch.equal(lens, (t_idx - prev_end)[d])
        assert torch.equal(k, (torch.cumsum(d.int(), 0) - 1)[d].long())
        assert torch.equal(e, torch.arange(E).view(1, E).expand(T, E)[d])
    for p in (torch.nn.functional.one_hot(torch.randint(0, 4, (5000,)), 4).float(),        # 4 facings: one round of four
              torch.randint(0, 7, (5000, 1)).float().repeat(1, 4),                          # 7 distinct rows: two rounds
              torch.randn(300, 4)):                                                         # all distinct: torch.unique fallback
        rows, inv = _few_distinct_rows(p)
        assert torch.equal(rows[inv], p) and rows.shape[0] <= max(8, torch.unique(p, dim=0).shape[0])
