"""K3 (mm_gae) against the numpy oracle restatement of PPO.get_GAEs and the recorded reference outputs: bit-exact."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN
from oracle import ppo_oracle as po

pytestmark = pytest.mark.gpu


def test_gae_kernel_reference_episodes_bit_exact():
    from marl_maze_b200 import gae
    Z = np.load(os.path.join(GOLDEN, "ppo_kats.npz"))
    for k in range(int(Z["gae/n"])):
        rew, val, adv = Z[f"gae/{k}/rew"], Z[f"gae/{k}/val"], Z[f"gae/{k}/adv"]
        L = len(rew)
        done = np.zeros(L, np.uint8); done[-1] = 1
        got = gae(torch.from_numpy(rew).cuda().view(L, 1).contiguous(), torch.from_numpy(val).cuda().view(L, 1).contiguous(),
                  torch.from_numpy(done).cuda().view(L, 1).contiguous(), None)
        assert np.array_equal(got.cpu().numpy()[:, 0].view(np.uint32), adv.view(np.uint32)), k


@pytest.mark.parametrize("T,E", [(128, 4096), (1, 7), (37, 1001)])
def test_gae_kernel_fixed_horizon_vs_oracle(T, E):
    from marl_maze_b200 import gae
    rng = np.random.default_rng(T * 1000 + E)
    rew = rng.choice([0.0, 0.5, 1.0], size=(T, E), p=[.9, .05, .05]).astype(np.float32)
    val = rng.standard_normal((T, E)).astype(np.float32)
    done = (rng.random((T, E)) < 0.05).astype(np.uint8)
    vb = rng.standard_normal(E).astype(np.float32)
    adv, rtg = gae(torch.from_numpy(rew).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(done).cuda(), torch.from_numpy(vb).cuda(), with_rtg=True)
    want = po.gae_fixed_horizon(rew, val, done.astype(bool), vb)
    assert np.array_equal(adv.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(rtg.cpu().numpy(), (want + val).astype(np.float32))
