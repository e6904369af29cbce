"""Synthetic workload generators.  Host-side, CPU tensors, deterministic per index.

The reference ships no datasets (`datasets` is a 23-byte placeholder), so tests and bench.py use these
shapes: ETH/UCY-like scenes (`obs_traj [N,2,T_p]`, `pred_traj_gt [N,2,T_f]` as utils/dataloader.py yields
them) and NBA-like batches (`{'past_traj': [B,11,T_p,2], 'future_traj': [B,11,T_f,2]}` as
data/dataloader_nba.py yields them).
"""
import torch

ETH_AGENT_CYCLE = (2, 3, 5, 7, 10)
ZARA_AGENT_CYCLE = (10, 20, 30, 40, 50, 60)


def eth_scene(index, num_agents, past_length=8, future_length=12, extent=10.0, speed=0.4):
    """One scene: positions = p0 + cumsum(v), p0 ~ U(-extent, extent)^2, v ~ N(0, speed^2) per frame."""
    g = torch.Generator().manual_seed(1000 + index)
    T = past_length + future_length
    p0 = (torch.rand(num_agents, 1, 2, generator=g) * 2 - 1) * extent
    v = torch.randn(num_agents, T, 2, generator=g) * speed
    traj = p0 + torch.cumsum(v, dim=1)                                  # [N,T,2]
    obs = traj[:, :past_length].permute(0, 2, 1).contiguous()            # [N,2,T_p]
    gt = traj[:, past_length:].permute(0, 2, 1).contiguous()             # [N,2,T_f]
    return obs, gt


def eth_scenes(count, cycle=ETH_AGENT_CYCLE, past_length=8, future_length=12, start=0):
    return [eth_scene(start + i, cycle[(start + i) % len(cycle)], past_length, future_length) for i in range(count)]


def nba_batch(index, batch_size, num_agents=11, past_length=10, future_length=20, speed=0.3):
    """One reference batch of NBA-like scenes: p0 ~ U([0,28]x[0,15]) metres, v ~ N(0, speed^2)."""
    g = torch.Generator().manual_seed(5000 + index)
    T = past_length + future_length
    p0 = torch.rand(batch_size, num_agents, 1, 2, generator=g) * torch.tensor([28.0, 15.0])
    v = torch.randn(batch_size, num_agents, T, 2, generator=g) * speed
    traj = p0 + torch.cumsum(v, dim=2)
    return {"past_traj": traj[:, :, :past_length].contiguous(), "future_traj": traj[:, :, past_length:].contiguous()}


def latent(rows, sample_k=20, zdim=32, seed=7):
    """The injected prior sample z ~ N(0, I) [rows*K, zdim] (model/STTODE.py:609-617 draws it with randn_like)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows * sample_k, zdim, generator=g)
