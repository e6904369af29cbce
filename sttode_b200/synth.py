"""Synthetic workload generators.  Host-side, CPU tensors, deterministic per index.

The reference ships no datasets (`datasets` is a 23-byte placeholder), so tests and bench.py use these
shapes: ETH/UCY-like scenes (`obs_traj [N,2,T_p]`, `pred_traj_gt [N,2,T_f]` as utils/dataloader.py yields
them) and NBA-like batches (`{'past_traj': [B,11,T_p,2], 'future_traj': [B,11,T_f,2]}` as
data/dataloader_nba.py yields them).
"""
import argparse
import math
from collections import OrderedDict

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


# ---------------------------------------------------------------------------------------------------
# synthetic weights with the reference's state_dict keys and shapes (no checkpoints ship with the reference)
# ---------------------------------------------------------------------------------------------------
D_MODEL, N_HEAD, FF_DIM, ZDIM, GRU_H, CONV_C = 64, 8, 1024, 32, 96, 32


def state_dict_spec(past_length=8, future_length=12):
    """(key, shape, kind) for all 90 parameters + 2 buffers of STTODENet (model/STTODE.py:349-366)."""
    spec = []

    def lin(prefix, out_f, in_f):
        spec.append((prefix + ".weight", (out_f, in_f), "w"))
        spec.append((prefix + ".bias", (out_f,), "b"))

    def ode(prefix):
        p = prefix + "ODE_Encoder.odeblock.odefunc.layers.0."
        a = p + "self_attn.temporal_attention_before."
        spec.append((a + "in_proj_weight", (3 * D_MODEL, D_MODEL), "w"))
        spec.append((a + "in_proj_bias", (3 * D_MODEL,), "b"))
        lin(a + "out_proj", D_MODEL, D_MODEL)
        lin(p + "self_attn.temporal_info", D_MODEL, D_MODEL)
        lin(p + "self_attn.temporal_gate", D_MODEL, D_MODEL)
        lin(p + "linear1", FF_DIM, D_MODEL)
        lin(p + "linear2", D_MODEL, FF_DIM)
        for n in ("norm1", "norm2"):
            spec.append((p + n + ".weight", (D_MODEL,), "g"))
            spec.append((p + n + ".bias", (D_MODEL,), "b"))

    def encoder(prefix, length):
        lin(prefix + "input_fc", D_MODEL, 4)
        lin(prefix + "input_fc2", D_MODEL, D_MODEL * length)
        lin(prefix + "input_fc3", D_MODEL, D_MODEL + 3)
        ode(prefix)
        spec.append((prefix + "pos_encoder.pe", (200, D_MODEL), "pe"))
        lin(prefix + "pos_encoder.fc", D_MODEL, 2 * D_MODEL)

    encoder("past_encoder.", past_length)
    lin("pz_layer", 2 * ZDIM, 4 * D_MODEL)
    encoder("future_encoder.", future_length)
    lin("future_encoder.out_mlp.affine_layers.0", 128, 4 * D_MODEL)
    lin("future_encoder.qz_layer", 2 * ZDIM, 128)
    for i in range(2):
        p = f"decoder.decompose.{i}."
        spec.append((p + "conv_past.weight", (CONV_C, 2, 3), "w"))
        spec.append((p + "conv_past.bias", (CONV_C,), "b"))
        spec.append((p + "encoder_past.weight_ih_l0", (3 * GRU_H, CONV_C), "w"))
        spec.append((p + "encoder_past.weight_hh_l0", (3 * GRU_H, GRU_H), "w"))
        spec.append((p + "encoder_past.bias_ih_l0", (3 * GRU_H,), "b"))
        spec.append((p + "encoder_past.bias_hh_l0", (3 * GRU_H,), "b"))
        for name, out in (("decoder_y", 2 * future_length), ("decoder_x", 2 * past_length)):
            lin(p + name + ".layers.0", 512, 4 * D_MODEL)
            lin(p + name + ".layers.1", 256, 512)
            lin(p + name + ".layers.2", out, 256)
    return spec


def positional_table(max_len=200, d_model=D_MODEL, dtype=torch.float32):
    """Sinusoid table (model/STTODE.py:149-155)."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def make_state_dict(seed=1, past_length=8, future_length=12, dtype=torch.float32):
    """Deterministic synthetic weights with the reference's keys/shapes.

    Not the reference's init: every bias / LayerNorm affine is non-trivial so that bias paths are
    exercised (the reference zero-inits in_proj_bias, hyptransformerlib.py:388-390).  Scales follow
    fan-in so activations stay O(1) like a default-initialised reference model.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in state_dict_spec(past_length, future_length):
        if kind == "pe":
            t = positional_table(shape[0], shape[1])
        elif kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = (3.0 / fan_in) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "g":
            t = 1.0 + (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        sd[key] = t.to(dtype)
    return sd




def make_args(dataset="eth", past_length=8, future_length=12, sample_k=20):
    """The Namespace `STTODENet(args, device)` reads (train.py:21-52 defaults)."""
    return argparse.Namespace(hidden_dim=64, zdim=32, hyper_scales=[5, 11], past_length=past_length,
                              future_length=future_length, num_decompose=2, learn_prior=False,
                              ztype="gaussian", sample_k=sample_k, min_clip=2.0, dataset=dataset,
                              max_train_agent=100, rand_rot_scene=True, discrete_rot=False, traj_scale=1)
