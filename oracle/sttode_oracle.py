"""CPU oracle for the STTODE forward hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker / the timed CPU baseline.  Nothing under `sttode_b200/`
imports it; the product path fails loudly when the CUDA library is missing.

What it is: a functional (no nn.Module) PyTorch-CPU restatement of the reference algorithm, dtype-generic
so the same code runs in float32 (the reference precision) and float64 (the arbiter).  Each function
cites the reference file:line it follows (paths relative to /root/reference).

Parity status: the reference ships no tests / golden vectors, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the unmodified reference under
`oracle/ref_shim.py`, loads the same state_dict into it, and commits its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this oracle against them (bit-exact where the op sequence is
identical, <=2e-6 relative otherwise).  The only third-party arithmetic on the path, torchdiffeq 0.2.3's
fixed-grid Euler, is restated in `ode_euler` (and in the shim) — that boundary is "parity unpinned" by
any upstream test.
"""
from collections import OrderedDict
import math

import torch
import torch.nn.functional as F

D_MODEL = 64
N_HEAD = 8
FF_DIM = 1024
ZDIM = 32
GRU_H = 96
CONV_C = 32
ODE_TIME = 12.0          # model/STTODE.py:195,254
OBLIQUE_EPS = {torch.float32: 1e-4, torch.float64: 1e-7}   # core/manifolds/oblique.py:7


# --------------------------------------------------------------------------------------------------
# weights: the deterministic synthetic state_dict lives in the package (bench.py's product leg needs it and
# must not import oracle/); re-exported here for the tests.
# --------------------------------------------------------------------------------------------------
from sttode_b200.synth import make_state_dict, positional_table, state_dict_spec      # noqa: E402,F401


def cast_state_dict(sd, dtype):
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


# --------------------------------------------------------------------------------------------------
# manifold + MHGSA  (hyptransformerlib.py, core/manifolds/oblique.py)
# --------------------------------------------------------------------------------------------------
def linear(x, w, b=None):
    """hyptransformerlib.py:7-26 — addmm only for 2-D input, else matmul then in-place bias add."""
    if x.dim() == 2 and b is not None:
        return torch.addmm(b, x, w.t())
    out = x.matmul(w.t())
    if b is not None:
        out += b
    return out


def oblique_proj(p):
    """core/manifolds/oblique.py:15-16 — row-normalise, no epsilon."""
    return p / p.norm(dim=-1, keepdim=True)


def oblique_dist(p1, p2):
    """core/manifolds/oblique.py:36-45 — acos(clamp(p2 @ p1^T)); note the argument order."""
    eps = OBLIQUE_EPS[p1.dtype]
    inner = (p2 @ p1.transpose(-2, -1)).clamp(-1 + eps, 1 - eps)
    return torch.acos(inner)


def hyp_mhsa(query, key, value, in_w, in_b, out_w, out_b, num_heads=N_HEAD, need_weights=True):
    """hyptransformerlib.py:29-311, live branches only (no masks, no bias_kv, dropout p=0).

    query [L,bsz,E]; key/value [S,bsz,E].  Returns (out [L,bsz,E], head-averaged weights [bsz,L,S] | None).
    """
    L, bsz, E = query.shape
    S = key.shape[0]
    hd = E // num_heads
    q = linear(query, in_w[:E], in_b[:E])                       # :115 (self-attn: one fused projection)
    k = linear(key, in_w[E:2 * E], in_b[E:2 * E])
    v = linear(value, in_w[2 * E:], in_b[2 * E:])
    q = q * (float(hd) ** -0.5)                                  # :191
    q = q.contiguous().view(L, bsz * num_heads, hd).transpose(0, 1)   # :214-218
    k = k.contiguous().view(S, bsz * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(S, bsz * num_heads, hd).transpose(0, 1)
    w = -oblique_dist(oblique_proj(q), oblique_proj(k))          # :252-254 -> [bsz*H, S, L]
    if list(w.shape) != [bsz * num_heads, L, S]:                 # :261-265 (no transpose when L == S)
        w = w.transpose(1, 2)
    w = torch.softmax(w, dim=-1)                                 # :294
    o = torch.bmm(w, v)                                          # :300
    o = o.transpose(0, 1).contiguous().view(L, bsz, E)           # :304
    o = linear(o, out_w, out_b)                                  # :305
    if need_weights:
        return o, w.view(bsz, num_heads, L, S).sum(dim=1) / num_heads   # :308-309
    return o, None


def hyp_mhsa_stages(x, in_w, in_b, num_heads=N_HEAD):
    """Per-stage dump for kernel debugging: q_hat, k_hat, v [bsz*H, L, hd], scores, probs."""
    L, bsz, E = x.shape
    hd = E // num_heads
    qkv = linear(x, in_w, in_b)
    q, k, v = qkv.chunk(3, dim=-1)
    q = q * (float(hd) ** -0.5)
    sp = lambda t: t.contiguous().view(L, bsz * num_heads, hd).transpose(0, 1)
    qh, kh, vv = oblique_proj(sp(q)), oblique_proj(sp(k)), sp(v)
    sc = -oblique_dist(qh, kh)
    return dict(q_hat=qh, k_hat=kh, v=vv, scores=sc, probs=torch.softmax(sc, dim=-1))


# --------------------------------------------------------------------------------------------------
# geodesic transformer layer + tensor-ODE block  (hypertransformer.py, ode_demo.py)
# --------------------------------------------------------------------------------------------------
def _attn_prefix(p):
    return p + "self_attn.temporal_attention_before."


def hypattention(x4, sd, p):
    """hypertransformer.py:55-89.  x4 [T, N, sn, D] -> [T, N, sn, D]."""
    T, N, sn, D = x4.shape
    x3 = x4.reshape(T, N * sn, D)
    a = _attn_prefix(p)
    m, w = hyp_mhsa(x3, x3, x3, sd[a + "in_proj_weight"], sd[a + "in_proj_bias"],
                    sd[a + "out_proj.weight"], sd[a + "out_proj.bias"])
    info = torch.tanh(F.linear(m, sd[p + "self_attn.temporal_info.weight"], sd[p + "self_attn.temporal_info.bias"]))
    gate = torch.sigmoid(F.linear(m, sd[p + "self_attn.temporal_gate.weight"], sd[p + "self_attn.temporal_gate.bias"]))
    return (info * gate).reshape(T, N, sn, D), w


def encoder_layer(x4, sd, p):
    """hypertransformer.py:134-153 (dropouts p=0, model/STTODE.py:192,251)."""
    D = x4.shape[-1]
    s = x4 + hypattention(x4, sd, p)[0]
    s = F.layer_norm(s, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    ff = F.linear(torch.relu(F.linear(s, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                  sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(s + ff, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)


def ode_euler(func, y0, t_grid):
    """torchdiffeq 0.2.3 odeint(method='euler') with step_size=None: the grid IS `t`; one RHS eval per
    interval; dt*f(t0,y0) increments (third-party, un-vendored: requirement.txt:195; ode_demo.py:188)."""
    y = y0
    for i in range(len(t_grid) - 1):
        y = y + (t_grid[i + 1] - t_grid[i]) * func(t_grid[i], y)
    return y


def ode_encoder(x4, sd, prefix, time=ODE_TIME, grid=None):
    """ODEG_Encoder.forward (ode_demo.py:217-231): relu(odeint(f, x, [0,time], euler)[1]).

    prefix e.g. 'past_encoder.'; `grid` overrides the 2-point time grid (generalised solver loop)."""
    p = prefix + "ODE_Encoder.odeblock.odefunc.layers.0."
    t = torch.tensor([0, time]).type_as(x4) if grid is None else torch.as_tensor(grid).type_as(x4)
    return torch.relu(ode_euler(lambda _t, y: encoder_layer(y, sd, p), x4, t))


# --------------------------------------------------------------------------------------------------
# encoders  (model/STTODE.py:137-236)
# --------------------------------------------------------------------------------------------------
def add_category(x):
    """model/STTODE.py:199-210 — only the last agent's third flag is set."""
    B, N, _ = x.shape
    cat = torch.zeros(N, 3).type_as(x)
    cat[N - 1, 2] = 1
    return torch.cat((x, cat.repeat(B, 1, 1)), dim=-1)


def encoder_prologue(inputs, B, N, sd, prefix, drop_mask=None):
    """PastEncoder.forward up to the ODE block (model/STTODE.py:214-223).  drop_mask None = eval mode (PE dropout off);
    a 0/1 mask [B*N,T,64] = train mode with that Bernoulli draw (nn.Dropout(0.1): x * mask / 0.9, :140,176)."""
    T = inputs.shape[1]
    tf_in = F.linear(inputs, sd[prefix + "input_fc.weight"], sd[prefix + "input_fc.bias"]).view(B * N, T, D_MODEL)
    pe = sd[prefix + "pos_encoder.pe"][:T][None].repeat(B * N, 1, 1)
    tf_pos = F.linear(torch.cat([tf_in, pe], dim=-1), sd[prefix + "pos_encoder.fc.weight"], sd[prefix + "pos_encoder.fc.bias"])
    if drop_mask is not None:
        tf_pos = tf_pos * (drop_mask.to(tf_pos.dtype) / (1.0 - 0.1))
    tf_pos = tf_pos.view(B, N, T, D_MODEL)
    x = F.linear(tf_pos.contiguous().view(B, N, T * D_MODEL), sd[prefix + "input_fc2.weight"], sd[prefix + "input_fc2.bias"])
    return F.linear(add_category(x), sd[prefix + "input_fc3.weight"], sd[prefix + "input_fc3.bias"])   # [B,N,64]


def past_encoder(inputs, B, N, sd, prefix="past_encoder.", drop_mask=None):
    """model/STTODE.py:214-236 -> [B*N, 128] = cat(pre-ODE feature, ODE output)."""
    x0 = encoder_prologue(inputs, B, N, sd, prefix, drop_mask)
    y = ode_encoder(x0.unsqueeze(2), sd, prefix).squeeze(2)
    return torch.cat((x0, y), dim=-1).view(B * N, -1)


def future_encoder(inputs_post, B, N, past_feature, sd, prefix="future_encoder.", drop_mask=None):
    """FutureEncoder.forward (model/STTODE.py:276-300) -> q(z) parameters [B*N, 2*zdim]."""
    x0 = encoder_prologue(inputs_post, B, N, sd, prefix, drop_mask)
    y = ode_encoder(x0.unsqueeze(2), sd, prefix).squeeze(2)
    ff = torch.cat((x0, y), dim=-1).view(B * N, -1)
    h = torch.relu(F.linear(torch.cat((past_feature, ff), dim=-1), sd[prefix + "out_mlp.affine_layers.0.weight"],
                            sd[prefix + "out_mlp.affine_layers.0.bias"]))
    return F.linear(h, sd[prefix + "qz_layer.weight"], sd[prefix + "qz_layer.bias"])


def posterior_inputs(past_traj, future_traj):
    """inputs_for_posterior (model/STTODE.py:477-481): future positions and their velocities (first step from the last past position)."""
    vel = future_traj - torch.cat([past_traj[:, [-1]], future_traj[:, :-1, :]], dim=1)
    return torch.cat((future_traj, vel), dim=-1)


def training_objective(sd, past_traj, future_traj, B, N, eps_q, eps_p, min_clip=2.0, masks=None, features=None, tensors=False):
    """STTODENet.forward (model/STTODE.py:553-568) with the two live rsample draws injected (eps_q [R,32] for q(z|x,y);
    eps_p [R*20,32] for the prior samples of the best-of-20 term).  masks None = eval mode; (mask_past, mask_future) = train
    mode with those PE-dropout draws.  features = (inputs, inputs_for_posterior) when they are not the NBA-style ones derived
    from the trajectories (ETH staging keeps un-normalised velocities, model/STTODE.py:432-457).
    Returns (total, pred, recover, kl, diverse) as python floats plus the q(z) parameters; tensors=True keeps `total` a
    tensor with autograd history (sd entries that require grad receive the reference's gradients)."""
    inputs, inputs_post = features if features is not None else (velocity_features(past_traj), posterior_inputs(past_traj, future_traj))
    mp_, mf_ = masks if masks is not None else (None, None)
    pf = past_encoder(inputs, B, N, sd, drop_mask=mp_)
    qz = future_encoder(inputs_post, B, N, pf, sd, drop_mask=mf_)
    mu, logvar = torch.chunk(qz, 2, dim=-1)
    zq = mu + eps_q * torch.exp(0.5 * logvar)
    cur = past_traj[:, [-1]]
    pred, rec = decoder(pf, zq, past_traj, cur, 1, sd, mode="train")
    Tf, Tp = future_traj.shape[1], past_traj.shape[1]
    loss_pred = (future_traj - pred).pow(2).sum() / B / Tf
    loss_rec = (past_traj - rec).pow(2).sum() / B / Tp
    kl = 0.5 * ((mu / (1 + 1e-8)) ** 2 + (torch.exp(0.5 * logvar) / (1 + 1e-8)) ** 2) - 0.5 - torch.log(torch.exp(0.5 * logvar) / (1 + 1e-8))
    loss_kl = (kl.sum() / (B * N)).clamp_min(min_clip)
    div, _ = decoder(pf.repeat_interleave(20, dim=0), eps_p, past_traj, cur, 20, sd, mode="inference")      # [R,20,Tf,2]
    loss_div = (future_traj.unsqueeze(1) - div).pow(2).sum(dim=-1).sum(dim=-1).min(dim=1)[0].mean()
    total = loss_pred + loss_rec + loss_kl + loss_div
    return dict(total=total if tensors else float(total.detach()), pred=float(loss_pred.detach()), recover=float(loss_rec.detach()), kl=float(loss_kl.detach()),
                diverse=float(loss_div.detach()), qz=qz)


# --------------------------------------------------------------------------------------------------
# decoder  (model/STTODE.py:16-77, 302-347; model/utils.py:67-95)
# --------------------------------------------------------------------------------------------------
def gru_last_state(x, w_ih, w_hh, b_ih, b_hh, fast=False):
    """nn.GRU(batch_first=True) final hidden state, h0 = 0 (model/STTODE.py:31,68-69).  x [rows,T,C]."""
    if fast:       # the ATen entry point nn.GRU itself dispatches to (used for CPU-baseline timing)
        h0 = torch.zeros(1, x.shape[0], w_hh.shape[1], dtype=x.dtype, device=x.device)
        return torch._VF.gru(x, h0, [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, False, False, True)[1].squeeze(0)
    Hn = w_hh.shape[1]
    h = torch.zeros(x.shape[0], Hn, dtype=x.dtype, device=x.device)
    for t in range(x.shape[1]):
        gi = F.linear(x[:, t], w_ih, b_ih)
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :Hn] + gh[:, :Hn])
        zt = torch.sigmoid(gi[:, Hn:2 * Hn] + gh[:, Hn:2 * Hn])
        n = torch.tanh(gi[:, 2 * Hn:] + r * gh[:, 2 * Hn:])
        h = (1 - zt) * n + zt * h
    return h


def mlp3(x, sd, p):
    """MLP(hidden_size=(512,256)) (model/utils.py:67-95): Linear-ReLU-Linear-ReLU-Linear."""
    h = torch.relu(F.linear(x, sd[p + "layers.0.weight"], sd[p + "layers.0.bias"]))
    h = torch.relu(F.linear(h, sd[p + "layers.1.weight"], sd[p + "layers.1.bias"]))
    return F.linear(h, sd[p + "layers.2.weight"], sd[p + "layers.2.bias"])


def decompose_block(x_true, x_hat, f, sd, p, past_len, future_len, fast=False):
    """DecomposeBlock.forward (model/STTODE.py:51-77)."""
    x_ = torch.transpose(x_true - x_hat, 1, 2)
    emb = torch.relu(F.conv1d(x_, sd[p + "conv_past.weight"], sd[p + "conv_past.bias"], stride=1, padding=1))
    emb = torch.transpose(emb, 1, 2)
    state = gru_last_state(emb, sd[p + "encoder_past.weight_ih_l0"], sd[p + "encoder_past.weight_hh_l0"],
                           sd[p + "encoder_past.bias_ih_l0"], sd[p + "encoder_past.bias_hh_l0"], fast=fast)
    feat = torch.cat((f, state), dim=1)
    x_after = mlp3(feat, sd, p + "decoder_x.").contiguous().view(-1, past_len, 2)
    y_hat = mlp3(feat, sd, p + "decoder_y.").contiguous().view(-1, future_len, 2)
    return x_after, y_hat, state


def decoder(past_feature_rep, z, past_traj, cur_location, sample_num, sd, mode="inference", fast=False,
            dump=None):
    """Decoder.forward (model/STTODE.py:320-347).  past_feature_rep / z are [rows*K, .]."""
    past_len, future_len = past_traj.shape[1], sd["decoder.decompose.0.decoder_y.layers.2.bias"].shape[0] // 2
    rows_k = past_feature_rep.shape[0]
    x_true = past_traj.repeat_interleave(sample_num, dim=0)
    hidden = torch.cat((past_feature_rep.view(-1, sample_num, past_feature_rep.shape[-1]),
                        z.view(-1, sample_num, z.shape[-1])), dim=-1).view(rows_k, -1)
    x_hat = torch.zeros_like(x_true)
    pred = torch.zeros(rows_k, future_len, 2, dtype=x_true.dtype, device=x_true.device)
    rec = torch.zeros(rows_k, past_len, 2, dtype=x_true.dtype, device=x_true.device)
    for i in range(2):
        x_hat, y_hat, state = decompose_block(x_true, x_hat, hidden, sd, f"decoder.decompose.{i}.",
                                              past_len, future_len, fast=fast)
        if dump is not None:
            dump[f"state{i}"], dump[f"x_hat{i}"], dump[f"y_hat{i}"] = state, x_hat, y_hat
        pred += y_hat
        rec += x_hat
    out = pred + cur_location.repeat_interleave(sample_num, dim=0)
    if mode == "inference":
        out = out.view(-1, sample_num, *out.shape[1:])
    return out, rec


# --------------------------------------------------------------------------------------------------
# model API  (model/STTODE.py:397-486, 574-623)
# --------------------------------------------------------------------------------------------------
def eth_scene_normalise(obs_traj):
    """set_data, eval mode (model/STTODE.py:397-461): obs_traj [N,2,T_p] -> (past_traj [N,T_p,2], scene_orig [2])."""
    pre_motion = obs_traj.permute(2, 0, 1)                      # [T,N,2]
    scene_orig = pre_motion[[-1]].view(-1, 2).mean(dim=0)
    return (pre_motion - scene_orig).permute(1, 0, 2), scene_orig


def eth_train_staging(obs_traj, gt_traj, theta):
    """STTODENet.set_data in train() mode (model/STTODE.py:397-461) without agent subsampling: rotation by `theta` about the
    scene origin (rotation_2d_torch, :6-14).  obs [N,2,T_p], gt [N,2,T_f] -> (inputs, inputs_for_posterior, past, future)."""
    pre, fut = obs_traj.permute(2, 0, 1), gt_traj.permute(2, 0, 1)
    orig = pre[[-1]].view(-1, 2).mean(dim=0)

    def rot(x):
        nx = x - orig
        out = torch.zeros_like(x)
        out[..., 0] = nx[..., 0] * torch.cos(theta) - nx[..., 1] * torch.sin(theta)
        out[..., 1] = nx[..., 0] * torch.sin(theta) + nx[..., 1] * torch.cos(theta)
        return out + orig, out
    pre, pre_n = rot(pre)
    fut, fut_n = rot(fut)
    pre_vel = pre[1:] - pre[:-1]
    pre_vel = torch.cat([pre_vel[[0]], pre_vel], dim=0)
    fut_vel = fut - torch.cat([pre[[-1]], fut[:-1]])
    return (torch.cat([pre_n, pre_vel], dim=-1).permute(1, 0, 2), torch.cat([fut_n, fut_vel], dim=-1).permute(1, 0, 2),
            pre_n.permute(1, 0, 2), fut_n.permute(1, 0, 2))


def velocity_features(past_traj):
    """model/STTODE.py:582-596 — velocity from the (normalised) positions, first step duplicated."""
    vel = past_traj[:, 1:] - past_traj[:, :-1, :]
    vel = torch.cat([vel[:, [0]], vel], dim=1)
    return torch.cat((past_traj, vel), dim=-1)


def inference(sd, past_traj, batch_size, agent_num, z, sample_k=20, scene_orig=None, fast=False, dump=None):
    """STTODENet.inference (model/STTODE.py:574-623) with the latent `z` [rows*K, 32] injected.

    past_traj [B*N, T_p, 2] (already scene-normalised for ETH/SDD).  Returns [K, B*N, T_f, 2]."""
    inputs = velocity_features(past_traj)
    cur_location = past_traj[:, [-1]]
    pf = past_encoder(inputs, batch_size, agent_num, sd)
    if dump is not None:
        dump["past_feature"] = pf
    pf_rep = pf.repeat_interleave(20, dim=0)                    # sample_num hard-coded 20 (:600)
    out, _ = decoder(pf_rep, z, past_traj, cur_location, sample_k, sd, mode="inference", fast=fast, dump=dump)
    out = out.permute(1, 0, 2, 3)
    if scene_orig is not None:                                  # dataset != 'nba' (:621-622)
        out = out + scene_orig
    return out


# --------------------------------------------------------------------------------------------------
# metrics (utils/metrics.py:7-26), vectorised over agents
# --------------------------------------------------------------------------------------------------
def min_ade_fde(pred_nkt2, gt_nt2):
    """pred [N,K,T,2], gt [N,T,2] -> (mean over agents of min-over-K ADE, same for FDE)."""
    dist = torch.linalg.norm(pred_nkt2 - gt_nt2[:, None], dim=-1)     # [N,K,T]
    ade = dist.mean(dim=-1).min(dim=1)[0].mean()
    fde = dist[..., -1].min(dim=1)[0].mean()
    return float(ade), float(fde)
