"""GPU parity of the backward entry points (sttode_mhgsa_bwd, sttode_ode_rhs_bwd, sttode_ode_block_bwd) through the
Python autograd nodes, against torch autograd over the float64 CPU oracle (the reference differentiates the same
graph with torch autograd) and against gradients of the reference itself (tests/golden/backward.npz).
Tolerance: <= 1e-4 relative (max-norm per tensor), as for the forward."""
import numpy as np
import pytest
import torch

from oracle import sttode_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
LAYER = "past_encoder.ODE_Encoder.odeblock.odefunc.layers.0."
ATT = LAYER + "self_attn.temporal_attention_before."
ATT_KEYS = ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias")


def rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(dev):
    from oracle.ref_shim import make_args
    from sttode_b200.net import STTODENet
    sd = O.make_state_dict(1, 8, 12)
    m = STTODENet(make_args("eth", 8, 12), dev)
    m.load_state_dict(sd, strict=True)
    m.set_device(dev)
    return m.eval(), sd


def oracle_leaves(sd, prefix):
    """float64 leaf copies (requires_grad) of every tensor under `prefix`, inside a full float64 state_dict."""
    sd64 = O.cast_state_dict(sd, torch.float64)
    leaves = {}
    for k in sd64:
        if k.startswith(prefix) and sd64[k].is_floating_point():
            sd64[k] = sd64[k].clone().requires_grad_(True)
            leaves[k] = sd64[k]
    return sd64, leaves


def zero_grads(mod):
    for p in mod.parameters():
        p.grad = None


def compare_param_grads(mod, prefix, leaves, skip=()):
    worst = 0.0
    for name, p in mod.named_parameters():
        ref = leaves[prefix + name].grad
        if name in skip or ref is None:
            continue
        assert p.grad is not None, name
        e = rel(p.grad, ref)
        assert e < TOL, (name, e)
        worst = max(worst, e)
    return worst


# ------------------------------------------------------------------ MHGSA
@pytest.mark.parametrize("L,S,bsz", [(1, 1, 4), (5, 5, 3), (33, 33, 2), (128, 128, 11), (4, 6, 3), (40, 7, 5), (1, 9, 2)])
def test_mhgsa_backward_vs_oracle_autograd(model, dev, L, S, bsz):
    m, sd = model
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    g = torch.Generator().manual_seed(100 + L * 7 + S)
    q = torch.randn(L, bsz, 64, generator=g)
    kv = q if L == S else torch.randn(S, bsz, 64, generator=g)
    dout = torch.randn(L, bsz, 64, generator=g)
    # oracle, float64
    sd64, leaves = oracle_leaves(sd, ATT)
    q64 = q.double().requires_grad_(True)
    kv64 = q64 if L == S else kv.double().requires_grad_(True)
    out64, _ = O.hyp_mhsa(q64, kv64, kv64, *(sd64[ATT + k] for k in ATT_KEYS))
    out64.backward(dout.double())
    # CUDA
    zero_grads(attn)
    qd = q.to(dev).requires_grad_(True)
    kvd = qd if L == S else kv.to(dev).requires_grad_(True)
    out, w = attn(qd, kvd, kvd)
    assert rel(out, out64) < TOL
    assert not w.requires_grad
    out.backward(dout.to(dev))
    assert rel(qd.grad, q64.grad) < TOL
    if L != S:
        assert rel(kvd.grad, kv64.grad) < TOL
    compare_param_grads(attn, ATT, leaves)


def test_mhgsa_backward_distinct_key_value_tensors(model, dev):
    """key and value are different tensors (cross-attention branch): each gets its own gradient."""
    m, sd = model
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    g = torch.Generator().manual_seed(5)
    q, k, v, dout = (torch.randn(n, 3, 64, generator=g) for n in (6, 9, 9, 6))
    sd64, leaves = oracle_leaves(sd, ATT)
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    O.hyp_mhsa(q64, k64, v64, *(sd64[ATT + n] for n in ATT_KEYS))[0].backward(dout.double())
    zero_grads(attn)
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    attn(qd, kd, vd)[0].backward(dout.to(dev))
    assert rel(qd.grad, q64.grad) < TOL and rel(kd.grad, k64.grad) < TOL and rel(vd.grad, v64.grad) < TOL
    compare_param_grads(attn, ATT, leaves)


def test_mhgsa_backward_clamp_blocks_gradient(dev):
    """Outside [-1+eps, 1-eps] torch.clamp passes no gradient (hyptransformerlib.py:254 via oblique.py:37): with identity
    projections and q = k rows pointing the same way per head, every Gram entry is clamped, so the score carries no gradient
    and only the value path contributes."""
    from sttode_b200.mhgsa import Hyp_mhsa
    attn = Hyp_mhsa(64, 8).to(dev)
    with torch.no_grad():
        eye = torch.eye(64)
        attn.in_proj_weight.copy_(torch.cat([eye, eye, eye]))
        attn.in_proj_bias.zero_()
        attn.out_proj.weight.copy_(eye)
        attn.out_proj.bias.zero_()
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 2, 64, generator=g)
    scale = torch.rand(5, 2, 1, generator=g) + 0.5
    x = (base * scale)                                   # all 5 rows of an agent are parallel in every head: g_ij = 1 > 1 - eps
    dout = torch.randn(5, 2, 64, generator=g)
    sd = {k: v.detach().cpu().double().requires_grad_(True) for k, v in attn.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    O.hyp_mhsa(x64, x64, x64, *(sd[k] for k in ATT_KEYS))[0].backward(dout.double())
    xd = x.to(dev).requires_grad_(True)
    attn(xd, xd, xd)[0].backward(dout.to(dev))
    assert rel(xd.grad, x64.grad) < TOL
    # uniform softmax (all scores equal) -> d x_j = mean_i dout_i through the value path only
    assert rel(xd.grad, dout.mean(dim=0, keepdim=True).expand(5, 2, 64)) < TOL
    assert rel(attn.in_proj_weight.grad, sd["in_proj_weight"].grad) < TOL


def test_backward_is_deterministic(model, dev):
    m, _ = model
    layer = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
    g = torch.Generator().manual_seed(8)
    x, df = torch.randn(9, 11, 1, 64, generator=g).to(dev), torch.randn(9, 11, 1, 64, generator=g).to(dev)
    res = []
    for _ in range(2):
        zero_grads(layer)
        xd = x.clone().requires_grad_(True)
        layer(xd).backward(df)
        res.append([xd.grad.clone()] + [p.grad.clone() for p in layer.parameters()])
    for a, b in zip(*res):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ Hypattention / encoder layer (the ODE right-hand side)
@pytest.mark.parametrize("L,bsz", [(1, 7), (6, 5), (32, 11), (130, 3)])
def test_encoder_layer_backward_vs_oracle_autograd(model, dev, L, bsz):
    m, sd = model
    layer = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
    g = torch.Generator().manual_seed(200 + L)
    x, df = torch.randn(L, bsz, 1, 64, generator=g), torch.randn(L, bsz, 1, 64, generator=g)
    sd64, leaves = oracle_leaves(sd, LAYER)
    x64 = x.double().requires_grad_(True)
    O.encoder_layer(x64, sd64, LAYER).backward(df.double())
    zero_grads(layer)
    xd = x.to(dev).requires_grad_(True)
    layer(xd).backward(df.to(dev))
    assert rel(xd.grad, x64.grad) < TOL
    # with one key (L == 1) softmax is constant: q/k projections get exactly zero gradient on both sides
    compare_param_grads(layer, LAYER, leaves)


@pytest.mark.parametrize("L,bsz", [(1, 4), (12, 5)])
def test_hypattention_backward_vs_oracle_autograd(model, dev, L, bsz):
    m, sd = model
    att = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn
    g = torch.Generator().manual_seed(300 + L)
    x, df = torch.randn(L, bsz, 1, 64, generator=g), torch.randn(L, bsz, 1, 64, generator=g)
    sd64, leaves = oracle_leaves(sd, LAYER + "self_attn.")
    x64 = x.double().requires_grad_(True)
    O.hypattention(x64, sd64, LAYER)[0].backward(df.double())
    zero_grads(att)
    xd = x.to(dev).requires_grad_(True)
    att(xd, xd, xd)[0].backward(df.to(dev))
    assert rel(xd.grad, x64.grad) < TOL
    compare_param_grads(att, LAYER + "self_attn.", leaves)


# ------------------------------------------------------------------ ODE block (solver loop + final ReLU)
@pytest.mark.parametrize("L,bsz,steps,time", [(1, 7, 1, 12.0), (6, 5, 1, 12.0), (32, 11, 1, 12.0), (5, 4, 3, 0.6)])
def test_ode_block_backward_vs_oracle_autograd(model, dev, L, bsz, steps, time):
    from sttode_b200.layers import ODEG_Encoder
    m, sd = model
    layer = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
    enc = ODEG_Encoder(layer, 1, time, n_steps=steps).to(dev)
    enc.odeblock.odefunc.layers[0].load_state_dict(layer.state_dict())
    lay = enc.odeblock.odefunc.layers[0]
    g = torch.Generator().manual_seed(400 + L)
    x, dy = torch.randn(L, bsz, 1, 64, generator=g) * 0.7, torch.randn(L, bsz, 1, 64, generator=g)
    sd64, leaves = oracle_leaves(sd, LAYER)
    x64 = x.double().requires_grad_(True)
    grid = [time * i / steps for i in range(steps + 1)]
    y64 = O.ode_encoder(x64, sd64, "past_encoder.", grid=grid)
    y64.backward(dy.double())
    zero_grads(lay)
    xd = x.to(dev).requires_grad_(True)
    y = enc(xd)
    assert rel(y, y64) < TOL
    y.backward(dy.to(dev))
    assert rel(xd.grad, x64.grad) < TOL
    compare_param_grads(lay, LAYER, leaves)


def test_backward_error_paths(model, dev):
    """Unsupported solver for the backward raises (no silent fallback)."""
    from sttode_b200 import _lib
    from sttode_b200.layers import ODEG_Encoder
    m, _ = model
    layer = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
    enc = ODEG_Encoder(layer, 1, 0.3, method="midpoint", n_steps=2).to(dev)
    x = torch.randn(3, 2, 1, 64, device=dev, requires_grad=True)
    y = enc(x)
    with pytest.raises(_lib.SttodeError):
        y.sum().backward()


# ------------------------------------------------------------------ gradients of the reference itself
@pytest.mark.parametrize("name", ["sq5", "x46"])
def test_mhgsa_backward_vs_reference_golden(model, dev, golden, name):
    m, _ = model
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    gi, gb = golden("mhgsa"), golden("backward")
    q = torch.from_numpy(gi[f"{name}_q"]).to(dev).requires_grad_(True)
    square = gi[f"{name}_q"].shape == gi[f"{name}_kv"].shape
    kv = q if square else torch.from_numpy(gi[f"{name}_kv"]).to(dev).requires_grad_(True)
    zero_grads(attn)
    attn(q, kv, kv)[0].backward(torch.from_numpy(gb[f"mhgsa_{name}_dout"]).to(dev))
    assert rel(q.grad, gb[f"mhgsa_{name}_dq"]) < TOL
    if not square:
        assert rel(kv.grad, gb[f"mhgsa_{name}_dkv"]) < TOL
    for pname, p in attn.named_parameters():
        assert rel(p.grad, gb[f"mhgsa_{name}_g_{pname}"]) < TOL, pname


@pytest.mark.parametrize("name", ["L1", "L6"])
def test_ode_block_backward_vs_reference_golden(model, dev, golden, name):
    m, _ = model
    enc = m.past_encoder.ODE_Encoder
    gi, gb = golden("ode_block"), golden("backward")
    x = torch.from_numpy(gi[f"{name}_x"]).to(dev).requires_grad_(True)
    zero_grads(enc)
    enc(x).backward(torch.from_numpy(gb[f"ode_{name}_dy"]).to(dev))
    assert rel(x.grad, gb[f"ode_{name}_dx"]) < TOL
    for pname, p in enc.named_parameters():
        assert rel(p.grad, gb[f"ode_{name}_g_{pname}"]) < TOL, pname


# ------------------------------------------------------------------ decoder (Conv1d + GRU through time + MLPs)
@pytest.mark.parametrize("tp,tf,R,K,with_recon", [(8, 12, 7, 1, True), (10, 20, 6, 20, False), (5, 10, 3, 4, True)])
def test_decoder_backward_vs_torch_autograd(dev, tp, tf, R, K, with_recon):
    """sttode_decoder_bwd (through the train()-mode Decoder) against torch autograd over the same blocks written with torch ops
    (DecomposeBlock.forward: F.conv1d, an unrolled fp32 GRU, nn.Linear), in float64 on the CPU."""
    from oracle.ref_shim import make_args
    from sttode_b200.net import Decoder
    torch.manual_seed(11)
    dec = Decoder(make_args("nba", tp, tf)).to(dev).train()
    g = torch.Generator().manual_seed(12)
    P = R * K
    pf, z = torch.randn(R, 128, generator=g), torch.randn(P, 32, generator=g)
    traj = torch.cumsum(torch.randn(R, tp, 2, generator=g) * 0.3, dim=1)
    d_out, d_rec = torch.randn(P, tf, 2, generator=g), torch.randn(P, tp, 2, generator=g)
    # float64 CPU reference: the same module's torch-op blocks
    ref = Decoder(make_args("nba", tp, tf)).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in dec.state_dict().items()})
    pf64, z64 = pf.double().requires_grad_(True), z.double().requires_grad_(True)
    x_true = traj.double().repeat_interleave(K, dim=0)
    hidden = torch.cat((pf64.repeat_interleave(K, dim=0), z64), dim=-1)
    x_hat, pred, rec = torch.zeros_like(x_true), 0, 0
    for blk in ref.decompose:
        x_hat, y_hat = blk(x_true, x_hat, hidden)
        pred, rec = pred + y_hat, rec + x_hat
    out64 = pred + x_true[:, [-1]]
    loss64 = (out64 * d_out.double()).sum() + ((rec * d_rec.double()).sum() if with_recon else 0)
    loss64.backward()
    # CUDA
    pfd, zd = pf.to(dev).requires_grad_(True), z.to(dev).requires_grad_(True)
    trajd = traj.to(dev)
    out, rec_d = dec(pfd.repeat_interleave(K, dim=0), zd, 1, R, trajd, trajd[:, [-1]], sample_num=K, mode="inference" if K > 1 else "train")
    assert rel(out.reshape(P, tf, 2), out64) < TOL and rel(rec_d, rec) < TOL
    loss = (out.reshape(P, tf, 2) * d_out.to(dev)).sum() + ((rec_d * d_rec.to(dev)).sum() if with_recon else 0)
    loss.backward()
    assert rel(pfd.grad, pf64.grad) < TOL and rel(zd.grad, z64.grad) < TOL
    refp = dict(ref.named_parameters())
    for n, p in dec.named_parameters():
        if refp[n].grad is None:                      # block-1 decoder_x without the reconstruction: no gradient on either side
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        e = rel(p.grad, refp[n].grad)
        assert e < TOL, (n, e)
