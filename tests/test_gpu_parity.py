"""GPU parity: every C-ABI path of libsttode_sm100.so against the CPU oracle and the committed golden
vectors (which came from the reference itself).  Tolerance per north_star: <= 1e-4 relative (max-norm)."""
import numpy as np
import pytest
import torch

from oracle import sttode_oracle as O
from oracle import pmath_oracle as PO
from sttode_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4
ATT = "past_encoder.ODE_Encoder.odeblock.odefunc.layers.0.self_attn.temporal_attention_before."
LAYER = "past_encoder.ODE_Encoder.odeblock.odefunc.layers.0."


def rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def build_model(dataset, tp, tf, seed, dev, tensor_cores=True):
    from oracle.ref_shim import make_args
    from sttode_b200.net import STTODENet
    sd = O.make_state_dict(seed, tp, tf)
    m = STTODENet(make_args(dataset, tp, tf), dev)
    m.load_state_dict(sd, strict=True)
    m.set_device(dev)
    m.set_tensor_cores(tensor_cores)
    return m.eval(), sd


@pytest.fixture(scope="module", params=["tc", "ffma"])
def eth(dev, request):
    """Every test that uses this fixture runs twice: tensor-core (tcgen05) decoder and FP32-FFMA decoder."""
    m, sd = build_model("eth", 8, 12, 1, dev, tensor_cores=request.param == "tc")
    yield m, sd
    assert m.tc_status() == 0, "a tcgen05 barrier timed out"


# ------------------------------------------------------------------ MHGSA
@pytest.mark.parametrize("name", ["sq5", "sq1", "sq33", "x46"])
def test_mhgsa_vs_reference_golden(golden, eth, dev, name):
    m, _ = eth
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    g = golden("mhgsa")
    q = torch.from_numpy(g[f"{name}_q"]).to(dev)
    kv = q if name.startswith("sq") else torch.from_numpy(g[f"{name}_kv"]).to(dev)
    out, w = attn(q, kv, kv)
    assert rel(out, g[f"{name}_out"]) < TOL
    assert rel(w, g[f"{name}_w"]) < TOL
    out2, w2 = attn(q, kv, kv, need_weights=False)
    assert w2 is None and torch.equal(out, out2)


@pytest.mark.parametrize("L,bsz", [(16, 11), (128, 11), (200, 3), (512, 2)])
def test_mhgsa_sweep_vs_oracle(eth, dev, L, bsz):
    m, sd = eth
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    x = torch.randn(L, bsz, 64, generator=torch.Generator().manual_seed(L))
    ref, wref = O.hyp_mhsa(x, x, x, sd[ATT + "in_proj_weight"], sd[ATT + "in_proj_bias"], sd[ATT + "out_proj.weight"], sd[ATT + "out_proj.bias"])
    xd = x.to(dev)
    out, w = attn(xd, xd, xd)
    assert rel(out, ref) < TOL
    assert rel(w, wref) < TOL


def test_mhgsa_distinct_equal_tensors_square(eth, dev):
    """query/key/value passed as three different tensors with equal values: the reference still takes the
    self-attention branch (torch.equal) and L == S keeps the un-transposed scores."""
    m, sd = eth
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    x = torch.randn(7, 3, 64, generator=torch.Generator().manual_seed(2))
    ref, _ = O.hyp_mhsa(x, x, x, sd[ATT + "in_proj_weight"], sd[ATT + "in_proj_bias"], sd[ATT + "out_proj.weight"], sd[ATT + "out_proj.bias"])
    out, _ = attn(x.to(dev), x.clone().to(dev), x.clone().to(dev))
    assert rel(out, ref) < TOL


def test_mhgsa_clamp_and_nan_rows(eth, dev):
    """Adversarial rows: k = q direction (upper clamp), k = -q (lower clamp), a zero row -> NaN
    exactly where the reference produces NaN (Oblique.proj has no epsilon)."""
    m, sd = eth
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    W, b = sd[ATT + "in_proj_weight"].clone(), sd[ATT + "in_proj_bias"].clone()
    W[64:128] = W[:64]; b[64:128] = b[:64]              # k == q  -> diagonal hits +clamp
    x = torch.randn(6, 2, 64, generator=torch.Generator().manual_seed(5))
    x[3] = -x[2]
    b[:128] = 0                                          # then k(-x) = -k(x): lower clamp between rows 2 and 3
    ref, wref = O.hyp_mhsa(x, x, x, W, b, sd[ATT + "out_proj.weight"], sd[ATT + "out_proj.bias"])
    from sttode_b200.mhgsa import hyp_mhsa
    args = lambda t: (t, t, t, 64, 8, W.to(dev), b.to(dev), None, None, False, 0.0, sd[ATT + "out_proj.weight"].to(dev), sd[ATT + "out_proj.bias"].to(dev))
    out, w = hyp_mhsa(*args(x.to(dev)), training=False)
    assert rel(out, ref) < TOL and rel(w, wref) < TOL
    x0 = x.clone(); x0[1, 0] = 0                         # zero token: q = k = 0 for (l=1, agent 0)
    ref0, _ = O.hyp_mhsa(x0, x0, x0, W, b, sd[ATT + "out_proj.weight"], sd[ATT + "out_proj.bias"])
    out0, _ = hyp_mhsa(*args(x0.to(dev)), training=False)
    assert torch.equal(torch.isnan(out0).cpu(), torch.isnan(ref0))
    assert torch.isnan(ref0).any() and not torch.isnan(ref0).all()


def test_mhgsa_poincare_score_mode(eth, dev):
    m, sd = eth
    from sttode_b200.mhgsa import hyp_mhsa, SCORE_POINCARE_DIST
    L, bsz, c = 9, 3, 0.7
    # modest activations keep expmap0(q), expmap0(k) inside the ball, where the distance is well conditioned
    x = torch.randn(L, bsz, 64, generator=torch.Generator().manual_seed(8)) * 0.3
    W, b, Wo, bo = (sd[ATT + k] for k in ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias"))
    # oracle: same pipeline with the score replaced by -dist_matrix(project(expmap0(k)), project(expmap0(q)))
    qkv = O.linear(x, W, b); q, k, v = qkv.chunk(3, -1); q = q * 8 ** -0.5
    sp = lambda t: t.contiguous().view(L, bsz * 8, 8).transpose(0, 1)
    qP, kP, vv = PO.project(PO.expmap0(sp(q), c), c), PO.project(PO.expmap0(sp(k), c), c), sp(v)
    sc = torch.stack([-PO.dist_matrix(kP[i], qP[i], c) for i in range(bsz * 8)])
    o = torch.bmm(torch.softmax(sc, -1), vv).transpose(0, 1).contiguous().view(L, bsz, 64)
    ref = O.linear(o, Wo, bo)
    out, _ = hyp_mhsa(x.to(dev), x.to(dev), x.to(dev), 64, 8, W.to(dev), b.to(dev), None, None, False, 0.0, Wo.to(dev), bo.to(dev),
                      training=False, score_mode=SCORE_POINCARE_DIST, curvature=c)
    assert rel(out, ref) < TOL


def test_oblique_and_linear(dev):
    from sttode_b200.mhgsa import Oblique, linear
    g = torch.Generator().manual_seed(3)
    p1, p2 = torch.randn(5, 7, 8, generator=g), torch.randn(5, 9, 8, generator=g)
    M = Oblique()
    a, b = M.proj(p1.to(dev)), M.proj(p2.to(dev))
    assert rel(a, O.oblique_proj(p1)) < 1e-6
    assert rel(M.dist(a, b), O.oblique_dist(O.oblique_proj(p1), O.oblique_proj(p2))) < TOL
    x, W, bb = torch.randn(4, 3, 64, generator=g), torch.randn(192, 64, generator=g) * 0.1, torch.randn(192, generator=g)
    assert rel(linear(x.to(dev), W.to(dev), bb.to(dev)), O.linear(x, W, bb)) < 1e-5


# ------------------------------------------------------------------ ODE block
@pytest.mark.parametrize("name", ["L1", "L6", "L40"])
def test_ode_block_vs_reference_golden(golden, eth, dev, name):
    m, _ = eth
    g = golden("ode_block")
    y = m.past_encoder.ODE_Encoder(torch.from_numpy(g[f"{name}_x"]).to(dev))
    assert y.shape == g[f"{name}_y"].shape
    assert rel(y, g[f"{name}_y"]) < TOL


def test_layer_and_hypattention_vs_oracle(eth, dev):
    m, sd = eth
    layer = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
    x = torch.randn(12, 5, 1, 64, generator=torch.Generator().manual_seed(4))
    assert rel(layer(x.to(dev)), O.encoder_layer(x, sd, LAYER)) < TOL
    xd = x.to(dev)
    got, w = layer.self_attn(xd, xd, xd, need_weights=True)
    ref, wref = O.hypattention(x, sd, LAYER)
    assert rel(got, ref) < TOL and rel(w, wref) < TOL
    assert rel(m.past_encoder.ODE_Encoder.odeblock.odefunc(0.0, xd), O.encoder_layer(x, sd, LAYER)) < TOL


@pytest.mark.parametrize("method,steps", [("euler", 3), ("midpoint", 2), ("rk4_classic", 2), ("rk4", 1)])
def test_solver_loop_generalisation(eth, dev, method, steps):
    """Fixed-step solvers beyond the reference's single Euler step, against the same RHS on the CPU."""
    from sttode_b200.layers import ODEG_Encoder
    m, sd = eth
    src = m.past_encoder.ODE_Encoder
    enc = ODEG_Encoder(src.odeblock.odefunc.layers[0], 1, 0.3, method=method, n_steps=steps).to(dev).eval()
    x = torch.randn(5, 4, 1, 64, generator=torch.Generator().manual_seed(6)) * 0.5
    f = lambda y: O.encoder_layer(y, sd, LAYER)
    h, y = 0.3 / steps, x
    for _ in range(steps):
        if method == "euler":
            y = y + h * f(y)
        elif method == "midpoint":
            y = y + h * f(y + 0.5 * h * f(y))
        elif method == "rk4_classic":
            k1 = f(y); k2 = f(y + 0.5 * h * k1); k3 = f(y + 0.5 * h * k2); k4 = f(y + h * k3)
            y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        else:   # torchdiffeq's 'rk4' is the 3/8 rule
            k1 = f(y); k2 = f(y + h * k1 / 3); k3 = f(y + h * (k2 - k1 / 3)); k4 = f(y + h * (k1 - k2 + k3))
            y = y + h * (k1 + 3 * (k2 + k3) + k4) / 8
    assert rel(enc(x.to(dev)), torch.relu(y)) < TOL


# ------------------------------------------------------------------ encoder / decoder / inference
@pytest.mark.parametrize("idx,n", [(0, 2), (1, 5), (2, 10)])
def test_inference_eth_vs_reference_golden(golden, eth, dev, idx, n):
    m, _ = eth
    g = golden("inference_eth")
    obs, gt = synth.eth_scene(idx, n)
    m.z_override = synth.latent(n, seed=7 + idx)
    m.set_data(None, obs.to(dev), gt.to(dev), torch.ones(n, 8), torch.ones(n, 12))
    pred = m.inference(None)
    assert pred.shape == (20, n, 12, 2) and m.agent_num == n
    assert rel(m.past_feature, g[f"eth{idx}_pf"]) < TOL
    # displacement (prediction minus the pass-through current position), SURVEY §8d tolerance (ii)
    cur = (m.cur_location + m.scene_orig).cpu().numpy()[None]
    assert rel(pred.cpu().numpy() - cur, g[f"eth{idx}_pred"] - cur) < TOL
    # ADE/FDE delta vs the reference's predictions
    a0, f0 = O.min_ade_fde(torch.from_numpy(g[f"eth{idx}_pred"]).permute(1, 0, 2, 3), gt.permute(0, 2, 1))
    a1, f1 = O.min_ade_fde(pred.permute(1, 0, 2, 3).cpu(), gt.permute(0, 2, 1))
    assert abs(a0 - a1) <= TOL * a0 and abs(f0 - f1) <= TOL * f0
    m.z_override = None


@pytest.mark.parametrize("tcores", [True, False])
@pytest.mark.parametrize("name,tp,tf", [("nba1020", 10, 20), ("nba0510", 5, 10)])
def test_inference_nba_vs_reference_golden(golden, dev, name, tp, tf, tcores):
    g = golden("inference_nba")
    B = int(g[f"{name}_B"])
    m, _ = build_model("nba", tp, tf, 2, dev, tensor_cores=tcores)
    data = synth.nba_batch(0, B, 11, tp, tf)
    m.z_override = synth.latent(B * 11, seed=9)
    m.set_data_nba(data)
    pred = m.inference(data)
    ref = g[f"{name}_pred"]
    cur = data["past_traj"].view(B * 11, tp, 2)[:, -1].numpy()[None, :, None, :]
    assert rel(pred.cpu().numpy() - cur, ref - cur) < TOL
    assert m.tc_status() == 0


def test_inference_nba_batch128_vs_oracle(dev):
    """The reference's test batch size (test.py:618): L = S = 128 scenes attend to each other per agent."""
    B, tp, tf = 128, 10, 20
    m, sd = build_model("nba", tp, tf, 4, dev)
    data = synth.nba_batch(3, B, 11, tp, tf)
    z = synth.latent(B * 11, seed=11)
    m.z_override = z
    pred = m.inference(data)
    dump = {}
    ref = O.inference(sd, data["past_traj"].view(B * 11, tp, 2), B, 11, z, dump=dump, fast=True)
    assert rel(m.past_feature, dump["past_feature"]) < TOL
    cur = data["past_traj"].view(B * 11, tp, 2)[:, -1][None, :, None, :]
    assert rel(pred.cpu() - cur, ref - cur) < TOL


def test_batched_scenes_equal_per_scene_calls(eth, dev):
    """forecast_scenes (many scenes per launch) must reproduce set_data + inference scene by scene."""
    m, sd = eth
    scenes = synth.eth_scenes(9, cycle=(2, 3, 5, 7, 10, 60))
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes])
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    z = synth.latent(sum(counts), seed=21)
    m.z_override = z
    pred = m.forecast_scenes(obs.to(dev), sp)          # [R,K,T,2]
    pf_all = m.past_feature.clone()
    r0 = 0
    for (o, gt), n in zip(scenes, counts):
        past, orig = O.eth_scene_normalise(o)
        dump = {}
        ref = O.inference(sd, past, 1, n, z[r0 * 20:(r0 + n) * 20], scene_orig=orig, dump=dump, fast=True)   # [K,n,T,2]
        assert rel(pf_all[r0:r0 + n], dump["past_feature"]) < TOL
        cur = (past[:, -1] + orig)[:, None, None, :]
        assert rel(pred[r0:r0 + n].cpu() - cur, ref.permute(1, 0, 2, 3) - cur) < TOL
        r0 += n
    m.z_override = None


def test_module_level_apis_vs_oracle(eth, dev):
    """PastEncoder.forward and Decoder.forward with the reference's own signatures."""
    m, sd = eth
    n = 6
    obs, _ = synth.eth_scene(40, n)
    past, orig = O.eth_scene_normalise(obs)
    inputs = O.velocity_features(past)
    pf = m.past_encoder(inputs.to(dev), 1, n)
    pf_ref = O.past_encoder(inputs, 1, n, sd)
    assert rel(pf, pf_ref) < TOL
    z = synth.latent(n, seed=3)
    out, rec = m.decoder(pf_ref.repeat_interleave(20, 0).to(dev), z.to(dev), 1, n, past.to(dev), past[:, [-1]].to(dev), sample_num=20, mode="inference")
    out_ref, rec_ref = O.decoder(pf_ref.repeat_interleave(20, 0), z, past, past[:, [-1]], 20, sd, mode="inference")
    assert out.shape == out_ref.shape == (n, 20, 12, 2)
    cur = past[:, -1][:, None, None, :]
    assert rel(out.cpu() - cur, out_ref - cur) < TOL
    assert rel(rec, rec_ref) < TOL


@pytest.mark.parametrize("what", ["latent", "past_feature", "one_agent"])
def test_decoder_beyond_fp16_range_falls_back_to_fp32(dev, what):
    """The tensor-core decoder splits operands into fp16 hi/lo pairs (|v| < 65504).  Tiles that stage larger
    values must be recomputed by the FP32 kernel, so the result still matches the oracle (the training path
    feeds posterior latents of magnitude ~1e5 through the decoder with random-init weights)."""
    m, sd = build_model("eth", 8, 12, 1, dev, tensor_cores=True)
    n = 9
    obs, _ = synth.eth_scene(41, n)
    past, _ = O.eth_scene_normalise(obs)
    pf = O.past_encoder(O.velocity_features(past), 1, n, sd)
    z = synth.latent(n, seed=4)
    if what == "latent":
        z = z * 9.0e4
    elif what == "past_feature":
        pf = pf * 3.0e5
    else:                                  # a single (agent, sample) pair out of range: only its tile is redone
        z[3 * 20 + 7] *= 2.0e5
    out, rec = m.decoder(pf.repeat_interleave(20, 0).to(dev), z.to(dev), 1, n, past.to(dev), past[:, [-1]].to(dev), sample_num=20, mode="inference")
    out_ref, rec_ref = O.decoder(pf.repeat_interleave(20, 0), z, past, past[:, [-1]], 20, sd, mode="inference")
    assert torch.isfinite(out).all()
    assert rel(out, out_ref) < TOL
    assert rel(rec, rec_ref) < TOL
    assert m.tc_status() == 0


def test_in_kernel_latent_is_standard_normal_and_seeded(eth, dev):
    """With z not injected the decoder draws z ~ N(0, I) itself (the reference's randn_like).  This test checks the module-level
    behaviour only: deterministic per torch seed, different across seeds, samples of an agent differ.  The draw's statistics (moments,
    KS, independence) and oracle parity of the un-injected path are in tests/test_gpu_parity2.py (via sttode_latent_draw)."""
    m, _ = eth
    obs, gt = synth.eth_scene(7, 10)
    m.z_override = None
    m.set_data(None, obs.to(dev), gt.to(dev), None, None)
    torch.manual_seed(5); a = m.inference(None).clone()
    torch.manual_seed(5); b = m.inference(None).clone()
    torch.manual_seed(6); c = m.inference(None).clone()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert torch.isfinite(a).all()
    # the K samples of an agent differ (latent actually sampled per (agent, sample))
    assert (a[0] - a[1]).abs().max() > 1e-4


def test_min_ade_fde_kernel(eth, dev):
    import ctypes as C
    import sttode_b200._lib as L
    g = torch.Generator().manual_seed(9)
    pred = torch.randn(37, 20, 12, 2, generator=g)
    gt = torch.randn(37, 2, 12, generator=g)
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    L.check(L.lib().sttode_min_ade_fde(L.dptr(pred.to(dev)), L.dptr(gt.to(dev)), 1, 37, 20, 12, L.dptr(sums), L.stream_ptr()))
    ade, fde = O.min_ade_fde(pred, gt.permute(0, 2, 1))
    s = sums.cpu().numpy()
    assert s[2] == 37 and abs(s[0] / 37 - ade) < 1e-5 and abs(s[1] / 37 - fde) < 1e-5


def test_host_buffer_entry_point(eth, dev):
    """sttode_forecast_host: host in, host out (H2D + D2H inside the call); same seed => same result as the
    device-buffer entry point."""
    import ctypes as C
    import sttode_b200._lib as L
    m, _ = eth
    scenes = synth.eth_scenes(5, cycle=(3, 8, 20))
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes]).contiguous()
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    rows = obs.shape[0]
    pred_h = torch.empty(rows, 20, 12, 2)
    mw = m._weights()
    L.check(L.lib().sttode_forecast_host(obs.data_ptr(), 1, sp.data_ptr(), len(counts), rows, 1, 1, 1234, 20, C.byref(mw), pred_h.data_ptr(), L.stream_ptr()))
    m._seed = lambda: 1234
    pred_d = m.forecast_scenes(obs.to(dev), sp)
    del m._seed
    assert torch.equal(pred_h, pred_d.cpu())


def test_host_entry_point_chunked_copy_overlap(dev):
    """Large batches are forecast in several launches so the D2H copy of one chunk overlaps the kernels of the next; the latent
    draw is keyed by the pair's index in the WHOLE batch, so the result is bit-identical to the single-launch device entry point.
    ETH mode (ragged scenes, chunk cuts on scene boundaries) and NBA mode (cuts on whole attention groups), pageable and pinned output."""
    import ctypes as C
    import sttode_b200._lib as L
    m, _ = build_model("eth", 8, 12, 1, dev)
    scenes = synth.eth_scenes(460, cycle=(10, 20, 30, 40, 50, 60, 7))
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes]).contiguous()
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    rows = obs.shape[0]
    assert rows * 20 * 12 * 2 * 4 > 2 * (12 << 20)            # at least two chunks
    mw = m._weights()
    m._seed = lambda: 77
    want = m.forecast_scenes(obs.to(dev), sp).cpu()
    # both output modes (sttode_host_output_mode): sliced D2H copies and decoder kernels storing into the page-locked buffer itself
    prev = L.lib().sttode_host_output_mode(-1)
    try:
        for mode in (0, 1):
            assert L.lib().sttode_host_output_mode(mode) in (0, 1) and L.lib().sttode_host_output_mode(-1) == mode
            for pinned in (False, True):
                pred_h = torch.full((rows, 20, 12, 2), float("nan")).pin_memory() if pinned else torch.full((rows, 20, 12, 2), float("nan"))
                L.check(L.lib().sttode_forecast_host(obs.data_ptr(), 1, sp.data_ptr(), len(counts), rows, 1, 1, 77, 20, C.byref(mw), pred_h.data_ptr(), L.stream_ptr()))
                assert torch.equal(pred_h, want), (mode, pinned)
    finally:
        L.lib().sttode_host_output_mode(prev)
    # NBA mode: 48 groups of (L = 16 scenes x 11 agents), 10/20 frames
    mn, _ = build_model("nba", 10, 20, 2, dev)
    Lg, groups = 16, 48
    past = torch.cat([synth.nba_batch(i, Lg, 11, 10, 20)["past_traj"] for i in range(groups)]).contiguous()      # [groups*L, 11, 10, 2]
    rows = groups * Lg * 11
    mwn = mn._weights()
    pred_h = torch.empty(rows, 20, 20, 2)
    L.check(L.lib().sttode_forecast_host(past.data_ptr(), 0, None, 0, rows, Lg, 11, 5, 20, C.byref(mwn), pred_h.data_ptr(), L.stream_ptr()))
    pred_d = torch.empty(rows, 20, 20, 2, device=dev)
    ws = L.workspace(L.lib().sttode_forecast_workspace_bytes(rows, Lg, 20, 10), dev)
    pd = past.to(dev)
    L.check(L.lib().sttode_forecast(L.dptr(pd), 0, None, None, 0, rows, Lg, 11, None, 5, 20, C.byref(mwn), L.dptr(pred_d), None, None, None,
                                    L.dptr(ws), ws.numel(), L.stream_ptr()))
    assert torch.equal(pred_h, pred_d.cpu())
    assert m.tc_status() == 0 and mn.tc_status() == 0


# ------------------------------------------------------------------ pmath
@pytest.mark.parametrize("c", [1.0, 0.5])
def test_pmath_vs_reference_golden(golden, dev, c):
    from sttode_b200 import pmath as P
    g = golden("pmath")
    x, y, mat = (torch.from_numpy(g[k]).to(dev) for k in ("x", "y", "mat"))
    t = f"c{c}"
    px, py = P.project(x, c=c), P.project(y, c=c)
    got = {
        "project": px, "lambda_x": P.lambda_x(px, c=c), "mobius_add": P.mobius_add(px, py, c=c),
        "dist": P.dist(px, py, c=c), "dist0": P.dist0(px, c=c), "expmap0": P.expmap0(x, c=c),
        "logmap0": P.logmap0(px, c=c), "expmap": P.expmap(px, y, c=c), "logmap": P.logmap(px, py, c=c),
        "mobius_matvec": P.mobius_matvec(mat, px, c=c), "dist_matrix": P.dist_matrix(px, py, c=c),
    }
    for k, v in got.items():
        ref = g[f"{k}_{t}"]
        assert tuple(v.shape) == ref.shape, k
        got_np = v.cpu().numpy()
        if k in ("dist", "dist0", "dist_matrix"):
            # d = 2/sqrt(c) artanh(u): near the ball boundary d'(u) = 1/(1-u^2) ~ 5e4, so one fp32 ulp of u moves d by
            # ~3e-3 (the reference's own fp32 noise).  Compare the well-conditioned u = tanh(sqrt(c) d / 2) everywhere
            # and d itself away from the boundary.
            u_got, u_ref = np.tanh(np.sqrt(c) * got_np / 2), np.tanh(np.sqrt(c) * ref / 2)
            np.testing.assert_allclose(u_got, u_ref, rtol=0, atol=2e-6, err_msg=k)
            inner = u_ref < 0.99
            np.testing.assert_allclose(got_np[inner], ref[inner], rtol=2e-4, atol=2e-6, err_msg=k)
        else:
            np.testing.assert_allclose(got_np, ref, rtol=2e-4, atol=2e-6, err_msg=k, equal_nan=True)


def test_pmath_scalar_fns_and_large(golden, dev):
    from sttode_b200 import pmath as P
    g = golden("pmath")
    np.testing.assert_allclose(P.tanh(torch.linspace(-20, 20, 41).to(dev)).cpu().numpy(), g["tanh"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(P.artanh(torch.from_numpy(g["artanh_in"]).to(dev)).cpu().numpy(), g["artanh"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(P.arsinh(torch.from_numpy(g["arsinh_in"]).to(dev)).cpu().numpy(), g["arsinh"], rtol=1e-5, atol=1e-6)
    gen = torch.Generator().manual_seed(12)
    x = PO.project(torch.randn(300, 64, generator=gen) * 0.1)
    y = PO.project(torch.randn(200, 64, generator=gen) * 0.1)
    np.testing.assert_allclose(P.dist_matrix(x.to(dev), y.to(dev)).cpu().numpy(), PO.dist_matrix(x, y).numpy(), rtol=2e-4, atol=2e-6)
    # linearity-free property at size: logmap0(expmap0(u)) == u inside the ball
    u = torch.randn(4096, 64, generator=gen).to(dev) * 0.05
    back = P.logmap0(P.expmap0(u, c=1.0), c=1.0)
    assert rel(back, u) < 1e-4


# ------------------------------------------------------------------ edge cases (ragged, extreme sizes, option errors)
def test_single_agent_scene_and_ragged_batch(eth, dev):
    """A scene with one agent (tile of 20 pairs), ragged scene sizes that do not fill row / pair tiles."""
    m, sd = eth
    scenes = [synth.eth_scene(100 + i, n) for i, n in enumerate((1, 4, 1, 13, 2))]
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes])
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    z = synth.latent(sum(counts), seed=31)
    m.z_override = z
    pred = m.forecast_scenes(obs.to(dev), sp)
    r0 = 0
    for (o, _), n in zip(scenes, counts):
        past, orig = O.eth_scene_normalise(o)
        ref = O.inference(sd, past, 1, n, z[r0 * 20:(r0 + n) * 20], scene_orig=orig, fast=True)
        cur = (past[:, -1] + orig)[:, None, None, :]
        assert rel(pred[r0:r0 + n].cpu() - cur, ref.permute(1, 0, 2, 3) - cur) < TOL
        r0 += n
    m.z_override = None


def test_sdd_large_scene_512_agents(eth, dev):
    """SDD-shape sweep end point (BASELINE.json configs[3]): one scene with 512 agents, coordinates U(0,1500)/50."""
    m, sd = eth
    g = torch.Generator().manual_seed(77)
    n = 512
    p0 = torch.rand(n, 1, 2, generator=g) * 30.0
    traj = p0 + torch.cumsum(torch.randn(n, 8, 2, generator=g) * 0.4, dim=1)
    obs = traj.permute(0, 2, 1).contiguous()
    z = synth.latent(n, seed=5)
    m.z_override = z
    m.set_data(None, obs.to(dev), None, None, None)
    pred = m.inference(None)
    past, orig = O.eth_scene_normalise(obs)
    ref = O.inference(sd, past, 1, n, z, scene_orig=orig, fast=True)
    cur = (past[:, -1] + orig)[None, :, None, :]
    assert rel(pred.cpu() - cur, ref - cur) < TOL
    m.z_override = None


@pytest.mark.parametrize("k", [1, 7])
def test_other_sample_counts(eth, dev, k):
    """sample_k other than 20 through the C ABI (the reference hard-codes 20 in `inference`, model/STTODE.py:600, but
    `Decoder.forward` takes any sample_num, e.g. 1 in decoder_future_0)."""
    m, sd = eth
    n = 9
    obs, _ = synth.eth_scene(55, n)
    past, orig = O.eth_scene_normalise(obs)
    pf = O.past_encoder(O.velocity_features(past), 1, n, sd)
    z = synth.latent(n, sample_k=k, seed=4)
    out_ref, rec_ref = O.decoder(pf.repeat_interleave(k, 0), z, past, past[:, [-1]], k, sd, mode="inference")
    m.z_override = z
    sp = torch.tensor([0, n], dtype=torch.int32)
    pred = m.forecast_scenes(obs.to(dev), sp, sample_k=k)
    cur = (past[:, -1])[:, None, None, :]
    assert pred.shape == (n, k, 12, 2)
    assert rel(pred.cpu() - orig - cur, out_ref - cur) < TOL
    m.z_override = None


def test_ode_encoder_large_attention_axis(eth, dev):
    """Gram-tile stress of SURVEY §8d C4(ii): L = S = 256 with 64 agents through the whole ODE block."""
    m, sd = eth
    x = torch.randn(256, 64, 1, 64, generator=torch.Generator().manual_seed(12)) * 0.7
    y = m.past_encoder.ODE_Encoder(x.to(dev))
    assert rel(y, O.ode_encoder(x, sd, "past_encoder.")) < TOL


def test_c_abi_error_paths(eth, dev):
    import ctypes as C
    import sttode_b200._lib as L
    m, _ = eth
    lib = L.lib()
    x = torch.randn(4, 2, 64, device=dev)
    out = torch.empty_like(x)
    st = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
    from sttode_b200.layers import _ode_struct
    w = _ode_struct(st)
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    # workspace too small
    rc = lib.sttode_ode_block_fwd(L.dptr(x), C.byref(w), L.dptr(out), 64, 4, 2, 1, 12.0, 0, 1, L.dptr(ws), 16, L.stream_ptr())
    assert rc == -3 and b"workspace" in lib.sttode_last_error()
    # bad method / bad shape / unsupported head config
    big = L.workspace(lib.sttode_ode_block_workspace_bytes(4, 2, 3), dev)
    assert lib.sttode_ode_block_fwd(L.dptr(x), C.byref(w), L.dptr(out), 64, 4, 2, 1, 12.0, 9, 1, L.dptr(big), big.numel(), L.stream_ptr()) == -1
    assert lib.sttode_ode_block_fwd(L.dptr(x), C.byref(w), L.dptr(out), 32, 4, 2, 1, 12.0, 0, 1, L.dptr(big), big.numel(), L.stream_ptr()) == -1
    mh = w.attn
    assert lib.sttode_mhgsa_fwd(L.dptr(x), L.dptr(x), L.dptr(x), C.byref(mh), L.dptr(out), None, 4, 4, 2, 64, 4, 0, 1e-4, 1.0, L.dptr(big), big.numel(), L.stream_ptr()) == -2
    # the stream is still healthy afterwards
    y = m.past_encoder.ODE_Encoder(x.unsqueeze(2))
    assert torch.isfinite(y).all()


def test_cuda_graph_forecast(eth, dev):
    """The whole inference as one CUDA-graph launch: same predictions as the eager call for an injected-free run with the
    same RNG stream, fresh samples on every replay, and parity of the deterministic part (past_feature)."""
    m, sd = eth
    scenes = synth.eth_scenes(4, cycle=(5, 9))
    counts = [o.shape[0] for o, _ in scenes]
    obs = torch.cat([o for o, _ in scenes]).to(dev)
    sp = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    g = m.capture_forecast(rows=obs.shape[0], L=1, agents=1, n_scenes=len(counts))
    a = g(obs, sp).clone()
    b = g(obs, sp).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and not torch.equal(a, b)          # the graph bumps its own RNG counter
    # deterministic part against the oracle
    r0 = 0
    for (o, _), n in zip(scenes, counts):
        past, orig = O.eth_scene_normalise(o)
        pf = O.past_encoder(O.velocity_features(past), 1, n, sd)
        assert rel(g.past_feature[r0:r0 + n], pf) < TOL
        r0 += n
    # replay i of the graph == eager call with seed + i through the same entry point
    import ctypes as C
    import sttode_b200._lib as L
    mw = m._weights()
    pred = torch.empty_like(a)
    ws = L.workspace(L.lib().sttode_forecast_workspace_bytes(obs.shape[0], 1, 20, 8), dev)
    rs = g.row_scene
    L.check(L.lib().sttode_forecast(L.dptr(obs), 1, L.dptr(sp.to(dev)), L.dptr(rs), len(counts), obs.shape[0], 1, 1, None, g.seed + 2, 20,
                                    C.byref(mw), L.dptr(pred), None, None, None, L.dptr(ws), ws.numel(), L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(pred, b)          # warm-up used counter 0, capture does not execute, replays use 1 and 2


def test_future_encoder_and_training_objective_vs_reference(golden, dev):
    """FutureEncoder.forward and the training objective STTODENet.forward() in eval() mode (no autograd history) against the reference's own
    forward() output (tests/golden/training_objective.npz: eval mode, injected rsample draws)."""
    g = golden("training_objective")
    B, tp, tf = int(g["B"]), 10, 20
    m, sd = build_model("nba", tp, tf, 2, dev)
    data = synth.nba_batch(1, B, 11, tp, tf)
    gq = torch.Generator().manual_seed(21)
    eps_q, _, eps_p = torch.randn(B * 11, 32, generator=gq), torch.randn(B * 11, 32, generator=gq), torch.randn(B * 11 * 20, 32, generator=gq)
    m.eps_override = (eps_q, eps_p)
    m.set_data_nba(data)
    total, lp, lr, lk, ld = m.forward()
    assert rel(m.qz_param, g["qz"]) < TOL
    for got, key in ((float(total), "total"), (lp, "pred"), (lr, "recover"), (lk, "kl"), (ld, "diverse")):
        assert abs(got - float(g[key])) <= TOL * abs(float(g[key])), key
    # the Decoder K = 1 path returns the reconstruction the reference calls recover_pre_seq
    past = data["past_traj"].view(B * 11, tp, 2)
    ref = O.training_objective(sd, past, data["future_traj"].view(B * 11, tf, 2), B, 11, eps_q, eps_p)
    assert abs(float(total) - ref["total"]) <= TOL * abs(ref["total"])


def test_pmath_autograd_rules(dev):
    """Artanh / Arsinh / RiemannianGradient carry the reference's custom backward rules (hyptorch/pmath.py:16-60)."""
    from sttode_b200 import pmath as P
    u = torch.tensor([-1.5, -0.99999, -0.5, 0.0, 0.3, 0.99998, 2.0], device=dev, requires_grad=True)
    g = torch.linspace(0.5, 2.0, 7, device=dev)
    P.artanh(u).backward(g)
    uc = u.detach().clamp(-1 + 1e-5, 1 - 1e-5)
    assert torch.allclose(u.grad, g / (1 - uc ** 2), rtol=1e-6)             # passes through the clamp, as upstream
    s = torch.linspace(-20, 20, 9, device=dev, requires_grad=True)
    gs = torch.ones(9, device=dev)
    P.arsinh(s).backward(gs)
    assert torch.allclose(s.grad, 1 / torch.sqrt(1 + s.detach() ** 2), rtol=1e-6)
    x = (torch.randn(5, 8, generator=torch.Generator().manual_seed(2)) * 0.2).to(dev).requires_grad_(True)
    w = torch.randn(5, 8, generator=torch.Generator().manual_seed(3)).to(dev)
    (P.RiemannianGradient.apply(x) * w).sum().backward()
    assert torch.allclose(x.grad, w * (1 - x.detach().pow(2).sum(-1, keepdim=True)).pow(2) / 4, rtol=1e-6)
