"""Size-independent properties of the path, checked on the oracle (CPU, float64 where the property is exact up to rounding).  They hold for the reference by
construction and are the invariants a correct kernel must keep: attention rows are probability vectors; agents are a BATCH axis of MHGSA
(permuting them permutes the output, the attention axis is the scene axis); Oblique.proj is idempotent with unit rows; the swap quirk for L == S; K samples
are independent given the latent; ETH scenes are translation-equivariant (scene centring + velocities); the solver grid composes; min-ADE/FDE ignores the
sample order and is zero when one sample is the ground truth."""
import pytest
import torch

from oracle import sttode_oracle as O
from sttode_b200 import synth

P = "past_encoder.ODE_Encoder.odeblock.odefunc.layers.0."
A = P + "self_attn.temporal_attention_before."


@pytest.fixture(scope="module")
def sd64():
    return O.make_state_dict(1, 8, 12, torch.float64)


def mhsa(sd, q, kv=None):
    kv = q if kv is None else kv
    return O.hyp_mhsa(q, kv, kv, sd[A + "in_proj_weight"], sd[A + "in_proj_bias"], sd[A + "out_proj.weight"], sd[A + "out_proj.bias"])


def test_attention_rows_are_probability_vectors_and_agents_are_a_batch_axis(sd64):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(9, 6, 64, generator=g, dtype=torch.float64)                   # [L scenes, bsz agents, E]
    out, w = mhsa(sd64, x)
    assert w.shape == (6, 9, 9) and torch.allclose(w.sum(-1), torch.ones(6, 9, dtype=torch.float64), atol=1e-12) and (w > 0).all()
    perm = torch.tensor([3, 0, 5, 1, 4, 2])
    out_p, w_p = mhsa(sd64, x[:, perm])
    assert torch.allclose(out_p, out[:, perm], atol=1e-12) and torch.allclose(w_p, w[perm], atol=1e-12)
    # one scene on the attention axis: softmax over a single key is 1, the block reduces to out_proj(v_proj(x))  (ETH / SDD mode)
    out1, w1 = mhsa(sd64, x[:1])
    E = 64
    v = O.linear(x[:1], sd64[A + "in_proj_weight"][2 * E:], sd64[A + "in_proj_bias"][2 * E:])
    assert torch.equal(w1, torch.ones(6, 1, 1, dtype=torch.float64))
    assert torch.allclose(out1, O.linear(v, sd64[A + "out_proj.weight"], sd64[A + "out_proj.bias"]), atol=1e-12)


def test_square_case_uses_the_untransposed_score_and_cross_case_the_transposed_one(sd64):
    """hyptransformerlib.py:251-265: Oblique.dist returns [.., S, L]; the reference transposes it only when its shape is not already [.., L, S], i.e. never
    for L == S.  So out_i = sum_j softmax_j(-acos(k_i . q_j)) v_j in the square case, while a cross-attention call with the same tensors padded by one key
    uses -acos(q_i . k_j).  The two differ unless the score matrix is symmetric."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(5, 2, 64, generator=g, dtype=torch.float64)
    st = O.hyp_mhsa_stages(x, sd64[A + "in_proj_weight"], sd64[A + "in_proj_bias"])
    qh, kh, v = st["q_hat"], st["k_hat"], st["v"]
    eps = O.OBLIQUE_EPS[torch.float64]                                                                      # core/manifolds/oblique.py:7: 1e-4 in fp32, 1e-7 in fp64
    w_ref = torch.softmax(-torch.acos((kh @ qh.transpose(1, 2)).clamp(-1 + eps, 1 - eps)), dim=-1)          # [bsz*H, i (key index!), j (query index)]
    assert torch.allclose(st["probs"], w_ref, atol=1e-12)
    w_textbook = torch.softmax(-torch.acos((qh @ kh.transpose(1, 2)).clamp(-1 + eps, 1 - eps)), dim=-1)
    assert (w_ref - w_textbook).abs().max() > 1e-3                                                          # the quirk is observable
    _, w = mhsa(sd64, x)
    assert torch.allclose(w, w_ref.view(2, 8, 5, 5).mean(1), atol=1e-12)


def test_oblique_proj_is_idempotent_with_unit_rows_and_dist_is_clamped():
    g = torch.Generator().manual_seed(6)
    p = torch.randn(3, 7, 8, generator=g, dtype=torch.float64)
    u = O.oblique_proj(p)
    assert torch.allclose(u.norm(dim=-1), torch.ones(3, 7, dtype=torch.float64), atol=1e-15) and torch.allclose(O.oblique_proj(u), u, atol=1e-15)
    # the diagonal is acos(1 - eps), never 0: the reference's clamp epsilon (core/manifolds/oblique.py:7: 1e-4 for float32, 1e-7 for float64)
    for dt, eps in ((torch.float64, 1e-7), (torch.float32, 1e-4)):
        assert O.OBLIQUE_EPS[dt] == eps
        ud = O.oblique_proj(p.to(dt))
        hi, lo = torch.acos(torch.tensor(1 - eps, dtype=dt)), torch.acos(torch.tensor(-1 + eps, dtype=dt))
        assert torch.allclose(O.oblique_dist(ud, ud).diagonal(dim1=-2, dim2=-1), hi.expand(3, 7), rtol=1e-6)
        assert torch.allclose(O.oblique_dist(ud, -ud).diagonal(dim1=-2, dim2=-1), lo.expand(3, 7), rtol=1e-6)
    assert torch.allclose(O.oblique_dist(u[:, :4], u[:, 2:]), O.oblique_dist(u[:, 2:], u[:, :4]).transpose(1, 2), atol=1e-15)      # [.., S, L] <-> [.., L, S]
    z = O.oblique_proj(torch.zeros(1, 8))                                          # no epsilon in proj: a zero row is NaN, as upstream
    assert torch.isnan(z).all()


def test_solver_grid_composes(sd64):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 4, 1, 64, generator=g, dtype=torch.float64)
    f = lambda y: O.encoder_layer(y, sd64, P)
    assert torch.allclose(O.ode_encoder(x, sd64, "past_encoder."), torch.relu(x + 12.0 * f(x)), atol=1e-12)
    two = O.ode_encoder(x, sd64, "past_encoder.", grid=[0.0, 5.0, 12.0])
    y1 = x + 5.0 * f(x)
    assert torch.allclose(two, torch.relu(y1 + 7.0 * f(y1)), atol=1e-12)
    assert torch.allclose(O.ode_encoder(x, sd64, "past_encoder.", grid=[0.0, 12.0]), O.ode_encoder(x, sd64, "past_encoder."), atol=0)


def test_samples_are_independent_given_the_latent_and_agents_given_the_scene(sd64):
    """model/STTODE.py:600-619: row r*K + k of the repeated features belongs to (agent r, sample k).  Permuting the K latent rows of an agent permutes that
    agent's K predictions and nothing else; in ETH mode (attention over ONE scene) agents do not interact at all, so dropping agents does not change the rest."""
    N, K = 5, 20
    obs, _ = synth.eth_scene(11, N)
    past, orig = O.eth_scene_normalise(obs.double())
    z = synth.latent(N, seed=3).double()
    pred = O.inference(sd64, past, 1, N, z, scene_orig=orig)                      # [K, N, T_f, 2]
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(1))
    z2 = z.view(N, K, 32).clone()
    z2[2] = z2[2][perm]
    pred2 = O.inference(sd64, past, 1, N, z2.view(N * K, 32), scene_orig=orig)
    assert torch.allclose(pred2[:, 2], pred[perm, 2], atol=1e-10)
    keep = [0, 1, 3, 4]
    assert torch.allclose(pred2[:, keep], pred[:, keep], atol=0)
    # the category flag marks the LAST agent of a scene (model/STTODE.py:199-210), so a prefix of the agents (not containing it) is unaffected by the others
    sub = O.inference(sd64, past[:3], 1, 3, z.view(N, K, 32)[:3].reshape(-1, 32), scene_orig=orig)
    assert torch.allclose(sub[:, :2], pred[:, :2], atol=1e-10)
    assert (sub[:, 2] - pred[:, 2]).abs().max() > 1e-6                             # agent 2 became the last one: its flag changed


def test_eth_scene_is_translation_equivariant(sd64):
    N = 6
    obs, _ = synth.eth_scene(12, N)
    obs = obs.double()
    z = synth.latent(N, seed=4).double()
    shift = torch.tensor([37.5, -12.25], dtype=torch.float64)
    p0, o0 = O.eth_scene_normalise(obs)
    p1, o1 = O.eth_scene_normalise(obs + shift[None, :, None])
    assert torch.allclose(p0, p1, atol=1e-12) and torch.allclose(o1 - o0, shift, atol=1e-12)
    a = O.inference(sd64, p0, 1, N, z, scene_orig=o0)
    b = O.inference(sd64, p1, 1, N, z, scene_orig=o1)
    assert torch.allclose(b - a, shift.expand_as(a), atol=1e-9)


def test_min_ade_fde_properties():
    g = torch.Generator().manual_seed(8)
    pred = torch.randn(13, 20, 12, 2, generator=g, dtype=torch.float64)
    gt = torch.randn(13, 12, 2, generator=g, dtype=torch.float64)
    ade, fde = O.min_ade_fde(pred, gt)
    perm = torch.randperm(20, generator=g)
    assert O.min_ade_fde(pred[:, perm], gt) == (ade, fde)
    assert fde <= float(torch.linalg.norm(pred[:, 0, -1] - gt[:, -1], dim=-1).mean()) and ade <= float(torch.linalg.norm(pred[:, 0] - gt, dim=-1).mean())
    pred[:, 7] = gt
    assert O.min_ade_fde(pred, gt) == (0.0, 0.0)
    s = 3.5                                                                        # test.py:187-188 scales prediction and ground truth by traj_scale: the metrics scale with it
    a2, f2 = O.min_ade_fde(s * torch.randn(4, 20, 12, 2, generator=torch.Generator().manual_seed(2), dtype=torch.float64),
                           s * torch.zeros(4, 12, 2, dtype=torch.float64))
    a1, f1 = O.min_ade_fde(torch.randn(4, 20, 12, 2, generator=torch.Generator().manual_seed(2), dtype=torch.float64), torch.zeros(4, 12, 2, dtype=torch.float64))
    assert abs(a2 - s * a1) < 1e-12 and abs(f2 - s * f1) < 1e-12
