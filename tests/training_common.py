"""Shared by the CPU oracle pin and the GPU parity test of one training step (tests/golden/training_step.npz)."""
import math

import torch

from sttode_b200 import synth

CASES = {"nba": ("nba", 10, 20, 8, 11), "eth": ("eth", 8, 12, 1, 7)}       # dataset, T_p, T_f, B, N  (oracle/make_golden.py)


def training_draws(R, tp, tf, seed):
    """Every random draw of one training step; identical to oracle/make_golden.py:training_draws."""
    g = torch.Generator().manual_seed(seed)
    return dict(mask_past=(torch.rand(R, tp, 64, generator=g) >= 0.1).float(), mask_future=(torch.rand(R, tf, 64, generator=g) >= 0.1).float(),
                eps_q=torch.randn(R, 32, generator=g), eps_mid=torch.randn(R, 32, generator=g), eps_p=torch.randn(R * 20, 32, generator=g),
                theta_u=torch.rand(1, generator=g))


def theta_of(d):
    return d["theta_u"] * math.pi * 2            # model/STTODE.py:423: torch.rand(1) * np.pi * 2


def grad_sample_index(numel, count=512):
    return torch.linspace(0, numel - 1, min(numel, count)).round().long()


def case_inputs(name):
    dataset, tp, tf, B, N = CASES[name]
    d = training_draws(B * N, tp, tf, seed=40 + B)
    if dataset == "nba":
        return d, synth.nba_batch(2, B, N, tp, tf)
    return d, synth.eth_scene(3, N)


def grad_errors(got, ref, ref_max=None):
    """(L2-relative, max-norm-relative) error of a gradient tensor (float64 CPU tensors)."""
    d = got - ref
    return float(d.norm() / ref.norm().clamp_min(1e-30)), float(d.abs().max() / (ref.abs().max().clamp_min(1e-30) if ref_max is None else max(ref_max, 1e-30)))


def assert_grad_close(got, ref, tol, what, ref_max=None):
    """Gradient tolerance.  The loss is piecewise smooth: every ReLU (4.7 M of them per NBA step at B = 8), the best-of-20 min,
    the KL clamp and the acos clamp is a discrete decision, and a pre-activation within fp32 rounding of its threshold is decided
    differently by two correct fp32 evaluations (the CPU fp32 oracle shows the same against float64).
    One such flip moves one row of one weight gradient by a few 1e-4 of the tensor's max.  So: L2-relative error <= tol (a flip
    barely registers), max-norm error <= 10 * tol."""
    l2, mx = grad_errors(got, ref, ref_max)
    assert l2 < tol and mx < 10 * tol, (what, l2, mx)
    return l2


def check_against_golden(g, name, losses, named_grads, tol):
    """losses: (total, pred, recover, kl, diverse); named_grads: iterable of (param name, grad tensor or None)."""
    ref = g[f"{name}_losses"]
    for got, want, what in zip(losses, ref, ("total", "pred", "recover", "kl", "diverse")):
        got = float(got.detach()) if torch.is_tensor(got) else float(got)
        assert abs(got - want) <= tol * abs(want), (what, got, want)
    worst = 0.0
    for (pname, grad), has in zip(named_grads, g[f"{name}_has_grad"]):
        norm, gmax = g[f"{name}_gstat_{pname}"]
        if not has:
            assert grad is None or float(grad.abs().max()) == 0.0, pname       # pz_layer: no gradient (learn_prior=False)
            continue
        assert grad is not None, pname
        gr = grad.detach().flatten().cpu().double()
        want = torch.from_numpy(g[f"{name}_g_{pname}"]).double()
        worst = max(worst, assert_grad_close(gr[grad_sample_index(gr.numel())], want, tol, pname, ref_max=gmax))
        assert abs(float(gr.norm()) - norm) <= 4 * tol * max(norm, 1e-30), pname
    return worst
