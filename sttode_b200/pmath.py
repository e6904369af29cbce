"""Poincare-ball math on B200: drop-in for the reference's `hyptorch/pmath.py` (forward).

Same function names and keyword-only curvature `c`; each call is one kernel of csrc/pmath.cu through the
C ABI.  The reference never calls these on a live path (SURVEY.md finding 3); they are provided as the
parity-checked library `north_star` names plus the optional Poincare score mode of the MHGSA kernel.
float32 CUDA tensors only.  `artanh` / `arsinh` carry the reference's custom autograd rules (`Artanh`, `Arsinh`, pmath.py:16-27,51-60) and
`RiemannianGradient` (:30-45) is provided.  Every other function is differentiable like the reference's torch code: the forward is the CUDA
kernel; the backward re-evaluates the same formula with device-side torch ops (clamps and the Artanh / Arsinh gradient rules included) and
differentiates that (`_Diff`) — these functions are dead in the reference, so their backward is not a hot path.  `_mobius_addition_batch`,
`_hyperbolic_softmax`, `p2k`, `k2p`, `lorenz_factor`, `poincare_mean` (pmath.py:416-480) are device-side compositions.
"""
import numpy as np
import torch

from . import _lib
from .mhgsa import _f32c


def _rows(x):
    d = x.shape[-1]
    return x.numel() // d, d


def _row_op(fn_name, x, y=None, *, c, out_vec=True, keepdim=False):
    x = _f32c(x, "x")
    rows, d = _rows(x)
    if y is not None:
        y = _f32c(y, "y")
        if y.shape != x.shape:
            x, y = torch.broadcast_tensors(x, y)
            x, y = x.contiguous(), y.contiguous()
            rows, d = _rows(x)
    out = torch.empty_like(x) if out_vec else torch.empty(x.shape[:-1] + ((1,) if keepdim else ()), device=x.device, dtype=torch.float32)
    fn = getattr(_lib.lib(), fn_name)
    args = [_lib.dptr(x)] + ([_lib.dptr(y)] if y is not None else []) + [_lib.dptr(out), rows, d, float(c), _lib.stream_ptr()]
    _lib.check(fn(*args))
    return out


def _elem(fn_name, x, *extra):
    x = _f32c(x, "x")
    out = torch.empty_like(x)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.dptr(x), _lib.dptr(out), x.numel(), *extra, _lib.stream_ptr()))
    return out


def tanh(x, clamp=15):
    return _elem("sttode_pmath_tanh", x, float(clamp))


class Artanh(torch.autograd.Function):
    """artanh(clamp(x, +-(1 - 1e-5))) with the reference's backward g / (1 - x_clamped^2) (pmath.py:16-27): the gradient passes through the
    clamp, unlike torch.clamp's own rule."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x.detach().clamp(-1 + 1e-5, 1 - 1e-5))
        return _elem("sttode_pmath_artanh", x)

    @staticmethod
    def backward(ctx, grad_output):
        (xc,) = ctx.saved_tensors
        return grad_output / (1 - xc ** 2)


class Arsinh(torch.autograd.Function):
    """log(clamp_min(x + sqrt(1 + x^2), 1e-5)) with backward g / sqrt(1 + x^2) (pmath.py:51-60)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x.detach())
        return _elem("sttode_pmath_arsinh", x)

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        return grad_output / (1 + x ** 2) ** 0.5


class RiemannianGradient(torch.autograd.Function):
    """Identity whose backward rescales the Euclidean gradient to the Riemannian one: g * (1 - c |x|^2)^2 / 4 (pmath.py:30-45)."""
    c = 1

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        return grad_output * ((1 - RiemannianGradient.c * x.pow(2).sum(-1, keepdim=True)).pow(2) / 4)


def artanh(x):
    return Artanh.apply(x)


def arsinh(x):
    return Arsinh.apply(x)


# ---- differentiable device-side restatements used by the backward passes (torch ops on the GPU; same clamps, same custom gradient rules) -------
class _ArtanhT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = x.clamp(-1 + 1e-5, 1 - 1e-5)
        ctx.save_for_backward(xc)
        return 0.5 * (torch.log1p(xc) - torch.log1p(-xc))

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        return g / (1 - xc * xc)


def _norm(v):
    return v.norm(dim=-1, p=2, keepdim=True)


def _t_tanh(v):
    return v.clamp(-15, 15).tanh()


def _t_project(x, c):
    n = _norm(x).clamp_min(1e-5)
    mx = (1 - 1e-3) / c ** 0.5
    return torch.where(n > mx, x / n * mx, x)


def _t_lambda(x, c):
    return 2 / (1 - c * (x * x).sum(-1, keepdim=True))


def _t_madd(x, y, c):
    x2, y2, xy = (x * x).sum(-1, keepdim=True), (y * y).sum(-1, keepdim=True), (x * y).sum(-1, keepdim=True)
    return ((1 + 2 * c * xy + c * y2) * x + (1 - c * x2) * y) / (1 + 2 * c * xy + c * c * x2 * y2 + 1e-5)


def _t_dist(x, y, c):
    return _ArtanhT.apply(c ** 0.5 * _norm(_t_madd(-x, y, c))) * (2 / c ** 0.5)


def _t_dist0(x, c):
    return _ArtanhT.apply(c ** 0.5 * _norm(x)) * (2 / c ** 0.5)


def _t_expmap0(u, c):
    n = _norm(u).clamp_min(1e-5)
    return _t_tanh(c ** 0.5 * n) * u / (c ** 0.5 * n)


def _t_logmap0(y, c):
    n = _norm(y).clamp_min(1e-5)
    return y / n / c ** 0.5 * _ArtanhT.apply(c ** 0.5 * n)


def _t_expmap(x, u, c):
    n = _norm(u).clamp_min(1e-5)
    return _t_madd(x, _t_tanh(c ** 0.5 / 2 * _t_lambda(x, c) * n) * u / (c ** 0.5 * n), c)


def _t_logmap(x, y, c):
    sub = _t_madd(-x, y, c)
    n = _norm(sub)
    return 2 / c ** 0.5 / _t_lambda(x, c) * _ArtanhT.apply(c ** 0.5 * n) * sub / n


def _t_matvec(m, x, c):
    xn = _norm(x).clamp_min(1e-5)
    mx = x @ m.transpose(-1, -2)
    mn = _norm(mx)
    res = _t_tanh(mn / xn * _ArtanhT.apply(c ** 0.5 * xn)) * mx / (mn * c ** 0.5)
    zero = (mx == 0).all(dim=-1, keepdim=True)
    return _t_project(torch.where(zero, torch.zeros_like(res), res), c)


def _mobius_addition_batch(x, y, c):
    """All-pairs (x_i) (+)_c (y_j) -> [B, C, D] (pmath.py:416-427), device-side torch ops (differentiable)."""
    xy = x @ y.t()
    x2, y2 = (x * x).sum(-1, keepdim=True), (y * y).sum(-1, keepdim=True).t()
    num = (1 + 2 * c * xy + c * y2).unsqueeze(2) * x.unsqueeze(1) + (1 - c * x2).unsqueeze(2) * y
    return num / ((1 + 2 * c * xy + c * c * x2 * y2).unsqueeze(2) + 1e-5)


def _t_dist_matrix(x, y, c):
    return 2 / c ** 0.5 * _ArtanhT.apply(c ** 0.5 * _mobius_addition_batch(-x, y, c).norm(dim=-1))


class _Diff(torch.autograd.Function):
    """forward: the CUDA kernel `fwd(*tensors)`; backward: d/d(tensors) of the device-side torch restatement `ref(*tensors)` of the same formula."""

    @staticmethod
    def forward(ctx, fwd, ref, *tensors):
        ctx.ref = ref
        ctx.save_for_backward(*tensors)
        return fwd(*tensors)

    @staticmethod
    def backward(ctx, g):
        with torch.enable_grad():
            ins = [t.detach().requires_grad_(True) for t in ctx.saved_tensors]
            out = ctx.ref(*ins)
            need = [i for i, n in enumerate(ctx.needs_input_grad[2:]) if n]
            grads = torch.autograd.grad(out, [ins[i] for i in need], g.reshape(out.shape).to(out.dtype), allow_unused=True)
        full = [None] * len(ins)
        for i, gr in zip(need, grads):
            full[i] = gr
        return (None, None, *full)


def _diff(fwd, ref, *tensors):
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        return _Diff.apply(fwd, ref, *tensors)
    return fwd(*tensors)


def _bc(x, y):
    if x.shape != y.shape:
        x, y = torch.broadcast_tensors(x, y)
    return x, y


def project(x, *, c=1.0):
    return _diff(lambda a: _row_op("sttode_pmath_project", a, c=c), lambda a: _t_project(a, c), x)


def lambda_x(x, *, c=1.0, keepdim=False):
    return _diff(lambda a: _row_op("sttode_pmath_lambda_x", a, c=c, out_vec=False, keepdim=keepdim),
                 lambda a: _t_lambda(a, c) if keepdim else _t_lambda(a, c).squeeze(-1), x)


def mobius_add(x, y, *, c=1.0):
    x, y = _bc(x, y)
    return _diff(lambda a, b: _row_op("sttode_pmath_mobius_add", a, b, c=c), lambda a, b: _t_madd(a, b, c), x, y)


def dist(x, y, *, c=1.0, keepdim=False):
    x, y = _bc(x, y)
    return _diff(lambda a, b: _row_op("sttode_pmath_dist", a, b, c=c, out_vec=False, keepdim=keepdim),
                 lambda a, b: _t_dist(a, b, c) if keepdim else _t_dist(a, b, c).squeeze(-1), x, y)


def dist0(x, *, c=1.0, keepdim=False):
    return _diff(lambda a: _row_op("sttode_pmath_dist0", a, c=c, out_vec=False, keepdim=keepdim),
                 lambda a: _t_dist0(a, c) if keepdim else _t_dist0(a, c).squeeze(-1), x)


def expmap0(u, *, c=1.0):
    return _diff(lambda a: _row_op("sttode_pmath_expmap0", a, c=c), lambda a: _t_expmap0(a, c), u)


def logmap0(y, *, c=1.0):
    return _diff(lambda a: _row_op("sttode_pmath_logmap0", a, c=c), lambda a: _t_logmap0(a, c), y)


def expmap(x, u, *, c=1.0):
    x, u = _bc(x, u)
    return _diff(lambda a, b: _row_op("sttode_pmath_expmap", a, b, c=c), lambda a, b: _t_expmap(a, b, c), x, u)


def logmap(x, y, *, c=1.0):
    x, y = _bc(x, y)
    return _diff(lambda a, b: _row_op("sttode_pmath_logmap", a, b, c=c), lambda a, b: _t_logmap(a, b, c), x, y)


def _matvec_fwd(m, x, c):
    m, x = _f32c(m, "m"), _f32c(x, "x")
    if m.dim() != 2 or m.shape[1] != x.shape[-1]:
        raise _lib.SttodeError("mobius_matvec: m must be [d_out, d]")
    rows, d = _rows(x)
    out = torch.empty(x.shape[:-1] + (m.shape[0],), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sttode_pmath_mobius_matvec(_lib.dptr(m), _lib.dptr(x), _lib.dptr(out), rows, d, m.shape[0], float(c), _lib.stream_ptr()))
    return out


def mobius_matvec(m, x, *, c=1.0):
    return _diff(lambda a, b: _matvec_fwd(a, b, c), lambda a, b: _t_matvec(a, b, c), m, x)


def _dist_matrix_fwd(x, y, c):
    x, y = _f32c(x, "x"), _f32c(y, "y")
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise _lib.SttodeError("dist_matrix: x [b,d], y [n,d]")
    out = torch.empty(x.shape[0], y.shape[0], device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sttode_pmath_dist_matrix(_lib.dptr(x), _lib.dptr(y), _lib.dptr(out), x.shape[0], y.shape[0], x.shape[1], float(c), _lib.stream_ptr()))
    return out


def dist_matrix(x, y, c=1.0):
    return _diff(lambda a, b: _dist_matrix_fwd(a, b, c), lambda a, b: _t_dist_matrix(a, b, c), x, y)


# ---- the remaining helpers of pmath.py:430-480: compositions of device-side torch ops and the functions above (differentiable) -------------------
def _hyperbolic_softmax(X, A, P, c):
    """Hyperbolic MLR logits (pmath.py:430-437): X [B,D] points, A [C,D] normals, P [C,D] offsets -> [B, C]."""
    c = torch.as_tensor(c, dtype=X.dtype, device=X.device)
    lam = 2 / (1 - c * (P * P).sum(dim=1))
    k = lam * A.norm(dim=1) / torch.sqrt(c)
    ma = _mobius_addition_batch(-P, X, c)
    num = 2 * torch.sqrt(c) * (ma * A.unsqueeze(1)).sum(dim=-1)
    den = A.norm(dim=1, keepdim=True) * (1 - c * (ma * ma).sum(dim=2))
    return (k.unsqueeze(1) * arsinh(num / den)).permute(1, 0)


def p2k(x, c):
    """Poincare -> Klein (pmath.py:440-442)."""
    return 2 * x / (1 + c * (x * x).sum(-1, keepdim=True))


def k2p(x, c):
    """Klein -> Poincare (pmath.py:445-447)."""
    return x / (1 + torch.sqrt(1 - c * (x * x).sum(-1, keepdim=True)))


def lorenz_factor(x, *, c=1.0, dim=-1, keepdim=False):
    return 1 / torch.sqrt(1 - c * (x * x).sum(dim=dim, keepdim=keepdim))


def poincare_mean(x, dim=0, c=1.0):
    """Einstein midpoint in the Klein model mapped back to the ball (pmath.py:472-480)."""
    xk = p2k(x, c)
    lam = lorenz_factor(xk, c=c, keepdim=True)
    return k2p((lam * xk).sum(dim=dim, keepdim=True) / lam.sum(dim=dim, keepdim=True), c).squeeze(dim)


def auto_select_c(d):
    """Host-side scalar helper (pmath.py:496-505): ball radius such that the d-ball has volume pi."""
    import math
    dim2 = d / 2.0
    R = math.gamma(dim2 + 1) / (np.pi ** (dim2 - 1))
    R = R ** (1 / float(d))
    return 1 / (R ** 2)
