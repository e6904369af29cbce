"""Poincare-ball math on B200: drop-in for the reference's `hyptorch/pmath.py` (forward).

Same function names and keyword-only curvature `c`; each call is one kernel of csrc/pmath.cu through the
C ABI.  The reference never calls these on a live path (SURVEY.md finding 3); they are provided as the
parity-checked library `north_star` names plus the optional Poincare score mode of the MHGSA kernel.
float32 CUDA tensors only.  `artanh` / `arsinh` carry the reference's custom autograd rules (`Artanh`, `Arsinh`, pmath.py:16-27,51-60) and
`RiemannianGradient` (:30-45) is provided; the other functions are forward-only.
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


def project(x, *, c=1.0):
    return _row_op("sttode_pmath_project", x, c=c)


def lambda_x(x, *, c=1.0, keepdim=False):
    return _row_op("sttode_pmath_lambda_x", x, c=c, out_vec=False, keepdim=keepdim)


def mobius_add(x, y, *, c=1.0):
    return _row_op("sttode_pmath_mobius_add", x, y, c=c)


def dist(x, y, *, c=1.0, keepdim=False):
    return _row_op("sttode_pmath_dist", x, y, c=c, out_vec=False, keepdim=keepdim)


def dist0(x, *, c=1.0, keepdim=False):
    return _row_op("sttode_pmath_dist0", x, c=c, out_vec=False, keepdim=keepdim)


def expmap0(u, *, c=1.0):
    return _row_op("sttode_pmath_expmap0", u, c=c)


def logmap0(y, *, c=1.0):
    return _row_op("sttode_pmath_logmap0", y, c=c)


def expmap(x, u, *, c=1.0):
    return _row_op("sttode_pmath_expmap", x, u, c=c)


def logmap(x, y, *, c=1.0):
    return _row_op("sttode_pmath_logmap", x, y, c=c)


def mobius_matvec(m, x, *, c=1.0):
    m, x = _f32c(m, "m"), _f32c(x, "x")
    if m.dim() != 2 or m.shape[1] != x.shape[-1]:
        raise _lib.SttodeError("mobius_matvec: m must be [d_out, d]")
    rows, d = _rows(x)
    out = torch.empty(x.shape[:-1] + (m.shape[0],), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sttode_pmath_mobius_matvec(_lib.dptr(m), _lib.dptr(x), _lib.dptr(out), rows, d, m.shape[0], float(c), _lib.stream_ptr()))
    return out


def dist_matrix(x, y, c=1.0):
    x, y = _f32c(x, "x"), _f32c(y, "y")
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise _lib.SttodeError("dist_matrix: x [b,d], y [n,d]")
    out = torch.empty(x.shape[0], y.shape[0], device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sttode_pmath_dist_matrix(_lib.dptr(x), _lib.dptr(y), _lib.dptr(out), x.shape[0], y.shape[0], x.shape[1], float(c), _lib.stream_ptr()))
    return out


def auto_select_c(d):
    """Host-side scalar helper (pmath.py:496-505): ball radius such that the d-ball has volume pi."""
    import math
    dim2 = d / 2.0
    R = math.gamma(dim2 + 1) / (np.pi ** (dim2 - 1))
    R = R ** (1 / float(d))
    return 1 / (R ** 2)
