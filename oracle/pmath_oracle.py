"""CPU oracle for the Poincare-ball math library (rows P1-P8 of SURVEY.md §8a).  TEST INFRASTRUCTURE.

A forward-only PyTorch-CPU restatement of hyptorch/pmath.py (no autograd Functions), dtype-generic.
Pinned against the reference's own `hyptorch.pmath` outputs by `oracle/make_golden.py` ->
`tests/golden/pmath_*.npz` (the reference has no tests for these; pmath.py marks several
`# pragma: no cover`).  Same import rule as sttode_oracle.py: never imported by `sttode_b200/`.
"""
import torch

MIN_NORM = 1e-5          # clamp_min on norms (pmath.py:99,270,302,367,400)
BALL_EPS = 1e-3          # project(): maxnorm = (1 - 1e-3)/sqrt(c)  (pmath.py:100)
ARTANH_EPS = 1e-5        # Artanh clamp (pmath.py:19)
DENOM_EPS = 1e-5         # mobius_add denominator (pmath.py:177,426)


def _c(c, x):
    return torch.as_tensor(c).type_as(x)


def tanh(x, clamp=15):
    """pmath.py:11-12"""
    return x.clamp(-clamp, clamp).tanh()


def artanh(x):
    """pmath.py:16-22 forward: clamp to +-(1-1e-5) then 0.5*(log(1+x) - log(1-x))."""
    x = x.clamp(-1 + ARTANH_EPS, 1 - ARTANH_EPS)
    return (torch.log(1 + x) - torch.log(1 - x)) * 0.5


def arsinh(x):
    """pmath.py:51-55"""
    return (x + torch.sqrt(1 + x.pow(2))).clamp_min(MIN_NORM).log()


def project(x, c=1.0):
    """pmath.py:76-103"""
    c = _c(c, x)
    norm = torch.clamp_min(x.norm(dim=-1, keepdim=True, p=2), MIN_NORM)
    maxnorm = (1 - BALL_EPS) / (c ** 0.5)
    return torch.where(norm > maxnorm, x / norm * maxnorm, x)


def lambda_x(x, c=1.0, keepdim=False):
    """pmath.py:106-129"""
    c = _c(c, x)
    return 2 / (1 - c * x.pow(2).sum(-1, keepdim=keepdim))


def mobius_add(x, y, c=1.0):
    """pmath.py:132-177"""
    c = _c(c, x)
    x2 = x.pow(2).sum(dim=-1, keepdim=True)
    y2 = y.pow(2).sum(dim=-1, keepdim=True)
    xy = (x * y).sum(dim=-1, keepdim=True)
    num = (1 + 2 * c * xy + c * y2) * x + (1 - c * x2) * y
    denom = 1 + 2 * c * xy + c ** 2 * x2 * y2
    return num / (denom + DENOM_EPS)


def dist(x, y, c=1.0, keepdim=False):
    """pmath.py:180-208"""
    c = _c(c, x)
    sqrt_c = c ** 0.5
    return artanh(sqrt_c * mobius_add(-x, y, c).norm(dim=-1, p=2, keepdim=keepdim)) * 2 / sqrt_c


def dist0(x, c=1.0, keepdim=False):
    """pmath.py:211-234"""
    c = _c(c, x)
    sqrt_c = c ** 0.5
    return artanh(sqrt_c * x.norm(dim=-1, p=2, keepdim=keepdim)) * 2 / sqrt_c


def expmap0(u, c=1.0):
    """pmath.py:280-304"""
    c = _c(c, u)
    sqrt_c = c ** 0.5
    u_norm = torch.clamp_min(u.norm(dim=-1, p=2, keepdim=True), MIN_NORM)
    return tanh(sqrt_c * u_norm) * u / (sqrt_c * u_norm)


def logmap0(y, c=1.0):
    """pmath.py:342-368"""
    c = _c(c, y)
    sqrt_c = c ** 0.5
    y_norm = torch.clamp_min(y.norm(dim=-1, p=2, keepdim=True), MIN_NORM)
    return y / y_norm / sqrt_c * artanh(sqrt_c * y_norm)


def expmap(x, u, c=1.0):
    """pmath.py:237-277"""
    c = _c(c, x)
    sqrt_c = c ** 0.5
    u_norm = torch.clamp_min(u.norm(dim=-1, p=2, keepdim=True), MIN_NORM)
    second = tanh(sqrt_c / 2 * lambda_x(x, c, keepdim=True) * u_norm) * u / (sqrt_c * u_norm)
    return mobius_add(x, second, c)


def logmap(x, y, c=1.0):
    """pmath.py:307-339"""
    c = _c(c, x)
    sub = mobius_add(-x, y, c)
    sub_norm = sub.norm(dim=-1, p=2, keepdim=True)
    lam = lambda_x(x, c, keepdim=True)
    sqrt_c = c ** 0.5
    return 2 / sqrt_c / lam * artanh(sqrt_c * sub_norm) * sub / sub_norm


def mobius_matvec(m, x, c=1.0):
    """pmath.py:371-408"""
    c = _c(c, x)
    x_norm = torch.clamp_min(x.norm(dim=-1, keepdim=True, p=2), MIN_NORM)
    sqrt_c = c ** 0.5
    mx = x @ m.transpose(-1, -2)
    mx_norm = mx.norm(dim=-1, keepdim=True, p=2)
    res_c = tanh(mx_norm / x_norm * artanh(sqrt_c * x_norm)) * mx / (mx_norm * sqrt_c)
    cond = (mx == 0).all(dim=-1, keepdim=True)
    res = torch.where(cond, torch.zeros(1, dtype=res_c.dtype), res_c)
    return project(res, c)


def mobius_addition_batch(x, y, c=1.0):
    """pmath.py:416-427 — all-pairs x_i (+) y_j, materialised [B,C,D]."""
    c = _c(c, x)
    xy = x @ y.t()
    x2 = x.pow(2).sum(-1, keepdim=True)
    y2 = y.pow(2).sum(-1, keepdim=True)
    num = (1 + 2 * c * xy + c * y2.permute(1, 0)).unsqueeze(2) * x.unsqueeze(1)
    num = num + (1 - c * x2).unsqueeze(2) * y
    denom = 1 + 2 * c * xy + c ** 2 * x2 * y2.permute(1, 0)
    return num / (denom.unsqueeze(2) + DENOM_EPS)


def dist_matrix(x, y, c=1.0):
    """pmath.py:482-493"""
    c = _c(c, x)
    sqrt_c = c ** 0.5
    return 2 / sqrt_c * artanh(sqrt_c * torch.norm(mobius_addition_batch(-x, y, c), dim=-1))


def dist_matrix_closed_form(x, y, c=1.0):
    """The Gram-epilogue closed form the CUDA kernel uses, restated on CPU so the
    algebra itself is checked against `dist_matrix` in float64:
        g=<x,y>  a=1-2cg+c*y2  b=1-c*x2  |num|^2=a^2*x2-2ab*g+b^2*y2  den=1-2cg+c^2*x2*y2+1e-5
    """
    c = _c(c, x)
    g = x @ y.t()
    x2 = x.pow(2).sum(-1, keepdim=True)
    y2 = y.pow(2).sum(-1, keepdim=True).t()
    a = 1 - 2 * c * g + c * y2
    b = 1 - c * x2
    n2 = (a * a * x2 - 2 * a * b * g + b * b * y2).clamp_min(0)
    den = 1 - 2 * c * g + c * c * x2 * y2 + DENOM_EPS
    sqrt_c = c ** 0.5
    return 2 / sqrt_c * artanh(sqrt_c * torch.sqrt(n2) / den)


def p2k(x, c=1.0):
    """pmath.py:440-442"""
    return 2 * x / (1 + c * x.pow(2).sum(-1, keepdim=True))


def k2p(x, c=1.0):
    """pmath.py:445-447"""
    return x / (1 + torch.sqrt(1 - c * x.pow(2).sum(-1, keepdim=True)))


def poincare_mean(x, dim=0, c=1.0):
    """pmath.py:472-479"""
    xk = p2k(x, c)
    lamb = 1 / torch.sqrt(1 - c * xk.pow(2).sum(dim=-1, keepdim=True))
    mean = torch.sum(lamb * xk, dim=dim, keepdim=True) / torch.sum(lamb, dim=dim, keepdim=True)
    return k2p(mean, c).squeeze(dim)
