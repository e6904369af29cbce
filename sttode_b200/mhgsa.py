"""Multi-Head Geodesic Self-Attention on B200: drop-in for the reference's `hyptransformerlib.py`
(`linear`, `hyp_mhsa`, `Hyp_mhsa`) and `core/manifolds/oblique.py` (`Oblique.proj/dist/inner`).

Same names, argument meaning, state_dict keys (`in_proj_weight`, `in_proj_bias`, `out_proj.weight`,
`out_proj.bias`) and error behaviour; the arithmetic runs in libsttode_sm100.so (csrc/encoder.cu:
k_proj -> k_geo_attn -> k_proj).  Options the reference accepts but whose code path is dead or
commented out there (masks, bias_kv, zero_attn, dropout > 0 in training) raise instead of silently
differing.  Differentiable: forward and backward are one autograd node over sttode_mhgsa_fwd / sttode_mhgsa_bwd.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib

SCORE_OBLIQUE_ACOS = 0
SCORE_POINCARE_DIST = 1
OBLIQUE_EPS = {torch.float32: 1e-4, torch.float64: 1e-7}       # core/manifolds/oblique.py:7


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise _lib.SttodeError(f"{name}: only float32 is implemented on the B200 path (got {t.dtype})")
    return t.detach().contiguous()


def linear(input, weight, bias=None):
    """y = x A^T + b (hyptransformerlib.py:7-26).  Runs as one k_proj launch per 64 output features when
    in_features == 64 and out_features % 64 == 0 (every call site on the path); otherwise raises."""
    x = _f32c(input, "input")
    out_f, in_f = weight.shape
    if in_f != 64 or out_f % 64 != 0 or bias is None:
        raise _lib.SttodeError("linear(): the B200 path implements the 64 -> 64k projections of the MHGSA block only")
    rows = x.numel() // 64
    out = torch.empty(*x.shape[:-1], out_f, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sttode_linear64(_lib.dptr(x), _lib.dptr(_f32c(weight, "weight")), _lib.dptr(_f32c(bias, "bias")),
                                          _lib.dptr(out), rows, out_f, _lib.stream_ptr()))
    return out


class Oblique:
    """Oblique manifold ops used by the geodesic score (core/manifolds/oblique.py:10-45)."""
    name = "Oblique"

    def proj(self, p):
        x = _f32c(p, "p")
        out = torch.empty_like(x)
        d = x.shape[-1]
        _lib.check(_lib.lib().sttode_oblique_proj(_lib.dptr(x), _lib.dptr(out), x.numel() // d, d, _lib.stream_ptr()))
        return out

    def inner(self, p, u, v=None, keepdim=False):
        # the reference's `inner` is a plain batched Gram (oblique.py:40-45); kept for API parity
        v = u if v is None else v
        return u @ v.transpose(-2, -1)

    def dist(self, p1, p2, keepdim=False):
        """acos(clamp(p2 @ p1^T, -1+eps, 1-eps)) — note the argument order (oblique.py:36-38)."""
        a, b = _f32c(p1, "p1"), _f32c(p2, "p2")
        if a.shape[:-2] != b.shape[:-2] or a.shape[-1] != b.shape[-1]:
            raise _lib.SttodeError("Oblique.dist: batch dims and feature dim of p1, p2 must match")
        n1, n2, d = a.shape[-2], b.shape[-2], a.shape[-1]
        batch = a.numel() // (n1 * d)
        out = torch.empty(*a.shape[:-2], n2, n1, device=a.device, dtype=torch.float32)
        _lib.check(_lib.lib().sttode_oblique_dist(_lib.dptr(a), _lib.dptr(b), _lib.dptr(out), batch, n1, n2, d,
                                                  OBLIQUE_EPS[torch.float32], _lib.stream_ptr()))
        return out


def _mhgsa_struct(in_w, in_b, out_w, out_b):
    s = _lib.MhgsaWeights()
    s.in_proj_w, s.in_proj_b = _lib.dptr(in_w, "in_proj_weight"), _lib.dptr(in_b, "in_proj_bias")
    s.out_proj_w, s.out_proj_b = _lib.dptr(out_w, "out_proj.weight"), _lib.dptr(out_b, "out_proj.bias")
    return s


def hyp_mhsa(query, key, value, embed_dim_to_check, num_heads, in_proj_weight, in_proj_bias, bias_k, bias_v,
             add_zero_attn, dropout_p, out_proj_weight, out_proj_bias, training=True, key_padding_mask=None,
             need_weights=True, attn_mask=None, use_separate_proj_weight=False, q_proj_weight=None,
             k_proj_weight=None, v_proj_weight=None, static_k=None, static_v=None, sparse_gate=None,
             cross_range=0, interaction_mask=None, seq_mask=False, score_mode=SCORE_OBLIQUE_ACOS, curvature=1.0):
    """Functional MHGSA with the reference's signature (hyptransformerlib.py:29-57) plus `score_mode`.

    query [L,N,E], key/value [S,N,E] -> (out [L,N,E], head-averaged weights [N,L,S] or None).
    L == S reproduces the reference's un-transposed score matrix (SURVEY.md finding 4)."""
    L, bsz, E = query.shape
    if E != embed_dim_to_check:
        raise AssertionError("embed_dim mismatch")                      # hyptransformerlib.py:105
    if key.shape != value.shape:
        raise AssertionError("key and value must have the same shape")  # :106
    unsupported = [n for n, v in (("bias_k", bias_k), ("bias_v", bias_v), ("key_padding_mask", key_padding_mask),
                                  ("attn_mask", attn_mask), ("static_k", static_k), ("static_v", static_v),
                                  ("sparse_gate", sparse_gate)) if v is not None]
    if add_zero_attn:
        unsupported.append("add_zero_attn")
    if use_separate_proj_weight:
        unsupported.append("use_separate_proj_weight")
    if dropout_p > 0.0 and training:
        unsupported.append("dropout>0 in training")
    if unsupported:
        raise NotImplementedError("hyp_mhsa on B200: unsupported option(s) " + ", ".join(unsupported) +
                                  " (dead or commented out on the reference path, hyptransformerlib.py:266-296)")
    if in_proj_bias is None or out_proj_bias is None:
        raise NotImplementedError("hyp_mhsa on B200: bias=False is not implemented")
    S = key.shape[0]
    self_attn = (key is query or (key.data_ptr() == query.data_ptr() and key.shape == query.shape)) and \
                (value is key or (value.data_ptr() == key.data_ptr() and value.shape == key.shape))
    out, w = _MhgsaFn.apply(query, None if self_attn else key, None if self_attn else value, in_proj_weight, in_proj_bias,
                            out_proj_weight, out_proj_bias, num_heads, bool(need_weights), score_mode, float(curvature))
    return out, (w if need_weights else None)


def _clamp_eps():
    return float(np.float32(OBLIQUE_EPS[torch.float32]))


class _MhgsaFn(torch.autograd.Function):
    """sttode_mhgsa_fwd / sttode_mhgsa_bwd as one autograd node (the reference differentiates hyptransformerlib.py:29-311
    with torch autograd).  key/value None = the self-attention branch (:113-115).  The backward re-materialises the forward
    from the inputs, so nothing but the inputs is saved.  The head-averaged weights output is not differentiable here."""

    @staticmethod
    def forward(ctx, query, key, value, in_w, in_b, out_w, out_b, num_heads, need_weights, score_mode, curvature):
        L, bsz, E = query.shape
        q = _f32c(query, "query")
        self_attn = key is None
        k = q if self_attn else _f32c(key, "key")
        v = q if self_attn else _f32c(value, "value")
        S = k.shape[0]
        out = torch.empty(L, bsz, E, device=q.device, dtype=torch.float32)
        w = torch.empty(bsz, L, S, device=q.device, dtype=torch.float32) if need_weights else None
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_mhgsa_workspace_bytes(L, S, bsz), q.device)
        params = [_f32c(t, n) for t, n in ((in_w, "in_proj_weight"), (in_b, "in_proj_bias"), (out_w, "out_proj.weight"), (out_b, "out_proj.bias"))]
        st = _mhgsa_struct(*params)
        _lib.check(lib.sttode_mhgsa_fwd(_lib.dptr(q), _lib.dptr(k), _lib.dptr(v), C.byref(st), _lib.dptr(out), _lib.dptr(w),
                                        L, S, bsz, E, num_heads, score_mode, _clamp_eps(), curvature, _lib.dptr(ws), ws.numel(),
                                        _lib.stream_ptr()))
        ctx.save_for_backward(q, k, v, *params)
        ctx.cfg = (self_attn, num_heads, score_mode)
        if w is None:
            w = out.new_empty(0)
        ctx.mark_non_differentiable(w)
        return out, w

    @staticmethod
    def backward(ctx, dout, _dw):
        q, k, v, in_w, in_b, out_w, out_b = ctx.saved_tensors
        self_attn, num_heads, score_mode = ctx.cfg
        L, bsz, E = q.shape
        S = k.shape[0]
        dout = _f32c(dout, "grad_output")
        dq = torch.empty_like(q)
        dk = None if self_attn else torch.empty_like(k)
        dv = None if self_attn else torch.empty_like(v)
        grads = [torch.zeros_like(t) for t in (in_w, in_b, out_w, out_b)]       # the library accumulates (+=)
        gs = _lib.MhgsaGrads()
        gs.in_proj_w, gs.in_proj_b, gs.out_proj_w, gs.out_proj_b = (_lib.dptr(g) for g in grads)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_mhgsa_bwd_workspace_bytes(L, S, bsz), q.device)
        st = _mhgsa_struct(in_w, in_b, out_w, out_b)
        _lib.check(lib.sttode_mhgsa_bwd(_lib.dptr(q), _lib.dptr(k), _lib.dptr(v), C.byref(st), _lib.dptr(dout), _lib.dptr(dq), _lib.dptr(dk),
                                        _lib.dptr(dv), C.byref(gs), L, S, bsz, E, num_heads, score_mode, _clamp_eps(), _lib.dptr(ws),
                                        ws.numel(), _lib.stream_ptr()))
        return (dq, dk, dv, *grads, None, None, None, None)


class Hyp_mhsa(nn.Module):
    """Module form (hyptransformerlib.py:314-454): same constructor, parameters and forward signature."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None,
                 vdim=None, sparse_gate_class=None, score_mode=SCORE_OBLIQUE_ACOS, curvature=1.0):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.kdim = embed_dim if kdim is None else kdim
        self.vdim = embed_dim if vdim is None else vdim
        self._qkv_same_embed_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        if not self._qkv_same_embed_dim or add_bias_kv or not bias:
            raise NotImplementedError("Hyp_mhsa on B200: kdim/vdim != embed_dim, add_bias_kv and bias=False are not implemented")
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.register_parameter("q_proj_weight", None)
        self.register_parameter("k_proj_weight", None)
        self.register_parameter("v_proj_weight", None)
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        self.sparse_attn_gate = sparse_gate_class
        self.score_mode, self.curvature = score_mode, curvature
        self._reset_parameters()

    def _reset_parameters(self):          # hyptransformerlib.py:379-394
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, cross_range=0,
                interaction_mask=None, seq_mask=False):
        return hyp_mhsa(query, key, value, self.embed_dim, self.num_heads, self.in_proj_weight, self.in_proj_bias,
                        self.bias_k, self.bias_v, self.add_zero_attn, self.dropout, self.out_proj.weight,
                        self.out_proj.bias, training=self.training, key_padding_mask=key_padding_mask,
                        need_weights=need_weights, attn_mask=attn_mask, sparse_gate=self.sparse_attn_gate,
                        cross_range=cross_range, interaction_mask=interaction_mask, seq_mask=seq_mask,
                        score_mode=self.score_mode, curvature=self.curvature)
