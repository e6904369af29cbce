"""Geodesic transformer layer and tensor-ODE block on B200: drop-ins for the reference's
`hypertransformer.py` (`Hypattention`, `TransformerEncoderLayer`) and `ode_demo.py`
(`TransformerEncoder_ode`, `ODEblock_en`, `ODEG_Encoder`).

Same class names, constructor signatures, attribute side effects and state_dict keys
(`self_attn.temporal_attention_before.*`, `self_attn.temporal_info.*`, `self_attn.temporal_gate.*`,
`linear1/2.*`, `norm1/2.*`; prefix `odeblock.odefunc.layers.0.`).  One `ODEG_Encoder.forward` is one
call of `sttode_ode_block_fwd` (csrc/encoder.cu): the torchdiffeq solver loop, the right-hand side and the
final ReLU all run in the library; each forward is one autograd node whose backward is the library's
`sttode_ode_rhs_bwd` / `sttode_ode_block_bwd`.
"""
import copy
import ctypes as C

import torch
from torch import nn

from . import _lib
from .mhgsa import Hyp_mhsa, _f32c, _mhgsa_struct

ODE_EULER, ODE_MIDPOINT, ODE_RK4, ODE_RK4_38 = 0, 1, 2, 3
_METHODS = {"euler": ODE_EULER, "midpoint": ODE_MIDPOINT, "rk4_classic": ODE_RK4, "rk4": ODE_RK4_38}


_ODE_FIELDS = ("info_w", "info_b", "gate_w", "gate_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "norm1_g", "norm1_b", "norm2_g", "norm2_b")


def _ode_params(layer):
    """The 16 tensors of sttode_ode_weights in struct order (only the first 8 exist for a bare Hypattention)."""
    att = layer.self_attn if isinstance(layer, TransformerEncoderLayer) else layer
    mh = att.temporal_attention_before
    ps = [mh.in_proj_weight, mh.in_proj_bias, mh.out_proj.weight, mh.out_proj.bias,
          att.temporal_info.weight, att.temporal_info.bias, att.temporal_gate.weight, att.temporal_gate.bias]
    if isinstance(layer, TransformerEncoderLayer):
        if layer.linear1.weight.shape != (1024, 64):
            raise NotImplementedError("TransformerEncoderLayer on B200: d_model=64, dim_feedforward=1024 only (model/STTODE.py:190-194)")
        ps += [layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias,
               layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias]
    return ps


def _ode_struct_from(tensors, cls=_lib.OdeWeights):
    """sttode_ode_weights (or sttode_ode_grads) over a list laid out like `_ode_params`."""
    s = cls()
    s.attn.in_proj_w, s.attn.in_proj_b, s.attn.out_proj_w, s.attn.out_proj_b = (_lib.dptr(t) for t in tensors[:4])
    for name, t in zip(_ODE_FIELDS, tensors[4:]):
        setattr(s, name, _lib.dptr(t))
    return s


def _ode_struct(layer, tensor_cores=None):
    """sttode_ode_weights for a TransformerEncoderLayer (or just its Hypattention when `layer` is one).  With tensor_cores (default: the
    layer is in eval mode and `use_tensor_cores` is set) the struct carries the fp16 hi/lo operand images of the weights and the forward
    runs on the tcgen05 kernels."""
    s = _ode_struct_from(_ode_params(layer))
    if tensor_cores is None:
        tensor_cores = isinstance(layer, TransformerEncoderLayer) and layer.use_tensor_cores and not layer.training
    if tensor_cores:
        s.packed = _lib.dptr(layer._packed.get(layer))
    return s


class _PackedOde:
    """fp16 hi/lo operand images of one ODE block's weights for k_qkv_tc / k_ode_tail_tc (sttode_ode_pack): one persistent buffer,
    rewritten in place when the weights change (same staleness rules as net._PackedDecoder)."""

    def __init__(self):
        self.key, self.buf, self.packs = None, None, 0

    def invalidate(self):
        self.key = None

    def __deepcopy__(self, memo):
        return _PackedOde()

    def get(self, layer):
        params = _ode_params(layer)
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self.key:
            lib = _lib.lib()
            if self.buf is None or self.buf.device != params[0].device:
                self.buf = torch.empty(lib.sttode_ode_pack_bytes(), dtype=torch.uint8, device=params[0].device)
            st = _ode_struct_from(params)
            _lib.check(lib.sttode_ode_pack(C.byref(st), _lib.dptr(self.buf), self.buf.numel(), _lib.stream_ptr()))
            self.key = key
            self.packs += 1
        return self.buf


class _OdeRhsFn(torch.autograd.Function):
    """f(x) = TransformerEncoderLayer.forward (gate_only = 0) or Hypattention.forward (gate_only = 1) with its backward
    (sttode_ode_rhs_fwd / sttode_hypattention_fwd / sttode_ode_rhs_bwd).  x [L, bsz, 64]."""

    @staticmethod
    def forward(ctx, x, gate_only, packed, *params):
        L, bsz = x.shape[0], x.shape[1]
        params_in = params
        params = [_f32c(p, "parameter") for p in params]
        out = torch.empty_like(x)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_ode_block_workspace_bytes(L, bsz, ODE_EULER), x.device)
        st = _ode_struct_from(params)
        st.packed = _lib.dptr(packed)               # tensor-core operand images (eval mode) or None -> FP32 kernels
        ctx.params_in, ctx.sinks = params_in, _lib.grad_sinks(params_in)
        fn = lib.sttode_hypattention_fwd if gate_only else lib.sttode_ode_rhs_fwd
        _lib.check(fn(_lib.dptr(x), C.byref(st), _lib.dptr(out), L, bsz, _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        ctx.save_for_backward(x, *params)
        ctx.gate_only = gate_only
        return out

    @staticmethod
    def backward(ctx, df):
        x, *params = ctx.saved_tensors
        L, bsz = x.shape[0], x.shape[1]
        df = _f32c(df, "grad_output")
        dx = torch.empty_like(x)
        direct = _lib.sinks_still_valid(ctx.params_in, ctx.sinks)
        grads = ctx.sinks if direct else [torch.zeros_like(p) for p in params]                # the library accumulates (+=)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_ode_rhs_bwd_workspace_bytes(L, bsz), x.device)
        st, gs = _ode_struct_from(params), _ode_struct_from(grads, _lib.OdeGrads)
        _lib.check(lib.sttode_ode_rhs_bwd(_lib.dptr(x), _lib.dptr(df), C.byref(st), _lib.dptr(dx), C.byref(gs), L, bsz, int(ctx.gate_only),
                                          _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        return (dx, None, None, *([None] * len(grads) if direct else grads))


class _OdeBlockFn(torch.autograd.Function):
    """y = [relu](odeint(f, x, grid)[-1]) with its backward (sttode_ode_block_fwd / sttode_ode_block_bwd): the reference
    back-propagates through torchdiffeq's unrolled solver (adjoint=False, ode_demo.py:7); the library recomputes the
    solver states and runs the discrete adjoint of the same grid.  x [L, bsz, 64]."""

    @staticmethod
    def forward(ctx, x, n_steps, h, method, final_relu, packed, *params):
        L, bsz = x.shape[0], x.shape[1]
        ctx.params_in, ctx.sinks = params, _lib.grad_sinks(params)
        params = [_f32c(p, "parameter") for p in params]
        y = torch.empty_like(x)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_ode_block_workspace_bytes(L, bsz, method), x.device)
        st = _ode_struct_from(params)
        st.packed = _lib.dptr(packed)
        _lib.check(lib.sttode_ode_block_fwd(_lib.dptr(x), C.byref(st), _lib.dptr(y), 64, L, bsz, n_steps, h, method, int(final_relu),
                                            _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        ctx.save_for_backward(x, *params)
        ctx.cfg = (n_steps, h, method, final_relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, *params = ctx.saved_tensors
        n_steps, h, method, final_relu = ctx.cfg
        L, bsz = x.shape[0], x.shape[1]
        dy = _f32c(dy, "grad_output")
        dx = torch.empty_like(x)
        direct = _lib.sinks_still_valid(ctx.params_in, ctx.sinks)
        grads = ctx.sinks if direct else [torch.zeros_like(p) for p in params]
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_ode_block_bwd_workspace_bytes(L, bsz, n_steps), x.device)
        st, gs = _ode_struct_from(params), _ode_struct_from(grads, _lib.OdeGrads)
        _lib.check(lib.sttode_ode_block_bwd(_lib.dptr(x), _lib.dptr(dy), 64, C.byref(st), _lib.dptr(dx), C.byref(gs), L, bsz, n_steps, h, method,
                                            int(final_relu), _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        return (dx, None, None, None, None, None, *([None] * len(grads) if direct else grads))


def _grad_view(t, name):
    """Contiguous fp32 view that keeps the autograd history (unlike `_f32c`, which detaches)."""
    if t.dtype != torch.float32:
        raise _lib.SttodeError(f"{name}: only float32 is implemented on the B200 path (got {t.dtype})")
    return t.contiguous()


def _check_eval_dropout(p, training, what):
    if p > 0.0 and training:
        raise NotImplementedError(f"{what}: dropout > 0 in training mode is not implemented on the B200 path "
                                  "(the reference builds these layers with dropout = 0, model/STTODE.py:192,251)")


class Hypattention(nn.Module):
    """MHGSA followed by the tanh * sigmoid gate (hypertransformer.py:19-89)."""

    def __init__(self, d_model, nhead, dropout=0., motion_only=True, cross_range=0, num_conv_layer=3):
        super().__init__()
        self.model_dim = d_model
        self.temporal_attention_before = Hyp_mhsa(d_model, nhead, dropout=dropout)
        self.temporal_info = nn.Linear(d_model, d_model)
        self.temporal_gate = nn.Linear(d_model, d_model)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=False, attn_mask=None, seq_mask=False):
        # [T, N, sample_num, D]; the reference ignores every mask argument here (hypertransformer.py:69-79)
        assert query.dim() == key.dim() == value.dim() == 4
        assert query.shape[1] == key.shape[1] == value.shape[1] and query.shape[2] == key.shape[2] == value.shape[2]
        if not (key is query and value is query):
            raise NotImplementedError("Hypattention on B200: self-attention only (query is key is value), as on the reference path")
        _check_eval_dropout(self.temporal_attention_before.dropout, self.training, "Hypattention")
        T, N, sn, D = query.shape
        x = _grad_view(query, "query").reshape(T, N * sn, D)
        out = _OdeRhsFn.apply(x, True, None, *_ode_params(self))
        weights = None
        if need_weights:
            weights = self.temporal_attention_before(x, x, x, need_weights=True)[1]
        return out.reshape(T, N, sn, D), weights


class TransformerEncoderLayer(nn.Module):
    """s = LN1(x + Hypattention(x)); LN2(s + W2 relu(W1 s)) (hypertransformer.py:91-153) = the ODE right-hand side."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu"):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("TransformerEncoderLayer on B200: relu only (the reference default)")
        self.self_attn = Hypattention(d_model, nhead, dropout=dropout, motion_only=False, cross_range=2, num_conv_layer=7)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        # eval-mode forwards run the in-projection / out-proj / gate / FFN on the tcgen05 tensor cores (3xFP16 split); False -> FP32 FFMA
        self.use_tensor_cores = True
        self.tensor_cores_in_training = True        # train-mode forward of the ODE block on the tensor cores as well (its backward stays FP32)
        self._packed = _PackedOde()

    def _packed_or_none(self):
        """The tensor-core operand images when this forward should run on the tcgen05 kernels: always in eval mode; in train mode too when
        `tensor_cores_in_training` is set (the weights change every step, so every training forward re-splits them: one 4 us launch)."""
        on = self.use_tensor_cores and self.linear1.weight.is_cuda and (not self.training or self.tensor_cores_in_training)
        return self._packed.get(self) if on else None

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        _check_eval_dropout(self.dropout.p, self.training, "TransformerEncoderLayer")
        shape = src.shape
        x = _grad_view(src, "src").reshape(shape[0], -1, shape[-1])
        out = _OdeRhsFn.apply(x, False, self._packed_or_none(), *_ode_params(self))
        return out.reshape(shape)


class TransformerEncoder_ode(nn.Module):
    """The ODE function f(t, y): a stack of encoder layers; ignores t and the stored masks (ode_demo.py:25-72)."""

    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers, self.norm = num_layers, norm
        self.src = self.mask = self.src_key_padding_mask = None
        self.num_agent = 1

    def forward(self, t, src):
        out = src
        for mod in self.layers:
            out = mod(out)
        if self.norm is not None:
            out = self.norm(out)
        return out


class ODEblock_en(nn.Module):
    """z = odeint(f, x, t, method)[1] with the solver loop inside the library (ode_demo.py:169-190).

    `n_steps` / `method` generalise the reference's single Euler step over t = [0, time]."""

    def __init__(self, odefunc, t=torch.tensor([0, 1]), method="euler", n_steps=1):
        super().__init__()
        self.t, self.odefunc, self.method, self.n_steps = t, odefunc, method, n_steps

    def set_x0(self, tgt):
        self.odefunc.tgt = tgt.detach()          # the reference clones; nothing reads it (ode_demo.py:175-177)

    def setparrament(self, src, mask, src_key_padding_mask, num_agent=1):
        self.odefunc.src, self.odefunc.mask = src, (mask,)
        self.odefunc.src_key_padding_mask, self.odefunc.num_agent = src_key_padding_mask, num_agent

    def solve(self, x, final_relu):
        if self.odefunc.num_layers != 1 or self.odefunc.norm is not None:
            raise NotImplementedError("ODEblock_en on B200: one encoder layer, no final norm (model/STTODE.py:193-195)")
        layer = self.odefunc.layers[0]
        _check_eval_dropout(layer.dropout.p, layer.training, "ODEblock_en")
        shape = x.shape
        xs = _grad_view(x, "x").reshape(shape[0], -1, shape[-1])
        span = float(self.t[-1]) - float(self.t[0])
        method = _METHODS[self.method]
        y = _OdeBlockFn.apply(xs, self.n_steps, span / self.n_steps, method, bool(final_relu), layer._packed_or_none(), *_ode_params(layer))
        return y.reshape(shape)

    def forward(self, x):
        return self.solve(x, final_relu=False)


class ODEG_Encoder(nn.Module):
    """relu(ODEblock_en(src)) (ode_demo.py:217-231)."""

    def __init__(self, encoder_layer, nlayer, time, method="euler", n_steps=1):
        super().__init__()
        self.odeblock = ODEblock_en(TransformerEncoder_ode(encoder_layer, nlayer), t=torch.tensor([0, time]),
                                    method=method, n_steps=n_steps)

    def forward(self, src, mask=None, src_key_padding_mask=None, num_agent=1):
        self.odeblock.set_x0(src)
        self.odeblock.setparrament(src=src, mask=mask, src_key_padding_mask=src_key_padding_mask, num_agent=num_agent)
        return self.odeblock.solve(src, final_relu=True)
