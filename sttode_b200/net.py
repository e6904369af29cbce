"""STTODENet on B200: drop-in for the reference's `model/STTODE.py` model API.

`STTODENet(args, device)` keeps the reference's constructor, parameter names (90 parameters + 2 buffers, so
reference checkpoints load with `strict=True`), `set_data` / `set_data_nba` / `inference` call sequence
(test.py:164-199, 495-526) and the attributes callers read (`agent_num`, `scene_orig`, `past_feature`,
`past_traj`, `cur_location`).  `inference` is one `sttode_forecast` call: scene normalisation, velocity
features, past encoder (front end + tensor-ODE block with MHGSA), K-sample decoder and the latent draw all
run inside libsttode_sm100.so.  `forecast_scenes` is the batched form (many ETH/SDD scenes per launch) the
reference does not have.

Training (`train()` mode; train.py:55-95): `forward()` returns the reference's loss tuple with autograd history.  Every layer of
the step runs forward and backward in libsttode_sm100 behind autograd nodes: the MHGSA / tensor-ODE blocks of both encoders
(sttode_ode_block_fwd / _bwd), the thin Linear layers around them (sttode_dense_fwd / _bwd) and the whole decoder
(sttode_decoder_fwd / _bwd: Conv1d, GRU through time, MLPs).  torch supplies the graph bookkeeping, concatenations, the
Bernoulli / Gaussian draws and the scalar loss arithmetic.  `eval()` + `forward()` evaluates the same objective forward-only.
"""
import ctypes as C
import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import _lib
from .layers import ODEG_Encoder, TransformerEncoderLayer, _ode_struct
from .mhgsa import _f32c

TRAJ_RT2, TRAJ_R2T = 0, 1


def initialize_weights(modules):
    """model/utils.py:11-21 (Linear branch is the only one reached by this model)."""
    for m in modules:
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0, 0.01)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class MLP(nn.Module):
    """Linear-ReLU stack, last layer linear (model/utils.py:67-95).  Parameter container: the arithmetic of
    the decoder MLPs runs inside k_dec_pair."""

    def __init__(self, input_dim, output_dim, hidden_size=(1024, 512), activation="relu", discrim=False, dropout=-1):
        super().__init__()
        dims = [input_dim, *hidden_size, output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):                        # train() mode only (model/utils.py:88-95)
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i != len(self.layers) - 1:
                x = torch.relu(x)
        return x


class MLP2(nn.Module):
    """model/STTODE.py:111-133 (FutureEncoder.out_mlp)."""

    def __init__(self, input_dim, hidden_dims=(128, 128), activation="tanh"):
        super().__init__()
        self.out_dim = hidden_dims[-1]
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh
        initialize_weights(self.affine_layers.modules())


class Normal:
    """Diagonal Gaussian (model/STTODE.py:79-109); used by callers of the sampler stage."""

    def __init__(self, mu=None, logvar=None, params=None):
        if params is not None:
            self.mu, self.logvar = torch.chunk(params, chunks=2, dim=-1)
        else:
            self.mu, self.logvar = mu, logvar
        self.sigma = torch.exp(0.5 * self.logvar)

    def rsample(self):
        return self.mu + torch.randn_like(self.sigma) * self.sigma

    sample = rsample

    def kl(self, p=None):
        if p is None:
            return -0.5 * (1 + self.logvar - self.mu.pow(2) - self.logvar.exp())
        t1 = (self.mu - p.mu) / (p.sigma + 1e-8)
        t2 = self.sigma / (p.sigma + 1e-8)
        return 0.5 * (t1 * t1 + t2 * t2) - 0.5 - torch.log(t2)

    def mode(self):
        return self.mu


class PositionalAgentEncoding(nn.Module):
    """Sinusoid table + Linear(2d -> d) on [x | pe] (model/STTODE.py:137-176)."""

    def __init__(self, d_model, dropout=0.1, max_t_len=200, concat=True):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.concat, self.d_model = concat, d_model
        if concat:
            self.fc = nn.Linear(2 * d_model, d_model)
        pe = torch.zeros(max_t_len, d_model)
        pos = torch.arange(0, max_t_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe)


class _FoldedFrontEnd:
    """The eval-mode encoder front end (input_fc, pos_encoder.fc, input_fc2, input_fc3: Linear layers with nothing nonlinear between them)
    collapsed to one affine map of the 4 T input features (sttode_encoder_fold, float64 accumulation).  One persistent buffer, rebuilt in
    place when a front-end weight changes (same staleness rules as _PackedDecoder)."""

    def __init__(self):
        self.key, self.buf, self.packs = None, None, 0

    def invalidate(self):
        self.key = None

    def __deepcopy__(self, memo):
        return _FoldedFrontEnd()

    def get(self, enc, T):
        params = [enc.input_fc.weight, enc.input_fc.bias, enc.pos_encoder.pe, enc.pos_encoder.fc.weight, enc.pos_encoder.fc.bias,
                  enc.input_fc2.weight, enc.input_fc2.bias, enc.input_fc3.weight, enc.input_fc3.bias]
        key = tuple((p.data_ptr(), p._version) for p in params) + (T,)
        if key != self.key:
            lib = _lib.lib()
            nbytes = lib.sttode_encoder_fold_bytes(T)
            if self.buf is None or self.buf.device != params[0].device or self.buf.numel() * 4 != nbytes:
                self.buf = torch.empty(nbytes // 4, dtype=torch.float32, device=params[0].device)
            st = _encoder_struct(enc, tensor_cores=False)
            _lib.check(lib.sttode_encoder_fold(C.byref(st), T, _lib.dptr(self.buf), nbytes, _lib.stream_ptr()))
            self.key = key
            self.packs += 1
        return self.buf


def _encoder_struct(enc, tensor_cores=None):
    """sttode_encoder_weights of a Past/FutureEncoder.  tensor_cores (default: eval mode, `use_tensor_cores` set, sequence length <= 30): the
    struct also carries the ODE block's tensor-core operand images and the folded front-end table, which selects the tcgen05 encoder kernels."""
    s = _lib.EncoderWeights()
    s.fc_w, s.fc_b = _lib.dptr(enc.input_fc.weight), _lib.dptr(enc.input_fc.bias)
    s.pe = _lib.dptr(enc.pos_encoder.pe)
    s.pefc_w, s.pefc_b = _lib.dptr(enc.pos_encoder.fc.weight), _lib.dptr(enc.pos_encoder.fc.bias)
    s.fc2_w, s.fc2_b = _lib.dptr(enc.input_fc2.weight), _lib.dptr(enc.input_fc2.bias)
    s.fc3_w, s.fc3_b = _lib.dptr(enc.input_fc3.weight), _lib.dptr(enc.input_fc3.bias)
    layer = enc.ODE_Encoder.odeblock.odefunc.layers[0]
    if tensor_cores is None:
        tensor_cores = not enc.training and layer.use_tensor_cores
    s.ode = _ode_struct(layer, tensor_cores=bool(tensor_cores))
    T = enc.input_fc2.weight.shape[1] // 64
    if tensor_cores and getattr(enc, "fold_front_end", True) and T <= 30:
        s.folded = _lib.dptr(enc._packed.get(enc, T))
    return s


class _EncoderBase(nn.Module):
    def __init__(self, args, length, in_dim=4):
        super().__init__()
        self.args = args
        self.model_dim = args.hidden_dim
        if self.model_dim != 64:
            raise NotImplementedError("B200 path: hidden_dim = 64 only (train.py:38 default)")
        self.input_fc = nn.Linear(in_dim, self.model_dim)
        self.input_fc2 = nn.Linear(self.model_dim * length, self.model_dim)
        self.input_fc3 = nn.Linear(self.model_dim + 3, self.model_dim)
        self.fold_front_end = True                # eval mode: front end folded into one affine map inside the tensor-core kernels
        self._packed = _FoldedFrontEnd()
        self.nhead, self.ff_dim, self.dropout, self.nlayer = 8, 1024, 0., 1
        self.ODE_Encoder = ODEG_Encoder(TransformerEncoderLayer(self.model_dim, self.nhead, self.ff_dim, self.dropout), self.nlayer, 12)
        self.pos_encoder = PositionalAgentEncoding(self.model_dim, 0.1, concat=True)
        self.agent_mask = None

    def get_agent_mask(self, agent_mask):
        self.agent_mask = agent_mask


class PastEncoder(_EncoderBase):
    """model/STTODE.py:178-236.  forward(inputs [B*N,T_p,4], B, N) -> [B*N,128]."""

    def __init__(self, args, in_dim=4):
        super().__init__(args, args.past_length, in_dim)
        self.scale_number = len(args.hyper_scales)

    def forward(self, inputs, batch_size, agent_num):
        return _encoder_forward(self, inputs, batch_size, agent_num)


def _encoder_forward(enc, inputs, batch_size, agent_num):
    """Front end + tensor-ODE block on explicit [x, y, vx, vy] features -> [B*N, 128].
    eval(): one library call (sttode_encoder_fwd), no autograd history.  train(): differentiable (see `_encoder_forward_train`)."""
    if enc.training:
        return _encoder_forward_train(enc, inputs, batch_size, agent_num)
    feat = _f32c(inputs, "inputs")
    rows, T = feat.shape[0], feat.shape[1]
    if feat.shape[-1] != 4 or rows != batch_size * agent_num:
        raise _lib.SttodeError("encoder inputs must be [batch_size*agent_num, T, 4]")
    st = _encoder_struct(enc)
    out = torch.empty(rows, 128, device=feat.device, dtype=torch.float32)
    lib = _lib.lib()
    ws = _lib.workspace(lib.sttode_past_encoder_workspace_bytes(rows, batch_size), feat.device)
    _lib.check(lib.sttode_encoder_fwd(_lib.dptr(feat), rows, batch_size, agent_num, T, C.byref(st), _lib.dptr(out), _lib.dptr(ws), ws.numel(),
                                      _lib.stream_ptr()))
    return out


class _DenseFn(torch.autograd.Function):
    """y = act(x W^T + b) with its backward, both in libsttode_sm100 (sttode_dense_fwd / sttode_dense_bwd): the train()-mode form of the
    thin Linear layers around the tensor-ODE block.  x [..., K] (any leading shape), W [N, K], b [N]."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        if not x.is_cuda:
            raise _lib.SttodeError("dense layer input must be a CUDA tensor; this path has no CPU implementation")
        xs = _f32c(x, "x").reshape(-1, x.shape[-1])
        w, b = _f32c(weight, "weight"), _f32c(bias, "bias")
        rows, K, N = xs.shape[0], xs.shape[1], w.shape[0]
        y = torch.empty(rows, N, device=xs.device, dtype=torch.float32)
        _lib.check(_lib.lib().sttode_dense_fwd(_lib.dptr(xs), K, _lib.dptr(w), _lib.dptr(b), _lib.dptr(y), rows, K, N, int(relu), _lib.stream_ptr()))
        ctx.save_for_backward(xs, w, y if relu else None)
        ctx.meta = (x.shape, bool(relu), ctx.needs_input_grad[0])
        ctx.params_in, ctx.sinks = (weight, bias), _lib.grad_sinks((weight, bias))
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        xs, w, y = ctx.saved_tensors
        shape, relu, need_dx = ctx.meta
        rows, K, N = xs.shape[0], xs.shape[1], w.shape[0]
        dy = dy.reshape(rows, N)
        dy = (dy * (y > 0)).contiguous() if relu else _f32c(dy, "grad_output")
        dx = torch.empty_like(xs) if need_dx else None
        direct = _lib.sinks_still_valid(ctx.params_in, ctx.sinks)
        if direct:                                  # the kernels accumulate straight into the parameters' .grad (flat gradient buffer)
            dw, db = ctx.sinks
        else:
            dw, db = torch.zeros_like(w), torch.zeros(N, device=w.device, dtype=torch.float32)      # the library accumulates (+=)
        lib = _lib.lib()
        ws = _lib.scratch(lib.sttode_dense_bwd_workspace_bytes(), xs.device, "splitk")
        _lib.check(lib.sttode_dense_bwd(_lib.dptr(xs), K, _lib.dptr(w), _lib.dptr(dy), _lib.dptr(dx), K, _lib.dptr(dw), _lib.dptr(db), rows, K, N,
                                        _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        return (dx.view(shape) if need_dx else None), (None if direct else dw), (None if direct else db), None


def _dense(x, layer, relu=False):
    return _DenseFn.apply(x, layer.weight, layer.bias, relu)


def _encoder_forward_train(enc, inputs, batch_size, agent_num):
    """train() mode (model/STTODE.py:214-236 with the PE dropout of :140,176 live).  Every layer runs forward AND backward in
    libsttode_sm100: the thin Linear layers as `_DenseFn` nodes (sttode_dense_fwd / _bwd), the MHGSA / tensor-ODE block as one node
    (sttode_ode_block_fwd / _bwd).  torch only concatenates, reshapes and applies the Bernoulli mask.
    `enc.dropout_mask_override` ([R,T,64] of 0/1) replaces the draw (tests)."""
    rows, T = inputs.shape[0], inputs.shape[1]
    if inputs.shape[-1] != 4 or rows != batch_size * agent_num:
        raise _lib.SttodeError("encoder inputs must be [batch_size*agent_num, T, 4]")
    h = _dense(inputs, enc.input_fc)
    pe = enc.pos_encoder.pe[:T].unsqueeze(0).expand(rows, T, 64)
    h = _dense(torch.cat((h, pe), dim=-1), enc.pos_encoder.fc)
    mask = getattr(enc, "dropout_mask_override", None)
    p = enc.pos_encoder.dropout.p
    h = F.dropout(h, p, True) if mask is None else h * (mask.to(h) / (1.0 - p))
    h = _dense(h.reshape(batch_size, agent_num, T * 64), enc.input_fc2)
    cat = torch.zeros(agent_num, 3, device=h.device, dtype=h.dtype)
    cat[agent_num - 1:, 2:].fill_(1.0)                                          # model/STTODE.py:199-210 (fill_: capturable, no H2D scalar copy)
    x0 = _dense(torch.cat((h, cat.expand(batch_size, agent_num, 3)), dim=-1), enc.input_fc3)
    y = enc.ODE_Encoder(x0.unsqueeze(2), mask=None, num_agent=agent_num).squeeze(2)
    return torch.cat((x0, y), dim=-1).view(rows, 128)


def _linear(x, layer, relu):
    out = torch.empty(x.shape[0], layer.out_features, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().sttode_linear(_lib.dptr(x), _lib.dptr(layer.weight), _lib.dptr(layer.bias), _lib.dptr(out), x.shape[0],
                                        layer.in_features, layer.out_features, int(relu), _lib.stream_ptr()))
    return out


class FutureEncoder(_EncoderBase):
    """Posterior encoder (model/STTODE.py:238-300)."""

    def __init__(self, args, in_dim=4):
        super().__init__(args, args.future_length, in_dim)
        scale_num = 2 + len(args.hyper_scales)
        self.out_mlp = MLP2(scale_num * self.model_dim, [128], "relu")
        self.qz_layer = nn.Linear(self.out_mlp.out_dim, 2 * args.zdim)
        initialize_weights(self.qz_layer.modules())

    def forward(self, inputs, batch_size, agent_num, past_feature):
        """-> q(z | past, future) parameters [B*N, 2*zdim] (model/STTODE.py:276-300)."""
        ff = _encoder_forward(self, inputs, batch_size, agent_num)
        if self.training:
            h = _dense(torch.cat((past_feature, ff), dim=-1), self.out_mlp.affine_layers[0], relu=True)
            return _dense(h, self.qz_layer)
        h = torch.cat((_f32c(past_feature, "past_feature"), ff), dim=-1)
        h = _linear(h, self.out_mlp.affine_layers[0], relu=True)
        return _linear(h, self.qz_layer, relu=False)


class DecomposeBlock(nn.Module):
    """Conv1d(2->32,k3)+GRU(32->96)+two MLPs (model/STTODE.py:16-77).  Parameter container."""

    def __init__(self, past_len, future_len, input_dim):
        super().__init__()
        self.past_len, self.future_len = past_len, future_len
        self.conv_past = nn.Conv1d(2, 32, 3, stride=1, padding=1)
        self.encoder_past = nn.GRU(32, 96, 1, batch_first=True)
        self.decoder_y = MLP(96 + input_dim, future_len * 2, hidden_size=(512, 256))
        self.decoder_x = MLP(96 + input_dim, past_len * 2, hidden_size=(512, 256))
        nn.init.kaiming_normal_(self.conv_past.weight)
        nn.init.kaiming_normal_(self.encoder_past.weight_ih_l0)
        nn.init.kaiming_normal_(self.encoder_past.weight_hh_l0)
        nn.init.zeros_(self.conv_past.bias)
        nn.init.zeros_(self.encoder_past.bias_ih_l0)
        nn.init.zeros_(self.encoder_past.bias_hh_l0)

    def forward(self, x_true, x_hat, f):
        """Differentiable form used in train() mode (model/STTODE.py:51-77); inference runs k_dec_row_tc / k_dec_pair_tc.
        The GRU recurrence is unrolled with fp32 matmuls (cuDNN's RNN path drops to TF32 on sm_100, which costs ~2e-4 in the
        gradients); same gate order and arithmetic as nn.GRU (r, z, n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn)))."""
        emb = torch.relu(F.conv1d((x_true - x_hat).transpose(1, 2), self.conv_past.weight, self.conv_past.bias, padding=1)).transpose(1, 2)
        g = self.encoder_past
        gi_all = F.linear(emb, g.weight_ih_l0, g.bias_ih_l0)                      # [rows, T, 288]
        h = emb.new_zeros(emb.shape[0], 96)
        for t in range(emb.shape[1]):
            gi, gh = gi_all[:, t], F.linear(h, g.weight_hh_l0, g.bias_hh_l0)
            r = torch.sigmoid(gi[:, :96] + gh[:, :96])
            zt = torch.sigmoid(gi[:, 96:192] + gh[:, 96:192])
            n = torch.tanh(gi[:, 192:] + r * gh[:, 192:])
            h = (1 - zt) * n + zt * h
        feat = torch.cat((f, h), dim=1)
        return self.decoder_x(feat).view(-1, self.past_len, 2), self.decoder_y(feat).view(-1, self.future_len, 2)


def _mlp3_struct(mlp):
    s = _lib.Mlp3Weights()
    for i in range(3):
        setattr(s, f"w{i}", _lib.dptr(mlp.layers[i].weight))
        setattr(s, f"b{i}", _lib.dptr(mlp.layers[i].bias))
    return s


def _decompose_struct(blk):
    s = _lib.DecomposeWeights()
    s.conv_w, s.conv_b = _lib.dptr(blk.conv_past.weight), _lib.dptr(blk.conv_past.bias)
    g = blk.encoder_past
    s.gru_w_ih, s.gru_w_hh = _lib.dptr(g.weight_ih_l0), _lib.dptr(g.weight_hh_l0)
    s.gru_b_ih, s.gru_b_hh = _lib.dptr(g.bias_ih_l0), _lib.dptr(g.bias_hh_l0)
    s.dec_y, s.dec_x = _mlp3_struct(blk.decoder_y), _mlp3_struct(blk.decoder_x)
    return s


class _PackedDecoder:
    """fp16 hi/lo operand images of the decoder weights for the tcgen05 kernels.

    ONE buffer per (device, horizon), allocated once and rewritten in place (a single `k_pack_all` launch) whenever the weights
    changed, so a pointer baked into a CUDA graph (ForecastGraph, GraphedTrainStep) stays valid for the life of the model and always
    sees the newest pack.  Staleness is detected by (data_ptr, _version) of every decoder parameter — torch optimizers,
    `load_state_dict` and in-place tensor ops all bump `_version` — plus an explicit `invalidate()` for updates that do not:
    CUDA-graph replays of an optimizer step and the library's fused all-reduce + Adam kernel (parallel.py call it)."""

    def __init__(self):
        self.key, self.buf, self.shape = None, None, None
        self.packs = 0                      # number of pack launches issued (tests)

    def invalidate(self):
        self.key = None

    def get(self, decoder, past_length, future_length):
        params = [p for blk in decoder.decompose for p in blk.parameters()]
        key = tuple((p.data_ptr(), p._version) for p in params)
        shape = (past_length, future_length, params[0].device)
        if key != self.key or shape != self.shape:
            lib = _lib.lib()
            if self.buf is None or shape != self.shape:
                self.buf = torch.empty(lib.sttode_decoder_pack_bytes(past_length, future_length), dtype=torch.uint8, device=params[0].device)
            mw = _lib.ModelWeights()
            mw.past_length, mw.future_length = past_length, future_length
            mw.dec[0], mw.dec[1] = _decompose_struct(decoder.decompose[0]), _decompose_struct(decoder.decompose[1])
            _lib.check(lib.sttode_decoder_pack(C.byref(mw), _lib.dptr(self.buf), self.buf.numel(), _lib.stream_ptr()))
            self.key, self.shape = key, shape
            self.packs += 1
        return self.buf

    def status(self, past_length, future_length):
        """0 unless a tensor-core barrier ever timed out (synchronises)."""
        if self.buf is None:
            return 0
        st = C.c_int(0)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().sttode_decoder_pack_status(_lib.dptr(self.buf), past_length, future_length, C.byref(st)))
        return st.value


class Decoder(nn.Module):
    """K-sample residual decoder (model/STTODE.py:302-347)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.model_dim = args.hidden_dim
        self.decode_way = "RES"
        self.num_decompose = args.num_decompose
        if self.num_decompose != 2:
            raise NotImplementedError("B200 path: num_decompose = 2 only (train.py:40 default)")
        input_dim = 2 * self.model_dim + args.zdim
        if args.zdim != 32:
            raise NotImplementedError("B200 path: zdim = 32 only (train.py:37 default)")
        self.past_length, self.future_length = args.past_length, args.future_length
        self.decompose = nn.ModuleList(DecomposeBlock(args.past_length, args.future_length, input_dim) for _ in range(2))
        # tcgen05 decoder tile (k_dec_pair_tc); False -> FP32-FFMA kernels (also chosen for horizons whose layer-3 block does not fit
        # a ring stage / the TMEM column budget: 2 * max(T_p, T_f) > 48)
        self.use_tensor_cores = 2 * max(args.past_length, args.future_length) <= 48
        self._packed = _PackedDecoder()

    def forward(self, past_feature, z, batch_size_curr, agent_num_perscene, past_traj, cur_location, sample_num, mode="train"):
        """Reference signature: past_feature / z are [rows*sample_num, .] (already repeated per sample).
        Returns (out_seq, recover_pre_seq).  Every (agent, sample) pair is treated as its own row (K = 1 inside
        the library) so arbitrary per-pair features are honoured; STTODENet.inference uses the de-duplicated
        path instead."""
        if self.training:
            return self._forward_train(past_feature, z, past_traj, cur_location, sample_num, mode)
        pf = _f32c(past_feature, "past_feature")
        zz = _f32c(z, "z").reshape(-1, 32)
        pairs = pf.shape[0]
        traj = _f32c(past_traj, "past_traj").repeat_interleave(sample_num, dim=0)
        # the library adds past_traj[:, -1] itself (every reference call site passes exactly that as cur_location); only the shape is checked
        # here — comparing values would force a device->host sync (the class of hidden sync SURVEY K1 removes) and break graph capture
        if cur_location.reshape(-1, 2).shape[0] != past_traj.shape[0]:
            raise NotImplementedError("Decoder on B200: cur_location must be past_traj[:, -1] (as at every reference call site)")
        mw = _lib.ModelWeights()
        mw.past_length, mw.future_length = self.past_length, self.future_length
        mw.dec[0], mw.dec[1] = _decompose_struct(self.decompose[0]), _decompose_struct(self.decompose[1])
        mw.dec_packed = _lib.dptr(self._packed.get(self, self.past_length, self.future_length)) if self.use_tensor_cores else None
        pred = torch.empty(pairs, 1, self.future_length, 2, device=pf.device, dtype=torch.float32)
        recon = torch.empty(pairs, self.past_length, 2, device=pf.device, dtype=torch.float32)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_decoder_workspace_bytes(pairs, 1), pf.device)
        _lib.check(lib.sttode_decoder_fwd(_lib.dptr(pf), _lib.dptr(traj), _lib.dptr(zz), 0, None, None, pairs, 1, C.byref(mw),
                                          _lib.dptr(pred), _lib.dptr(recon), _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        out = pred.view(pairs, self.future_length, 2)
        if mode == "inference":
            out = out.view(-1, sample_num, self.future_length, 2)
        return out, recon


def _decompose_params(decoder):
    """The 36 decoder tensors in sttode_decompose_weights order, block 0 then block 1."""
    ps = []
    for blk in decoder.decompose:
        g = blk.encoder_past
        ps += [blk.conv_past.weight, blk.conv_past.bias, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0]
        for mlp in (blk.decoder_y, blk.decoder_x):
            for layer in mlp.layers:
                ps += [layer.weight, layer.bias]
    return ps


def _decompose_struct_from(t18, cls):
    """sttode_decompose_weights / sttode_decompose_grads over 18 tensors laid out like `_decompose_params`."""
    s = cls()
    s.conv_w, s.conv_b, s.gru_w_ih, s.gru_w_hh, s.gru_b_ih, s.gru_b_hh = (_lib.dptr(t) for t in t18[:6])
    for mlp, off in ((s.dec_y, 6), (s.dec_x, 12)):
        mlp.w0, mlp.b0, mlp.w1, mlp.b1, mlp.w2, mlp.b2 = (_lib.dptr(t) for t in t18[off:off + 6])
    return s


class _DecoderTrainFn(torch.autograd.Function):
    """Decoder.forward in train() mode as one autograd node over sttode_decoder_fwd (tcgen05 kernels, every pair its own row) and
    sttode_decoder_bwd (Conv1d / GRU-through-time / MLP backward in libsttode_sm100).  past_feature [P,128], z [P,32],
    traj [P,T_p,2] = x_true per pair -> (out_seq [P,T_f,2] incl. the current position, recover_pre_seq [P,T_p,2])."""

    @staticmethod
    def forward(ctx, past_feature, z, traj, Tp, Tf, packed, *params):
        pf, zz, tj = _f32c(past_feature, "past_feature"), _f32c(z, "z"), _f32c(traj, "past_traj")
        ctx.params_in, ctx.sinks = params, _lib.grad_sinks(params)
        params = [_f32c(p, "parameter") for p in params]
        P = pf.shape[0]
        mw = _lib.ModelWeights()
        mw.past_length, mw.future_length = Tp, Tf
        mw.dec[0], mw.dec[1] = _decompose_struct_from(params[:18], _lib.DecomposeWeights), _decompose_struct_from(params[18:], _lib.DecomposeWeights)
        mw.dec_packed = _lib.dptr(packed)              # fp16 hi/lo images of the CURRENT weights -> tcgen05 forward; None -> FP32 kernels
        pred = torch.empty(P, 1, Tf, 2, device=pf.device, dtype=torch.float32)
        recon = torch.empty(P, Tp, 2, device=pf.device, dtype=torch.float32)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_decoder_workspace_bytes(P, 1), pf.device)
        _lib.check(lib.sttode_decoder_fwd(_lib.dptr(pf), _lib.dptr(tj), _lib.dptr(zz), 0, None, None, P, 1, C.byref(mw), _lib.dptr(pred),
                                          _lib.dptr(recon), _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        ctx.save_for_backward(pf, zz, tj, *params)
        ctx.shape = (Tp, Tf)
        return pred.view(P, Tf, 2), recon

    @staticmethod
    def backward(ctx, d_pred, d_recon):
        pf, zz, tj, *params = ctx.saved_tensors
        Tp, Tf = ctx.shape
        P = pf.shape[0]
        d_pred = _f32c(d_pred, "grad_output")
        d_recon = None if d_recon is None else _f32c(d_recon, "grad_output")
        mw = _lib.ModelWeights()
        mw.past_length, mw.future_length = Tp, Tf
        mw.dec[0], mw.dec[1] = _decompose_struct_from(params[:18], _lib.DecomposeWeights), _decompose_struct_from(params[18:], _lib.DecomposeWeights)
        direct = _lib.sinks_still_valid(ctx.params_in, ctx.sinks)
        grads = ctx.sinks if direct else [torch.zeros_like(p) for p in params]         # the library accumulates (+=)
        gs = (_lib.DecomposeGrads * 2)(_decompose_struct_from(grads[:18], _lib.DecomposeGrads), _decompose_struct_from(grads[18:], _lib.DecomposeGrads))
        d_pf, d_z = torch.empty_like(pf), torch.empty_like(zz)
        lib = _lib.lib()
        ws = _lib.workspace(lib.sttode_decoder_bwd_workspace_bytes(P, Tp, Tf), pf.device)
        _lib.check(lib.sttode_decoder_bwd(_lib.dptr(pf), _lib.dptr(zz), _lib.dptr(tj), _lib.dptr(d_pred), _lib.dptr(d_recon), P, C.byref(mw), gs,
                                          _lib.dptr(d_pf), _lib.dptr(d_z), _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        return (d_pf, d_z, None, None, None, None, *([None] * len(grads) if direct else grads))


def _decoder_forward_train(self, past_feature, z, past_traj, cur_location, sample_num, mode):
    """train() mode (model/STTODE.py:320-347) with autograd history: forward and backward both run in libsttode_sm100."""
    if not past_feature.is_cuda:
        raise _lib.SttodeError("Decoder inputs must be CUDA tensors; this path has no CPU implementation")
    if cur_location.shape[0] != past_traj.shape[0]:            # the library adds past_traj[:, -1] itself (every reference call site passes it);
        raise NotImplementedError("Decoder on B200: cur_location must be past_traj[:, -1]")      # no value check here: it would force a device sync
    x_true = past_traj.detach().repeat_interleave(sample_num, dim=0)
    pf = past_feature.reshape(x_true.shape[0], -1)
    # the tensor-core forward needs the weights re-split into fp16 hi/lo images whenever they change (every optimizer step): one
    # repack serves both decoder passes of a step (the cache is keyed on the parameters' version counters)
    packed = self._packed.get(self, self.past_length, self.future_length) if self.use_tensor_cores else None
    out, rec = _DecoderTrainFn.apply(pf, z.reshape(-1, 32), x_true, self.past_length, self.future_length, packed, *_decompose_params(self))
    if mode == "inference":
        out = out.view(-1, sample_num, *out.shape[1:])
    return out, rec


Decoder._forward_train = _decoder_forward_train


def _rotation_2d(x, theta, origin):
    """model/STTODE.py:6-14: rotate about `origin`; returns (rotated, rotated - origin)."""
    nx = x - origin
    c, sn = torch.cos(theta), torch.sin(theta)
    rot = torch.stack((nx[..., 0] * c - nx[..., 1] * sn, nx[..., 0] * sn + nx[..., 1] * c), dim=-1)
    return rot + origin, rot


class ForecastGraph:
    """One `sttode_forecast` call captured as a CUDA graph for a fixed shape (the whole K-sample inference is then a single
    graph launch: no per-kernel launch latency, no host work between kernels).  Inputs are copied into static buffers; the
    latent draw reads a device counter that the graph itself bumps, so every replay samples fresh z (sttode.h: rng_state)."""

    def __init__(self, model, rows, L, agents, n_scenes=0, layout=TRAJ_R2T, sample_k=None):
        self.model, self.rows, self.L, self.agents, self.n_scenes, self.layout = model, rows, L, agents, n_scenes, layout
        self.K = model.args.sample_k if sample_k is None else sample_k
        dev = next(model.parameters()).device
        Tp, Tf = model.args.past_length, model.args.future_length
        lib = _lib.lib()
        self.traj = torch.zeros(rows, 2, Tp, device=dev) if layout == TRAJ_R2T else torch.zeros(rows, Tp, 2, device=dev)
        self.scene_ptr = torch.zeros(n_scenes + 1, dtype=torch.int32, device=dev) if n_scenes else None
        self.row_scene = torch.zeros(rows, dtype=torch.int32, device=dev) if n_scenes else None
        self.pred = torch.empty(rows, self.K, Tf, 2, device=dev)
        self.past_feature = torch.empty(rows, 128, device=dev)
        self.traj_norm = torch.empty(rows, Tp, 2, device=dev)
        self.origin = torch.empty(max(n_scenes, 1), 2, device=dev)
        self.ws = _lib.workspace(lib.sttode_forecast_workspace_bytes(rows, L, self.K, Tp), dev)
        self.rng_state = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seed = model._seed()
        self.graph = None

    def _launch(self):
        mw = self.model._weights()
        mw.rng_state = self.rng_state.data_ptr()
        self._mw = mw
        self._packed_ptrs = (mw.dec_packed, mw.past.ode.packed, mw.past.folded)
        _lib.check(_lib.lib().sttode_forecast(_lib.dptr(self.traj), self.layout, _lib.dptr(self.scene_ptr), _lib.dptr(self.row_scene),
                                              self.n_scenes, self.rows, self.L, self.agents, None, self.seed, self.K, C.byref(mw),
                                              _lib.dptr(self.pred), _lib.dptr(self.past_feature), _lib.dptr(self.traj_norm),
                                              _lib.dptr(self.origin), _lib.dptr(self.ws), self.ws.numel(), _lib.stream_ptr()))

    def capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._launch()                       # warm-up outside the capture (packs the decoder weights, sets smem attributes)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._launch()
        return self

    @torch.no_grad()
    def __call__(self, traj, scene_ptr=None):
        # the graph holds the pointers of the model's persistent pack buffers: refresh their contents (in place) if the weights changed since
        if self.model._refresh_packed() != self._packed_ptrs:
            raise _lib.SttodeError("ForecastGraph: the model's packed-weight buffers moved or the tensor-core switches changed; capture again")
        self.traj.copy_(traj.reshape(self.traj.shape), non_blocking=True)
        if self.n_scenes:
            sp = scene_ptr.to(device=self.traj.device, dtype=torch.int32)
            self.scene_ptr.copy_(sp, non_blocking=True)
            counts = (sp[1:] - sp[:-1]).long()
            self.row_scene.copy_(torch.repeat_interleave(torch.arange(self.n_scenes, device=sp.device, dtype=torch.int32), counts,
                                                         output_size=self.rows))
        self.graph.replay()
        return self.pred


class STTODENet(nn.Module):
    """model/STTODE.py:349-623."""

    def __init__(self, args, device):
        super().__init__()
        self.device, self.args = device, args
        self.max_train_agent = args.max_train_agent
        self.rand_rot_scene, self.discrete_rot = args.rand_rot_scene, args.discrete_rot
        scale_num = 2 + len(args.hyper_scales)
        self.past_encoder = PastEncoder(args)
        self.pz_layer = nn.Linear(scale_num * args.hidden_dim, 2 * args.zdim)
        if args.learn_prior:
            raise NotImplementedError("learn_prior=True crashes in the reference too (pz_layer expects 256 features, gets 128; model/STTODE.py:361,513)")
        self.future_encoder = FutureEncoder(args)
        self.decoder = Decoder(args)
        self.param_annealers = nn.ModuleList()
        self.z_override = None          # test hook: inject the latent instead of drawing it
        self._struct_cache = None

    def set_device(self, device):
        self.device = device
        self.to(device)

    # ---- weights ---------------------------------------------------------------------------------
    def _weights(self):
        mw = _lib.ModelWeights()
        mw.past_length, mw.future_length = self.args.past_length, self.args.future_length
        mw.past = _encoder_struct(self.past_encoder)
        mw.dec[0], mw.dec[1] = _decompose_struct(self.decoder.decompose[0]), _decompose_struct(self.decoder.decompose[1])
        if self.decoder.use_tensor_cores:
            mw.dec_packed = _lib.dptr(self.decoder._packed.get(self.decoder, self.args.past_length, self.args.future_length))
        return mw

    def _refresh_packed(self):
        """Bring the tensor-core operand images used by `inference` up to date; -> (decoder pack ptr, past-encoder ODE pack ptr), None = FFMA."""
        dec = self.decoder
        d = dec._packed.get(dec, self.args.past_length, self.args.future_length).data_ptr() if dec.use_tensor_cores else None
        layer = self.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0]
        pk = layer._packed_or_none()
        enc = self.past_encoder
        fd = enc._packed.get(enc, self.args.past_length).data_ptr() if (pk is not None and enc.fold_front_end and self.args.past_length <= 30) else None
        return (d, pk.data_ptr() if pk is not None else None, fd)

    def set_tensor_cores(self, on):
        """Select the tcgen05 tensor-core kernels (default) or the FP32-FFMA kernels for the decoder and both ODE blocks."""
        self.decoder.use_tensor_cores = bool(on) and 2 * max(self.args.past_length, self.args.future_length) <= 48
        for m in self.modules():
            if isinstance(m, TransformerEncoderLayer):
                m.use_tensor_cores = bool(on)

    def invalidate_packed(self):
        """Call after changing weights in a way torch cannot see (writes through raw pointers, CUDA-graph replays of an update):
        the next forward re-splits them into the tensor-core operand images (decoder and both ODE blocks)."""
        for m in self.modules():
            pk = getattr(m, "_packed", None)
            if pk is not None:
                pk.invalidate()

    def tc_status(self):
        return self.decoder._packed.status(self.args.past_length, self.args.future_length)

    # ---- data staging (reference call sequence) ---------------------------------------------------
    def set_data(self, batch, pre_motion, fut_motion, pre_motion_mask, fut_motion_mask):
        """ETH/UCY/SDD: pre_motion [N,2,T_p], fut_motion [N,2,T_f] as the loaders yield them (model/STTODE.py:397-461).
        Stages the explicit tensors the reference stages (`inputs`, `inputs_for_posterior`, `past_traj`, `future_traj`,
        `cur_location`, `scene_orig`); in train() mode with the reference's agent subsampling (np.random.choice, > max_train_agent)
        and random scene rotation.  `inference()` itself only needs the raw scene (normalisation and velocities run inside
        sttode_forecast), so fut_motion may be None there."""
        self.batch_size = 1
        self._mode = "scene"
        dev = self.device
        pre = pre_motion.to(dev)
        self._obs = _f32c(pre, "pre_motion")                       # inference() hands the raw scene to sttode_forecast
        self.pre_motion_mask, self.fut_motion_mask = pre_motion_mask, fut_motion_mask
        if fut_motion is None:                                     # inference-only staging
            self.agent_num = self._obs.shape[0]
            self.fut_motion_orig = None
            self._features_staged = False
            return
        fut = fut_motion.to(dev)
        fut_orig = fut.transpose(1, 2)
        pre, fut = pre.permute(2, 0, 1), fut.permute(2, 0, 1)                         # [T,N,2]
        if self.training and pre.shape[1] > self.max_train_agent:                     # model/STTODE.py:405-413
            ind = torch.tensor(np.random.choice(pre.shape[1], self.max_train_agent).tolist(), device=dev)
            pre, fut = pre.index_select(1, ind).contiguous(), fut.index_select(1, ind).contiguous()
            pre_motion_mask, fut_motion_mask = pre_motion_mask.to(dev).index_select(0, ind), fut_motion_mask.to(dev).index_select(0, ind)
            self.pre_motion_mask, self.fut_motion_mask = pre_motion_mask, fut_motion_mask
            fut_orig = fut_orig.index_select(0, ind).contiguous()
            self._obs = pre.permute(1, 2, 0).contiguous()
        self.agent_num = pre.shape[1]
        self.scene_orig = pre[-1:].view(-1, 2).mean(dim=0)
        if self.rand_rot_scene and self.training:                                     # :419-426
            theta = getattr(self, "theta_override", None)
            if theta is None:
                theta = (torch.randint(high=24, size=(1,)).to(dev) * (np.pi / 12)) if self.discrete_rot else torch.rand(1).to(dev) * np.pi * 2
            pre, pre_norm = _rotation_2d(pre, theta, self.scene_orig)
            fut, fut_norm = _rotation_2d(fut, theta, self.scene_orig)
            fut_orig, _ = _rotation_2d(fut_orig, theta, self.scene_orig)
            self._obs = pre.permute(1, 2, 0).contiguous()
        else:
            pre_norm, fut_norm = pre - self.scene_orig, fut - self.scene_orig
        pre_vel = pre[1:] - pre[:-1]
        pre_vel = torch.cat([pre_vel[:1], pre_vel], dim=0)
        fut_vel = fut - torch.cat([pre[-1:], fut[:-1]])
        self.fut_motion_orig = fut_orig
        self.inputs = torch.cat([pre_norm, pre_vel], dim=-1).permute(1, 0, 2).contiguous()
        self.inputs_for_posterior = torch.cat([fut_norm, fut_vel], dim=-1).permute(1, 0, 2).contiguous()
        self.past_traj = pre_norm.permute(1, 0, 2).contiguous()
        self.future_traj = fut_norm.permute(1, 0, 2).contiguous()
        self.cur_location = self.past_traj[:, -1:]
        self._features_staged = True

    def set_data_nba(self, data):
        """NBA: data['past_traj'] [B,N,T_p,2] (model/STTODE.py:463-486)."""
        self.data = data
        self.batch_size, self.agent_num = data["past_traj"].shape[0], data["past_traj"].shape[1]
        self._mode = "nba"

    def _seed(self):
        return int(torch.randint(0, 2 ** 62, (1,)).item())      # follows torch.manual_seed like the reference's randn_like

    def _forecast(self, traj, layout, scene_ptr, row_scene, n_scenes, rows, L, agents, K):
        lib = _lib.lib()
        Tp, Tf = self.args.past_length, self.args.future_length
        dev = traj.device
        pred = torch.empty(rows, K, Tf, 2, device=dev, dtype=torch.float32)
        pf = torch.empty(rows, 128, device=dev, dtype=torch.float32)
        tn = torch.empty(rows, Tp, 2, device=dev, dtype=torch.float32)
        org = torch.empty(max(n_scenes, 1), 2, device=dev, dtype=torch.float32)
        ws = _lib.workspace(lib.sttode_forecast_workspace_bytes(rows, L, K, Tp), dev)
        z = self.z_override
        if z is not None:
            z = _f32c(z.to(dev), "z").reshape(-1, 32)
            if z.shape[0] != rows * K:
                raise _lib.SttodeError(f"z_override has {z.shape[0]} rows, expected rows*K = {rows * K}")
        mw = self._weights()
        _lib.check(lib.sttode_forecast(_lib.dptr(traj), layout, _lib.dptr(scene_ptr), _lib.dptr(row_scene), n_scenes, rows, L, agents,
                                       _lib.dptr(z), self._seed() if z is None else 0, K, C.byref(mw), _lib.dptr(pred), _lib.dptr(pf),
                                       _lib.dptr(tn), _lib.dptr(org), _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        self.past_feature, self.past_traj, self._origin = pf, tn, org
        return pred

    @torch.no_grad()
    def inference(self, data=None):
        """-> [K, B*N, T_f, 2] like the reference (model/STTODE.py:574-623)."""
        K = self.args.sample_k
        Tp = self.args.past_length
        if self.args.dataset == "nba":
            src = data if data is not None else self.data
            past = _f32c(src["past_traj"].to(self.device), "past_traj")
            B, N = past.shape[0], past.shape[1]
            rows = B * N
            pred = self._forecast(past.view(rows, Tp, 2), TRAJ_RT2, None, None, 0, rows, B, N, K)
            self.scene_orig = self.past_traj                               # model/STTODE.py:473
        else:
            obs = self._obs
            n = obs.shape[0]
            sp = torch.tensor([0, n], dtype=torch.int32, device=obs.device)
            rs = torch.zeros(n, dtype=torch.int32, device=obs.device)
            pred = self._forecast(obs, TRAJ_R2T, sp, rs, 1, n, 1, n, K)
            self.scene_orig = self._origin[0]
        self.cur_location = self.past_traj[:, -1:]
        return pred.permute(1, 0, 2, 3)

    @torch.no_grad()
    def forecast_scenes(self, obs_traj, scene_ptr, layout=TRAJ_R2T, sample_k=None):
        """Batched ETH/UCY/SDD inference: many scenes per launch (no reference equivalent; each scene is
        processed exactly as one `set_data` + `inference` call).  obs_traj [sum N_i, 2, T_p] (or [.., T_p, 2]
        with layout=TRAJ_RT2), scene_ptr int32 [n_scenes+1] CSR offsets.  -> pred [sum N_i, K, T_f, 2]."""
        K = self.args.sample_k if sample_k is None else sample_k
        obs = _f32c(obs_traj.to(self.device), "obs_traj")
        sp = scene_ptr.to(device=obs.device, dtype=torch.int32).contiguous()
        n_scenes, rows = sp.numel() - 1, obs.shape[0]
        counts = (sp[1:] - sp[:-1]).long()
        rs = torch.repeat_interleave(torch.arange(n_scenes, device=obs.device, dtype=torch.int32), counts).contiguous()
        if rs.numel() != rows:
            raise _lib.SttodeError("scene_ptr does not cover obs_traj")
        return self._forecast(obs, layout, sp, rs, n_scenes, rows, 1, 1, K)

    @torch.no_grad()
    def evaluate_host(self, obs_traj, gt_traj, scene_ptr=None, seed=None, sample_k=None):
        """The body of the reference's evaluation loop (test.py:171-208 / :495-552) for a batch of scenes as ONE library call on HOST (CPU) tensors:
        H2D, set_data + inference, min-over-K ADE/FDE on the device, three doubles back — the K samples never leave HBM (sttode_evaluate_host).
        ETH/SDD: obs_traj [sum N_i, 2, T_p], gt_traj [sum N_i, 2, T_f] as the loader yields them, scene_ptr int32 [n_scenes+1] CSR offsets.
        NBA (scene_ptr None): obs_traj [B, N, T_p, 2], gt_traj [B, N, T_f, 2], one reference batch (the B scenes attend to each other).
        -> (sum of minADE over agents, sum of minFDE, agents): `ade_meter.update(sum / agents, n=agents)` reproduces test.py:204-208."""
        K = self.args.sample_k if sample_k is None else sample_k
        Tp, Tf = self.args.past_length, self.args.future_length
        if obs_traj.is_cuda or gt_traj.is_cuda:
            raise _lib.SttodeError("evaluate_host takes host tensors; for device tensors use forecast_scenes / inference + sttode_min_ade_fde")
        obs = obs_traj.detach().to(torch.float32).contiguous()
        gt = gt_traj.detach().to(torch.float32).contiguous()
        if scene_ptr is not None:
            sp = scene_ptr.detach().to(device="cpu", dtype=torch.int32).contiguous()
            rows, n_scenes, L, agents, layout, gt_layout = obs.shape[0], sp.numel() - 1, 1, 1, TRAJ_R2T, 1
            if tuple(obs.shape) != (rows, 2, Tp) or tuple(gt.shape) != (rows, 2, Tf):
                raise _lib.SttodeError(f"evaluate_host: expected obs [rows,2,{Tp}] and gt [rows,2,{Tf}], got {tuple(obs.shape)} / {tuple(gt.shape)}")
        else:
            sp = None
            B, N = obs.shape[0], obs.shape[1]
            rows, n_scenes, L, agents, layout, gt_layout = B * N, 0, B, N, TRAJ_RT2, 0
            if tuple(obs.shape) != (B, N, Tp, 2) or tuple(gt.shape) != (B, N, Tf, 2):
                raise _lib.SttodeError(f"evaluate_host: expected obs [B,N,{Tp},2] and gt [B,N,{Tf},2], got {tuple(obs.shape)} / {tuple(gt.shape)}")
        sums = (C.c_double * 3)()
        mw = self._weights()
        with torch.cuda.device(next(self.parameters()).device):
            _lib.check(_lib.lib().sttode_evaluate_host(obs.data_ptr(), gt.data_ptr(), layout, gt_layout, sp.data_ptr() if sp is not None else None, n_scenes,
                                                       rows, L, agents, self._seed() if seed is None else int(seed), K, C.byref(mw), sums, _lib.stream_ptr()))
        return sums[0], sums[1], int(sums[2])

    def capture_forecast(self, rows, L, agents, n_scenes=0, layout=TRAJ_R2T, sample_k=None):
        """CUDA-graph version of `forecast_scenes` (n_scenes > 0, L = 1) or of NBA-style inference (n_scenes = 0) for a fixed shape."""
        return ForecastGraph(self, rows, L, agents, n_scenes, layout, sample_k).capture()

    # ---- training objective: the reference's stage methods (model/STTODE.py:488-568), also driven by sampler.py ------------------
    def _stage_features(self):
        """Make `inputs`, `inputs_for_posterior`, `past_traj`, `future_traj`, `cur_location` current for the staged batch."""
        if getattr(self, "_mode", None) == "nba":                                   # model/STTODE.py:463-481
            dev, B, N = self.device, self.batch_size, self.agent_num
            past = _f32c(self.data["past_traj"].to(dev), "past_traj").view(B * N, self.args.past_length, 2)
            fut = _f32c(self.data["future_traj"].to(dev), "future_traj").view(B * N, self.args.future_length, 2)
            vel = past[:, 1:] - past[:, :-1]
            self.inputs = torch.cat((past, torch.cat([vel[:, :1], vel], dim=1)), dim=-1)
            fvel = fut - torch.cat([past[:, -1:], fut[:, :-1]], dim=1)
            self.inputs_for_posterior = torch.cat((fut, fvel), dim=-1)
            self.scene_orig = self.past_traj = past
            self.future_traj, self.cur_location = fut, past[:, -1:]
        elif not (getattr(self, "_mode", None) == "scene" and getattr(self, "_features_staged", False)):
            raise NotImplementedError("STTODENet: stage the batch with set_data_nba, or with set_data(..., fut_motion) first")

    def _grad_ctx(self):
        # autograd history exists wherever a sub-module is in train() mode (train.py trains everything; trainsampler.py:135 puts
        # only the decoder in train() so gradients reach the sampler through z)
        on = self.training or self.decoder.training
        return torch.set_grad_enabled(on and torch.is_grad_enabled())

    def encode_history(self):
        """model/STTODE.py:488-496."""
        with self._grad_ctx():
            self._stage_features()
            self.past_feature = self.past_encoder(self.inputs, self.batch_size, self.agent_num)

    def fu_encoder(self):
        """model/STTODE.py:498-527: q(z | past, future), its sample, and the N(0, I) prior with its (discarded) sample."""
        args, dev = self.args, self.device
        with self._grad_ctx():
            self.qz_param = self.future_encoder(self.inputs_for_posterior, self.batch_size, self.agent_num, self.past_feature)
            self.qz_dis = Normal(params=self.qz_param.repeat_interleave(20, dim=0))
            self.qz_distribution = Normal(params=self.qz_param)
            eps = getattr(self, "eps_override", None)
            sigma = self.qz_distribution.sigma
            self.qz_sampled = self.qz_distribution.mu + (eps[0].to(dev) if eps is not None else torch.randn_like(sigma)) * sigma
            R = self.past_feature.shape[0]
            self.pz_distribution = Normal(mu=torch.zeros(R, args.zdim, device=dev), logvar=torch.zeros(R, args.zdim, device=dev))
            self.pz_sampled = torch.randn_like(sigma) if eps is None else None       # drawn and overwritten in the reference too (:525,549)

    def decoder_future_0(self, qz_sampled):
        """model/STTODE.py:534-551: K = 1 decode (prediction + reconstruction), then the prior samples of the best-of-20 term."""
        args, dev = self.args, self.device
        with self._grad_ctx():
            self.pred_traj, self.recover_traj = self.decoder(self.past_feature, qz_sampled, self.batch_size, self.agent_num, self.past_traj,
                                                             self.cur_location, sample_num=1)
            self.past_feature_repeat = self.past_feature.repeat_interleave(20, dim=0)
            R20 = self.past_feature_repeat.shape[0]
            self.pz_dis = Normal(mu=torch.zeros(R20, args.zdim, device=dev), logvar=torch.zeros(R20, args.zdim, device=dev))
            eps = getattr(self, "eps_override", None)
            self.pz_sampled = eps[1].to(dev) if eps is not None else torch.randn(R20, args.zdim, device=dev)

    def decoder_future_1(self, pz_sampled):
        """model/STTODE.py:529-532: the K = 20 decode."""
        with self._grad_ctx():
            self.diverse_pred_traj, self.attn_weights = self.decoder(self.past_feature_repeat, pz_sampled, self.batch_size, self.agent_num,
                                                                     self.past_traj, self.cur_location, sample_num=20, mode="inference")

    def objective(self):
        """The reference's training objective (model/STTODE.py:553-568) on the staged batch as tensors, with no host sync (usable under
        CUDA-graph capture): (total_loss, stack of [loss_pred, loss_recover, loss_kl, loss_diverse])."""
        args = self.args
        self.encode_history()
        self.fu_encoder()
        self.decoder_future_0(self.qz_sampled)
        with self._grad_ctx():
            B, N, Tp, Tf = self.batch_size, self.agent_num, args.past_length, args.future_length
            loss_pred = (self.future_traj - self.pred_traj).pow(2).sum() / B / Tf                  # :372-376
            loss_recover = (self.past_traj - self.recover_traj).pow(2).sum() / B / Tp              # :384-387
            loss_kl = (self.qz_distribution.kl(self.pz_distribution).sum() / (B * N)).clamp_min(args.min_clip)     # :378-382
        self.decoder_future_1(self.pz_sampled)
        with self._grad_ctx():
            diff = self.future_traj.unsqueeze(1) - self.diverse_pred_traj                          # :389-395
            loss_diverse = diff.pow(2).sum(dim=-1).sum(dim=-1).min(dim=1)[0].mean()
            total = loss_pred + loss_recover + loss_kl + loss_diverse
            parts = torch.stack((loss_pred.detach(), loss_recover.detach(), loss_kl.detach(), loss_diverse.detach()))
        return total, parts

    def forward(self):
        """(total_loss, loss_pred, loss_recover, loss_kl, loss_diverse) like the reference.  train(): total_loss carries autograd history
        (call .backward() on it as train.py:65 does).  eval(): forward-only, every stage inside the library.
        `self.eps_override = (eps_q [R,zdim], eps_p [R*20,zdim])` injects the two live rsample draws (tests)."""
        total, parts = self.objective()
        return (total, *parts.tolist())                                                            # one sync for the four floats

    def step_annealer(self):
        for anl in self.param_annealers:
            anl.step()
