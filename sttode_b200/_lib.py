"""ctypes binding of libsttode_sm100.so (include/sttode.h).  No CPU fallback: a missing library or a
non-sm_100 device raises, it never silently routes elsewhere."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsttode_sm100.so")

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)


class MhgsaWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b")]


class OdeWeights(C.Structure):
    _fields_ = [("attn", MhgsaWeights)] + [(n, C.c_void_p) for n in (
        "info_w", "info_b", "gate_w", "gate_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
        "norm1_g", "norm1_b", "norm2_g", "norm2_b", "packed")]


class MhgsaGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b")]


class OdeGrads(C.Structure):
    _fields_ = [("attn", MhgsaGrads)] + [(n, C.c_void_p) for n in (
        "info_w", "info_b", "gate_w", "gate_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
        "norm1_g", "norm1_b", "norm2_g", "norm2_b")]


class EncoderWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("fc_w", "fc_b", "pe", "pefc_w", "pefc_b", "fc2_w", "fc2_b", "fc3_w", "fc3_b")] + [("ode", OdeWeights), ("folded", C.c_void_p)]


class Mlp3Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w1", "b1", "w2", "b2")]


class DecomposeWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv_w", "conv_b", "gru_w_ih", "gru_w_hh", "gru_b_ih", "gru_b_hh")] + \
               [("dec_y", Mlp3Weights), ("dec_x", Mlp3Weights)]


class Mlp3Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w1", "b1", "w2", "b2")]


class DecomposeGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv_w", "conv_b", "gru_w_ih", "gru_w_hh", "gru_b_ih", "gru_b_hh")] + \
               [("dec_y", Mlp3Grads), ("dec_x", Mlp3Grads)]


class ModelWeights(C.Structure):
    _fields_ = [("past_length", C.c_int), ("future_length", C.c_int), ("past", EncoderWeights), ("dec", DecomposeWeights * 2),
                ("dec_packed", C.c_void_p), ("rng_state", C.c_void_p)]


# every symbol include/sttode.h declares: name -> (restype, argtypes)
_V, _I, _F, _Z, _U64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_ulonglong
_ROW1 = [_V, _V, _I, _I, _F, _V]
_ROW2 = [_V, _V, _V, _I, _I, _F, _V]
SIGNATURES = {
    "sttode_version": (_I, []),
    "sttode_last_error": (C.c_char_p, []),
    "sttode_device_check": (_I, []),
    "sttode_launch_count": (_U64, []),
    "sttode_profile_enable": (_I, [_I]),
    "sttode_profile_read": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "sttode_umma_selftest": (_I, [_V, _V, _V, _I, _I, _I, _I, _V, _V]),
    "sttode_umma_bench": (_I, [_I, _I, _V, _V]),
    "sttode_umma_pattern": (_I, [_I, _I, _I, _I, _V, _V]),
    "sttode_ode_pack_bytes": (_Z, []),
    "sttode_ode_pack": (_I, [C.POINTER(OdeWeights), _V, _Z, _V]),
    "sttode_encoder_fold_bytes": (_Z, [_I]),
    "sttode_encoder_fold": (_I, [C.POINTER(EncoderWeights), _I, _V, _Z, _V]),
    "sttode_mhgsa_workspace_bytes": (_Z, [_I, _I, _I]),
    "sttode_mhgsa_fwd": (_I, [_V, _V, _V, C.POINTER(MhgsaWeights), _V, _V, _I, _I, _I, _I, _I, _I, _F, _F, _V, _Z, _V]),
    "sttode_linear64": (_I, [_V, _V, _V, _V, _I, _I, _V]),
    "sttode_ode_block_workspace_bytes": (_Z, [_I, _I, _I]),
    "sttode_ode_block_fwd": (_I, [_V, C.POINTER(OdeWeights), _V, _I, _I, _I, _I, _F, _I, _I, _V, _Z, _V]),
    "sttode_ode_rhs_fwd": (_I, [_V, C.POINTER(OdeWeights), _V, _I, _I, _V, _Z, _V]),
    "sttode_hypattention_fwd": (_I, [_V, C.POINTER(OdeWeights), _V, _I, _I, _V, _Z, _V]),
    "sttode_mhgsa_bwd_workspace_bytes": (_Z, [_I, _I, _I]),
    "sttode_mhgsa_bwd": (_I, [_V, _V, _V, C.POINTER(MhgsaWeights), _V, _V, _V, _V, C.POINTER(MhgsaGrads), _I, _I, _I, _I, _I, _I, _F, _V, _Z, _V]),
    "sttode_ode_rhs_bwd_workspace_bytes": (_Z, [_I, _I]),
    "sttode_ode_rhs_bwd": (_I, [_V, _V, C.POINTER(OdeWeights), _V, C.POINTER(OdeGrads), _I, _I, _I, _V, _Z, _V]),
    "sttode_ode_block_bwd_workspace_bytes": (_Z, [_I, _I, _I]),
    "sttode_ode_block_bwd": (_I, [_V, _V, _I, C.POINTER(OdeWeights), _V, C.POINTER(OdeGrads), _I, _I, _I, _F, _I, _I, _V, _Z, _V]),
    "sttode_past_encoder_workspace_bytes": (_Z, [_I, _I]),
    "sttode_past_encoder_fwd": (_I, [_V, _I, _V, _V, _I, _I, _I, _I, C.POINTER(ModelWeights), _V, _V, _V, _V, _Z, _V]),
    "sttode_encoder_fwd": (_I, [_V, _I, _I, _I, _I, C.POINTER(EncoderWeights), _V, _V, _Z, _V]),
    "sttode_linear": (_I, [_V, _V, _V, _V, _I, _I, _I, _I, _V]),
    "sttode_decoder_workspace_bytes": (_Z, [_I, _I]),
    "sttode_decoder_fwd": (_I, [_V, _V, _V, _U64, _V, _V, _I, _I, C.POINTER(ModelWeights), _V, _V, _V, _Z, _V]),
    "sttode_dense_fwd": (_I, [_V, _I, _V, _V, _V, _I, _I, _I, _I, _V]),
    "sttode_gemm": (_I, [_I, _I, _I, _V, C.c_long, C.c_long, _V, C.c_long, C.c_long, _V, C.c_long, _V, _V, C.c_long, _V, C.c_long, _I, _I, _I, _V, _Z, _V]),
    "sttode_dense_bwd_workspace_bytes": (_Z, []),
    "sttode_dense_bwd": (_I, [_V, _I, _V, _V, _V, _I, _V, _V, _I, _I, _I, _V, _Z, _V]),
    "sttode_decoder_bwd_workspace_bytes": (_Z, [_I, _I, _I]),
    "sttode_decoder_bwd": (_I, [_V, _V, _V, _V, _V, _I, C.POINTER(ModelWeights), C.POINTER(DecomposeGrads), _V, _V, _V, _Z, _V]),
    "sttode_decoder_pack_bytes": (_Z, [_I, _I]),
    "sttode_decoder_pack": (_I, [C.POINTER(ModelWeights), _V, _Z, _V]),
    "sttode_decoder_pack_status": (_I, [_V, _I, _I, C.POINTER(C.c_int)]),
    "sttode_decoder_timing": (_I, [_V, _I, _I, _I, C.POINTER(C.c_longlong)]),
    "sttode_forecast_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "sttode_forecast": (_I, [_V, _I, _V, _V, _I, _I, _I, _I, _V, _U64, _I, C.POINTER(ModelWeights), _V, _V, _V, _V, _V, _Z, _V]),
    "sttode_forecast_host": (_I, [_V, _I, _V, _I, _I, _I, _I, _U64, _I, C.POINTER(ModelWeights), _V, _V]),
    "sttode_evaluate_host": (_I, [_V, _V, _I, _I, _V, _I, _I, _I, _I, _U64, _I, C.POINTER(ModelWeights), C.POINTER(C.c_double), _V]),
    "sttode_release_host_arena": (None, []),
    "sttode_host_output_mode": (_I, [_I]),
    "sttode_peer_alloc": (_I, [_Z, C.POINTER(C.c_void_p), _V]),
    "sttode_peer_open": (_I, [_V, C.POINTER(C.c_void_p)]),
    "sttode_peer_close": (_I, [_V]),
    "sttode_peer_free": (_I, [_V]),
    "sttode_dp_adam_step": (_I, [_V, _V, _I, _I, _V, _V, _V, _V, _Z, _V, _F, _F, _F, _V]),
    "sttode_peer_status": (_I, [_V, _Z, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "sttode_latent_draw": (_I, [_U64, _U64, _V, _Z, _V, _V]),
    "sttode_min_ade_fde": (_I, [_V, _V, _I, _I, _I, _I, _V, _V]),
    "sttode_oblique_proj": (_I, [_V, _V, _I, _I, _V]),
    "sttode_oblique_dist": (_I, [_V, _V, _V, _I, _I, _I, _I, _F, _V]),
    "sttode_pmath_project": (_I, _ROW1), "sttode_pmath_lambda_x": (_I, _ROW1), "sttode_pmath_dist0": (_I, _ROW1),
    "sttode_pmath_expmap0": (_I, _ROW1), "sttode_pmath_logmap0": (_I, _ROW1),
    "sttode_pmath_mobius_add": (_I, _ROW2), "sttode_pmath_dist": (_I, _ROW2),
    "sttode_pmath_expmap": (_I, _ROW2), "sttode_pmath_logmap": (_I, _ROW2),
    "sttode_pmath_mobius_matvec": (_I, [_V, _V, _V, _I, _I, _I, _F, _V]),
    "sttode_pmath_dist_matrix": (_I, [_V, _V, _V, _I, _I, _I, _F, _V]),
    "sttode_pmath_tanh": (_I, [_V, _V, _Z, _F, _V]),
    "sttode_pmath_artanh": (_I, [_V, _V, _Z, _V]),
    "sttode_pmath_arsinh": (_I, [_V, _V, _Z, _V]),
}

_lib = None
_device_ok = False


class SttodeError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every declared symbol (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SttodeError(f"{LIB_PATH} is missing: build it with `python -m sttode_b200.build` "
                              "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def lib():
    """The bound library, after checking that the current device is a B200-class (sm_100) GPU."""
    global _device_ok
    l = load()
    if not _device_ok:
        if not torch.cuda.is_available():
            raise SttodeError("libsttode_sm100 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
        check(l.sttode_device_check())
        _device_ok = True
    return l


def check(rc):
    if rc != 0:
        raise SttodeError(f"libsttode_sm100 error {rc}: {load().sttode_last_error().decode()}")


def dptr(t, name="tensor"):
    """Device pointer of a contiguous fp32 (or int32) CUDA tensor; None -> NULL."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SttodeError(f"{name} must be a CUDA tensor (got {t.device}); this path has no CPU implementation")
    if t.dtype not in (torch.float32, torch.int32, torch.float64, torch.uint8):
        raise SttodeError(f"{name} has dtype {t.dtype}")
    if not t.is_contiguous():
        raise SttodeError(f"{name} must be contiguous")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


_scratch = {}


def scratch(nbytes, device, tag="default"):
    """A per-(device, tag) scratch buffer that only grows: kernels that need a fixed-size workspace on every call (split-K partial sums) share it, and
    its pointer stays valid across CUDA-graph replays.  Single-stream use only (the library's calls are stream-ordered)."""
    key = (str(device), tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def launch_count():
    return int(load().sttode_launch_count())


# ---- direct gradient sinks -------------------------------------------------------------------------------------------------------------
# The library's backward kernels ACCUMULATE parameter gradients (+=).  When the parameters' `.grad` tensors already exist (parallel.FlatGradReducer
# makes them views into one flat buffer) the autograd nodes let the kernels write straight into them and return no parameter gradient, instead of
# filling a zero tensor per parameter and having autograd add it to `.grad` afterwards (two extra launches per parameter per node).
DIRECT_GRAD_SINKS = False


def grad_sinks(params):
    """The `.grad` tensors of `params` if every one is a preallocated contiguous fp32 CUDA tensor of the parameter's shape (and sinks are enabled)."""
    if not DIRECT_GRAD_SINKS or os.environ.get("STTODE_NO_GRAD_SINKS"):
        return None
    sinks = []
    for p in params:
        g = getattr(p, "grad", None)
        if g is None or not g.is_cuda or g.dtype != torch.float32 or g.shape != p.shape or not g.is_contiguous() or not p.requires_grad:
            return None
        sinks.append(g)
    return sinks


def sinks_still_valid(params, sinks):
    """True if every parameter's `.grad` is still the tensor recorded at forward time (zero_grad(set_to_none=True) would break the link)."""
    return sinks is not None and all(p.grad is not None and p.grad.data_ptr() == s.data_ptr() for p, s in zip(params, sinks))
