"""tcgen05 building blocks (csrc/tc.cuh): operand layout, descriptors, TMEM round trip and the 3xFP16 split."""
import pytest
import torch

import sttode_b200._lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(16, 16), (64, 32), (256, 128), (48, 96), (192, 128)])
def test_umma_selftest(N, K):
    dev = torch.device("cuda:0")
    lib = L.lib()
    g = torch.Generator().manual_seed(N * 1000 + K)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g) * 0.1
    Ad, Bd = A.to(dev), B.to(dev)
    for passes, tol in ((1, 2e-6), (3, 2e-6)):
        C = torch.full((128, N), float("nan"), device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        L.check(lib.sttode_umma_selftest(L.dptr(Ad), L.dptr(Bd), L.dptr(C), N, K, passes, 0, L.dptr(st), L.stream_ptr()))
        torch.cuda.synchronize()
        assert int(st.item()) == 0, "tcgen05.mma never completed"
        ref = (A.half().double() @ B.half().double().T) if passes == 1 else (A.double() @ B.double().T)
        err = ((C.cpu().double() - ref).abs().max() / ref.abs().max()).item()
        assert err < tol, (passes, err)


@pytest.mark.parametrize("N,K", [(16, 16), (256, 32), (32, 128), (96, 64)])
def test_umma_selftest_a_from_tmem(N, K):
    """tcgen05.mma with the A operand in tensor memory (lane = row, two consecutive-k halves per 32-bit column, written with tcgen05.st): the
    form the decoder's layer 2 / layer 3 use so that relu(D + b) never leaves TMEM."""
    dev = torch.device("cuda:0")
    lib = L.lib()
    g = torch.Generator().manual_seed(N * 1000 + K + 7)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g) * 0.1
    Ad, Bd = A.to(dev), B.to(dev)
    for passes, tol in ((1, 2e-6), (3, 2e-6)):
        C = torch.full((128, N), float("nan"), device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        L.check(lib.sttode_umma_selftest(L.dptr(Ad), L.dptr(Bd), L.dptr(C), N, K, passes, 2, L.dptr(st), L.stream_ptr()))
        torch.cuda.synchronize()
        assert int(st.item()) == 0, "tcgen05.mma never completed"
        ref = (A.half().double() @ B.half().double().T) if passes == 1 else (A.double() @ B.double().T)
        err = ((C.cpu().double() - ref).abs().max() / ref.abs().max()).item()
        assert err < tol, (passes, err)


# ------------------------------------------------------------------ k_gemm_tc: the training backward's GEMM on the tensor cores (bf16x3 split)
@pytest.mark.parametrize("M,N,K,form", [(7040, 512, 256, "nt"), (7040, 256, 512, "nn"), (512, 256, 7040, "tn"), (352, 1024, 64, "nt"), (130, 24, 257, "nt"),
                                        (64, 640, 3520, "tn"), (1000, 96, 100, "nn")])
@pytest.mark.parametrize("scale", [1.0, 1e-6])
def test_gemm_tc_vs_float64(M, N, K, form, scale):
    """sttode_gemm(mode = tensor cores) against a float64 product, NT / NN / TN operand strides (forward, input-gradient and weight-gradient
    forms), ragged sizes, the full epilogue (bias, ReLU mask, residual, accumulate), split-K with a workspace — and operands of magnitude
    1e-6 (gradient-like): the bf16x3 split keeps fp32-level relative accuracy there, which an fp16 hi/lo split would not."""
    import ctypes as C
    import sttode_b200._lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g) * scale
    b = torch.randn(K, N, generator=g)
    bias = torch.randn(N, generator=g) * scale
    res = torch.randn(M, N, generator=g) * scale
    mask = torch.randn(M, N, generator=g)
    c0 = torch.randn(M, N, generator=g) * scale
    if form == "nt":       # A [M,K] row-major, B stored as W [N,K]
        A, sam, sak = a.contiguous(), K, 1
        Bm, sbk, sbn = b.t().contiguous(), 1, K
    elif form == "nn":     # A row-major, B [K,N] row-major
        A, sam, sak = a.contiguous(), K, 1
        Bm, sbk, sbn = b.contiguous(), N, 1
    else:                  # A stored transposed [K,M] (dy^T), B [K,N] row-major
        A, sam, sak = a.t().contiguous(), 1, M
        Bm, sbk, sbn = b.contiguous(), N, 1
    lib = L.lib()
    ws = L.scratch(lib.sttode_dense_bwd_workspace_bytes(), dev, "splitk")
    ref_plain = a.double() @ b.double()
    ref_full = ref_plain + bias.double()
    ref_full = torch.where(mask.double() > 0, ref_full, torch.zeros_like(ref_full)) + res.double() + c0.double()
    for mode in (1, 0):
        Ad, Bd = A.to(dev), Bm.to(dev)
        out = torch.empty(M, N, device=dev)
        L.check(lib.sttode_gemm(M, N, K, L.dptr(Ad), sam, sak, L.dptr(Bd), sbk, sbn, L.dptr(out), N, None, None, 0, None, 0, 0, 0, mode,
                                L.dptr(ws), ws.numel(), L.stream_ptr()))
        err = float((out.cpu().double() - ref_plain).abs().max() / ref_plain.abs().max())
        assert err < (1e-5 if mode == 1 else 2e-6), (mode, "plain", err)          # the tensor cores accumulate with truncation: ~2e-6 at K = 256
        out2, bias_d, res_d, mask_d = c0.to(dev).clone(), bias.to(dev), res.to(dev), mask.to(dev)      # keep the device tensors alive across the call
        L.check(lib.sttode_gemm(M, N, K, L.dptr(Ad), sam, sak, L.dptr(Bd), sbk, sbn, L.dptr(out2), N, L.dptr(bias_d), L.dptr(res_d), N,
                                L.dptr(mask_d), N, 0, 1, mode, L.dptr(ws), ws.numel(), L.stream_ptr()))
        err2 = float((out2.cpu().double() - ref_full).abs().max() / ref_full.abs().max())
        assert err2 < (1e-5 if mode == 1 else 2e-6), (mode, "epilogue", err2)
