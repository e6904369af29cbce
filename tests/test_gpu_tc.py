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
