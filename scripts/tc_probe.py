"""Probe the tcgen05 self-test over several shapes (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sttode_b200._lib as L

dev = torch.device("cuda:0")
lib = L.lib()
g = torch.Generator().manual_seed(0)
for (N, K) in [(16, 16), (64, 32), (256, 128), (32, 128), (96, 96), (192, 128), (48, 64)]:
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g) * 0.1
    Ad, Bd = A.to(dev), B.to(dev)          # keep the device tensors alive across the call
    for passes in (1, 3):
        C = torch.full((128, N), float("nan"), device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        L.check(lib.sttode_umma_selftest(L.dptr(Ad), L.dptr(Bd), L.dptr(C), N, K, passes, 0, L.dptr(st), L.stream_ptr()))
        torch.cuda.synchronize()
        ref = (A.half().double() @ B.half().double().T) if passes == 1 else (A.double() @ B.double().T)
        err = ((C.cpu().double() - ref).abs().max() / ref.abs().max()).item()
        print(f"N={N:3d} K={K:3d} passes={passes} status={int(st.item())} rel_err={err:.3e}", flush=True)
