"""Cycles per tcgen05.mma (M=128, K=16, kind::f16, no-swizzle K-major operands) as a function of N: run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sttode_b200._lib as L
lib = L.lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for n in (16, 32, 48, 64, 96, 128, 192, 256):
    for reps in (64, 256):
        L.check(lib.sttode_umma_bench(n, reps, out.data_ptr(), L.stream_ptr()))
        torch.cuda.synchronize()
        tot, iss = out.tolist()
        print(f"N={n:3d} reps={reps:3d}: total {tot:7d} cyc = {tot / reps:6.1f}/mma   issue loop {iss:7d} = {iss / reps:6.1f}/mma   (ideal {n / 2:.0f})")
for mode, name in ((0x1000, "A from tensor memory (.ts)"), (0x2000, "A from shared memory, 3 warps on tcgen05.ld/st"), (0x3000, ".ts, 3 warps on tcgen05.ld/st")):
    for n in (128, 256):
        L.check(lib.sttode_umma_bench(n | mode, 256, out.data_ptr(), L.stream_ptr()))
        torch.cuda.synchronize()
        tot, iss = out.tolist()
        print(f"{name:48s} N={n:3d}: total {tot / 256:6.1f}/mma   issue loop {iss / 256:6.1f}/mma")
print("GRU-slab pattern probe (cycles per slab group of 3 x N=192 + 3 x N=96; 448 expected from the per-N rates):")
for poll, pname in ((0, "no bystanders"), (1, "16 warps, every lane polls an mbarrier"), (2, "16 warps, one lane polls"), (3, "16 warps, one lane polls with nanosleep(100)")):
    for mode, name in ((0, "rz rz rz n n n"), (1, "rz n rz n rz n")):
        L.check(lib.sttode_umma_pattern(mode, 288, 128, poll, out.data_ptr(), L.stream_ptr()))
        torch.cuda.synchronize()
        tot, iss = out.tolist()
        print(f"  {pname:46s} order {name}: {tot / 128:7.1f} cycles per group")
