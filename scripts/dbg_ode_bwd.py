import sys; sys.path.insert(0, '/root/repo')
import torch
from oracle import sttode_oracle as O
from sttode_b200 import synth
from sttode_b200.net import STTODENet
dev = torch.device('cuda:0')
torch.manual_seed(3)
m = STTODENet(synth.make_args("nba", 10, 20), torch.device("cpu"))
sd = {k: v.clone() for k, v in m.state_dict().items()}
m.set_device(dev)
enc = m.past_encoder.ODE_Encoder
P = "past_encoder.ODE_Encoder."
L, bsz = 8, 11
for scale in (1.0, 5.0, 20.0):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(L, bsz, 1, 64, generator=g) * scale
    dy = torch.randn(L, bsz, 1, 64, generator=g)
    res = {}
    for dt in (torch.float64, torch.float32):
        s = {k: v.to(dt) for k, v in sd.items()}
        names = [k for k in s if k.startswith(P)]
        for k in names: s[k].requires_grad_(True)
        xx = x.clone().to(dt).requires_grad_(True)
        O.ode_encoder(xx, s, "past_encoder.").backward(dy.to(dt))
        res[dt] = {k[len(P):]: s[k].grad for k in names}; res[dt]["x"] = xx.grad
    enc.zero_grad(set_to_none=True)
    xd = x.to(dev).requires_grad_(True)
    enc(xd).backward(dy.to(dev))
    got = {n: p.grad for n, p in enc.named_parameters()}; got["x"] = xd.grad
    print("scale", scale)
    for n in got:
        r64 = res[torch.float64][n]
        if r64 is None: continue
        den = r64.abs().max().clamp_min(1e-30)
        eg = float((got[n].cpu().double() - r64).abs().max() / den)
        ec = float((res[torch.float32][n].double() - r64).abs().max() / den)
        if eg > 2e-5 or "linear1" in n: print(f"  {n:60s} gpu {eg:.2e} cpu32 {ec:.2e}")
