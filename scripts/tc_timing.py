"""Per-phase cycle accounting of one k_dec_pair_tc tile (CTA 0): run on the GPU box."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sttode_b200._lib as L
from sttode_b200 import synth
from sttode_b200.net import STTODENet
dev = torch.device("cuda:0")
m = STTODENet(synth.make_args("eth", 8, 12), dev); m.load_state_dict(synth.make_state_dict(1, 8, 12)); m.set_device(dev); m.eval()
scenes = synth.eth_scenes(384, cycle=(10, 20, 30, 40, 50, 60))
obs = torch.cat([o for o, _ in scenes]).to(dev)
sp = torch.tensor(np.concatenate([[0], np.cumsum([o.shape[0] for o, _ in scenes])]), dtype=torch.int32)
for _ in range(2): m.forecast_scenes(obs, sp)
torch.cuda.synchronize()
buf = m.decoder._packed.buf
lib = L.lib()
out = (C.c_longlong * 31)()
L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 2, None))
m.forecast_scenes(obs, sp); torch.cuda.synchronize()
m.forecast_scenes(obs, sp); torch.cuda.synchronize()
L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 0, out))
v = list(out)
def rel(idx, names):
    base = min(v[i] for i in idx if v[i])
    return {n: v[i] - base for i, n in zip(idx, names) if v[i]}
base = min(x for x in v[0:29] if x)
R = lambda xs: [x - base if x else None for x in xs]
print("issuer ch4 [L1 start, ring(4) ok, MMA1(4) issued, hfull(3) ok, MMA2(3) issued]:", R(v[0:5]))
print("issuer ch5 [L1 start, ring(5) ok, MMA1(5) issued, hfull(4) ok, MMA2(4) issued]:", R(v[5:10]))
print("epi ch4 grp0 [before d1full, d1full ok, tmem ld done, staged+signalled]:", R(v[10:14]))
print("epi ch5 grp1 [before d1full, d1full ok, tmem ld done, staged+signalled]:", R(v[15:19]))
print("producer blk A of chunk 5:", [x - base if x else None for x in v[23:26]], "(loop top, wfree ok, arrived)")
print("producer blk B of chunk 5:", [x - base if x else None for x in v[26:29]])
ph = [v[14], v[19], v[20], v[21], v[22], v[29]]
print("epilogue phases (cycles): x0+prep", ph[1] - ph[0], " gru", ph[2] - ph[1], " y0", ph[3] - ph[2], " y1", ph[4] - ph[3], " output", ph[5] - ph[4], " total", ph[5] - ph[0])
