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
L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 1, None))
m.forecast_scenes(obs, sp); torch.cuda.synchronize()
out = (C.c_longlong * 31)()
L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 0, out))
names = {0: "issuer: L1 mma issue code", 1: "issuer: wait wfree", 2: "issuer: wait wfull", 3: "issuer: wait hfull", 4: "issuer: wait xfull", 5: "issuer: wait d3free",
         6: "issuer: mlp x0", 7: "issuer: gru", 8: "issuer: mlp y0", 9: "issuer: mlp y1", 10: "epi: wait d1full", 11: "epi: wait hfree", 12: "epi: wait acc",
         13: "epi: gru gate math", 20: "issuer: L2 mma issue code", 21: "issuer: gru mma issue code", 16: "epi: mlp x0", 17: "epi: gru", 18: "epi: mlp y0", 19: "epi: mlp y1", 29: "epi thread total", 30: "other waits"}
for i in range(31):
    if out[i]: print(f"{i:2d} {names.get(i, ''):28s} {out[i]:>10d}")

# timeline of chunks 4,5 of MLP y0
L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 2, None))
m.forecast_scenes(obs, sp); torch.cuda.synchronize()
L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 0, out))
ev = ["I0 iter start", "I1 acquired", "I2 MMA1 issued", "I3 hfull(pc) done", "I4 MMA2(pc) issued", "I5 refilled", "E6 before d1full", "E7 d1full done", "E8 tmem ld done", "E9 -", "E10 staged+signalled"]
t0 = min(v for v in out[:22] if v)
for ch in range(2):
    print("chunk", 4 + ch, {ev[e]: out[ch * 11 + e] - t0 for e in range(11) if out[ch * 11 + e]})
