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
d = lambda a, b: (v[b] - v[a]) if v[a] and v[b] else None
# phase order of the final kernel at these shapes: x0, y0, conv + GRU, y1, output (stamps 24/25 x0, 26/27 y0, 27..6 conv input + GRU, 6..28 y1, 28/29 output)
print("epilogue phases (cycles): x0", d(24, 25), " y0", d(26, 27), " conv input + gru", d(27, 6), " y1", d(6, 28), " output", d(28, 29), " total", d(24, 29))
print("gru step t=2, epilogue thread 0: conv image t+1", d(0, 1), " wait half A", d(1, 2), " gates A + stage (-> HA signalled)", d(2, 3))
print("gru step t=2, issuer (relative to the epilogue's iteration start): before GREAD wait", d(0, 20), " GREAD ok", d(0, 21), " HA ok", d(0, 22), " HB ok", d(0, 7))
print("gru step t=2 epilogue stamps from iteration start: conv", d(0, 1), " A ok", d(0, 2), " q0 ld", d(0, 10), " q0 math", d(0, 11), " q1 ld", d(0, 12), " q1 math(+GREAD)", d(0, 13),
      " B ok", d(0, 16), " hA stored", d(0, 17), " HA signalled", d(0, 3), " q2 ld", d(0, 14), " q2 math", d(0, 15), " HB signalled", d(0, 18))
print("GRU phase from the end of y0 (27): conv input built", d(27, 4), " h zeroed / loop start", d(27, 5), " 8 steps done", d(27, 6))
print("y0 MLP from phase start (26): chunk loop done", d(26, 8), " D2 complete", d(26, 9), " pieces staged", d(26, 19), " D3 complete", d(26, 23), " read out + barrier", d(26, 30))
