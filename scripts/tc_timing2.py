"""Issuer / epilogue stamps of the y0 MLP of one k_dec_pair_tc tile (CTA 0); build with STTODE_TC_TIMELINE=2, run on the GPU box."""
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
for _ in range(3): m.forecast_scenes(obs, sp)
torch.cuda.synchronize()
buf = m.decoder._packed.buf
lib = L.lib()
out = (C.c_longlong * 31)()
import sys as _s
for flag, name in ((2, "with the weight stream"), (6, "weight copies skipped (results invalid)")):
  L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, flag, None))
  m.forecast_scenes(obs, sp); torch.cuda.synchronize()
  L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 0, out))
  v = list(out)
  print("==", name)
  t0 = v[19]
  print("y0 MLP, issuer lane, cycles from the issuer's phase start:")
  print("  chunk loop starts", v[20] - t0, " layer-2 of chunk ch issued at:", [v[c] - t0 for c in range(16)])
  print("  gaps between chunk issues:", [v[c + 1] - v[c] for c in range(15)])
  print("  sums over the 16 chunks: ring wait", v[16], " CFULL wait", v[17], " issue", v[18])
  print("  L3RDY half 0 ok", v[21] - t0, " half 1 ok", v[22] - t0, " all issued", v[23] - t0)
  print("y0 MLP, epilogue thread 0 (same origin): phase start", v[24] - t0, " D1FULL(group 0) ok", v[25] - t0, " chunk 0 converted + signalled", v[26] - t0,
        " own chunks done", v[27] - t0, " D2 complete", v[28] - t0, " L3RDY signalled", v[29] - t0, " D3 complete", v[30] - t0)
