"""Whole-tile accounting of k_dec_pair_tc (CTA 0 and its successor on the same SM); build with STTODE_TC_TIMELINE=3, run on the GPU box."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sttode_b200._lib as L
from sttode_b200 import synth
from sttode_b200.net import STTODENet
dev = torch.device("cuda:0")
m = STTODENet(synth.make_args("eth", 8, 12), dev); m.load_state_dict(synth.make_state_dict(1, 8, 12)); m.set_device(dev); m.eval()
scenes = synth.eth_scenes(1536, cycle=(10, 20, 30, 40, 50, 60))
obs = torch.cat([o for o, _ in scenes]).to(dev)
sp = torch.tensor(np.concatenate([[0], np.cumsum([o.shape[0] for o, _ in scenes])]), dtype=torch.int32)
for _ in range(3): m.forecast_scenes(obs, sp)
torch.cuda.synchronize()
buf = m.decoder._packed.buf
lib = L.lib()
out = (C.c_longlong * 31)()
for rep in range(2):
    L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 2, None))
    m.forecast_scenes(obs, sp); torch.cuda.synchronize()
    L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 0, out))
    v = list(out)
    t0 = v[0]
    print("sixth tile of CTA 0, cycles from the tile's start: prologue done", v[1] - t0, " epilogue: x0 starts", v[2] - t0, " x0 done", v[3] - t0, " gru done", v[4] - t0,
          " y0 done", v[5] - t0, " y1 done", v[6] - t0, " output written", v[7] - t0, " after the tile's last barrier", v[10] - t0,
          " | tile period (start of this tile -> start of the next one)", (v[8] - v[0]) if v[8] else None)
    print("   tile period of CTA 0 at its tile 5, 15, 25, 35, 45, 55:", [(v[13 + 2 * k] - v[12 + 2 * k]) if v[12 + 2 * k] and v[13 + 2 * k] else None for k in range(6)])
