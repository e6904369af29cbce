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
lib = L.lib()
def run(mode):
    m.forecast_scenes(obs, sp); torch.cuda.synchronize()
    buf = m.decoder._packed.buf
    L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, mode, None))
    lib.sttode_profile_enable(1)
    for _ in range(3): m.forecast_scenes(obs, sp)
    torch.cuda.synchronize()
    ms, n = C.c_double(), C.c_ulonglong()
    L.check(lib.sttode_profile_read(C.byref(ms), C.byref(n)))
    L.check(lib.sttode_decoder_timing(L.dptr(buf), 8, 12, 0, None))
    return ms.value / n.value
print("normal  k_dec_pair_tc ms:", run(0))
print("no-copy k_dec_pair_tc ms:", run(4))
