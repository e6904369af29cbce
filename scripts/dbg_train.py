import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from sttode_b200 import synth
from sttode_b200.net import STTODENet
g=np.load('/root/repo/tests/golden/training_objective.npz')
dev=torch.device('cuda:0')
B,tp,tf=int(g['B']),10,20
data=synth.nba_batch(1,B,11,tp,tf)
gq=torch.Generator().manual_seed(21)
eps_q,_,eps_p=torch.randn(B*11,32,generator=gq),torch.randn(B*11,32,generator=gq),torch.randn(B*11*20,32,generator=gq)
for tc in (True, False):
    m=STTODENet(synth.make_args('nba',tp,tf),dev); m.load_state_dict(synth.make_state_dict(2,tp,tf)); m.set_device(dev); m.eval()
    m.decoder.use_tensor_cores=tc
    m.eps_override=(eps_q,eps_p); m.set_data_nba(data)
    out=m.forward()
    print('tc' if tc else 'ffma', float(out[0]), out[1:], 'max|z|', float(m.qz_sampled.abs().max()))
print('ref ', float(g['total']), float(g['pred']), float(g['recover']), float(g['kl']), float(g['diverse']))
