"""Per-parameter gradient error of one training step: GPU path and CPU-fp32 oracle, both against the float64 oracle."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from oracle import sttode_oracle as O
from sttode_b200 import synth
from sttode_b200.net import STTODENet
from training_common import training_draws
dev = torch.device('cuda:0')
dataset, tp, tf, B, N = "nba", 10, 20, int(sys.argv[1]) if len(sys.argv) > 1 else 8, 11
torch.manual_seed(3)
m = STTODENet(synth.make_args(dataset, tp, tf), torch.device("cpu"))
sd = {k: v.clone() for k, v in m.state_dict().items()}
m.set_device(dev); m.train()
d = training_draws(B * N, tp, tf, seed=40 + B)
data = synth.nba_batch(2, B, N, tp, tf)
m.past_encoder.dropout_mask_override, m.future_encoder.dropout_mask_override = d["mask_past"], d["mask_future"]
m.eps_override = (d["eps_q"], d["eps_p"])
m.set_data_nba(data)
total = m.forward()[0]; total.backward()
res = {}
for dt in (torch.float64, torch.float32):
    s = {k: v.to(dt) for k, v in sd.items()}
    names = [k for k in s if not k.endswith("pos_encoder.pe")]
    for k in names: s[k].requires_grad_(True)
    out = O.training_objective(s, data["past_traj"].view(B * N, tp, 2).to(dt), data["future_traj"].view(B * N, tf, 2).to(dt), B, N,
                               d["eps_q"].to(dt), d["eps_p"].to(dt), masks=(d["mask_past"], d["mask_future"]), tensors=True)
    out["total"].backward()
    res[dt] = {k: s[k].grad for k in names}
    print(dt, float(out["total"]), float(total))
for n, p in m.named_parameters():
    r64 = res[torch.float64][n]
    if r64 is None: continue
    den = r64.abs().max().clamp_min(1e-30)
    eg = float((p.grad.cpu().double() - r64).abs().max() / den)
    ec = float((res[torch.float32][n].double() - r64).abs().max() / den)
    flag = " <<<" if eg > 1e-4 else ""
    print(f"{n:75s} gpu {eg:.2e}  cpu32 {ec:.2e}{flag}")
