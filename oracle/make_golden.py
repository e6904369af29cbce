"""Generate tests/golden/*.npz by running the UNMODIFIED reference (under oracle/ref_shim.py) here.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py [forward|backward|training|sampler|all]
The fixtures pin `oracle/sttode_oracle.py` and `oracle/pmath_oracle.py` to the reference's own outputs and
travel to the GPU box, where /root/reference does not exist.  Weights are NOT stored: both sides rebuild
them with `make_state_dict(seed)`; the reference side loads them with `load_state_dict(strict=True)`,
which also proves key/shape parity of the 90 parameters + 2 buffers.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim                      # noqa: E402
from oracle import sttode_oracle as O            # noqa: E402
from sttode_b200 import synth                    # noqa: E402

OUT = os.environ.get("STTODE_GOLDEN_OUT") or os.path.join(ROOT, "tests", "golden")      # override: tests/test_golden_reproducible.py


def ref_model(ref, dataset, tp, tf, seed, dtype=torch.float32):
    args = ref_shim.make_args(dataset, tp, tf)
    m = ref.sttode.STTODENet(args, torch.device("cpu")).to(dtype).eval()
    m.load_state_dict(O.make_state_dict(seed, tp, tf, dtype), strict=True)
    return m


def inject_z(ref, z):
    ref.sttode.Normal.rsample = lambda self: self.mu + z.to(self.mu.dtype) * self.sigma


def np32(t):
    return t.detach().to(torch.float32).numpy()


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.install()
    torch.manual_seed(0)

    # ---- MHGSA module: L == S (swap quirk), L != S (cross, transposed), adversarial rows ------------
    m = ref_model(ref, "eth", 8, 12, seed=1)
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    g = torch.Generator().manual_seed(11)
    rec = {}
    for name, (L, S, bsz) in {"sq5": (5, 5, 3), "sq1": (1, 1, 4), "sq33": (33, 33, 2), "x46": (4, 6, 3)}.items():
        q = torch.randn(L, bsz, 64, generator=g)
        kv = q if L == S else torch.randn(S, bsz, 64, generator=g)
        out, w = attn(q, kv, kv)
        rec[f"{name}_q"], rec[f"{name}_kv"], rec[f"{name}_out"], rec[f"{name}_w"] = map(np32, (q, kv, out, w))
    np.savez_compressed(os.path.join(OUT, "mhgsa.npz"), **rec)

    # ---- ODE block ---------------------------------------------------------------------------------
    rec = {}
    for name, (L, bsz) in {"L1": (1, 7), "L6": (6, 5), "L40": (40, 3)}.items():
        x = torch.randn(L, bsz, 1, 64, generator=g)
        m.past_encoder.agent_mask = torch.zeros(1)
        y = m.past_encoder.ODE_Encoder(x, mask=None, num_agent=bsz)
        rec[f"{name}_x"], rec[f"{name}_y"] = np32(x), np32(y)
    np.savez_compressed(os.path.join(OUT, "ode_block.npz"), **rec)

    # ---- full inference: ETH path (set_data + inference) -------------------------------------------
    rec = {}
    for idx, n in enumerate((2, 5, 10)):
        obs, gt = synth.eth_scene(idx, n)
        ones = torch.ones(n, 8)
        z = synth.latent(n, seed=7 + idx)
        inject_z(ref, z)
        m.set_data(None, obs, gt, ones, torch.ones(n, 12))
        pred = m.inference(None)                                    # [20,N,12,2]
        pf = m.past_encoder(torch.cat((m.past_traj, torch.cat([(m.past_traj[:, 1:] - m.past_traj[:, :-1])[:, [0]],
                            m.past_traj[:, 1:] - m.past_traj[:, :-1]], 1)), -1), 1, n)
        rec[f"eth{idx}_n"] = np.int64(n)
        rec[f"eth{idx}_pred"], rec[f"eth{idx}_pf"] = np32(pred), np32(pf)
        # float64 arbiter of the same reference code
        m64 = ref_model(ref, "eth", 8, 12, seed=1, dtype=torch.float64)
        m64.set_data(None, obs.double(), gt.double(), ones.double(), torch.ones(n, 12).double())
        rec[f"eth{idx}_pred64"] = m64.inference(None).numpy()
    np.savez_compressed(os.path.join(OUT, "inference_eth.npz"), **rec)

    # ---- full inference: NBA path (set_data_nba + inference), 10/20 and 5/10 -----------------------
    rec = {}
    for name, (B, tp, tf) in {"nba1020": (6, 10, 20), "nba0510": (9, 5, 10)}.items():
        mn = ref_model(ref, "nba", tp, tf, seed=2)
        data = synth.nba_batch(0, B, 11, tp, tf)
        z = synth.latent(B * 11, seed=9)
        inject_z(ref, z)
        mn.set_data_nba(data)
        pred = mn.inference(data)
        rec[f"{name}_pred"] = np32(pred)
        rec[f"{name}_B"] = np.int64(B)
    np.savez_compressed(os.path.join(OUT, "inference_nba.npz"), **rec)

    # ---- training objective (forward of train.py's step, eval mode so PE dropout is off): NBA shape, B = 4 -------------------
    rec = {}
    B, tp, tf = 4, 10, 20
    mn = ref_model(ref, "nba", tp, tf, seed=2)
    data = synth.nba_batch(1, B, 11, tp, tf)
    gq = torch.Generator().manual_seed(21)
    eps_q, eps_mid, eps_p = torch.randn(B * 11, 32, generator=gq), torch.randn(B * 11, 32, generator=gq), torch.randn(B * 11 * 20, 32, generator=gq)
    queue = [eps_q, eps_mid, eps_p]
    ref.sttode.Normal.rsample = lambda self: self.mu + queue.pop(0) * self.sigma
    mn.set_data_nba(data)
    total, lp, lr, lk, ld = mn.forward()
    rec.update(total=np.float64(float(total)), pred=np.float64(lp), recover=np.float64(lr), kl=np.float64(lk), diverse=np.float64(ld),
               qz=np32(mn.qz_param), B=np.int64(B))
    np.savez_compressed(os.path.join(OUT, "training_objective.npz"), **rec)

    # ---- pmath ---------------------------------------------------------------------------------------
    P = ref.pmath
    rec = {}
    x = torch.randn(16, 8, generator=g) * 0.3
    y = torch.randn(16, 8, generator=g) * 0.3
    x[0] = 0; y[1] = 0                                              # zero rows (clamp_min 1e-5 paths)
    x[2] = x[2] / x[2].norm() * 0.9995                              # beyond the ball margin -> project rescales
    x[3] = x[3] / x[3].norm() * 5.0                                 # far outside
    mat = torch.randn(6, 8, generator=g) * 0.5
    rec.update(x=np32(x), y=np32(y), mat=np32(mat))
    for c in (1.0, 0.5):
        t = f"c{c}"
        px, py = P.project(x, c=c), P.project(y, c=c)
        rec[f"project_{t}"] = np32(px)
        rec[f"lambda_x_{t}"] = np32(P.lambda_x(px, c=c))
        rec[f"mobius_add_{t}"] = np32(P.mobius_add(px, py, c=c))
        rec[f"dist_{t}"] = np32(P.dist(px, py, c=c))
        rec[f"dist0_{t}"] = np32(P.dist0(px, c=c))
        rec[f"expmap0_{t}"] = np32(P.expmap0(x, c=c))
        rec[f"logmap0_{t}"] = np32(P.logmap0(px, c=c))
        rec[f"expmap_{t}"] = np32(P.expmap(px, y, c=c))
        rec[f"logmap_{t}"] = np32(P.logmap(px, py, c=c))
        rec[f"mobius_matvec_{t}"] = np32(P.mobius_matvec(mat, px, c=c))
        rec[f"dist_matrix_{t}"] = np32(P.dist_matrix(px, py, c=c))
        rec[f"poincare_mean_{t}"] = np32(P.poincare_mean(px, dim=0, c=c))
    s = torch.linspace(-20, 20, 41)
    rec["tanh"] = np32(P.tanh(s))
    u = torch.tensor([-1.5, -1.0, -0.99999, -0.5, 0.0, 0.3, 0.99998, 1.0, 2.0])
    rec["artanh_in"], rec["artanh"] = np32(u), np32(P.artanh(u))
    rec["arsinh_in"], rec["arsinh"] = np32(s), np32(P.arsinh(s))
    np.savez_compressed(os.path.join(OUT, "pmath.npz"), **rec)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def grads_of(mod):
    return {n: np32(p.grad) for n, p in mod.named_parameters() if p.grad is not None}


def backward_golden():
    """Gradients of the reference itself (torch autograd over the unmodified modules) for the MHGSA module and the
    tensor-ODE block on the inputs of mhgsa.npz / ode_block.npz, with seeded upstream gradients."""
    ref = ref_shim.install()
    m = ref_model(ref, "eth", 8, 12, seed=1)
    attn = m.past_encoder.ODE_Encoder.odeblock.odefunc.layers[0].self_attn.temporal_attention_before
    gi = np.load(os.path.join(OUT, "mhgsa.npz"))
    g = torch.Generator().manual_seed(31)
    rec = {}
    for name in ("sq5", "x46"):
        q = torch.from_numpy(gi[f"{name}_q"]).requires_grad_(True)
        square = gi[f"{name}_q"].shape == gi[f"{name}_kv"].shape
        kv = q if square else torch.from_numpy(gi[f"{name}_kv"]).requires_grad_(True)
        dout = torch.randn(*q.shape, generator=g)
        m.zero_grad(set_to_none=True)
        attn(q, kv, kv)[0].backward(dout)
        rec[f"mhgsa_{name}_dout"], rec[f"mhgsa_{name}_dq"] = np32(dout), np32(q.grad)
        if not square:
            rec[f"mhgsa_{name}_dkv"] = np32(kv.grad)
        for n, v in grads_of(attn).items():
            rec[f"mhgsa_{name}_g_{n}"] = v
    go = np.load(os.path.join(OUT, "ode_block.npz"))
    enc = m.past_encoder.ODE_Encoder
    for name in ("L1", "L6"):
        x = torch.from_numpy(go[f"{name}_x"]).requires_grad_(True)
        dy = torch.randn(*x.shape, generator=g)
        m.zero_grad(set_to_none=True)
        m.past_encoder.agent_mask = torch.zeros(1)
        enc(x, mask=None, num_agent=x.shape[1]).backward(dy)
        rec[f"ode_{name}_dy"], rec[f"ode_{name}_dx"] = np32(dy), np32(x.grad)
        for n, p in enc.named_parameters():      # L == 1: q/k projections receive no gradient at all (softmax over one key)
            rec[f"ode_{name}_g_{n}"] = np32(p.grad if p.grad is not None else torch.zeros_like(p))
    np.savez_compressed(os.path.join(OUT, "backward.npz"), **rec)
    print("backward.npz", os.path.getsize(os.path.join(OUT, "backward.npz")))


class _MaskDropout(torch.nn.Module):
    """Stands in for PositionalAgentEncoding.dropout (nn.Dropout(0.1), model/STTODE.py:140): same arithmetic
    (x * mask / (1 - p)) with the Bernoulli mask injected so both sides use the same draw."""

    def __init__(self, mask, p=0.1):
        super().__init__()
        self.mask, self.p = mask, p

    def forward(self, x):
        return x * (self.mask.to(x.dtype) / (1.0 - self.p))


def training_draws(R, tp, tf, seed):
    """Every random draw of one training step, from one seeded generator (shared with tests/test_gpu_training.py)."""
    g = torch.Generator().manual_seed(seed)
    return dict(mask_past=(torch.rand(R, tp, 64, generator=g) >= 0.1).float(), mask_future=(torch.rand(R, tf, 64, generator=g) >= 0.1).float(),
                eps_q=torch.randn(R, 32, generator=g), eps_mid=torch.randn(R, 32, generator=g), eps_p=torch.randn(R * 20, 32, generator=g),
                theta_u=torch.rand(1, generator=g))


def grad_sample_index(numel, count=512):
    return torch.linspace(0, numel - 1, min(numel, count)).round().long()


def reference_training_step(ref, dataset, tp, tf, B, N, draws_seed, batch_index):
    """One training step of the reference itself in train() mode (train.py:61-65: set_data* -> forward -> backward) with every random draw injected.
    Weights are the reference's own default init under torch.manual_seed(3).  -> (model after backward, loss tuple as 5 floats)."""
    torch.manual_seed(3)
    m = ref.sttode.STTODENet(ref_shim.make_args(dataset, tp, tf), torch.device("cpu")).train()
    R = B * N
    d = training_draws(R, tp, tf, seed=draws_seed)
    m.past_encoder.pos_encoder.dropout = _MaskDropout(d["mask_past"])
    m.future_encoder.pos_encoder.dropout = _MaskDropout(d["mask_future"])
    queue = [d["eps_q"], d["eps_mid"], d["eps_p"]]
    ref.sttode.Normal.rsample = lambda self: self.mu + queue.pop(0) * self.sigma
    if dataset == "nba":
        m.set_data_nba(synth.nba_batch(batch_index, B, N, tp, tf))
    else:
        obs, gt = synth.eth_scene(batch_index, N)
        orig_rand = torch.rand
        torch.rand = lambda *a, **k: d["theta_u"].clone()        # the uniform draw behind the rotation angle (model/STTODE.py:423)
        try:
            m.set_data(None, obs, gt, torch.ones(N, tp), torch.ones(N, tf))
        finally:
            torch.rand = orig_rand
    m.zero_grad(set_to_none=True)
    total, lp, lr, lk, ld = m.forward()
    total.backward()
    return m, [float(total.detach()), lp, lr, lk, ld]


def training_golden():
    """One training step of the reference itself in train() mode with every random draw injected: loss tuple + the gradient of all
    parameters.  Weights are the reference's own default init under torch.manual_seed(3) (sttode_b200.net.STTODENet reproduces it
    bit for bit, tests/test_host_logic.py)."""
    ref = ref_shim.install()
    rec = {}
    for name, dataset, tp, tf, B, N in (("nba", "nba", 10, 20, 8, 11), ("eth", "eth", 8, 12, 1, 7)):
        m, losses = reference_training_step(ref, dataset, tp, tf, B, N, draws_seed=40 + B, batch_index=2 if dataset == "nba" else 3)
        rec[f"{name}_losses"] = np.array(losses, np.float64)
        for n, p_ in m.named_parameters():           # fixtures stay small: <= 512 evenly spaced entries + norm + max per tensor
            g_ = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).detach().flatten()
            rec[f"{name}_g_{n}"] = np32(g_[grad_sample_index(g_.numel())])
            rec[f"{name}_gstat_{n}"] = np.array([float(g_.double().norm()), float(g_.abs().max())], np.float64)
        rec[f"{name}_has_grad"] = np.array([p_.grad is not None for _, p_ in m.named_parameters()])
    np.savez_compressed(os.path.join(OUT, "training_step.npz"), **rec)
    print("training_step.npz", os.path.getsize(os.path.join(OUT, "training_step.npz")))


def sampler_args(dataset, tp, tf):
    a = ref_shim.make_args(dataset, tp, tf)
    a.nz, a.qnet_mlp, a.share_eps, a.train_w_mean = 32, [512, 256], True, True
    a.kld_weight, a.kld_min_clamp, a.recon_weight = 0.1, 10, 5.0        # trainsampler.py:59-62,90-92
    return a


def sampler_golden():
    """One step of the reference's second stage (trainsampler.py:133-149): Sampler.forward over the frozen STTODENet with only
    the decoder in train() mode, compute_sampler_loss_nba, backward: dec_motion, loss terms and the sampler's gradients.
    Both mean=True (the reference's default) and mean=False (eps injected)."""
    ref = ref_shim.install()
    import sampler as ref_sampler
    import samplerloss as ref_loss
    rec = {}
    g = torch.Generator().manual_seed(51)
    eps1 = torch.randn(1, 32, generator=g)
    # case "mean": NBA, B = 4, mean=True (the reference's default; its mean=False branch only shape-checks when batch_size == 1)
    # case "eps":  ETH scene of 7 agents, mean=False with the shared eps row injected
    for t, dataset, tp, tf, B, N, mean in (("mean", "nba", 5, 10, 4, 11, True), ("eps", "eth", 8, 12, 1, 7, False)):
        torch.manual_seed(3)
        net = ref.sttode.STTODENet(ref_shim.make_args(dataset, tp, tf), torch.device("cpu")).eval()
        torch.manual_seed(4)
        smp = ref_sampler.Sampler(sampler_args(dataset, tp, tf))
        d = training_draws(B * N, tp, tf, seed=50)
        queue = [d["eps_q"], d["eps_mid"], d["eps_p"]]
        ref.sttode.Normal.rsample = lambda self: self.mu + queue.pop(0) * self.sigma
        net.decoder.train()
        if dataset == "nba":
            data = synth.nba_batch(3, B, N, tp, tf)
            net.set_data_nba(data)
            fut = data["future_traj"].view(B * N, tf, 2)
        else:
            obs, gt = synth.eth_scene(5, N)
            net.set_data(None, obs, gt, torch.ones(N, tp), torch.ones(N, tf))
            fut = gt.transpose(1, 2)
        orig_randn = torch.randn
        torch.randn = lambda *a, **k: eps1.clone()             # Sampler's own draw (share_eps: one [1, nz] row, sampler.py:42-43)
        try:
            smp.zero_grad(set_to_none=True)
            dec_motion, sdist, vdist, _ = smp.forward(net, mean=mean)
        finally:
            torch.randn = orig_randn
        div_cfg = {"weight": 1, "scale": 1.0}                  # trainsampler.py:106-111 (eth and nba)
        if dataset == "nba":
            dm = dec_motion.permute(1, 0, 2, 3).reshape([-1, 20, 10, 2])       # the reference's reshuffle, kept as is (trainsampler.py:141-143)
            total, ld, _ = ref_loss.compute_sampler_loss_nba(sampler_args(dataset, tp, tf), fut, dm, 1, vdist, sdist, div_cfg)
        else:
            total, ld, _ = ref_loss.compute_sampler_loss(sampler_args(dataset, tp, tf), fut, dec_motion, 1, torch.ones(N, tf), vdist, sdist, div_cfg)
        total.backward()
        rec[f"{t}_dec_motion"] = np32(dec_motion)
        rec[f"{t}_losses"] = np.array([float(total), float(ld["kld"]), float(ld["diverse"])], np.float64)
        for n, p_ in smp.named_parameters():
            g_ = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).detach().flatten()
            rec[f"{t}_g_{n}"] = np32(g_[grad_sample_index(g_.numel())])
            rec[f"{t}_gstat_{n}"] = np.array([float(g_.double().norm()), float(g_.abs().max())], np.float64)
    rec["eps1"] = np32(eps1)
    np.savez_compressed(os.path.join(OUT, "sampler_step.npz"), **rec)
    print("sampler_step.npz", os.path.getsize(os.path.join(OUT, "sampler_step.npz")))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("sampler", "all"):
        sampler_golden()
    if what in ("training", "all"):
        training_golden()
    if what in ("forward", "all"):
        main()
    if what in ("backward", "all"):
        backward_golden()
