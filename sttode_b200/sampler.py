"""Second training stage on B200: drop-in for the reference's `sampler.py` (DLow-style `Sampler` over a frozen STTODENet) and
`samplerloss.py` (KL + diversity objective).  Same class / function names, constructor arguments (`args.sample_k, nz, share_eps,
train_w_mean, qnet_mlp`), parameter names (`q_mlp.affine_layers.*`, `q_A`, `q_b`, `q_c`, `linear`) and return values.

The sampler itself is four small Linear layers (<0.1 % of a step); it drives the STTODENet stage methods `encode_history`,
`fu_encoder`, `decoder_future_0`, `decoder_future_1` (sttode_b200/net.py), which is where the work is: the encoders (MHGSA /
tensor-ODE block) run in libsttode_sm100; with `net.decoder.train()` (trainsampler.py:135) the decoder keeps autograd history so
the gradient reaches the sampler through z.
"""
import torch
from torch import nn
from torch.nn import functional as F

from .net import Normal, initialize_weights


class MLP(nn.Module):
    """utils/mlp.py:5-28: Linear + activation stack (activation after every layer)."""

    def __init__(self, input_dim, hidden_dims=(128, 128), activation="tanh"):
        super().__init__()
        self.activation = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[activation]
        self.out_dim = hidden_dims[-1]
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh
        initialize_weights(self.affine_layers.modules())

    def forward(self, x):
        for affine in self.affine_layers:
            x = self.activation(affine(x))
        return x


class Sampler(nn.Module):
    """sampler.py:6-75."""

    def __init__(self, args):
        super().__init__()
        self.device = torch.device("cpu")
        self.args = args
        self.nk, self.nz = args.sample_k, args.nz
        self.share_eps, self.train_w_mean = args.share_eps, args.train_w_mean
        self.pred_model_dim = 64
        self.qnet_mlp = args.qnet_mlp
        self.q_mlp = MLP(self.pred_model_dim, self.qnet_mlp)
        self.q_A = nn.Linear(self.q_mlp.out_dim, self.nk * self.nz)
        self.q_b = nn.Linear(self.q_mlp.out_dim, self.nk * self.nz)
        self.q_c = nn.Linear(self.nk * self.nz, self.nz)
        self.linear = nn.Linear(128, 64)
        self.eps_override = None          # test hook: the N(0, I) draw of the non-mean path

    def set_device(self, device):
        self.device = device
        self.to(device)

    def _noise(self, agent_num):
        """The N(0, I) draw of the non-mean path (sampler.py:40-46): one row shared by every (agent, sample) when `share_eps`, else one per agent."""
        if self.eps_override is not None:
            return self.eps_override.to(self.device)
        if self.share_eps:
            return torch.randn([1, self.nz]).to(self.device).repeat((agent_num * self.nk, 1))
        return torch.randn([agent_num, self.nz]).to(self.device).repeat_interleave(self.nk, dim=0)

    def forward(self, net, mean=True, need_weights=False):
        """-> (dec_motion [R, nk, T_f, 2], sampler_dist, vae_dist, net.pred_traj) (sampler.py:32-73).

        Drives the frozen STTODENet stage by stage: both encoders, then the K = 1 decode from the merged latent `q_c(.)` and the K = nk
        decode from the nk per-agent latents themselves."""
        net.encode_history()
        net.fu_encoder()
        hidden = self.q_mlp(self.linear(net.past_feature))
        scale, shift = (head(hidden).view(-1, self.nz) for head in (self.q_A, self.q_b))              # A, b of z = A * eps + b
        latents = shift if mean else scale * self._noise(net.agent_num) + shift                      # [R * nk, nz]
        merged = self.q_c(latents.view(-1, self.nk * self.nz)).view(-1, self.nz)                      # [R, nz]
        sampler_dist = Normal(mu=shift, logvar=(scale ** 2 + 1e-8).log())
        net.decoder_future_0(merged)
        net.decoder_future_1(latents)
        dec_motion = net.diverse_pred_traj
        if self.args.dataset != "nba":
            self.scene_orig = net.scene_orig
            dec_motion = dec_motion + self.scene_orig
        return dec_motion, sampler_dist, net.pz_dis, net.pred_traj

    def step_annealer(self):
        pass


# ---- samplerloss.py ---------------------------------------------------------------------------------------------------
def compute_z_kld(q_z_dist_dlow, p_z_dist_infer, agent_num, min_clip, weight):
    """samplerloss.py:4-9."""
    loss_unweighted = (q_z_dist_dlow.kl(p_z_dist_infer).sum() / agent_num).clamp_min(min_clip)
    return loss_unweighted * weight, loss_unweighted


def diversity_loss(infer_dec_motion, agent_num, weight, scale):
    """samplerloss.py:12-20: mean over sample pairs of exp(-|y_i - y_j|^2 / scale), per agent."""
    loss_unweighted = 0
    fut_motions = infer_dec_motion.view(*infer_dec_motion.shape[:2], -1)
    for motion in fut_motions:
        dist = F.pdist(motion, 2) ** 2
        loss_unweighted = loss_unweighted + (-dist / scale).exp().mean()
    loss_unweighted = loss_unweighted / agent_num
    return loss_unweighted * weight, loss_unweighted


def recon_loss(fut_motion_orig, infer_dec_motion, fut_mask, weight):
    """samplerloss.py:23-31 (not part of the reference's total; kept for API parity)."""
    diff = (infer_dec_motion - fut_motion_orig.unsqueeze(1)) * fut_mask.unsqueeze(1).unsqueeze(-1)
    loss_unweighted = diff.pow(2).sum(dim=-1).sum(dim=-1).min(dim=1)[0].mean()
    return loss_unweighted * weight, loss_unweighted


def compute_sampler_loss(args, fut_motion_orig, infer_dec_motion, batch_size, fut_mask, p_z_dist, q_z_dist_dlow, div_cfg):
    """samplerloss.py:44-58: total = weighted clamped KL + weighted diversity (the reconstruction term is commented out upstream)."""
    agent_num = fut_motion_orig.shape[0]
    kld_loss, _ = compute_z_kld(q_z_dist_dlow, p_z_dist, agent_num, args.kld_min_clamp, args.kld_weight)
    div_loss, _ = diversity_loss(infer_dec_motion, agent_num, div_cfg["weight"], div_cfg["scale"])
    loss_dict = {"kld": kld_loss, "diverse": div_loss, "recon": 0}
    return kld_loss + div_loss, loss_dict, dict(loss_dict)


def compute_sampler_loss_nba(args, fut_motion_orig, infer_dec_motion, batch_size, p_z_dist, q_z_dist_dlow, div_cfg):
    """samplerloss.py:61-74."""
    return compute_sampler_loss(args, fut_motion_orig, infer_dec_motion, batch_size, None, p_z_dist, q_z_dist_dlow, div_cfg)
