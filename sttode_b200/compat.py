"""Register this package under the reference's flat module names, so the reference's own scripts
(`test.py`: `from model.STTODE import STTODENet`; `hypertransformer.py`: `from hyptransformerlib import Hyp_mhsa`)
pick up the B200 implementation without edits:

    import sttode_b200.compat as compat; compat.install()
"""
import sys
import types


def install():
    from . import layers, mhgsa, net, pmath, sampler
    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    mod("hyptransformerlib", linear=mhgsa.linear, hyp_mhsa=mhgsa.hyp_mhsa, Hyp_mhsa=mhgsa.Hyp_mhsa, Oblique=mhgsa.Oblique)
    mod("hypertransformer", Hypattention=layers.Hypattention, TransformerEncoderLayer=layers.TransformerEncoderLayer)
    mod("ode_demo", TransformerEncoder_ode=layers.TransformerEncoder_ode, ODEblock_en=layers.ODEblock_en,
        ODEG_Encoder=layers.ODEG_Encoder)
    core = mod("core")
    man = mod("core.manifolds", Oblique=mhgsa.Oblique)
    obl = mod("core.manifolds.oblique", Oblique=mhgsa.Oblique, EPS=mhgsa.OBLIQUE_EPS)
    core.manifolds, man.oblique = man, obl
    hyp = mod("hyptorch")
    hyp.pmath = pmath
    mods["hyptorch.pmath"] = pmath
    model = mod("model")
    st = mod("model.STTODE", **{k: getattr(net, k) for k in ("STTODENet", "PastEncoder", "FutureEncoder", "Decoder", "DecomposeBlock",
                                                             "Normal", "MLP2", "PositionalAgentEncoding")})
    ut = mod("model.utils", MLP=net.MLP, initialize_weights=net.initialize_weights)
    model.STTODE, model.utils = st, ut
    mod("sampler", Sampler=sampler.Sampler)                     # trainsampler.py:9-10, test_sampler.py
    mod("samplerloss", compute_sampler_loss=sampler.compute_sampler_loss, compute_sampler_loss_nba=sampler.compute_sampler_loss_nba,
        compute_z_kld=sampler.compute_z_kld, diversity_loss=sampler.diversity_loss, recon_loss=sampler.recon_loss)
    sys.modules.update(mods)
    return mods
