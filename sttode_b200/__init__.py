"""sttode_b200 — B200-native (sm_100a) implementation of the STTODE hot path (forward and backward).

Public surface mirrors the reference's modules:
    mhgsa   : linear, hyp_mhsa, Hyp_mhsa, Oblique            (hyptransformerlib.py, core/manifolds/oblique.py)
    layers  : Hypattention, TransformerEncoderLayer, TransformerEncoder_ode, ODEblock_en, ODEG_Encoder
              (hypertransformer.py, ode_demo.py)
    net     : PastEncoder, FutureEncoder, Decoder, DecomposeBlock, Normal, STTODENet  (model/STTODE.py)
    sampler : Sampler, compute_sampler_loss(_nba), diversity_loss, compute_z_kld      (sampler.py, samplerloss.py)
    pmath   : project, mobius_add, dist, expmap0, logmap0, dist_matrix, ...           (hyptorch/pmath.py)
    parallel: scene sharding, metric all-reduce, FlatGradReducer / train_step / GraphedTrainStep (data-parallel training)
    io      : the reference's disk formats (ETH csv, SDD pickle, NBA npy) and checkpoint dict
    compat  : install() registers the reference's module names so `from model.STTODE import STTODENet` works
All arithmetic runs in sttode_b200/lib/libsttode_sm100.so through the C ABI of include/sttode.h.
"""
__version__ = "0.1.0"
