/* sttode.h — C ABI of libsttode_sm100.so, the B200 (sm_100a) implementation of the STTODE hot path (forward and backward).
 *
 * The reference (joyecnu/STTODE) is pure Python/PyTorch and has no FFI; its boundary for this path is the
 * Python module API.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C: pointers + sizes only.  All `const float*` / `float*` are DEVICE pointers to contiguous fp32
 *     unless the name ends in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - ownership: the caller owns every buffer including the workspace; the library never allocates or frees
 *     device memory on these calls and keeps no pointer after return (exceptions: *_host entry points own a
 *     cached pinned/device staging arena, released by sttode_release_host_arena(); sttode_peer_alloc returns a
 *     cudaMalloc'ed arena because only such memory can be exported through CUDA IPC, freed by sttode_peer_free()).
 *   - errors: return 0 on success, negative STTODE_E* otherwise; text via sttode_last_error() (thread local).
 *     Unsupported shapes/options are errors, never fallbacks.  There is no CPU path.
 *   - work is only enqueued on `stream`; no implicit synchronisation; safe under CUDA-graph capture
 *     (except *_host, which synchronises the stream before returning).
 *   - weights use the reference's state_dict tensors as they are (nn.Linear layout W[out][in]).
 */
#ifndef STTODE_H_
#define STTODE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STTODE_VERSION 100          /* 0.1.0 */
#define STTODE_D 64                 /* model width  (model/STTODE.py:182, hidden_dim) */
#define STTODE_H 8                  /* heads        (model/STTODE.py:190) */
#define STTODE_FF 1024              /* FFN width    (model/STTODE.py:191) */
#define STTODE_ZDIM 32              /* latent width (train.py:37) */
#define STTODE_GRU 96               /* GRU state    (model/STTODE.py:26) */

enum {
    STTODE_OK = 0,
    STTODE_EINVAL = -1,     /* bad shape / option / null pointer */
    STTODE_EUNSUPPORTED = -2, /* option the reference accepts but this path does not implement */
    STTODE_EWORKSPACE = -3, /* workspace too small */
    STTODE_ECUDA = -4,      /* CUDA runtime error (launch, device mismatch: needs sm_100) */
};

/* score_mode of the geodesic attention */
enum {
    STTODE_SCORE_OBLIQUE_ACOS = 0,  /* -acos(clamp(<k^,q^>)) : the live reference path (core/manifolds/oblique.py:36-45) */
    STTODE_SCORE_POINCARE_DIST = 1, /* -dist_c(expmap0(q), expmap0(k)) via the Gram closed form (hyptorch/pmath.py:482-493) */
};

/* explicit fixed-step solvers for the tensor ODE (ode_demo.py:186-190 uses EULER with one step of h = 12) */
enum { STTODE_ODE_EULER = 0, STTODE_ODE_MIDPOINT = 1, STTODE_ODE_RK4 = 2 /* classic */, STTODE_ODE_RK4_38 = 3 /* torchdiffeq 'rk4' */ };

/* trajectory memory layouts */
enum { STTODE_TRAJ_RT2 = 0 /* [rows, T, 2]  (NBA: data['past_traj'].view(B*N,T,2)) */,
       STTODE_TRAJ_R2T = 1 /* [rows, 2, T]  (ETH/SDD loaders: obs_traj [N,2,T]) */,
       STTODE_FEAT_RT4 = 2 /* [rows, T, 4]  explicit [x, y, vx, vy] features (self.inputs / self.inputs_for_posterior):
                               used as they are, no centring and no velocity recomputation */ };

/* Hyp_mhsa parameters (hyptransformerlib.py:356-365): in_proj_weight [3E,E], in_proj_bias [3E], out_proj [E,E],[E] */
typedef struct sttode_mhgsa_weights {
    const float *in_proj_w, *in_proj_b, *out_proj_w, *out_proj_b;
} sttode_mhgsa_weights;

/* one TransformerEncoderLayer = the ODE right-hand side (hypertransformer.py:112-127) */
typedef struct sttode_ode_weights {
    sttode_mhgsa_weights attn;                  /* self_attn.temporal_attention_before.*            */
    const float *info_w, *info_b;               /* self_attn.temporal_info   [64,64],[64]           */
    const float *gate_w, *gate_b;               /* self_attn.temporal_gate   [64,64],[64]           */
    const float *lin1_w, *lin1_b;               /* linear1 [1024,64],[1024]                         */
    const float *lin2_w, *lin2_b;               /* linear2 [64,1024],[64]                           */
    const float *norm1_g, *norm1_b, *norm2_g, *norm2_b; /* LayerNorm(64), eps 1e-5                  */
    const void *packed;                         /* optional: buffer filled by sttode_ode_pack (fp16 hi/lo operand images of the
                                                   weights above) -> forward runs on the tensor cores (k_qkv_tc, k_ode_tail_tc);
                                                   NULL -> FP32-FFMA kernels.  Forward entry points only; ignored by *_bwd       */
} sttode_ode_weights;

/* PastEncoder / FutureEncoder front end (model/STTODE.py:179-197) */
typedef struct sttode_encoder_weights {
    const float *fc_w, *fc_b;                   /* input_fc  [64,4],[64]                            */
    const float *pe;                            /* pos_encoder.pe [200,64]                          */
    const float *pefc_w, *pefc_b;               /* pos_encoder.fc [64,128],[64]                     */
    const float *fc2_w, *fc2_b;                 /* input_fc2 [64,64*T],[64]                         */
    const float *fc3_w, *fc3_b;                 /* input_fc3 [64,67],[64]                           */
    sttode_ode_weights ode;                     /* ODE_Encoder.odeblock.odefunc.layers.0.*          */
    const float *folded;                        /* optional: table filled by sttode_encoder_fold (the eval-mode front end — a chain of
                                                   Linear layers — collapsed to one affine map of the 4 T input features).  With it AND
                                                   ode.packed the tensor-core kernels compute x0 in their input staging; NULL -> k_prologue */
} sttode_encoder_weights;

/* one DecomposeBlock (model/STTODE.py:16-38) */
typedef struct sttode_mlp3_weights { const float *w0, *b0, *w1, *b1, *w2, *b2; } sttode_mlp3_weights; /* 256->512->256->out */
typedef struct sttode_decompose_weights {
    const float *conv_w, *conv_b;               /* conv_past [32,2,3],[32]                          */
    const float *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh; /* encoder_past [288,32],[288,96],[288],[288] */
    sttode_mlp3_weights dec_y;                  /* decoder_y -> 2*T_f                               */
    sttode_mlp3_weights dec_x;                  /* decoder_x -> 2*T_p                               */
} sttode_decompose_weights;

typedef struct sttode_model_weights {
    int past_length, future_length;             /* args.past_length / future_length                 */
    sttode_encoder_weights past;                /* past_encoder.*                                   */
    sttode_decompose_weights dec[2];            /* decoder.decompose.{0,1}.*  (num_decompose = 2)   */
    const void *dec_packed;                     /* optional: buffer filled by sttode_decoder_pack -> tensor-core decoder;
                                                   NULL -> FP32-FFMA decoder kernels                */
    unsigned long long *rng_state;              /* optional device counter for CUDA-graph replay: when set and z == NULL the
                                                   in-kernel latent draw uses seed + *rng_state and the call bumps *rng_state
                                                   by one in-stream, so every replay of a captured call draws fresh samples   */
} sttode_model_weights;

/* ---- diagnostics ------------------------------------------------------------------------------- */
int sttode_version(void);
const char *sttode_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, else STTODE_ECUDA (load fails loudly elsewhere) */
int sttode_device_check(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long sttode_launch_count(void);

/* measurement hook: when enabled, every launch of the dominant kernel (k_dec_pair, the per-(agent,sample)
 * decoder tile) is bracketed by CUDA events on its own stream; _read waits for them, returns the summed
 * device time and the number of launches since the last read, and recycles the events. */
int sttode_profile_enable(int on);
int sttode_profile_read(double *total_ms, unsigned long long *launches);

/* tcgen05 building-block check: c[128,n] = a[128,k] . b[n,k]^T on the tensor cores, passes = 1 (fp16 inputs) or
 * 3 (fp16 hi/lo split, fp32-accurate).  swap_lbo_sbo: 0 = the library's descriptor convention, 1 = the swapped one (must fail), 2 = A operand
 * staged in tensor memory instead of shared memory.  status (device int) is set to 1 if the MMA never completed. */
int sttode_umma_selftest(const float *a, const float *b, float *c, int n, int k, int passes, int swap_lbo_sbo, int *status, void *stream);
/* Diagnostic: issue `reps` back-to-back tcgen05.mma (M = 128, K = 16, kind::f16, width n) on resident operands; cycles2_dev[0] = clock64 cycles from
 * first issue to completion, [1] = cycles spent in the issue loop (device pointer to two long long). */
int sttode_umma_bench(int n, int reps, long long *cycles2_dev, void *stream);   /* n | 0x1000: A operand from tensor memory; n | 0x2000: bystander warps on tcgen05.ld / st */
/* Diagnostic: `reps` GRU-slab-like groups of MMAs (mode 0: 3 x N=192 then 3 x N=96 with distinct hi/lo images, the decoder's order; 1: interleaved;
 * 2 / 3: only the N=192 / N=96 triples; 4: identical operands) with B images of b_rows rows; same outputs as sttode_umma_bench.
 * poll_mode > 0 adds 16 bystander warps waiting on an mbarrier meanwhile (1: every lane polls, 2: one lane per warp, 3: one lane + nanosleep). */
int sttode_umma_pattern(int mode, int b_rows, int reps, int poll_mode, long long *cycles2_dev, void *stream);

/* Tensor-core operand images of one ODE block's weights (tcgen05 3xFP16 path): call once per weight version, then put the buffer
 * in sttode_ode_weights.packed.  One launch, graph-capturable. */
size_t sttode_ode_pack_bytes(void);
int sttode_ode_pack(const sttode_ode_weights *w, void *packed, size_t packed_bytes, void *stream);

/* Eval-mode encoder front end (model/STTODE.py:214-223 without the PE dropout) folded into one affine map; T = the encoder's sequence
 * length (past_length for PastEncoder, future_length for FutureEncoder).  Call once per weight version.  One launch, graph-capturable. */
size_t sttode_encoder_fold_bytes(int T);
int sttode_encoder_fold(const sttode_encoder_weights *w, int T, float *folded, size_t folded_bytes, void *stream);

/* ---- MHGSA: replaces Hyp_mhsa.forward -> hyp_mhsa (hyptransformerlib.py:403-454 -> :29-311) ------
 * query [L,bsz,E], key/value [S,bsz,E] (key==value==query pointer for self-attention), out [L,bsz,E],
 * attn_weights (nullable) [bsz,L,S] = head-averaged softmax (need_weights=True, :306-309).
 * E must be 64 and H 8.  When L == S the reference's un-transposed score matrix is reproduced
 * (out_i = sum_j softmax_j(-acos(k^_i . q^_j)) v_j, SURVEY.md finding 4); L != S is textbook. */
size_t sttode_mhgsa_workspace_bytes(int L, int S, int bsz);
int sttode_mhgsa_fwd(const float *query, const float *key, const float *value, const sttode_mhgsa_weights *w,
                     float *out, float *attn_weights, int L, int S, int bsz, int E, int H, int score_mode,
                     float clamp_eps, float curvature, void *workspace, size_t workspace_bytes, void *stream);

/* `linear` (hyptransformerlib.py:7-26) for the path's shapes: out[rows,out_features] = x[rows,64] . W[out_features,64]^T + b,
 * out_features a multiple of 64 (QKV 192, out-proj / gates 64). */
int sttode_linear64(const float *x, const float *weight, const float *bias, float *out, int rows, int out_features, void *stream);

/* ---- tensor-ODE block: replaces ODEG_Encoder.forward (ode_demo.py:217-231) incl. the torchdiffeq solver
 * loop (ode_demo.py:186-190), TransformerEncoder_ode.forward (:53-72), TransformerEncoderLayer.forward
 * (hypertransformer.py:134-153) and Hypattention.forward (:55-89).
 * x [L, bsz, 64] (the reference's [L,bsz,1,64]); y rows have stride ldy floats (>= 64) so the result can
 * land in past_feature[:, 64:128].  y = relu(solve(x, t in [0, n_steps*h])) when final_relu, else raw.
 * Reference call: n_steps = 1, h = 12, method = EULER, final_relu = 1. */
size_t sttode_ode_block_workspace_bytes(int L, int bsz, int method);
int sttode_ode_block_fwd(const float *x, const sttode_ode_weights *w, float *y, int ldy, int L, int bsz,
                         int n_steps, float h, int method, int final_relu, void *workspace,
                         size_t workspace_bytes, void *stream);
/* one evaluation of the ODE right-hand side f(t, y) (TransformerEncoder_ode.forward, ode_demo.py:53-72) */
int sttode_ode_rhs_fwd(const float *x, const sttode_ode_weights *w, float *f, int L, int bsz, void *workspace,
                       size_t workspace_bytes, void *stream);

/* Hypattention.forward on its own (hypertransformer.py:55-89): out = tanh(W_i m + b_i) * sigmoid(W_g m + b_g),
 * m = MHGSA(x).  Only the attn / info / gate members of `w` are read.  Workspace as for the ODE block. */
int sttode_hypattention_fwd(const float *x, const sttode_ode_weights *w, float *out, int L, int bsz, void *workspace,
                            size_t workspace_bytes, void *stream);

/* ---- backward of the MHGSA / tensor-ODE block (training config: "forward+backward through ODE and MHGSA").
 * The reference obtains these from torch autograd over hyp_mhsa (hyptransformerlib.py:29-311), Hypattention.forward
 * (hypertransformer.py:55-89), TransformerEncoderLayer.forward (:134-153) and the Euler step (ode_demo.py:186-190,231);
 * here they are explicit entry points an autograd.Function calls.  Every backward call re-materialises the forward
 * intermediates from the block INPUT, so the forward saves nothing.  Parameter gradients are ACCUMULATED (+=) into the
 * buffers of the grads struct (same shapes as the weights; the struct or any member may be NULL to skip it); input
 * gradients are overwritten.  Summation order is fixed (no atomics): results are run-to-run deterministic. */
typedef struct sttode_mhgsa_grads { float *in_proj_w, *in_proj_b, *out_proj_w, *out_proj_b; } sttode_mhgsa_grads;
typedef struct sttode_ode_grads {
    sttode_mhgsa_grads attn;
    float *info_w, *info_b, *gate_w, *gate_b, *lin1_w, *lin1_b, *lin2_w, *lin2_b, *norm1_g, *norm1_b, *norm2_g, *norm2_b;
} sttode_ode_grads;

/* dout [L,bsz,E] -> dquery [L,bsz,E] (+ dkey, dvalue [S,bsz,E] when key/value are separate tensors; with
 * query == key == value the self-attention branch is differentiated and dquery receives the whole gradient, as in the
 * reference where only `query` feeds the packed projection, hyptransformerlib.py:113-115).  Oblique acos score only.
 * The head-averaged attention weights output is not differentiated. */
size_t sttode_mhgsa_bwd_workspace_bytes(int L, int S, int bsz);
int sttode_mhgsa_bwd(const float *query, const float *key, const float *value, const sttode_mhgsa_weights *w, const float *dout,
                     float *dquery, float *dkey, float *dvalue, const sttode_mhgsa_grads *grads, int L, int S, int bsz, int E, int H,
                     int score_mode, float clamp_eps, void *workspace, size_t workspace_bytes, void *stream);
/* df [L,bsz,64] = gradient of sttode_ode_rhs_fwd's output (gate_only = 0) or of sttode_hypattention_fwd's (gate_only = 1) */
size_t sttode_ode_rhs_bwd_workspace_bytes(int L, int bsz);
int sttode_ode_rhs_bwd(const float *x, const float *df, const sttode_ode_weights *w, float *dx, const sttode_ode_grads *grads, int L, int bsz,
                       int gate_only, void *workspace, size_t workspace_bytes, void *stream);
/* dy rows have stride lddy floats (>= 64): gradient of sttode_ode_block_fwd's y.  method must be STTODE_ODE_EULER. */
size_t sttode_ode_block_bwd_workspace_bytes(int L, int bsz, int n_steps);
int sttode_ode_block_bwd(const float *x, const float *dy, int lddy, const sttode_ode_weights *w, float *dx, const sttode_ode_grads *grads,
                         int L, int bsz, int n_steps, float h, int method, int final_relu, void *workspace, size_t workspace_bytes,
                         void *stream);

/* ---- PastEncoder.forward (model/STTODE.py:214-236), eval mode.
 * traj: raw positions in `layout`; n_batches groups of (L scenes x agents) rows, rows = n_batches*L*agents
 * ordered [batch][scene][agent] (NBA: L = B scenes attend to each other per agent).  For ETH/SDD (L = 1)
 * scenes of different sizes are concatenated: scene_ptr [n_scenes+1] and row_scene [rows] (int32, device)
 * give the CSR split; the scene origin (mean last-frame position, model/STTODE.py:417) is subtracted and
 * returned in origin [n_scenes,2].  scene_ptr == NULL means NBA semantics (no normalisation, category flag
 * on agent index agents-1 of every scene, model/STTODE.py:199-210).
 * Outputs: past_feature [rows,128], traj_norm [rows,T_p,2] (the reference's self.past_traj). */
size_t sttode_past_encoder_workspace_bytes(int rows, int L);
int sttode_past_encoder_fwd(const float *traj, int layout, const int *scene_ptr, const int *row_scene, int n_scenes,
                            int rows, int L, int agents, const sttode_model_weights *w, float *past_feature,
                            float *traj_norm, float *origin, void *workspace, size_t workspace_bytes, void *stream);

/* Generic encoder on explicit features: PastEncoder.forward (model/STTODE.py:214-236) or the shared front half of
 * FutureEncoder.forward (:276-295) — input_fc, positional fc, flatten-T fc, category fc, tensor-ODE block, concat.
 * feat4 [rows, T, 4]; rows ordered [batch][scene(L)][agent]; feature [rows,128] = [pre-ODE | ODE output]. */
int sttode_encoder_fwd(const float *feat4, int rows, int L, int agents, int T, const sttode_encoder_weights *w, float *feature,
                       void *workspace, size_t workspace_bytes, void *stream);
/* y[rows,N] = act(x[rows,K] . W[N,K]^T + b): the two small heads of FutureEncoder (out_mlp 256->128 ReLU, qz_layer 128->64;
 * model/STTODE.py:297-300).  K % 4 == 0, K <= 256, N % 64 == 0. */
int sttode_linear(const float *x, const float *weight, const float *bias, float *y, int rows, int K, int N, int relu, void *stream);

/* ---- Decoder.forward (model/STTODE.py:320-347) in 'inference' mode, both DecomposeBlocks (:51-77).
 * past_feature [rows,128] (one row per agent, NOT repeated K times), traj_norm [rows,T_p,2],
 * z [rows*K, 32] or NULL to draw z ~ N(0,I) in-kernel from (seed) with Philox (the reference's randn_like,
 * model/STTODE.py:616).  row_scene/origin (nullable) add the scene origin back (:621-622).
 * pred [rows, K, T_f, 2]  (== reference output [K,rows,T_f,2].permute(1,0,2,3), test.py:186).
 * recon (nullable) [rows*K, T_p, 2] = the reconstruction the reference returns as recover_pre_seq. */
size_t sttode_decoder_workspace_bytes(int rows, int sample_k);
int sttode_decoder_fwd(const float *past_feature, const float *traj_norm, const float *z, unsigned long long seed,
                       const int *row_scene, const float *origin, int rows, int sample_k,
                       const sttode_model_weights *w, float *pred, float *recon, void *workspace,
                       size_t workspace_bytes, void *stream);

/* ---- generic dense layer with its backward: the thin layers around the tensor-ODE block in train() mode (input_fc 4->64, PE fc 128->64,
 * input_fc2 64T->64, input_fc3 67->64: model/STTODE.py:214-223; FutureEncoder head 256->128->64: :297-300).  Any K, N; x rows have stride ldx.
 *   fwd: y[rows,N] = act(x[rows,K] . W[N,K]^T + b)          (relu != 0 -> ReLU)
 *   bwd: dx[rows,K] = dy . W (dx nullable, row stride lddx);  dW[N,K] += dy^T . x;  db[N] += column sums of dy   (dW / db nullable).
 * The caller masks dy with the ReLU derivative beforehand when the forward used one.  Deterministic summation order. */
int sttode_dense_fwd(const float *x, int ldx, const float *weight, const float *bias, float *y, int rows, int K, int N, int relu, void *stream);
/* General strided product with the backward kernels' epilogue, exported for tests and for callers that keep their own layers:
 *   C[m,n] = epi(sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]),  epi: + bias[n], ReLU, zero where mask[m,n] <= 0, + res[m,n], + C[m,n] (accumulate).
 * mode 0: FP32 FFMA (k_gemm: what the backward entry points use); 1: tensor cores (k_gemm_tc: tcgen05, bf16x3 split over the whole fp32 exponent
 * range, ~2e-6 relative error at K = 256 because the tensor cores truncate when they accumulate — 9 % faster training step, but outside the
 * 1e-4 gradient-parity bound at the end of the backward chain, hence opt-in); -1: as the environment says (STTODE_TC_GEMM=1 turns the tensor
 * cores on for products of >= 4 M multiply-adds inside the backward entry points too).  workspace (nullable, sttode_dense_bwd_workspace_bytes()): split-K scratch. */
int sttode_gemm(int M, int N, int K, const float *A, long sam, long sak, const float *B, long sbk, long sbn, float *C, long ldc,
                const float *bias, const float *res, long ldres, const float *mask, long ldmask, int relu, int accumulate, int mode,
                void *workspace, size_t workspace_bytes, void *stream);
/* workspace (nullable): scratch for the split-K form of the weight / bias gradients — with a few thousand rows and a 64-wide layer the
 * reduction runs over the rows, so without it one CTA walks the whole batch (0.2 ms per layer at 7040 rows); sttode_dense_bwd_workspace_bytes(). */
size_t sttode_dense_bwd_workspace_bytes(void);
int sttode_dense_bwd(const float *x, int ldx, const float *weight, const float *dy, float *dx, int lddx, float *dweight, float *dbias, int rows,
                     int K, int N, void *workspace, size_t workspace_bytes, void *stream);

/* ---- backward of Decoder.forward in 'train' mode (model/STTODE.py:320-347, both DecomposeBlocks :51-77); the reference gets it from
 * torch autograd (train.py:65).  One row per (agent, sample) pair: past_feature [pairs,128], z [pairs,32], traj_norm [pairs,T_p,2] (x_true,
 * already repeated per sample), d_pred [pairs,T_f,2] = gradient of out_seq, d_recon [pairs,T_p,2] = gradient of recover_pre_seq (NULL when
 * the reconstruction is unused, as in decoder_future_1).  The forward is re-materialised from the inputs (Conv1d as a GEMM over im2col,
 * the GRU recurrence step by step), then back-propagated through time; d_past_feature [pairs,128] and d_z [pairs,32] are overwritten,
 * parameter gradients are ACCUMULATED (+=) into `grads` (same shapes as the weights; member blocks may not be NULL). */
typedef struct sttode_mlp3_grads { float *w0, *b0, *w1, *b1, *w2, *b2; } sttode_mlp3_grads;
typedef struct sttode_decompose_grads {
    float *conv_w, *conv_b, *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh;
    sttode_mlp3_grads dec_y, dec_x;
} sttode_decompose_grads;
size_t sttode_decoder_bwd_workspace_bytes(int pairs, int past_length, int future_length);
int sttode_decoder_bwd(const float *past_feature, const float *z, const float *traj_norm, const float *d_pred, const float *d_recon,
                       int pairs, const sttode_model_weights *w, const sttode_decompose_grads *grads2, float *d_past_feature, float *d_z,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Tensor-core decoder: split the decoder weights once per weight version into the fp16 hi/lo operand images the
 * tcgen05 kernel streams (caller-owned buffer of sttode_decoder_pack_bytes).  The last 256 bytes are a status word
 * (int32, 0 = ok) that the kernel sets if an MMA/copy barrier ever times out; read it with sttode_decoder_pack_status
 * after synchronising. */
size_t sttode_decoder_pack_bytes(int past_length, int future_length);
int sttode_decoder_pack(const sttode_model_weights *w, void *packed, size_t packed_bytes, void *stream);
int sttode_decoder_pack_status(const void *packed, int past_length, int future_length, int *status_host);
/* debug: read (out31_host, nullable) then reset + enable/disable the per-phase cycle counters that CTA 0 of the tensor-core
 * decoder accumulates in the status tail (31 x int64). */
int sttode_decoder_timing(void *packed, int past_length, int future_length, int enable, long long *out31_host);

/* ---- STTODENet.inference (model/STTODE.py:574-623) incl. set_data / set_data_nba preprocessing
 * (:397-486): past encoder + K-sample decoder in one call, device buffers. */
size_t sttode_forecast_workspace_bytes(int rows, int L, int sample_k, int past_length);
int sttode_forecast(const float *traj, int layout, const int *scene_ptr, const int *row_scene, int n_scenes,
                    int rows, int L, int agents, const float *z, unsigned long long seed, int sample_k,
                    const sttode_model_weights *w, float *pred, float *past_feature_or_null,
                    float *traj_norm_or_null /* [rows,T_p,2] = self.past_traj */, float *origin_or_null /* [n_scenes,2] = self.scene_orig */,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Same with HOST trajectory / prediction buffers (the end-to-end call test.py's loop makes per batch:
 * H2D of the batch, inference, D2H of the K samples; test.py:175-194).  Weights stay device resident.
 * Stages through an internal pinned + device arena (grown on demand, freed by sttode_release_host_arena). */
int sttode_forecast_host(const float *traj_host, int layout, const int *scene_ptr_host, int n_scenes, int rows,
                         int L, int agents, unsigned long long seed, int sample_k, const sttode_model_weights *w,
                         float *pred_host, void *stream);
void sttode_release_host_arena(void);
/* How sttode_forecast_host returns the predictions: 0 = device buffer + device-to-host copies of slices overlapped with the decoder kernels,
 * 1 = the decoder kernels store into the page-locked host buffer directly (host-mapped memory; a pageable `pred_host` goes through the pinned
 * arena either way).  Returns the previous mode; any other argument only queries.  Initial mode: STTODE_HOST_OUTPUT=copy|direct, else the default. */
int sttode_host_output_mode(int mode);

/* The whole loop body of test.py:171-208 for a batch of scenes, HOST buffers in, metrics out: H2D of the trajectories and of the ground truth
 * (gt_layout 0 = [rows,T_f,2], 1 = [rows,2,T_f] as the ETH loader yields pred_traj_gt), set_data + inference (as sttode_forecast_host), then
 * compute_ADE / compute_FDE with the min over the K samples (utils/metrics.py:7-26) ON THE DEVICE, and a device-to-host copy of only the three
 * AverageMeter numerators (utils/utils.py:29-34): sums_host[0] = sum over agents of minADE, [1] = of minFDE, [2] = rows (overwritten, not accumulated).
 * The K samples never leave HBM, so the call moves rows*(T_p+T_f)*8 bytes in and 24 bytes out instead of rows*K*T_f*8 bytes out. */
int sttode_evaluate_host(const float *traj_host, const float *gt_host, int layout, int gt_layout, const int *scene_ptr_host, int n_scenes,
                         int rows, int L, int agents, unsigned long long seed, int sample_k, const sttode_model_weights *w,
                         double *sums_host, void *stream);

/* The latent the decoder kernels draw when `z` is NULL (the reference's `torch.randn_like`, model/STTODE.py:609-616), written out:
 * z[n_pairs, 32] for pairs pair_base .. pair_base + n_pairs - 1 of a batch, seed as passed to sttode_forecast / sttode_decoder_fwd
 * (plus *rng_state when given).  Same device function as the decoder kernels, so the un-injected path is verifiable:
 * forecast(z = NULL, seed) == forecast(z = this output), and the draw's statistics can be tested directly. */
int sttode_latent_draw(unsigned long long seed, unsigned long long pair_base, const unsigned long long *rng_state_or_null,
                       size_t n_pairs, float *z, void *stream);

/* ---- data-parallel training (BASELINE.json configs[4]; the reference itself is single-GPU, train.py:55-98,122-123) -----------------------
 * Gradient all-reduce fused with the Adam update over NVLink / NVSwitch peer memory, one process per GPU of one node, no NCCL call:
 * every rank owns an arena [n_floats gradients | flag words] (sttode_peer_alloc), exports it through CUDA IPC (the 64-byte handle travels by any
 * host channel, e.g. torch.distributed.all_gather_object) and maps its peers' arenas (sttode_peer_open).  sttode_dp_adam_step enqueues
 *   barrier -> [sum of all ranks' gradients in rank order, * 1/world, torch.optim.Adam update of param / exp_avg / exp_avg_sq] -> barrier
 * on `stream`: plain kernels, capturable into a CUDA graph together with the forward and backward.  grads_of_dev / flags_of_dev: DEVICE arrays of
 * `world` pointers (rank r's arena and arena + n_floats*4).  lr_dev: device scalar.  The step count lives in the arena (replay-safe).
 * A peer that never arrives raises the arena's error word after ~3 s (sttode_peer_status) instead of hanging the GPU. */
int sttode_peer_alloc(size_t n_floats, void **arena, void *ipc_handle_64_bytes);
int sttode_peer_open(const void *ipc_handle_64_bytes, void **arena);
int sttode_peer_close(void *arena);
int sttode_peer_free(void *arena);
int sttode_dp_adam_step(const void *grads_of_dev, const void *flags_of_dev, int world, int rank, float *param, float *exp_avg, float *exp_avg_sq,
                        float *grad_avg_or_null, size_t n_floats, const float *lr_dev, float beta1, float beta2, float eps, void *stream);
int sttode_peer_status(const void *arena, size_t n_floats, unsigned *err_host, unsigned *step_host);

/* ---- min-over-K ADE/FDE on device (utils/metrics.py:7-26; test.py:204-208): sums[0] += sum_agents minADE,
 * sums[1] += sum_agents minFDE, sums[2] += rows  (double[3], device; AverageMeter numerators). */
int sttode_min_ade_fde(const float *pred, const float *gt, int gt_layout, int rows, int sample_k, int future_length,
                       double *sums, void *stream);

/* ---- Oblique manifold stand-alone ops: replace Oblique.proj / Oblique.dist (core/manifolds/oblique.py:15-16, 36-45).
 * proj: out = x / ||x||_2 over the last dim (no epsilon).  dist(p1 [batch,n1,d], p2 [batch,n2,d]) ->
 * out [batch,n2,n1] = acos(clamp(p2 @ p1^T, -1+eps, 1-eps)) — note the reference's argument order. */
int sttode_oblique_proj(const float *x, float *out, int rows, int d, void *stream);
int sttode_oblique_dist(const float *p1, const float *p2, float *out, int batch, int n1, int n2, int d, float clamp_eps, void *stream);

/* ---- Poincare-ball library: replaces hyptorch/pmath.py (rows = product of leading dims, d = last dim) ---- */
int sttode_pmath_project(const float *x, float *out, int rows, int d, float c, void *stream);            /* :76-103  */
int sttode_pmath_lambda_x(const float *x, float *out, int rows, int d, float c, void *stream);           /* :106-129 */
int sttode_pmath_mobius_add(const float *x, const float *y, float *out, int rows, int d, float c, void *stream); /* :132-177 */
int sttode_pmath_dist(const float *x, const float *y, float *out, int rows, int d, float c, void *stream);       /* :180-208 */
int sttode_pmath_dist0(const float *x, float *out, int rows, int d, float c, void *stream);              /* :211-234 */
int sttode_pmath_expmap0(const float *u, float *out, int rows, int d, float c, void *stream);            /* :280-304 */
int sttode_pmath_logmap0(const float *y, float *out, int rows, int d, float c, void *stream);            /* :342-368 */
int sttode_pmath_expmap(const float *x, const float *u, float *out, int rows, int d, float c, void *stream);     /* :237-277 */
int sttode_pmath_logmap(const float *x, const float *y, float *out, int rows, int d, float c, void *stream);     /* :307-339 */
/* m [dout, d]; out [rows, dout] */
int sttode_pmath_mobius_matvec(const float *m, const float *x, float *out, int rows, int d, int dout, float c, void *stream); /* :371-408 */
/* x [b, d], y [n, d] -> out [b, n]; Gram-epilogue closed form, nothing of size b*n*d is materialised */
int sttode_pmath_dist_matrix(const float *x, const float *y, float *out, int b, int n, int d, float c, void *stream); /* :482-493 */
/* elementwise helpers (:11-12, :16-22, :51-55) */
int sttode_pmath_tanh(const float *x, float *out, size_t n, float clamp, void *stream);
int sttode_pmath_artanh(const float *x, float *out, size_t n, void *stream);
int sttode_pmath_arsinh(const float *x, float *out, size_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* STTODE_H_ */
