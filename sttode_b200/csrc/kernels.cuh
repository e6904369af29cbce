// Kernel declarations shared between the .cu translation units and the C-ABI layer.
#pragma once
#include "../../include/sttode.h"
#include "common.cuh"

namespace sttode {

// encoder.cu
__global__ void k_scene_origin(const float* traj, int layout, const int* scene_ptr, int n_scenes, int T, float* origin);
__global__ void k_prologue(const float* traj, int layout, const int* scene_ptr, const int* row_scene, const float* origin,
                           int rows, int agents, int T, sttode_encoder_weights w, float* x0, int ldx0, float* traj_norm);
__global__ void k_proj(const float* X, int ldx, int rows, const float* W, const float* b, int nblk, float* out, int ldo,
                       unsigned scale_mask, unsigned norm_mask, float scale, int score_mode, float curv, float* norms, int ldn,
                       const int* only_flagged);
__global__ void k_geo_attn(const float* rowop, int ldr, const float* colop, int ldc, const float* vals, int ldv, float* out,
                           int ldo, float* denom, int Lr, int Lc, int bsz, int score_mode, float lo, float hi, float curv);
__global__ void k_attn_weights(const float* rowop, int ldr, const float* colop, int ldc, const float* denom, float* wout,
                               int Lr, int Lc, int bsz, int score_mode, float lo, float hi, float curv);
__global__ void k_ode_tail(const float* x, int ldx, const float* attn_o, int ldo, const float* base, int ldb, float coef,
                           int relu, float* out, int ldout, int rows, sttode_ode_weights w, int gate_only, const int* only_flagged);
__global__ void k_lincomb(const float* base, int ldb, const float* k0, const float* k1, const float* k2, const float* k3,
                          float c0, float c1, float c2, float c3, int relu, float* out, int ldout, int rows);
__global__ void k_linear(const float* X, int rows, int K, const float* W, const float* b, int N, int relu, float* Y);
size_t linear_smem_bytes(int K);
size_t prologue_smem_bytes(int T);
size_t proj_smem_bytes();
size_t ode_tail_smem_bytes();

// encoder_tc.cu — tensor-core (tcgen05) in-projection and ODE right-hand side + weight packing
size_t ode_pack_bytes();
void launch_pack_ode(const sttode_ode_weights& w, unsigned char* pk, cudaStream_t st);
// folded encoder front end (k_fold_prologue): when `table` is set, the tensor-core kernels compute x0 from the raw trajectory in their input
// staging and write it to x0_out (the kernel's `x` with its ld) and the scene-normalised positions to traj_norm
struct FoldedIn {
    const float* table;         // [(4 T + 2), 64] from k_fold_prologue, or nullptr
    const float* traj; int layout; const int* scene_ptr; const int* row_scene; const float* origin; int agents; int T;
    float* x0_out; float* traj_norm;
};
size_t fold_table_bytes(int T);
void launch_fold_prologue(const sttode_encoder_weights& w, int T, float* out, cudaStream_t st);
__global__ void k_ode_tail_tc(const float* x, int ldx, const float* attn_o, int ldo, const float* base, int ldb, float coef, int relu, float* out,
                              int ldout, int rows, sttode_ode_weights w, const unsigned char* packed, int* tile_flags, const FoldedIn fold);
__global__ void k_qkv_tc(const float* x, int ldx, int rows, const float* in_b, float scale, const unsigned char* packed, float* out, int ldo,
                         int* tile_flags, const FoldedIn fold);
size_t ode_tail_tc_smem_bytes(int fold_T);
size_t qkv_tc_smem_bytes(int fold_T);
int ode_tail_tc_threads();

// backward.cu
__global__ void k_gemm(int M, int N, int K, const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                       const float* bias, const float* res, long ldres, const float* mask, long ldmask, int relu, int accumulate, int kchunk,
                       long slice_stride);
// gemm_tc.cu — the same product on the tcgen05 tensor cores (bf16x3 split, fp32-level accuracy, gradient-safe range); 128 x 128 tiles
__global__ void k_gemm_tc(int M, int N, int K, const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                          const float* bias, const float* res, long ldres, const float* mask, long ldmask, int relu, int accumulate, int kchunk,
                          long slice_stride);
size_t gemm_tc_smem_bytes();
int gemm_tc_threads();
__global__ void k_colsum_part(const float* X, long ldx, int rows, int N, int rchunk, float* part);
__global__ void k_splitk_epilogue(const float* part, int slices, int M, int N, float* C, long ldc, const float* bias, const float* res, long ldres,
                                  const float* mask, long ldmask, int relu, int accumulate);
__global__ void k_sum_slices(const float* part, int slices, long stride, long count, float* out);
__global__ void k_colsum(const float* X, long ldx, int rows, int N, float* out);
__global__ void k_gate_ln_fwd(const float* x, const float* ui, const float* ug, const float* gamma, const float* beta, float* z1,
                              float* s1, int rows);
__global__ void k_ln_bwd(const float* z, const float* gamma, const float* ds, long ldds, float* dz, float* xhat_ds, int rows);
__global__ void k_gate_bwd(const float* ui, const float* ug, const float* dg, long lddg, float* dui, float* dug, int rows);
__global__ void k_rows_axpb(const float* a, long lda, float alpha, const float* gate, long ldg, const float* b, long ldb, float* out,
                            long ldo, int rows);
__global__ void k_geo_attn_bwd(const float* qh, int ldq, const float* kh, int ldk, const float* v, int ldv, const float* nq,
                               const float* nk, int ldn, const float* dout, int ldo, float* dq, int lddq, float* dk, int lddk,
                               float* dv, int lddv, int L, int S, int bsz, int swap, float lo, float hi, float qscale);
size_t geo_attn_bwd_smem_bytes(int L, int S);
__global__ void k_conv_im2col(const float* x_true, const float* x_hat, int P, int T, float* col);
__global__ void k_conv_col2im_neg(const float* dcol, const float* base, int P, int T, float* out);
__global__ void k_gru_step_fwd(const float* gi, long ldgi, const float* gh, const float* h, int P, float* rzn, float* ghn, float* hout);
__global__ void k_gru_step_bwd(const float* dh, const float* rzn, const float* ghn, const float* hprev, int P, float* dgi, long lddgi, float* dgh,
                               float* dh_prev);
__global__ void k_gru_seq_fwd(const float* gi, long ldgi, const float* w_hh, const float* b_hh, int P, int T, float* h, float* rzn, float* ghn);
__global__ void k_gru_seq_bwd(const float* dh_T, long lddh, const float* w_hh, const float* rzn, const float* ghn, const float* h, int P, int T,
                              float* dgi, long lddgi, float* dgh);
size_t gru_seq_fwd_smem_bytes();
size_t gru_seq_bwd_smem_bytes();
int gru_seq_rows();
int gru_seq_threads();
__global__ void k_rows_copy(const float* src, long lds, float* dst, long ldd, int rows, int n);
__global__ void k_rows_add(const float* a, long lda, const float* b, long ldb, float* out, long ldo, int rows, int n);

// decoder.cu
constexpr int ROWPRE_LD = 2048;     // per-row layer-1 partial sums: [P0x | P0y | P1y | P1x], 512 each
// rowpre lives in HBM in tiles of 128 rows, [tile][column / 4][row % 128][4 floats]: the row kernels hold one row per thread, so a warp's 16-byte stores
// of one column group land in 512 contiguous bytes (row-major [R][2048] made every store instruction touch 32 half-used sectors: the tensor-core row
// kernel was bound by the store queue), and the pair kernels' 16-byte loads of neighbouring agents share sectors.  Rows are padded to a multiple of 128.
__host__ __device__ __forceinline__ size_t rowpre_off(int row, int col) {
    return (size_t)(row >> 7) * (128 * ROWPRE_LD) + (size_t)(col >> 2) * 512 + (size_t)(row & 127) * 4 + (col & 3);
}
__host__ __device__ constexpr size_t rowpre_col(int col0) { return (size_t)(col0 >> 2) * 512; }      // pointer offset of column col0 (a multiple of 4): rowpre_off is linear in it
inline size_t rowpre_floats(int rows) { return (size_t)((rows + 127) / 128) * 128 * ROWPRE_LD; }
__global__ void k_dec_row(const float* past_feature, const float* traj_norm, int rows, int Tp, sttode_decompose_weights w0,
                          sttode_decompose_weights w1, float* rowpre, const int* only_flagged);
__global__ void k_dec_pair(const float* traj_norm, const float* z, unsigned long long seed, unsigned long long pair_base, const unsigned long long* rng_state,
                           const float* rowpre,
                           const int* row_scene, const float* origin, int rows, int K, int rpt, int Tp, int Tf,
                           sttode_decompose_weights w0, sttode_decompose_weights w1, float* pred, float* recon, const int* only_flagged, int ntiles);
size_t dec_row_smem_bytes(int Tp);
size_t dec_pair_smem_bytes(int Tp, int Tf);

// decoder_tc.cu — tensor-core decoder tile + weight packing
constexpr int TC_NST = 4;            // weight ring stages
constexpr int TC_STAGE = 24576;      // bytes per ring stage
constexpr int TC_NCH = 32;           // ring blocks per MLP layer-1/2 sweep: 16 chunks of 32 hidden units x (hi block, lo block)
constexpr int TC_LEAD = 4;           // lead blocks of an MLP: layer-1 rows of hidden group 0, first K half (hi, lo), then [second K half | first K slice of group 1] (hi, lo)
constexpr int TC_MLP_BLOCKS = TC_LEAD + TC_NCH + 2;    // per MLP: lead blocks + TC_NCH chunk blocks + layer-3 pair
constexpr int TC_GRU_BLK = 18432;    // one K = 16 slab of [W_ih | W_hh]: 288 rows (hi + lo), split into an A image (units 0..63) and a B image (units 64..95)
constexpr int TC_GRU_A = 12288;      //   A slab image: 192 rows (hi 6144 B | lo 6144 B); x-slabs (conv input) in row order [n | r | z], h-slabs in [r | z | n]
constexpr int TC_GRU_B = 6144;       //   B slab image:  96 rows (hi 3072 B | lo 3072 B), same row orders
constexpr int TC_GRU_STEP_BLOCKS = 8;   // ring blocks per GRU step, two slabs each, in issue order: A x0x1, B x0x1 | B h0h1, B h2h3, A h0h1, A h2h3 | A h4h5, B h4h5
struct DecPackLayout {               // byte offsets inside the packed weight buffer; MLP order: x0, y0, y1, x1
    uint32_t mlp_off[4], lead_blk[4], slice_blk[4], mlp_blk[4], l3_off[4], l3_blk[4], n3p[4], n3[4], k1[4];
    uint32_t gru_off[2];             // decompose[0], decompose[1]: 4 A-blocks (2 slabs, 2 * TC_GRU_A each) followed by 4 B-blocks (2 * TC_GRU_B each)
    uint32_t rowp_off;               // k_dec_row_tc layer-1 partial-product slabs
    uint32_t total;                  // bytes of images; a 256-byte status tail follows
};
DecPackLayout dec_pack_layout(int Tp, int Tf);
__global__ void k_pack_image(const float* W, int ldw, int row0, int col0, int nvalid, int nrows, int K, unsigned char* hi,
                             unsigned char* lo);
void launch_pack_all(const sttode_decompose_weights& w0, const sttode_decompose_weights& w1, const DecPackLayout& L, unsigned char* pk, cudaStream_t st);
__global__ void k_dec_pair_tc(const float* traj_norm, const float* z, unsigned long long seed, unsigned long long pair_base, const unsigned long long* rng_state,
                              const float* rowpre,
                              const int* row_scene, const float* origin, int rows, int K, int rpt, int Tp, int Tf,
                              sttode_decompose_weights w0, sttode_decompose_weights w1, const unsigned char* packed,
                              const __grid_constant__ DecPackLayout lay, float* pred, float* recon, int* status, int* tile_flags, int ntiles, int* tile_counter, int y_first);
__global__ void k_dec_row_tc(const float* past_feature, const float* traj_norm, int rows, int Tp, sttode_decompose_weights w0,
                             sttode_decompose_weights w1, const unsigned char* packed, const __grid_constant__ DecPackLayout lay, float* rowpre, int* tile_flags);
size_t dec_row_tc_smem_bytes();
size_t dec_pair_tc_smem_bytes(int Tp, int Tf, int y_first);
int dec_pair_tc_threads();

__global__ void k_bump_counter(unsigned long long* p);
__global__ void k_latent_draw(unsigned long long seed, unsigned long long pair_base, const unsigned long long* rng_state, size_t n_pairs, float* z);

// comm.cu — gradient all-reduce fused with Adam over NVLink peer memory
constexpr int PEER_MAX_WORLD = 16, PEER_EPOCH = 16, PEER_ERR = 17, PEER_STEP = 18, PEER_FLAG_WORDS = 64;
void launch_peer_barrier(unsigned* const* flags_of, int world, int rank, int bump_step, cudaStream_t st);
void launch_allreduce_adam(const float* const* grads_of, int world, int rank, unsigned* const* flags_of, float* param, float* exp_avg, float* exp_avg_sq,
                           float* grad_avg, size_t n4, const float* lr_dev, float beta1, float beta2, float eps, cudaStream_t st);

// metrics.cu
__global__ void k_min_ade_fde(const float* pred, const float* gt, int gt_layout, int rows, int K, int Tf, double* sums);

}  // namespace sttode
