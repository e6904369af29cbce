// K-sample residual decoder (Decoder / DecomposeBlock, model/STTODE.py:16-77, 302-347), sm_100a, FP32 FFMA.
//
// Exact algebraic restructuring of the reference's op list (only summation order changes):
//   * block 0 sees x_true - 0 = past trajectory, identical for the K samples of an agent, so its
//     Conv1d+GRU runs once per agent row (k_dec_row), not once per (agent, sample);
//   * MLP layer 1 acts on [past_feature(128) | z(32) | gru_state(96)]; every part that does not depend
//     on the sample is pre-multiplied once per agent row into `rowpre` and enters the per-pair kernel as
//     a bias: block 0 keeps only the 32 z-columns per pair, block 1 the z- and state-columns (128);
//   * block 1's decoder_x is skipped unless the caller asks for the reconstruction (dead at inference,
//     model/STTODE.py:619 drops it).
// k_dec_pair then runs, for a tile of (agent, sample) pairs resident in shared memory:
//   decoder_x(block 0) -> x_hat ; Conv1d+GRU(block 1) on past - x_hat ; decoder_y(block 0) ; decoder_y(block 1)
//   ; pred = y0 + y1 + cur_location (+ scene origin).  The latent z is read or drawn in-kernel (Philox).
#include "kernels.cuh"

namespace sttode {

constexpr int LDZ = 128 + PAD;       // [z(32) | state(96)] per pair
constexpr int LDH = STTODE_GRU + PAD;
constexpr int LDX = 32 + PAD;        // conv output / 32-wide K chunks
constexpr int GRU_CHUNK_FLOATS = 48 * LDX + 48 * LDH;   // one 16-unit gate chunk of W_ih and W_hh
constexpr int GRU_SCRATCH = BM * LDH + BM * LDX + 2 * GRU_CHUNK_FLOATS;      // hB | xt | wch
constexpr int MLP_SCRATCH = BM * LDX + 2 * 256 * LDX + 2 * 32 * LDZ;         // h1c | w2c | w1c
constexpr int ROWP_SCRATCH = BM * LDZ + 2 * 64 * LDX;                         // pf_s | wbuf
constexpr int PAIR_REGION = MLP_SCRATCH > GRU_SCRATCH ? MLP_SCRATCH : GRU_SCRATCH;
constexpr int ROW_REGION = ROWP_SCRATCH > GRU_SCRATCH ? ROWP_SCRATCH : GRU_SCRATCH;

// ------------------------------------------------------------------------------------------------
// Conv1d(2->32,k3,pad1)+ReLU -> GRU(32->96) over T steps for 128 rows; d_s[m][t*2+c] is the conv input.
// The final state is left in hA (ld ldA).  hB/xt_s/wch are scratch.  Gate order r,z,n (torch nn.GRU).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gru_load_chunk(float* wch, const sttode_decompose_weights& w, int c) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        load_w_chunk_async(wch + g * 16 * LDX, LDX, w.gru_w_ih + (size_t)(g * 96 + c * 16) * 32, 32, 16, 32);
        load_w_chunk_async(wch + 48 * LDX + g * 16 * LDH, LDH, w.gru_w_hh + (size_t)(g * 96 + c * 16) * 96, 96, 16, 96);
    }
}

__device__ void gru_phase(const float* d_s, int ldd, int T, const sttode_decompose_weights& w, float* hA, int ldA,
                          float* hB, int ldB, float* xt_s, float* wch) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float* cur = (T & 1) ? hB : hA;
    int ldcur = (T & 1) ? ldB : ldA;
    float* nxt = (T & 1) ? hA : hB;
    int ldnxt = (T & 1) ? ldA : ldB;
    for (int idx = tid; idx < BM * 96; idx += NTHREADS) cur[(idx / 96) * ldcur + idx % 96] = 0.f;
    float cw[2][6], cb[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ch = tx + q * 16;
        cb[q] = __ldg(w.conv_b + ch);
#pragma unroll
        for (int e = 0; e < 6; ++e) cw[q][e] = __ldg(w.conv_w + ch * 6 + e);   // [ch][ci][k]
    }
    int wrow[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) wrow[g] = g * 16 + tx;
    const int Q = T * 6;
    gru_load_chunk(wch, w, 0);
    cp_async_commit();
    for (int q = 0; q < Q; ++q) {
        const int t = q / 6, c = q - t * 6;
        cp_async_wait<0>();
        __syncthreads();
        if (q + 1 < Q) {
            gru_load_chunk(wch + ((q + 1) & 1) * GRU_CHUNK_FLOATS, w, (c + 1) % 6);
            cp_async_commit();
        }
        if (c == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = i * 16 + ty;
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    float v = cb[qq];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int tt = t + k - 1;
                        if (tt >= 0 && tt < T) {
                            v = fmaf(cw[qq][k], d_s[m * ldd + tt * 2], v);
                            v = fmaf(cw[qq][3 + k], d_s[m * ldd + tt * 2 + 1], v);
                        }
                    }
                    xt_s[m * LDX + tx + qq * 16] = fmaxf(v, 0.f);
                }
            }
            __syncthreads();
        }
        const float* wb = wch + (q & 1) * GRU_CHUNK_FLOATS;
        float ax[8][3], ah[8][3];
        zero_acc(ax);
        zero_acc(ah);
        mma_tile<3>(ax, xt_s, LDX, wb, LDX, wrow, 32, ty);
        mma_tile<3>(ah, cur, ldcur, wb + 48 * LDX, LDH, wrow, 96, ty);
        const int u = c * 16 + tx;
        const float bir = __ldg(w.gru_b_ih + u), biz = __ldg(w.gru_b_ih + 96 + u), bin = __ldg(w.gru_b_ih + 192 + u);
        const float bhr = __ldg(w.gru_b_hh + u), bhz = __ldg(w.gru_b_hh + 96 + u), bhn = __ldg(w.gru_b_hh + 192 + u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = i * 16 + ty;
            const float r = sigmoidf_((ax[i][0] + bir) + (ah[i][0] + bhr));
            const float zg = sigmoidf_((ax[i][1] + biz) + (ah[i][1] + bhz));
            const float n = tanhf((ax[i][2] + bin) + r * (ah[i][2] + bhn));
            const float hp = cur[m * ldcur + u];
            nxt[m * ldnxt + u] = (hp - n) * zg + n;          // ATen gru_cell form of (1-z)*n + z*h
        }
        if (c == 5) {
            float* tp = cur; cur = nxt; nxt = tp;
            int tl = ldcur; ldcur = ldnxt; ldnxt = tl;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// per-agent-row precompute: block-0 GRU state and the sample-independent layer-1 partial sums.
// rowpre[r] = [ W0x[:, pf|st].[pf|st0]+b | W0y[:, pf|st].[pf|st0]+b | W1y[:, pf].pf+b | W1x[:, pf].pf+b ]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
k_dec_row(const float* __restrict__ past_feature, const float* __restrict__ traj_norm, int rows, int Tp,
          sttode_decompose_weights w0, sttode_decompose_weights w1, float* __restrict__ rowpre, const int* __restrict__ only_flagged) {
    extern __shared__ __align__(16) float smem[];
    if (only_flagged && !only_flagged[blockIdx.x]) return;      // fp32 redo of the tiles the tensor-core kernel flagged (fp16 range)
    const int ldd = 2 * Tp + 1;
    float* hA = smem;                          // [BM][LDH]  final state
    float* scratch = hA + BM * LDH;            // GRU: hB | xt | wch      P-phase: pf_s | wbuf
    float* hB = scratch;
    float* xt_s = hB + BM * LDH;
    float* wch = xt_s + BM * LDX;
    float* d_s = scratch + ROW_REGION;         // [BM][ldd]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * BM;
    for (int idx = tid; idx < BM * 2 * Tp; idx += NTHREADS) {
        const int m = idx / (2 * Tp), e = idx - m * 2 * Tp, r = row0 + m;
        d_s[m * ldd + e] = r < rows ? traj_norm[(size_t)r * 2 * Tp + e] : 0.f;
    }
    __syncthreads();
    gru_phase(d_s, ldd, Tp, w0, hA, LDH, hB, LDH, xt_s, wch);

    float* pf_s = scratch;                     // [BM][LDZ] (128 + pad)
    float* wbuf = pf_s + BM * LDZ;             // [2][64][36]
    for (int idx = tid; idx < BM * 32; idx += NTHREADS) {
        const int m = idx >> 5, c4 = (idx & 31) << 2, r = row0 + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) v = *reinterpret_cast<const float4*>(past_feature + (size_t)r * 128 + c4);
        *reinterpret_cast<float4*>(pf_s + m * LDZ + c4) = v;
    }
    // gemm_step's leading barrier orders the pf_s writes before the first read
    const float* W[4] = {w0.dec_x.w0, w0.dec_y.w0, w1.dec_y.w0, w1.dec_x.w0};
    const float* Bv[4] = {w0.dec_x.b0, w0.dec_y.b0, w1.dec_y.b0, w1.dec_x.b0};
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
#pragma unroll 1
        for (int nb = 0; nb < 8; ++nb) {
            float acc[8][4];
            zero_acc(acc);
            const float* wblk = W[p] + (size_t)nb * 64 * 256;
            gemm_step<4, 32>(acc, pf_s, LDZ, wblk, 256, 128, wbuf);
            if (p < 2) gemm_step<4, 32>(acc, hA, LDH, wblk + 160, 256, 96, wbuf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = row0 + i * 16 + ty;
                if (r >= rows) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = nb * 64 + j * 16 + tx;
                    rowpre[rowpre_off(r, p * 512 + n)] = acc[i][j] + __ldg(Bv[p] + n);
                }
            }
        }
    }
}

size_t dec_row_smem_bytes(int Tp) {
    return sizeof(float) * ((size_t)BM * LDH + ROW_REGION + (size_t)BM * (2 * Tp + 1));
}

// ------------------------------------------------------------------------------------------------
// 3-layer MLP on a 128-pair tile.  Layer 1: pair part A1[128][K1] . W1[:, 128:128+K1]^T + rowpre bias,
// ReLU; layer 2 (512->256) accumulated in registers over 16 hidden chunks of 32; layer 3 (256->n3) as a
// per-thread partial dot + 16-lane shuffle reduction.  Result (+= when accumulate) in out_s[m][n3].
// ------------------------------------------------------------------------------------------------
template <int K1>
__device__ void mlp_phase(const float* A1, int lda, const sttode_mlp3_weights& w, const float* __restrict__ rowpre,
                          const int* prow_s, int n3, float* out_s, int ldo, bool accumulate, float* h1c, float* w2c, float* w1c) {
    constexpr int LDW1 = K1 + PAD;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int wrow2[16], wrow1[2];
#pragma unroll
    for (int j = 0; j < 16; ++j) wrow2[j] = j * 16 + tx;
    wrow1[0] = tx; wrow1[1] = 16 + tx;
    float acc2[8][16];
    zero_acc(acc2);
    int prow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) prow[i] = prow_s[i * 16 + ty];

    load_w_chunk_async(w1c, LDW1, w.w0 + 128, 256, 32, K1);
    load_w_chunk_async(w2c, LDX, w.w1, 512, 256, 32);
    cp_async_commit();
#pragma unroll 1
    for (int hc = 0; hc < 16; ++hc) {
        cp_async_wait<0>();
        __syncthreads();
        if (hc + 1 < 16) {
            load_w_chunk_async(w1c + ((hc + 1) & 1) * 32 * LDW1, LDW1, w.w0 + (size_t)(hc + 1) * 32 * 256 + 128, 256, 32, K1);
            load_w_chunk_async(w2c + ((hc + 1) & 1) * 256 * LDX, LDX, w.w1 + (hc + 1) * 32, 512, 256, 32);
            cp_async_commit();
        }
        float acc1[8][2];
        zero_acc(acc1);
        mma_tile<2>(acc1, A1, lda, w1c + (hc & 1) * 32 * LDW1, LDW1, wrow1, K1, ty);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = i * 16 + ty;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = hc * 32 + j * 16 + tx;
                const float rb = prow[i] >= 0 ? __ldg(rowpre + rowpre_off(prow[i], n)) : 0.f;
                h1c[m * LDX + j * 16 + tx] = fmaxf(acc1[i][j] + rb, 0.f);
            }
        }
        __syncthreads();
        mma_tile<16>(acc2, h1c, LDX, w2c + (hc & 1) * 256 * LDX, LDX, wrow2, 32, ty);
    }
    // layer-2 bias + ReLU, then layer 3
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float b = __ldg(w.b1 + j * 16 + tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc2[i][j] = fmaxf(acc2[i][j] + b, 0.f);
    }
#pragma unroll 1
    for (int o = 0; o < n3; ++o) {
        float part[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float wv = __ldg(w.w2 + (size_t)o * 256 + j * 16 + tx);
#pragma unroll
            for (int i = 0; i < 8; ++i) part[i] = fmaf(acc2[i][j], wv, part[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = part[i];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            part[i] = v;
        }
        if (tx == (o & 15)) {
            const float b = __ldg(w.b2 + o);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float* dst = out_s + (i * 16 + ty) * ldo + o;
                *dst = accumulate ? *dst + (part[i] + b) : part[i] + b;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_dec_pair(const float* __restrict__ traj_norm, const float* __restrict__ z, unsigned long long seed, unsigned long long pair_base,
           const unsigned long long* __restrict__ rng_state,
           const float* __restrict__ rowpre, const int* __restrict__ row_scene, const float* __restrict__ origin, int rows,
           int K, int rpt, int Tp, int Tf, sttode_decompose_weights w0, sttode_decompose_weights w1,
           float* __restrict__ pred, float* __restrict__ recon, const int* __restrict__ only_flagged, int ntiles) {
    extern __shared__ __align__(16) float smem[];
    const int Tm = Tp > Tf ? Tp : Tf;
    const int ldxy = 2 * Tm + 1;
    float* zA = smem;                               // [BM][LDZ]: z | block-1 GRU state
    float* region = zA + BM * LDZ;                  // MLP: h1c | w2c | w1c      GRU: hB | xt | wch
    float* h1c = region;                            // [BM][36]
    float* w2c = h1c + BM * LDX;                    // [2][256][36]
    float* w1c = w2c + 2 * 256 * LDX;               // [2][32][132]
    float* hB = region;                             // [BM][LDH]
    float* xt_s = hB + BM * LDH;
    float* wch = xt_s + BM * LDX;
    float* xy_s = region + PAIR_REGION;             // [BM][ldxy]: x_hat0 -> conv input -> y accumulators
    int* prow_s = reinterpret_cast<int*>(xy_s + BM * ldxy);   // [BM]

    const int tid = threadIdx.x;
    if (rng_state) seed += *rng_state;
    // One tile per CTA when this is the decoder itself (gridDim.x == ntiles).  As the FP32 redo pass behind the tensor-core kernel it is launched
    // with a SMALL grid that walks the tile flags: launching one CTA per tile just to see a clear flag cost ~25 us per 8 960-tile forecast.
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (only_flagged && !only_flagged[tile]) continue;
    // a tile is 128 consecutive (agent, sample) pairs of the whole batch (same cut as k_dec_pair_tc); `rpt` = index of this launch's first tile
    const size_t pair0 = ((size_t)rpt + tile) * BM;
    const size_t P = (size_t)rows * K;
    const int np = (int)min((size_t)BM, P - pair0);          // valid pairs in this tile
    if (tid < BM) prow_s[tid] = tid < np ? (int)((pair0 + tid) / K) : -1;
    // latent z: one float4 per (pair, q)
    for (int idx = tid; idx < BM * 8; idx += NTHREADS) {
        const int m = idx >> 3, q = idx & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < np) {
            const size_t pair = pair0 + m;
            v = z ? *reinterpret_cast<const float4*>(z + pair * 32 + q * 4) : latent_normal4(seed, pair_base + pair, q);      // pair_base: index of this launch's first pair in the caller's batch
        }
        *reinterpret_cast<float4*>(zA + m * LDZ + q * 4) = v;
    }
    __syncthreads();

    // ---- block 0 decoder_x -> x_hat0
    mlp_phase<32>(zA, LDZ, w0.dec_x, rowpre + 0, prow_s, 2 * Tp, xy_s, ldxy, false, h1c, w2c, w1c);
    if (recon)
        for (int idx = tid; idx < np * 2 * Tp; idx += NTHREADS) {
            const int m = idx / (2 * Tp), e = idx - m * 2 * Tp;
            recon[(pair0 + m) * 2 * Tp + e] = xy_s[m * ldxy + e];
        }
    // conv input of block 1: x_true - x_hat0  (model/STTODE.py:62)
    for (int idx = tid; idx < BM * 2 * Tp; idx += NTHREADS) {
        const int m = idx / (2 * Tp), e = idx - m * 2 * Tp;
        const int r = prow_s[m];
        xy_s[m * ldxy + e] = r >= 0 ? traj_norm[(size_t)r * 2 * Tp + e] - xy_s[m * ldxy + e] : 0.f;
    }
    __syncthreads();
    // ---- block 1 Conv1d+GRU -> state1 in zA[:, 32:128]
    gru_phase(xy_s, ldxy, Tp, w1, zA + 32, LDZ, hB, LDH, xt_s, wch);
    // ---- decoder_y of both blocks, summed (prediction += y_hat, model/STTODE.py:338)
    mlp_phase<32>(zA, LDZ, w0.dec_y, rowpre + rowpre_col(512), prow_s, 2 * Tf, xy_s, ldxy, false, h1c, w2c, w1c);
    mlp_phase<128>(zA, LDZ, w1.dec_y, rowpre + rowpre_col(1024), prow_s, 2 * Tf, xy_s, ldxy, true, h1c, w2c, w1c);
    // pred = (y0 + y1) + cur_location (+ scene origin); the tile is one contiguous span of pred
    for (int idx = tid; idx < np * 2 * Tf; idx += NTHREADS) {
        const int m = idx / (2 * Tf), e = idx - m * 2 * Tf, c = e & 1;
        const int r = prow_s[m];
        float v = xy_s[m * ldxy + e] + traj_norm[((size_t)r * Tp + Tp - 1) * 2 + c];
        if (origin) v += origin[row_scene[r] * 2 + c];
        pred[(pair0 + m) * 2 * Tf + e] = v;
    }
    if (recon) {    // reconstruction += block-1 decoder_x (only when the caller wants recover_pre_seq)
        __syncthreads();
        mlp_phase<128>(zA, LDZ, w1.dec_x, rowpre + rowpre_col(1536), prow_s, 2 * Tp, xy_s, ldxy, false, h1c, w2c, w1c);
        for (int idx = tid; idx < np * 2 * Tp; idx += NTHREADS) {
            const int m = idx / (2 * Tp), e = idx - m * 2 * Tp;
            recon[(pair0 + m) * 2 * Tp + e] += xy_s[m * ldxy + e];
        }
    }
    __syncthreads();                                // the next tile of this CTA reuses every shared buffer
    }
}

__global__ void k_bump_counter(unsigned long long* p) { *p += 1ull; }

// The latent z the decoder kernels draw for (seed, pair) when no z is injected, written out: z[pair][4q..4q+3] = latent_normal4(seed', pair_base + pair, q),
// seed' = seed + *rng_state.  Same device function as k_dec_pair / k_dec_pair_tc call, so a test can check the draw itself (moments, independence)
// and feed exactly these values to the oracle (sttode_latent_draw).
__global__ void k_latent_draw(unsigned long long seed, unsigned long long pair_base, const unsigned long long* __restrict__ rng_state, size_t n_pairs,
                              float* __restrict__ z) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs * 8) return;
    if (rng_state) seed += *rng_state;
    const size_t pair = i >> 3;
    *reinterpret_cast<float4*>(z + i * 4) = latent_normal4(seed, pair_base + pair, (unsigned)(i & 7));
}

size_t dec_pair_smem_bytes(int Tp, int Tf) {
    const int Tm = Tp > Tf ? Tp : Tf;
    return sizeof(float) * ((size_t)BM * LDZ + PAIR_REGION + (size_t)BM * (2 * Tm + 1)) + sizeof(int) * BM;
}

}  // namespace sttode
