// Tensor-core (tcgen05 / TMEM) version of the per-(agent,sample) decoder tile, sm_100a.
//
// Same algebra as k_dec_pair (decoder.cu): decoder_x0 -> Conv1d+GRU1 -> decoder_y0 + decoder_y1 -> pred, for a tile of
// 128 pairs resident on one SM.  Every dense contraction runs as tcgen05.mma kind::f16 with the 3xFP16 split of
// tc.cuh (fp32-accurate), accumulators in TMEM:
//   MLP layer 1 : D1[128x16]   = [z|state](K1) . W1c^T          per 16-unit hidden chunk (double buffered in TMEM)
//   MLP layer 2 : D2[128x256] += relu(D1 + rowpre) . W2c^T      (the register epilogue of layer 1 re-stages the chunk as A)
//   MLP layer 3 : D3[128xn3p] += relu(D2 + b1)[:,16p..] . W3p^T
//   GRU step    : D_r|D_z|D_ni|D_nh [128x96 each] = conv(x)_t . W_ih^T (+) h . W_hh^T ; gates in the register epilogue
// Weights are pre-split into fp16 hi/lo images in the canonical no-swizzle K-major layout by k_pack_image (once per
// weight version) and streamed L2 -> shared memory with 1-D bulk copies (cp.async.bulk, UBLKCP) through a 4-stage ring
// guarded by mbarriers; tcgen05.commit releases ring stages and activation buffers.  One thread issues the copies and
// the MMAs; all 256 threads run the TMEM epilogues (warp w reads lanes 32*(w%4).., warps w and w+4 split the columns).
#include "kernels.cuh"
#include "tc.cuh"

namespace sttode {

__global__ void k_pack_image(const float* __restrict__ W, int ldw, int row0, int col0, int nvalid, int nrows, int K,
                             unsigned char* __restrict__ hi, unsigned char* __restrict__ lo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (K >> 3) * nrows) return;
    const int kc = idx / nrows, r = idx - kc * nrows;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = r < nvalid ? W[(size_t)(row0 + r) * ldw + col0 + kc * 8 + i] : 0.f;
    tc::store_split8(hi, lo, nrows, r, kc, v);
}

DecPackLayout dec_pack_layout(int Tp, int Tf) {
    DecPackLayout L{};
    const int n3[4] = {2 * Tp, 2 * Tf, 2 * Tf, 2 * Tp};      // x0, y0, y1, x1
    const int k1[4] = {32, 32, 128, 128};
    uint32_t off = 0;
    for (int m = 0; m < 4; ++m) {
        L.k1[m] = k1[m];
        L.n3[m] = n3[m];
        L.n3p[m] = (n3[m] + 15) / 16 * 16;
        L.mlp_blk[m] = 2 * 16 * k1[m] * 2 + 2 * 256 * 16 * 2;
        L.mlp_off[m] = off;
        off += TC_NCH * L.mlp_blk[m];
        L.l3_blk[m] = L.n3p[m] * 512;
        L.l3_off[m] = off;
        off += 2 * L.l3_blk[m];
    }
    for (int g = 0; g < 2; ++g) { L.gru_off[g] = off; off += 6 * TC_STAGE; }
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------------------
constexpr int SM_A1 = 0;                         // [hi 32 KB | lo 32 KB]   A1 = [z (k 0..31) | state (k 32..127)]
constexpr int SM_H = 65536;                      // 2 x [hi 4 KB | lo 4 KB] (MLP chunk operand)  /  GRU: conv image hi 8 KB | lo 8 KB
constexpr int SM_RING = SM_H + 16384;
constexpr int SM_XY = SM_RING + TC_NST * TC_STAGE;
constexpr int COL_D2 = 0, COL_D1 = 256, COL_D3 = 288;     // TMEM columns, MLP phases
constexpr int COL_GR = 0, COL_GZ = 96, COL_GNI = 192, COL_GNH = 288;   // GRU phase

struct TcCtx {
    unsigned char* sm;
    uint64_t *wfull, *wfree, *d1full, *hfree, *accfull;
    int* failed;
    uint32_t tmem;
    // per-thread (uniform) phase counters
    uint32_t n_d1[2], n_h[2], n_acc;
    // producer state (meaningful on thread 0)
    int ld, total_blocks;
    const unsigned char* packed;
    DecPackLayout lay;
    int Tp, recon;
};

__device__ __forceinline__ void tc_wait(TcCtx& c, uint64_t* bar, uint32_t parity) {
    if (*(volatile int*)c.failed) return;
    if (!tc::mbar_wait(bar, parity)) *(volatile int*)c.failed = 1;
}

// global block sequence: [mlp x0: 32 chunks, 2 L3] [gru1: 6*Tp] [mlp y0: 34] [mlp y1: 34] [mlp x1: 34 when recon]
__device__ __forceinline__ void block_src(const TcCtx& c, int i, const unsigned char*& src, uint32_t& bytes) {
    int m = -1;
    if (i < 34) m = 0;
    else {
        i -= 34;
        if (i < 6 * c.Tp) { src = c.packed + c.lay.gru_off[1] + (size_t)(i % 6) * TC_STAGE; bytes = TC_STAGE; return; }
        i -= 6 * c.Tp;
        if (i < 34) m = 1;
        else if (i < 68) { m = 2; i -= 34; }
        else { m = 3; i -= 68; }
    }
    if (i < TC_NCH) { src = c.packed + c.lay.mlp_off[m] + (size_t)i * c.lay.mlp_blk[m]; bytes = c.lay.mlp_blk[m]; }
    else { src = c.packed + c.lay.l3_off[m] + (size_t)(i - TC_NCH) * c.lay.l3_blk[m]; bytes = c.lay.l3_blk[m]; }
}

__device__ __forceinline__ void ring_load_next(TcCtx& c) {      // thread 0
    const int s = c.ld % TC_NST;
    const unsigned char* src; uint32_t bytes;
    block_src(c, c.ld, src, bytes);
    tc::mbar_arrive_expect_tx(c.wfull + s, bytes);
    tc::bulk_g2s(c.sm + SM_RING + s * TC_STAGE, src, bytes, c.wfull + s);
    ++c.ld;
}
__device__ __forceinline__ void ring_refill(TcCtx& c) {         // thread 0, never blocks
    while (c.ld < c.total_blocks) {
        const int s = c.ld % TC_NST;
        if (c.ld >= TC_NST && !tc::mbar_try_wait(c.wfree + s, (uint32_t)((c.ld / TC_NST - 1) & 1))) break;
        ring_load_next(c);
    }
}
// thread 0: make block i resident and return its stage base (shared-window address)
__device__ __forceinline__ uint32_t ring_acquire(TcCtx& c, int i) {
    while (c.ld <= i) {
        const int s = c.ld % TC_NST;
        if (c.ld >= TC_NST) tc_wait(c, c.wfree + s, (uint32_t)((c.ld / TC_NST - 1) & 1));
        ring_load_next(c);
    }
    tc_wait(c, c.wfull + (i % TC_NST), (uint32_t)((i / TC_NST) & 1));
    return tc::smem_u32(c.sm + SM_RING + (i % TC_NST) * TC_STAGE);
}

// stage 8 values of one row as kchunk `kc` of the chunk operand H[b], after the previous reader of H[b] has finished
__device__ __forceinline__ void stage_h(TcCtx& c, int b, int m, int kc, const float (&v)[8]) {
    if (c.n_h[b] >= 1) tc_wait(c, c.hfree + b, (c.n_h[b] - 1) & 1);
    ++c.n_h[b];
    tc::store_split8(c.sm + SM_H + b * 8192, c.sm + SM_H + b * 8192 + 4096, 128, m, kc, v);
}

// One 3-layer MLP on the tile.  blk0 = index of its first ring block.  Result (+=) into xy_s[m][0..n3).
__device__ void mlp_tc(TcCtx& c, int mi, int blk0, const float* __restrict__ rowpre, int prow, const float* __restrict__ b1,
                       const float* __restrict__ b2, float* xy_s, int ldxy, bool accumulate) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, wh = warp >> 2;
    const int m = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const int K1 = c.lay.k1[mi], n3 = c.lay.n3[mi], n3p = c.lay.n3p[mi];
    const uint32_t a1_hi = tc::smem_u32(c.sm + SM_A1), a1_lo = a1_hi + 32768;
    const uint32_t h_base = tc::smem_u32(c.sm + SM_H);
    const uint32_t w1_lo_off = 16 * K1 * 2, w2_off = 2 * w1_lo_off;

    if (tid == 0) {
        const uint32_t st = ring_acquire(c, blk0);
        tc::tc_fence_after();
        tc::mma_split3(c.tmem + COL_D1, a1_hi, a1_lo, st, st + w1_lo_off, 16, 0, 16, K1, false);
        tc::mma_commit(c.d1full + 0);
    }
#pragma unroll 1
    for (int ch = 0; ch < TC_NCH; ++ch) {
        const int b = ch & 1;
        if (tid == 0 && ch + 1 < TC_NCH) {          // look-ahead: keep the tensor pipe fed during this chunk's epilogue
            const uint32_t st = ring_acquire(c, blk0 + ch + 1);
            tc::tc_fence_after();
            tc::mma_split3(c.tmem + COL_D1 + (b ^ 1) * 16, a1_hi, a1_lo, st, st + w1_lo_off, 16, 0, 16, K1, false);
            tc::mma_commit(c.d1full + (b ^ 1));
        }
        float rb[8];
        if (prow >= 0) {
            const float4* p = reinterpret_cast<const float4*>(rowpre + (size_t)prow * ROWPRE_LD + ch * 16 + wh * 8);
            const float4 r0 = __ldg(p), r1 = __ldg(p + 1);
            rb[0] = r0.x; rb[1] = r0.y; rb[2] = r0.z; rb[3] = r0.w; rb[4] = r1.x; rb[5] = r1.y; rb[6] = r1.z; rb[7] = r1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) rb[i] = 0.f;
        }
        tc_wait(c, c.d1full + b, c.n_d1[b] & 1);
        ++c.n_d1[b];
        tc::tc_fence_after();
        float v[8];
        tc::tmem_ld8(c.tmem + lane_base + COL_D1 + b * 16 + wh * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] + rb[i], 0.f);
        stage_h(c, b, m, wh, v);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            const uint32_t st = tc::smem_u32(c.sm + SM_RING + ((blk0 + ch) % TC_NST) * TC_STAGE);
            tc::mma_split3(c.tmem + COL_D2, h_base + b * 8192, h_base + b * 8192 + 4096, st + w2_off, st + w2_off + 8192, 256, 0, 256, 16, ch > 0);
            tc::mma_commit(c.wfree + (blk0 + ch) % TC_NST);
            tc::mma_commit(c.hfree + b);
            ring_refill(c);
        }
    }
    // layer-2 epilogue feeds layer 3 sixteen columns at a time
    if (tid == 0) tc::mma_commit(c.accfull);
    tc_wait(c, c.accfull, c.n_acc & 1);
    ++c.n_acc;
    tc::tc_fence_after();
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
        const int b = p & 1;
        float v[8];
        tc::tmem_ld8(c.tmem + lane_base + COL_D2 + p * 16 + wh * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] + __ldg(b1 + p * 16 + wh * 8 + i), 0.f);
        stage_h(c, b, m, wh, v);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            const int blk = blk0 + TC_NCH + (p >> 3);
            const uint32_t st = ring_acquire(c, blk) + (uint32_t)(p & 7) * n3p * 64;
            tc::tc_fence_after();
            tc::mma_split3(c.tmem + COL_D3, h_base + b * 8192, h_base + b * 8192 + 4096, st, st + n3p * 32, n3p, 0, n3p, 16, p > 0);
            tc::mma_commit(c.hfree + b);
            if ((p & 7) == 7) tc::mma_commit(c.wfree + blk % TC_NST);
            ring_refill(c);
        }
    }
    if (tid == 0) tc::mma_commit(c.accfull);
    tc_wait(c, c.accfull, c.n_acc & 1);
    ++c.n_acc;
    tc::tc_fence_after();
    for (int c0 = 0; c0 < n3p; c0 += 16) {
        if (((c0 >> 4) & 1) != wh) continue;        // warp-uniform: the two column halves alternate 16-wide groups
        float v[16];
        tc::tmem_ld16(c.tmem + lane_base + COL_D3 + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int o = c0 + i;
            if (o < n3) {
                const float r = v[i] + __ldg(b2 + o);
                xy_s[m * ldxy + o] = accumulate ? xy_s[m * ldxy + o] + r : r;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// Conv1d+ReLU -> GRU over T steps; conv input d = xy_s; final state left in A1 k 32..127 (fp16 hi/lo images).
__device__ void gru_tc(TcCtx& c, int blk0, const sttode_decompose_weights& w, const float* xy_s, int ldxy, int T) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, wh = warp >> 2;
    const int m = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    unsigned char* a1_hi_p = c.sm + SM_A1;
    unsigned char* a1_lo_p = a1_hi_p + 32768;
    const uint32_t a1_hi = tc::smem_u32(a1_hi_p), a1_lo = a1_hi + 32768;
    const uint32_t x_hi = tc::smem_u32(c.sm + SM_H), x_lo = x_hi + 8192;
    // h0 = 0: clear k chunks 4..15 of both images
    for (int i = tid; i < 2 * 12 * 128; i += NTHREADS) {
        const int img = i / (12 * 128), r = i - img * 12 * 128;
        *reinterpret_cast<uint4*>(c.sm + SM_A1 + img * 32768 + 4 * 2048 + r * 16) = make_uint4(0, 0, 0, 0);
    }
    float hreg[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) hreg[i] = 0.f;
    const int u0 = wh * 48;

#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        // conv channels wh*16 .. +16 of row m at time t -> k chunks 2*wh, 2*wh+1 of the X image
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int chn = wh * 16 + half * 8 + i;
                float a = __ldg(w.conv_b + chn);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int tt = t + k - 1;
                    if (tt >= 0 && tt < T) {
                        a = fmaf(__ldg(w.conv_w + chn * 6 + k), xy_s[m * ldxy + tt * 2], a);
                        a = fmaf(__ldg(w.conv_w + chn * 6 + 3 + k), xy_s[m * ldxy + tt * 2 + 1], a);
                    }
                }
                v[i] = fmaxf(a, 0.f);
            }
            tc::store_split8(c.sm + SM_H, c.sm + SM_H + 8192, 128, m, wh * 2 + half, v);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
#pragma unroll 1
            for (int blk = 0; blk < 6; ++blk) {
                const int g = blk >> 1, hf = blk & 1;
                const int bi = blk0 + t * 6 + blk;
                const uint32_t st = ring_acquire(c, bi);
                tc::tc_fence_after();
                const uint32_t wih_hi = st, wih_lo = st + 3072, whh_hi = st + 6144, whh_lo = st + 6144 + 9216;
                if (g < 2) {
                    const uint32_t d = c.tmem + (g == 0 ? COL_GR : COL_GZ) + hf * 48;
                    tc::mma_split3(d, x_hi, x_lo, wih_hi, wih_lo, 48, 0, 48, 32, false);
                    tc::mma_split3(d, a1_hi + 4 * 2048, a1_lo + 4 * 2048, whh_hi, whh_lo, 48, 0, 48, 96, true);
                } else {
                    tc::mma_split3(c.tmem + COL_GNI + hf * 48, x_hi, x_lo, wih_hi, wih_lo, 48, 0, 48, 32, false);
                    tc::mma_split3(c.tmem + COL_GNH + hf * 48, a1_hi + 4 * 2048, a1_lo + 4 * 2048, whh_hi, whh_lo, 48, 0, 48, 96, false);
                }
                tc::mma_commit(c.wfree + bi % TC_NST);
                ring_refill(c);
            }
            tc::mma_commit(c.accfull);
        }
        tc_wait(c, c.accfull, c.n_acc & 1);
        ++c.n_acc;
        tc::tc_fence_after();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int u = u0 + q * 16;
            float gr[16], gz[16], gni[16], gnh[16];
            tc::tmem_ld16(c.tmem + lane_base + COL_GR + u, gr);
            tc::tmem_ld16(c.tmem + lane_base + COL_GZ + u, gz);
            tc::tmem_ld16(c.tmem + lane_base + COL_GNI + u, gni);
            tc::tmem_ld16(c.tmem + lane_base + COL_GNH + u, gnh);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float r = fast_sigmoid(gr[i] + (__ldg(w.gru_b_ih + u + i) + __ldg(w.gru_b_hh + u + i)));
                const float zg = fast_sigmoid(gz[i] + (__ldg(w.gru_b_ih + 96 + u + i) + __ldg(w.gru_b_hh + 96 + u + i)));
                const float n = fast_tanh((gni[i] + __ldg(w.gru_b_ih + 192 + u + i)) + r * (gnh[i] + __ldg(w.gru_b_hh + 192 + u + i)));
                const float hp = hreg[q * 16 + i];
                hreg[q * 16 + i] = (hp - n) * zg + n;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = hreg[q * 16 + j * 8 + i];
                tc::store_split8(a1_hi_p, a1_lo_p, 128, m, 4 + (u >> 3) + j, v);
            }
        }
        // the next step's barrier (after its conv image is written) orders these writes before the next MMAs
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_dec_pair_tc(const float* __restrict__ traj_norm, const float* __restrict__ z, unsigned long long seed,
              const float* __restrict__ rowpre, const int* __restrict__ row_scene, const float* __restrict__ origin, int rows, int K,
              int rpt, int Tp, int Tf, sttode_decompose_weights w0, sttode_decompose_weights w1,
              const unsigned char* __restrict__ packed, DecPackLayout lay, float* __restrict__ pred, float* __restrict__ recon,
              int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int Tm = Tp > Tf ? Tp : Tf;
    const int ldxy = 2 * Tm + 1;
    float* xy_s = reinterpret_cast<float*>(sm + SM_XY);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_XY + ((BM * ldxy * 4 + 15) / 16) * 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    int* failed = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, wh = warp >> 2;
    const int m = wq * 32 + lane;
    const int row0 = blockIdx.x * rpt;
    const int nrow = min(rpt, rows - row0);
    const int np = nrow * K;
    const int prow = m < np ? row0 + m / K : -1;
    const size_t pair0 = (size_t)row0 * K;

    TcCtx c;
    c.sm = sm;
    c.wfull = bars; c.wfree = bars + 4; c.d1full = bars + 8; c.hfree = bars + 10; c.accfull = bars + 12;
    c.failed = failed;
    c.n_d1[0] = c.n_d1[1] = c.n_h[0] = c.n_h[1] = c.n_acc = 0;
    c.ld = 0;
    c.total_blocks = 34 * 3 + 6 * Tp + (recon ? 34 : 0);
    c.packed = packed; c.lay = lay; c.Tp = Tp; c.recon = recon != nullptr;

    if (tid == 0) {
        for (int i = 0; i < 13; ++i) tc::mbar_init(bars + i, 1);
        tc::fence_mbar_init();
        *failed = 0;
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    // latent z -> A1 k chunks 0..3 (this thread: chunks 2*wh, 2*wh+1 of row m)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int kc = wh * 2 + j;
        float v[8];
        if (m < np) {
            const size_t pair = pair0 + m;
            float4 lo4, hi4;
            if (z) {
                lo4 = *reinterpret_cast<const float4*>(z + pair * 32 + kc * 8);
                hi4 = *reinterpret_cast<const float4*>(z + pair * 32 + kc * 8 + 4);
            } else {
                lo4 = latent_normal4(seed, pair, kc * 2);
                hi4 = latent_normal4(seed, pair, kc * 2 + 1);
            }
            v[0] = lo4.x; v[1] = lo4.y; v[2] = lo4.z; v[3] = lo4.w; v[4] = hi4.x; v[5] = hi4.y; v[6] = hi4.z; v[7] = hi4.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        tc::store_split8(sm + SM_A1, sm + SM_A1 + 32768, 128, m, kc, v);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    c.tmem = *tmem_slot;
    if (tid == 0) ring_refill(c);

    // ---- block 0 decoder_x -> x_hat0
    mlp_tc(c, 0, 0, rowpre + 0, prow, w0.dec_x.b1, w0.dec_x.b2, xy_s, ldxy, false);
    if (recon)
        for (int idx = tid; idx < np * 2 * Tp; idx += NTHREADS) {
            const int mm = idx / (2 * Tp), e = idx - mm * 2 * Tp;
            recon[(pair0 + mm) * 2 * Tp + e] = xy_s[mm * ldxy + e];
        }
    for (int idx = tid; idx < BM * 2 * Tp; idx += NTHREADS) {
        const int mm = idx / (2 * Tp), e = idx - mm * 2 * Tp;
        const int r = mm < np ? row0 + mm / K : -1;
        xy_s[mm * ldxy + e] = r >= 0 ? traj_norm[(size_t)r * 2 * Tp + e] - xy_s[mm * ldxy + e] : 0.f;
    }
    __syncthreads();
    // ---- block 1 Conv1d + GRU -> state1 (A1 k 32..127)
    gru_tc(c, 34, w1, xy_s, ldxy, Tp);
    // ---- decoder_y of both blocks
    const int by0 = 34 + 6 * Tp;
    mlp_tc(c, 1, by0, rowpre + 512, prow, w0.dec_y.b1, w0.dec_y.b2, xy_s, ldxy, false);
    mlp_tc(c, 2, by0 + 34, rowpre + 1024, prow, w1.dec_y.b1, w1.dec_y.b2, xy_s, ldxy, true);
    for (int idx = tid; idx < np * 2 * Tf; idx += NTHREADS) {
        const int mm = idx / (2 * Tf), e = idx - mm * 2 * Tf, cc = e & 1;
        const int r = row0 + mm / K;
        float v = xy_s[mm * ldxy + e] + traj_norm[((size_t)r * Tp + Tp - 1) * 2 + cc];
        if (origin) v += origin[row_scene[r] * 2 + cc];
        pred[(pair0 + mm) * 2 * Tf + e] = v;
    }
    if (recon) {
        __syncthreads();
        mlp_tc(c, 3, by0 + 68, rowpre + 1536, prow, w1.dec_x.b1, w1.dec_x.b2, xy_s, ldxy, false);
        for (int idx = tid; idx < np * 2 * Tp; idx += NTHREADS) {
            const int mm = idx / (2 * Tp), e = idx - mm * 2 * Tp;
            recon[(pair0 + mm) * 2 * Tp + e] += xy_s[mm * ldxy + e];
        }
    }
    __syncthreads();
    if (tid == 0 && *failed) *status = 1;
    if (warp == 1) tc::tmem_dealloc(c.tmem, 512);
}

size_t dec_pair_tc_smem_bytes(int Tp, int Tf) {
    const int Tm = Tp > Tf ? Tp : Tf;
    return (size_t)SM_XY + ((size_t)(BM * (2 * Tm + 1) * 4 + 15) / 16) * 16 + 16 * 8 + 16;
}

}  // namespace sttode
