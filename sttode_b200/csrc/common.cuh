// Shared device primitives for the STTODE sm_100a kernels.
//
// Layout conventions used by every fused row-tile kernel in this directory:
//   * a CTA owns a tile of BM = 128 rows (agent rows, or (agent,sample) pairs) and runs 256 threads;
//   * activations live in shared memory row-major with a +4 float pad (ld = K + 4) so that float4
//     reads along k by lanes that differ in row hit distinct banks;
//   * weights stay in the reference's nn.Linear layout W[N][K] (K contiguous) and are streamed from
//     L2 into shared memory in K-chunks with cp.async (16 B), double buffered;
//   * thread (tx = tid & 15, ty = tid >> 4) accumulates rows m = i*16 + ty (i < 8) and columns
//     n = j*16 + tx (j < TNC) in registers: an 8 x TNC register tile, 10.7 FFMA per LDS.128.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sttode {

constexpr int BM = 128;          // rows per CTA tile
constexpr int NTHREADS = 256;    // threads per CTA
constexpr int PAD = 4;           // float pad on every smem row

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Copy a weight chunk W[n0 .. n0+nrows)[k0 .. k0+kc) (row stride ldg floats, 16 B aligned) to
// Ws[nrows][lds].  kc % 4 == 0.  All NTHREADS threads participate; caller commits/waits.
__device__ __forceinline__ void load_w_chunk_async(float* Ws, int lds, const float* __restrict__ Wg, int ldg, int nrows, int kc) {
    const int vec_per_row = kc >> 2;
    const int total = nrows * vec_per_row;
    for (int v = threadIdx.x; v < total; v += NTHREADS) {
        int r = v / vec_per_row, c = (v - r * vec_per_row) << 2;
        cp_async16(Ws + r * lds + c, Wg + (size_t)r * ldg + c);
    }
}

// acc[i][j] += sum_k A[i*16+ty][k] * W[wrow[j]][k], k in [0, kc).  As/Ws are smem, k contiguous.
template <int TNC>
__device__ __forceinline__ void mma_tile(float (&acc)[8][TNC], const float* As, int lda, const float* Ws, int ldw,
                                         const int (&wrow)[TNC], int kc, int ty) {
#pragma unroll 2
    for (int k = 0; k < kc; k += 4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(As + (i * 16 + ty) * lda + k);
#pragma unroll
        for (int j = 0; j < TNC; ++j) {
            float4 w = *reinterpret_cast<const float4*>(Ws + wrow[j] * ldw + k);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][j] = fmaf(a[i].x, w.x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, w.y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, w.z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, w.w, acc[i][j]);
            }
        }
    }
}

// Full GEMM step: acc[8][TNC] += As[128][K] . Wg[16*TNC rows][K]^T with the weight streamed in
// KC-chunks through the double buffer Wbuf (2 * (16*TNC) * (KC+PAD) floats).
// Row j*16+tx of the weight block is output column j*16+tx.  Ends with a __syncthreads() so the
// caller may overwrite As / reuse Wbuf immediately.
template <int TNC, int KC>
__device__ __forceinline__ void gemm_step(float (&acc)[8][TNC], const float* As, int lda, const float* __restrict__ Wg,
                                          int ldg, int K, float* Wbuf) {
    constexpr int NROWS = 16 * TNC;
    constexpr int LDW = KC + PAD;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int wrow[TNC];
#pragma unroll
    for (int j = 0; j < TNC; ++j) wrow[j] = j * 16 + tx;
    const int nchunk = (K + KC - 1) / KC;
    load_w_chunk_async(Wbuf, LDW, Wg, ldg, NROWS, min(KC, K));
    cp_async_commit();
    for (int c = 0; c < nchunk; ++c) {
        const int k0 = c * KC;
        if (c + 1 < nchunk) {
            load_w_chunk_async(Wbuf + ((c + 1) & 1) * NROWS * LDW, LDW, Wg + k0 + KC, ldg, NROWS, min(KC, K - k0 - KC));
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        mma_tile<TNC>(acc, As + k0, lda, Wbuf + (c & 1) * NROWS * LDW, LDW, wrow, min(KC, K - k0), ty);
        __syncthreads();
    }
}

template <int TNC>
__device__ __forceinline__ void zero_acc(float (&acc)[8][TNC]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TNC; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LayerNorm over 64 features of each of `rows` smem rows (ld floats apart), in place, one warp per
// row, two features per lane, biased variance, eps inside the sqrt (torch native_layer_norm).
__device__ __forceinline__ void layernorm64_rows(float* S, int ld, int rows, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float g0 = __ldg(gamma + lane), g1 = __ldg(gamma + lane + 32);
    const float b0 = __ldg(beta + lane), b1 = __ldg(beta + lane + 32);
    for (int r = warp; r < rows; r += NTHREADS / 32) {
        float v0 = S[r * ld + lane], v1 = S[r * ld + lane + 32];
        float mean = warp_sum(v0 + v1) * (1.f / 64.f);
        float d0 = v0 - mean, d1 = v1 - mean;
        float var = warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f);
        float rstd = 1.f / sqrtf(var + eps);
        S[r * ld + lane] = d0 * rstd * g0 + b0;
        S[r * ld + lane + 32] = d1 * rstd * g1 + b1;
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// torch.clamp semantics: NaN propagates (fminf/fmaxf would swallow it).
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---- Philox4x32-10 + Box-Muller: the latent z ~ N(0, I) drawn inside the decoder kernel ----------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
    float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
    float u2 = (float)b * 2.3283064365386963e-10f;            // [0, 1)
    float r = sqrtf(-2.f * logf(u1));
    float s, c;
    sincospif(2.f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
// z[pair][4*q .. 4*q+3] for q in [0, 8): counter = (pair, q), key = seed.
__device__ __forceinline__ float4 latent_normal4(unsigned long long seed, unsigned long long pair, unsigned q) {
    uint4 r = philox4x32_10(make_uint4((unsigned)pair, (unsigned)(pair >> 32), q, 0x5774u),
                            make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
    return make_float4(n0.x, n0.y, n1.x, n1.y);
}

}  // namespace sttode
