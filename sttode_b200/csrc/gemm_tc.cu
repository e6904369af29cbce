// General fp32-in / fp32-out GEMM on the tcgen05 tensor cores for the training backward (sm_100a):
//     C[M,N] = epi( sum_k A(m,k) B(k,n) ),   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]   (any strides: NT, NN and TN products)
// with k_gemm's epilogue (bias, ReLU, ReLU-mask, residual, accumulate) and its split-K form (slice blockIdx.z -> C + z*slice_stride).
//
// Precision: gradients span many orders of magnitude, so the fp16 hi/lo split of the inference kernels (range 6e-8 .. 65504) is not safe
// here.  Operands are split into THREE bf16 terms, x = b0 + b1 + b2 (8 + 8 + 8 mantissa bits, fp32 exponent range), and the six products
// whose weight is >= 2^-16 of the leading one are accumulated in fp32 in TMEM:
//     x*y ~= b0c0 + (b0c1 + b1c0) + (b0c2 + b1c1 + b2c0)        -> six kind::f16 (bf16) MMAs per K = 16 step, fp32-level accuracy.
// Tile 128 x 128, K chunks of 64 through a 2-stage shared-memory ring.  Roles (288 threads): warps 0..7 load fp32 tiles from global, split them
// and write the canonical no-swizzle K-major operand images (tc.cuh layout), then run the epilogue out of TMEM; warp 8 issues the MMAs.
#include "kernels.cuh"
#include "tc.cuh"
#include <cuda_bf16.h>
#include <cstdio>

namespace sttode {
namespace gtc {
constexpr int TM = 128, TN = 128, TK = 64, THREADS = 288, LOADERS = 256, ISSUER = 8;
constexpr int IMG = 128 * TK * 2;                // one bf16 image of a 128-row x 64-k tile: 16 KB
constexpr int STAGE = 6 * IMG;                   // A0 A1 A2 B0 B1 B2
constexpr int NST = 2;
constexpr int SM_BARS = NST * STAGE;             // 196608
enum { B_FULL = 0, B_EMPTY = 2, B_DONE = 4, B_COUNT = 5 };

__device__ __forceinline__ void wait(uint64_t* bars, int bar, uint32_t parity) {
    if (!tc::mbar_wait(bars + bar, parity)) {
        printf("k_gemm_tc: mbarrier %d timed out (block %d,%d,%d thread %d)\n", bar, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
        __trap();
    }
}
__device__ __forceinline__ uint32_t bf16x2(float lo_elem, float hi_elem) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<const uint32_t*>(&p);
}
// 8 consecutive-k values of one row -> one 16-byte chunk in each of the three bf16 images (b0 = rn(x), b1 = rn(x - b0), b2 = rn(x - b0 - b1))
__device__ __forceinline__ void store_split8x3(unsigned char* img0, int row, int kchunk, const float (&v)[8]) {
    uint32_t w0[4], w1[4], w2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        w0[i] = bf16x2(a, b);
        const float a1 = a - __uint_as_float(w0[i] << 16), b1 = b - __uint_as_float(w0[i] & 0xffff0000u);
        w1[i] = bf16x2(a1, b1);
        const float a2 = a1 - __uint_as_float(w1[i] << 16), b2 = b1 - __uint_as_float(w1[i] & 0xffff0000u);
        w2[i] = bf16x2(a2, b2);
    }
    const size_t off = (size_t)kchunk * 128 * 16 + (size_t)row * 16;
    *reinterpret_cast<uint4*>(img0 + off) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    *reinterpret_cast<uint4*>(img0 + IMG + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    *reinterpret_cast<uint4*>(img0 + 2 * IMG + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
}  // namespace gtc

__global__ void __launch_bounds__(gtc::THREADS, 1)
k_gemm_tc(int M, int N, int K, const float* __restrict__ A, long sam, long sak, const float* __restrict__ B, long sbk, long sbn, float* __restrict__ C,
          long ldc, const float* __restrict__ bias, const float* __restrict__ res, long ldres, const float* __restrict__ mask, long ldmask, int relu,
          int accumulate, int kchunk, long slice_stride) {
    using namespace gtc;
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    const int nchunks = (kend - kbeg + TK - 1) / TK;
    C += (size_t)blockIdx.z * slice_stride;
    if (tid == 0) {
        tc::mbar_init(bars + B_FULL, LOADERS / 32); tc::mbar_init(bars + B_FULL + 1, LOADERS / 32);
        tc::mbar_init(bars + B_EMPTY, 1); tc::mbar_init(bars + B_EMPTY + 1, 1);
        tc::mbar_init(bars + B_DONE, 1);
        tc::fence_mbar_init();
    }
    if (warp == ISSUER) tc::tmem_alloc(tmem_slot, 128);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == ISSUER) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16(TN);
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                const int s = c & 1;
                wait(bars, B_FULL + s, (uint32_t)((c >> 1) & 1));
                tc::tc_fence_after();
                const uint32_t base = tc::smem_u32(sm + s * STAGE);
                uint64_t a[3], b[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) { a[i] = tc::op_desc(base + i * IMG, 128); b[i] = tc::op_desc(base + (3 + i) * IMG, 128); }
#pragma unroll
                for (int kk = 0; kk < TK / 16; ++kk) {
                    const uint32_t acc0 = (c > 0 || kk > 0) ? 1u : 0u;
                    const uint64_t o = (uint64_t)kk * 256;          // one K = 16 step = two k chunks of a 128-row image = 4096 B = 256 x 16 B
                    tc::mma_f16_nc(tmem, a[2] + o, b[0] + o, idesc, acc0);      // smallest terms first
                    tc::mma_f16_nc(tmem, a[1] + o, b[1] + o, idesc, 1u);
                    tc::mma_f16_nc(tmem, a[0] + o, b[2] + o, idesc, 1u);
                    tc::mma_f16_nc(tmem, a[1] + o, b[0] + o, idesc, 1u);
                    tc::mma_f16_nc(tmem, a[0] + o, b[1] + o, idesc, 1u);
                    tc::mma_f16_nc(tmem, a[0] + o, b[0] + o, idesc, 1u);
                }
                tc::mma_commit(bars + B_EMPTY + s);
            }
            tc::mma_commit(bars + B_DONE);
        }
        __syncwarp();
    } else {
        // ---- loaders: item = (operand, row r, k octet q); consecutive threads walk the operand's contiguous axis
        const bool a_kfast = sak == 1, b_kfast = sbk == 1;
        const bool a_vec = a_kfast && (sam & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
        const bool b_vec = b_kfast && (sbn & 3) == 0 && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const int s = c & 1, k0 = kbeg + c * TK;
            if (c >= NST) wait(bars, B_EMPTY + s, (uint32_t)(((c >> 1) - 1) & 1));
            unsigned char* stage = sm + s * STAGE;
            // all 16 loads of this thread's 8 items are issued before the first split: 512 B in flight per thread (a dependent load -> split -> store
            // chain per item left the SM with ~16 KB in flight and the kernel latency-bound at a tenth of the tensor rate)
            float4 x[8][2];
            int rr[8], qq[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int idx = tid + e * LOADERS;
                const bool isb = idx >= 1024;
                const int j = idx & 1023;
                const bool kfast = isb ? b_kfast : a_kfast;
                const int r = kfast ? j >> 3 : j & 127, q = kfast ? j & 7 : j >> 7;
                rr[e] = r; qq[e] = q;
                const int row = (isb ? n0 : m0) + r, k = k0 + q * 8;
                const bool row_ok = row < (isb ? N : M);
                const float* p = isb ? B + (long)row * sbn + (long)k * sbk : A + (long)row * sam + (long)k * sak;
                const long sk = isb ? sbk : sak;
                if (row_ok && k + 7 < kend && (isb ? b_vec : a_vec) && ((k & 3) == 0)) {
                    x[e][0] = __ldg(reinterpret_cast<const float4*>(p));
                    x[e][1] = __ldg(reinterpret_cast<const float4*>(p) + 1);
                } else {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = (row_ok && k + i < kend) ? __ldg(p + (long)i * sk) : 0.f;
                    x[e][0] = make_float4(v[0], v[1], v[2], v[3]);
                    x[e][1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float v[8] = {x[e][0].x, x[e][0].y, x[e][0].z, x[e][0].w, x[e][1].x, x[e][1].y, x[e][1].z, x[e][1].w};
                store_split8x3(stage + ((tid + e * LOADERS) >= 1024 ? 3 * IMG : 0), rr[e], qq[e], v);
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bars + B_FULL + s)) : "memory");
        }
        // ---- epilogue out of TMEM: thread = (row m0 + 32*(warp & 3) + lane, 64 columns starting at 64*(warp >> 2))
        wait(bars, B_DONE, 0);
        tc::tc_fence_after();
        const int m = m0 + (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll 1
        for (int piece = 0; piece < 4; ++piece) {
            const int nb = (warp >> 2) * 64 + piece * 16;
            float v[16];
            tc::tmem_ld16(tmem + lane_base + nb, v);
            if (m >= M) continue;
            if (nchunks <= 0) {                     // empty K range (trailing split-K slice): the partial product is zero, TMEM was never written
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (n0 + nb + 15 < N) && !mask && !res;
            if (vec) {                              // 64 contiguous bytes of row m per thread: four 16-byte stores instead of sixteen scattered 4-byte ones
                float4* cp = reinterpret_cast<float4*>(C + (long)m * ldc + n0 + nb);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float xv = v[q4 * 4 + i];
                        if (bias) xv += __ldg(bias + n0 + nb + q4 * 4 + i);
                        if (relu) xv = fmaxf(xv, 0.f);
                        o[i] = xv;
                    }
                    if (accumulate) { const float4 c4 = cp[q4]; o[0] += c4.x; o[1] += c4.y; o[2] += c4.z; o[3] += c4.w; }
                    cp[q4] = make_float4(o[0], o[1], o[2], o[3]);
                }
                continue;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = n0 + nb + i;
                if (n >= N) continue;
                float xv = v[i];
                if (bias) xv += __ldg(bias + n);
                if (relu) xv = fmaxf(xv, 0.f);
                if (mask && !(mask[(long)m * ldmask + n] > 0.f)) xv = 0.f;
                if (res) xv += res[(long)m * ldres + n];
                if (accumulate) xv += C[(long)m * ldc + n];
                C[(long)m * ldc + n] = xv;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ISSUER) tc::tmem_dealloc(tmem, 128);
}
size_t gemm_tc_smem_bytes() { return (size_t)gtc::SM_BARS + gtc::B_COUNT * 8 + 16; }
int gemm_tc_threads() { return gtc::THREADS; }

}  // namespace sttode
