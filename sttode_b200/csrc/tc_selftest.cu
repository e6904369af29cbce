// Stand-alone check of the tcgen05 building blocks in tc.cuh: C[128 x N] = A[128 x K] . B[N x K]^T.
// Used by tests/test_gpu_tc.py to pin the operand layout / descriptor conventions and the 3xFP16 accuracy
// before the decoder kernels rely on them.
#include "kernels.cuh"
#include "tc.cuh"

namespace sttode {

__global__ void __launch_bounds__(128, 1)
k_umma_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int N, int K, int passes,
                int swap_lbo_sbo, int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* a_hi = sm;
    unsigned char* a_lo = a_hi + (size_t)128 * K * 2;
    unsigned char* b_hi = a_lo + (size_t)128 * K * 2;
    unsigned char* b_lo = b_hi + (size_t)N * K * 2;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + (size_t)N * K * 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int kc = 0; kc < K / 8; ++kc) {
        float va[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) va[i] = A[(size_t)tid * K + kc * 8 + i];
        tc::store_split8(a_hi, a_lo, 128, tid, kc, va);
    }
    for (int r = tid; r < N; r += 128) {
        for (int kc = 0; kc < K / 8; ++kc) {
            float vb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) vb[i] = B[(size_t)r * K + kc * 8 + i];
            tc::store_split8(b_hi, b_lo, N, r, kc, vb);
        }
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const bool ts = swap_lbo_sbo == 2;              // A operand from tensor memory: hi image in columns 256.., lo image in columns 320..
    if (ts) {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int kc = 0; kc < K / 8; ++kc) {
            float va[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) va[i] = A[(size_t)tid * K + kc * 8 + i];
            uint32_t hw[4], lw[4];
            tc::split8_packed(va, hw, lw);
            tc::tmem_st4(tmem + lane_base + 256 + kc * 4, hw);
            tc::tmem_st4(tmem + lane_base + 320 + kc * 4, lw);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(N);
        const uint32_t a_k = 128 * 16, b_k = (uint32_t)N * 16;          // byte stride between 8-wide k chunks
        bool acc = false;
        for (int k = 0; k < K; k += 16) {
            const uint32_t ao = (uint32_t)(k >> 3) * a_k, bo = (uint32_t)(k >> 3) * b_k;
            auto D = [&](unsigned char* base, uint32_t off, uint32_t kstride) {
                return swap_lbo_sbo ? tc::make_desc(tc::smem_u32(base) + off, 128, kstride) : tc::make_desc(tc::smem_u32(base) + off, kstride, 128);
            };
            if (ts) {
                const uint32_t th = tmem + 256 + (k >> 1), tl = tmem + 320 + (k >> 1);
                if (passes == 3) {
                    tc::mma_f16_ts(tmem, tl, tc::make_desc(tc::smem_u32(b_hi) + bo, b_k, 128), idesc, acc ? 1u : 0u);
                    tc::mma_f16_ts(tmem, th, tc::make_desc(tc::smem_u32(b_lo) + bo, b_k, 128), idesc, 1u);
                    acc = true;
                }
                tc::mma_f16_ts(tmem, th, tc::make_desc(tc::smem_u32(b_hi) + bo, b_k, 128), idesc, acc ? 1u : 0u);
                acc = true;
                continue;
            }
            if (passes == 3) {
                tc::mma_f16(tmem, D(a_lo, ao, a_k), D(b_hi, bo, b_k), idesc, acc);
                tc::mma_f16(tmem, D(a_hi, ao, a_k), D(b_lo, bo, b_k), idesc, true);
                acc = true;
            }
            tc::mma_f16(tmem, D(a_hi, ao, a_k), D(b_hi, bo, b_k), idesc, acc);
            acc = true;
        }
        tc::mma_commit(bar);
    }
    const bool ok = tc::mbar_wait(bar, 0);
    tc::tc_fence_after();
    if (!ok) {
        if (tid == 0) *status = 1;
    } else {
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < N) C[(size_t)tid * N + c0 + i] = v[i];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Issue-rate probe: `reps` back-to-back tcgen05.mma (M = 128, K = 16, kind::f16) of width N on resident zero operands, timed with
// clock64 from the first issue to the mbarrier completion of the trailing commit.  cycles[0] = total, cycles[1] = issue loop only.
// mode bit 0: A operand from tensor memory (the `.ts` form the decoder's MLP layers 2 / 3 use); bit 1: warps 1..3 read / write other tensor-memory
// columns (tcgen05.ld / st) for the whole run, as the decoder's epilogue warps do while the MMAs execute.
__global__ void __launch_bounds__(128, 1) k_umma_bench(int N, int reps, long long* __restrict__ cycles, int mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (128 + 256) * 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    volatile int* stop = reinterpret_cast<volatile int*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 32 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); *stop = 0; }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(N);
        const uint64_t da = tc::op_desc(tc::smem_u32(sm), 128), db = tc::op_desc(tc::smem_u32(sm) + 128 * 32, N);
        const long long t0 = clock64();
        if (mode & 1) for (int i = 0; i < reps; ++i) tc::mma_f16_ts(tmem, tmem + 256, db, idesc, 1u);
        else          for (int i = 0; i < reps; ++i) tc::mma_f16_nc(tmem, da, db, idesc, 1u);
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        cycles[0] = t2 - t0;
        cycles[1] = t1 - t0;
        *stop = 1;
    } else if ((mode & 2) && warp >= 1) {
        const uint32_t col = tmem + ((uint32_t)(warp * 32) << 16) + 320 + (uint32_t)warp * 32;      // own lanes, columns no MMA touches
        float v[16];
        uint32_t w8[8];
        while (!*stop) {
            tc::tmem_ld16(col, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) w8[i] = __float_as_uint(v[i] + v[8 + i]);
            tc::tmem_st8(col, w8);
            tc::tmem_st_wait();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Pattern probe: `reps` GRU-like slabs of 6 MMAs (3 x N = 192 into D0, 3 x N = 96 into D1) with distinct hi/lo operand images, B images of
// `brows` rows (the decoder uses 288).  mode 0: rz rz rz n n n (the decoder's order); 1: interleaved rz n rz n rz n; 2: only the three N = 192;
// 3: only the three N = 96; 4: one N = 256 + ... (unused).  cycles[0] = total, cycles[1] = issue loop.
__global__ void __launch_bounds__(640, 1) k_umma_pattern(int mode, int brows, int reps, long long* __restrict__ cycles, int poll_mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int a_bytes = 2 * 128 * 32, b_bytes = 2 * 320 * 32;        // hi + lo images, K = 16
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + a_bytes + b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 2, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t i192 = tc::make_idesc_f16(192), i96 = tc::make_idesc_f16(96);
        const uint32_t a0 = tc::smem_u32(sm), b0 = a0 + a_bytes;
        const uint64_t ah = tc::op_desc(a0, 128), al = tc::op_desc(a0 + 128 * 32, 128);
        const uint64_t bh = tc::op_desc(b0, brows), bl = tc::op_desc(b0 + brows * 32, brows);
        const uint64_t nh = tc::op_desc(b0 + 192 * 16, brows), nl = tc::op_desc(b0 + brows * 32 + 192 * 16, brows);
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            if (mode == 0) {
                tc::mma_f16_nc(tmem, al, bh, i192, 1u); tc::mma_f16_nc(tmem, ah, bl, i192, 1u); tc::mma_f16_nc(tmem, ah, bh, i192, 1u);
                tc::mma_f16_nc(tmem + 192, al, nh, i96, 1u); tc::mma_f16_nc(tmem + 192, ah, nl, i96, 1u); tc::mma_f16_nc(tmem + 192, ah, nh, i96, 1u);
            } else if (mode == 1) {
                tc::mma_f16_nc(tmem, al, bh, i192, 1u); tc::mma_f16_nc(tmem + 192, al, nh, i96, 1u);
                tc::mma_f16_nc(tmem, ah, bl, i192, 1u); tc::mma_f16_nc(tmem + 192, ah, nl, i96, 1u);
                tc::mma_f16_nc(tmem, ah, bh, i192, 1u); tc::mma_f16_nc(tmem + 192, ah, nh, i96, 1u);
            } else if (mode == 2) {
                tc::mma_f16_nc(tmem, al, bh, i192, 1u); tc::mma_f16_nc(tmem, ah, bl, i192, 1u); tc::mma_f16_nc(tmem, ah, bh, i192, 1u);
            } else if (mode == 3) {
                tc::mma_f16_nc(tmem + 192, al, nh, i96, 1u); tc::mma_f16_nc(tmem + 192, ah, nl, i96, 1u); tc::mma_f16_nc(tmem + 192, ah, nh, i96, 1u);
            } else {    // same A / B / D every time at N = 192 (the plain issue-rate probe's pattern)
                tc::mma_f16_nc(tmem, ah, bh, i192, 1u); tc::mma_f16_nc(tmem, ah, bh, i192, 1u); tc::mma_f16_nc(tmem, ah, bh, i192, 1u);
            }
        }
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        cycles[0] = t2 - t0;
        cycles[1] = t1 - t0;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar + 2)) : "memory");     // release the pollers
    } else if (warp >= 4) {
        // bystander warps waiting on an mbarrier while the MMAs run, the way the epilogue warps wait for the accumulators:
        // poll_mode 1 = every lane polls (mbar_wait), 2 = one lane per warp polls, 3 = one lane polls with a nanosleep back-off
        if (poll_mode == 1) tc::mbar_wait(bar + 2, 0);
        else if (poll_mode == 2) { if ((tid & 31) == 0) tc::mbar_wait(bar + 2, 0); __syncwarp(); }
        else if (poll_mode == 3) { if ((tid & 31) == 0) while (!tc::mbar_try_wait(bar + 2, 0)) __nanosleep(100); __syncwarp(); }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int launch_umma_pattern(int mode, int brows, int reps, long long* cycles, cudaStream_t st, int poll_mode) {
    const size_t smem = 2 * 128 * 32 + 2 * 320 * 32 + 64;
    k_umma_pattern<<<1, poll_mode ? 640 : 128, smem, st>>>(mode, brows, reps, cycles, poll_mode);
    return (int)cudaGetLastError();
}

int launch_umma_bench(int N, int reps, long long* cycles, int mode, cudaStream_t st) {
    const size_t smem = (128 + 256) * 32 + 64;
    k_umma_bench<<<1, 128, smem, st>>>(N, reps, cycles, mode);
    return (int)cudaGetLastError();
}

int launch_umma_selftest(const float* A, const float* B, float* C, int N, int K, int passes, int swap, int* status, cudaStream_t st) {
    const size_t smem = (size_t)(128 + N) * K * 2 * 2 + 64;
    cudaError_t e = cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_umma_selftest<<<1, 128, smem, st>>>(A, B, C, N, K, passes, swap, status);
    return (int)cudaGetLastError();
}

}  // namespace sttode
