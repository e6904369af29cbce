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
    if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(N);
        const uint32_t a_k = 128 * 16, b_k = (uint32_t)N * 16;          // byte stride between 8-wide k chunks
        bool acc = false;
        for (int k = 0; k < K; k += 16) {
            const uint32_t ao = (uint32_t)(k >> 3) * a_k, bo = (uint32_t)(k >> 3) * b_k;
            auto D = [&](unsigned char* base, uint32_t off, uint32_t kstride) {
                return swap_lbo_sbo ? tc::make_desc(tc::smem_u32(base) + off, 128, kstride) : tc::make_desc(tc::smem_u32(base) + off, kstride, 128);
            };
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
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// Issue-rate probe: `reps` back-to-back tcgen05.mma (M = 128, K = 16, kind::f16) of width N on resident zero operands, timed with
// clock64 from the first issue to the mbarrier completion of the trailing commit.  cycles[0] = total, cycles[1] = issue loop only.
__global__ void __launch_bounds__(128, 1) k_umma_bench(int N, int reps, long long* __restrict__ cycles) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (128 + 256) * 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 32 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(N);
        const uint64_t da = tc::op_desc(tc::smem_u32(sm), 128), db = tc::op_desc(tc::smem_u32(sm) + 128 * 32, N);
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) tc::mma_f16_nc(tmem, da, db, idesc, 1u);
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        cycles[0] = t2 - t0;
        cycles[1] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

int launch_umma_bench(int N, int reps, long long* cycles, cudaStream_t st) {
    const size_t smem = (128 + 256) * 32 + 64;
    k_umma_bench<<<1, 128, smem, st>>>(N, reps, cycles);
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
