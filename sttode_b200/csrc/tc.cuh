// tcgen05 / TMEM / mbarrier / bulk-copy primitives (inline PTX, sm_100a) and the fp16 hi/lo split used to run
// fp32-accurate GEMMs on the 5th-generation tensor cores.
//
// Precision scheme ("3xFP16"): x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 mantissa bits together);
//   x*y ~= hi_x*hi_y + lo_x*hi_y + hi_x*lo_y   accumulated in fp32 in TMEM  (dropped lo*lo term ~ 2^-22).
// Three kind::f16 MMAs per k-step at the bf16/fp16 rate = 1/3 of the dense 16-bit peak, 2x the rate of 3xTF32.
//
// Shared-memory operand layout (no swizzle, K-major "interleave" canonical form, cute/arch/mma_sm100_desc.hpp):
//   element (row r, k) of an R-row operand lives at   base + (k/8) * (R*16) + r*16 + (k%8)*2   bytes
//   i.e. [k/8][row][8 halves]: a core matrix (8 rows x 16 B) is 128 contiguous bytes,
//   SBO (next 8-row group) = 128 B, LBO (next 8-wide k chunk) = R*16 B.  One MMA consumes K = 16 = two k chunks.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sttode {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a descriptor or protocol bug must never hang the GPU.  Returns false on timeout (~0.2 s).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 400000000LL) return false;
    }
    return true;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, async proxy), completion counted on an mbarrier.  bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {             // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread t of warp w reads TMEM lane 32*(w%4)+t, columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// four 8-column loads in flight, ONE wait: the GRU gate epilogue reads r, z, W_in x and W_hn h of the same 8 units from four column ranges
__device__ __forceinline__ void tmem_ld8x4(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, float (&a)[8], float (&b)[8], float (&c)[8], float (&d)[8]) {
    uint32_t r[32];
#define STTODE_LD8(T, O)                                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                         \
                 : "=r"(r[O]), "=r"(r[O + 1]), "=r"(r[O + 2]), "=r"(r[O + 3]), "=r"(r[O + 4]), "=r"(r[O + 5]), "=r"(r[O + 6]), "=r"(r[O + 7]) \
                 : "r"(T)                                                                                                          \
                 : "memory")
    STTODE_LD8(t0, 0);
    STTODE_LD8(t1, 8);
    STTODE_LD8(t2, 16);
    STTODE_LD8(t3, 24);
#undef STTODE_LD8
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); c[i] = __uint_as_float(r[16 + i]); d[i] = __uint_as_float(r[24 + i]); }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: thread t of warp w writes lane 32*(w%4)+t, columns [col, col+N)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
                 "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors + MMA -----------------------------------------------------------------------------------
// K-major, no swizzle.  lbo/sbo in bytes (multiples of 16).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
// kind::f16, A = B = fp16, D = fp32, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[128 x N] (+)= A[128 x K] . B[N x K]^T with the three split products.  a_hi/a_lo/b_hi/b_lo: smem byte addresses of
// operands in the layout above (A has 128 rows, B has n_rows_b rows, of which N starting at row b_row0 are used).
// Issued by ONE thread.  K % 16 == 0.
__device__ __forceinline__ void mma_split3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int n_rows_b,
                                           int b_row0, int N, int K, bool accumulate_first) {
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)n_rows_b * 16;
    const uint32_t boff = (uint32_t)b_row0 * 16;
    bool acc = accumulate_first;
    for (int k = 0; k < K; k += 16) {
        const uint32_t ao = (uint32_t)(k >> 3) * a_lbo, bo = (uint32_t)(k >> 3) * b_lbo + boff;
        const uint64_t dah = make_desc(a_hi + ao, a_lbo, 128), dal = make_desc(a_lo + ao, a_lbo, 128);
        const uint64_t dbh = make_desc(b_hi + bo, b_lbo, 128), dbl = make_desc(b_lo + bo, b_lbo, 128);
        mma_f16(d_tmem, dal, dbh, idesc, acc);      // small terms first
        mma_f16(d_tmem, dah, dbl, idesc, true);
        mma_f16(d_tmem, dah, dbh, idesc, true);
        acc = true;
    }
}

// Descriptor-based variant for hot issue loops: operand descriptors are built once; advancing by one K = 16 step adds
// `a_step` / `b_step` (in 16-byte units: 2 * rows of the operand image) to the 14-bit start-address field.
__device__ __forceinline__ void mma_f16_nc(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void mma_split3_d(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint32_t a_step, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t b_step, uint32_t idesc, int nk, bool accumulate_first) {
    uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll 2
    for (int k = 0; k < nk; ++k) {
        mma_f16_nc(d_tmem, a_lo, b_hi, idesc, acc);      // small terms first
        mma_f16_nc(d_tmem, a_hi, b_lo, idesc, 1u);
        mma_f16_nc(d_tmem, a_hi, b_hi, idesc, 1u);
        acc = 1u;
        a_hi += a_step; a_lo += a_step; b_hi += b_step; b_lo += b_step;
    }
}
// A operand in TENSOR MEMORY (".ts" form): row r of A = TMEM lane r, element k of a K = 16 step = half (k & 1) of 32-bit column a_tmem + k / 2
// (8 columns per step), K-major only.  B from shared memory as above.
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// three split products with A = (a_hi, a_lo) column blocks in TMEM (8 columns each per K = 16 step), advancing `a_step` columns per step
__device__ __forceinline__ void mma_split3_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_step, uint64_t b_hi, uint64_t b_lo, uint32_t b_step,
                                              uint32_t idesc, int nk, bool accumulate_first) {
    uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll 2
    for (int k = 0; k < nk; ++k) {
        mma_f16_ts(d_tmem, a_lo, b_hi, idesc, acc);      // small terms first
        mma_f16_ts(d_tmem, a_hi, b_lo, idesc, 1u);
        mma_f16_ts(d_tmem, a_hi, b_hi, idesc, 1u);
        acc = 1u;
        a_hi += a_step; a_lo += a_step; b_hi += b_step; b_lo += b_step;
    }
}
// descriptor of an R-row operand image starting at shared-window address `addr`
__device__ __forceinline__ uint64_t op_desc(uint32_t addr, int rows) { return make_desc(addr, (uint32_t)rows * 16, 128); }

// ---- fp16 hi/lo split --------------------------------------------------------------------------------------
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// packed conversion of two floats to f16x2 (one F2FP instruction; `lo_elem` lands in the low half)
__device__ __forceinline__ uint32_t cvt_f16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// 8 consecutive-k values of one row -> one 16-byte chunk each of the hi and lo operand images.
// Same roundings as split_f16 (hi = rn(x), lo = rn(x - hi)), two values per conversion instruction: the scalar F2F
// conversions share the 16-lane XU pipe with MUFU and were the epilogue's bottleneck.
__device__ __forceinline__ void store_split8(unsigned char* hi_base, unsigned char* lo_base, int rows, int row, int kchunk, const float (&v)[8]) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hw[i] = cvt_f16x2(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
        lw[i] = cvt_f16x2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    }
    const size_t off = (size_t)kchunk * rows * 16 + (size_t)row * 16;
    *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// 8 consecutive-k values of one row -> 4 packed hi words and 4 packed lo words (same roundings as store_split8), for a TMEM-resident A operand
__device__ __forceinline__ void split8_packed(const float (&v)[8], uint32_t (&hw)[4], uint32_t (&lw)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hw[i] = cvt_f16x2(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
        lw[i] = cvt_f16x2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    }
}

}  // namespace tc
}  // namespace sttode
