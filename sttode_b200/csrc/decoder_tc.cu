// Tensor-core (tcgen05 / TMEM) version of the per-(agent,sample) decoder tile, sm_100a.
//
// Same algebra as k_dec_pair (decoder.cu): decoder_x0 -> Conv1d+GRU1 -> decoder_y0 + decoder_y1 -> pred (model/STTODE.py:137-347), for tiles of 128
// consecutive (agent, sample) pairs, one persistent CTA per SM.  Every dense contraction runs as tcgen05.mma kind::f16 with the 3xFP16 split of
// tc.cuh (fp32-accurate), accumulators in TMEM:
//   MLP layer 1 : D1[128x128]  = [z|state](K1) . W1g^T          per GROUP of 128 hidden units (two 128-column D1 buffers; group g+1 in K slices behind group g's chunks)
//   MLP layer 2 : D2[128x256] += relu(D1 + rowpre)[:, 32c..] . W2c^T     per 32-unit chunk c; the epilogue converts the chunk to fp16 hi/lo IN PLACE in tensor memory
//                                                                (tcgen05.st) and the MMA reads it as a TMEM A operand: the hidden layer never touches shared memory
//   MLP layer 3 : D3[128xn3p] += relu(D2 + b1)[:, 32p..] . W3p^T          (same in-place conversion of D2; D3 aliases the second D1 buffer)
//   GRU step    : [W_in x | r | z | W_hn h] per N-half (units 0..63 / 64..95) = conv(x)_t . W_ih^T (+) h . W_hh^T ; gates in the register epilogue;
//                 pipelined across steps (half B and the conv image double-buffered)
// Phase order of a tile: x0-MLP, y0-MLP, Conv+GRU, y1-MLP, output (y0 needs only the latent; see RingCfg below for why it runs before the GRU).
// Weights are pre-split into fp16 hi/lo images in the canonical no-swizzle K-major layout by k_pack_all (once per weight version) and streamed
// L2 -> shared memory with 1-D bulk copies (cp.async.bulk, UBLKCP) through a ring of 2 or 3 slots guarded by mbarriers; tcgen05.commit releases ring
// slots and accumulators.  Roles: 16 epilogue warps (TMEM -> registers -> bias / ReLU / gates -> operands), one issuer warp (one elected lane issues
// every MMA), one producer warp (weight stream, barrier set-up, tile claiming).
#include "kernels.cuh"
#include "tc.cuh"
#include <cstdio>
#include <type_traits>

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
        // Layer 1 runs per GROUP of 128 hidden units (four 32-unit chunks: one N = 128 MMA per split product instead of four N = 32 ones at the
        // small-N issue floor).  Stream of one MLP: lead = two ring pairs: W1 rows of group 0, first K half [hi, lo], then [second K half | first K slice
        // of group 1] [hi, lo] (a ring pair costs one L2 round trip whatever its size: three small lead pairs were three serial round trips per MLP);
        // then per chunk c a (hi, lo) pair of [W2 chunk c image (256 rows x 32 k) | K slice (c+1) & 3 of W1 group ((c+1) >> 2) + 1]:
        // a group's layer 1 is complete one chunk before its first chunk is consumed.  K1 = 128: four slices of 32 k; K1 = 32: two of 16 k.
        L.lead_blk[m] = 128 * k1[m];            // 128 rows x K1/2 halves
        L.slice_blk[m] = 128 * 2 * (k1[m] == 32 ? 16 : k1[m] / 4);
        L.mlp_blk[m] = 16384 + L.slice_blk[m];
        L.mlp_off[m] = off;
        off += 4 * L.lead_blk[m] + 2 * L.slice_blk[m] + TC_NCH * L.mlp_blk[m];
        L.l3_blk[m] = L.n3p[m] * 512;           // 4 pieces of [hi n3p x 32 | lo n3p x 32]
        L.l3_off[m] = off;
        off += 2 * L.l3_blk[m];
    }
    for (int g = 0; g < 2; ++g) { L.gru_off[g] = off; off += 8 * TC_GRU_BLK; }
    L.rowp_off = off;                           // row kernel: 88 slabs of [256 rows x 16 k] hi 8 KB | lo 8 KB
    off += 88 * 16384;
    L.total = off;
    return L;
}

// ---- all weight images of one pack in ONE launch ------------------------------------------------------------------------------
// Job j (blockIdx.y) is one operand image; the job list is a closed-form function of the layout, so nothing but the weight pointers
// and the layout travels to the device (round 1 issued 264 single-image launches + 8 memsets per pack).
struct PackJob { const float* W; int ldw, row0, col0, nvalid, nrows, K; uint32_t hi, lo; int perm; };      // perm: 0 rows as stored, 1 / 2 GRU half A / B
constexpr int PACK_JOBS_MLP = 2 + 12 + 16 + 8, PACK_JOBS = 4 * PACK_JOBS_MLP + 32 + 88 + 1;
// GRU weight rows are stored [r 96 | z 96 | n 96]; half A (units 0..63) and half B (units 64..95) get their own images.  h-slabs (W_hh) are in
// gate order [r | z | n] (perm 1 / 2), x-slabs (W_ih) in [n | r | z] (perm 3 / 4): with the accumulator columns laid out [W_in x | r | z | W_hn h]
// both kinds of slab feed ONE contiguous column range (x: first three quarters, h: last three), i.e. one MMA per split product.
__device__ __forceinline__ int gru_perm_row(int perm, int i) {
    if (perm == 1) return i < 64 ? i : i < 128 ? 96 + (i - 64) : 192 + (i - 128);
    if (perm == 2) return i < 32 ? 64 + i : i < 64 ? 96 + 64 + (i - 32) : 192 + 64 + (i - 64);
    if (perm == 3) return i < 64 ? 192 + i : i < 128 ? (i - 64) : 96 + (i - 128);
    if (perm == 4) return i < 32 ? 192 + 64 + i : i < 64 ? 64 + (i - 32) : 96 + 64 + (i - 64);
    return i;
}
// byte offset (inside one GRU's image area) and size of ring block j of a step (TC_GRU_STEP_BLOCKS per step; the same 8 blocks every step)
__host__ __device__ __forceinline__ void gru_block(int j, uint32_t& off, uint32_t& bytes) {
    // slab-pair index: x0x1 -> 0, h0h1 -> 1, h2h3 -> 2, h4h5 -> 3;   order: A0 B0 | B1 B2 A1 A2 | A3 B3
    const bool isA = j == 0 || (j >= 4 && j <= 6);
    const int pi = j == 0 ? 0 : j == 1 ? 0 : j == 2 ? 1 : j == 3 ? 2 : j == 4 ? 1 : j == 5 ? 2 : 3;
    if (isA) { off = (uint32_t)pi * 2u * TC_GRU_A; bytes = 2u * TC_GRU_A; }
    else     { off = 8u * TC_GRU_A + (uint32_t)pi * 2u * TC_GRU_B; bytes = 2u * TC_GRU_B; }
}

__device__ __forceinline__ PackJob pack_job(int j, const sttode_decompose_weights& w0, const sttode_decompose_weights& w1, const DecPackLayout& L) {
    PackJob J{};
    if (j < 4 * PACK_JOBS_MLP) {
        const int m = j / PACK_JOBS_MLP, i = j - m * PACK_JOBS_MLP;
        const sttode_mlp3_weights& mw = m == 0 ? w0.dec_x : m == 1 ? w0.dec_y : m == 2 ? w1.dec_y : w1.dec_x;
        const int K1 = (int)L.k1[m], n3 = (int)L.n3[m], n3p = (int)L.n3p[m];
        const uint32_t lead = L.mlp_off[m], SB = L.slice_blk[m], chunk0 = lead + 4 * L.lead_blk[m] + 2 * SB;
        const int nsl = K1 == 32 ? 2 : 4, ksl = K1 / nsl;
        if (i < 2) {                    // group 0, K half i: hidden rows 0..127, pair columns 128 + i K1/2 ..  (lead area: [h0 hi | h0 lo | h1 hi | slice hi | h1 lo | slice lo])
            const uint32_t LB = L.lead_blk[m];
            J = PackJob{mw.w0, 256, 0, 128 + i * (K1 / 2), 128, 128, K1 / 2, lead + (i ? 2 * LB : 0u), lead + (i ? 3 * LB + SB : LB)};
        } else if (i < 14) {            // group g = 1..3, K slice s: rides in the lead (g = 1, s = 0) or behind W2 of chunk 4 (g - 1) + s - 1
            const int g = 1 + (i - 2) / 4, sl = (i - 2) & 3, cp = 4 * (g - 1) + sl;
            if (sl >= nsl) J = PackJob{mw.w0, 256, 0, 0, 0, 0, 8, 0, 0};          // no such slice: empty job
            else if (cp == 0) J = PackJob{mw.w0, 256, g * 128, 128 + sl * ksl, 128, 128, ksl, lead + 3 * L.lead_blk[m], lead + 4 * L.lead_blk[m] + SB};
            else {
                const uint32_t ba = chunk0 + (uint32_t)(2 * (cp - 1)) * L.mlp_blk[m] + 16384;
                J = PackJob{mw.w0, 256, g * 128, 128 + sl * ksl, 128, 128, ksl, ba, ba + L.mlp_blk[m]};
            }
        } else if (i < 30) {
            const int c = i - 14;
            const uint32_t ba = chunk0 + (uint32_t)(2 * c) * L.mlp_blk[m];
            J = PackJob{mw.w1, 512, 0, c * 32, 256, 256, 32, ba, ba + L.mlp_blk[m]};
        } else {
            const int p = i - 30;
            const uint32_t pc = L.l3_off[m] + (uint32_t)(p >> 2) * L.l3_blk[m] + (uint32_t)(p & 3) * n3p * 128;
            J = PackJob{mw.w2, 256, 0, p * 32, n3, n3p, 32, pc, pc + (uint32_t)n3p * 64};
        }
        return J;
    }
    j -= 4 * PACK_JOBS_MLP;
    if (j < 32) {
        const int g = j >> 4, half = (j >> 3) & 1, kb = j & 7;
        const sttode_decompose_weights& w = g ? w1 : w0;
        // slab kb lives in slab-pair block kb >> 1 of its half, slot kb & 1
        const uint32_t blk = L.gru_off[g] + (half ? 8u * TC_GRU_A + (uint32_t)(kb >> 1) * 2u * TC_GRU_B + (uint32_t)(kb & 1) * TC_GRU_B
                                                  : (uint32_t)(kb >> 1) * 2u * TC_GRU_A + (uint32_t)(kb & 1) * TC_GRU_A);
        const int nrows = half ? 96 : 192;
        if (kb < 2) J = PackJob{w.gru_w_ih, 32, 0, kb * 16, nrows, nrows, 16, blk, blk + (uint32_t)nrows * 32, 3 + half};
        else        J = PackJob{w.gru_w_hh, 96, 0, (kb - 2) * 16, nrows, nrows, 16, blk, blk + (uint32_t)nrows * 32, 1 + half};
        return J;
    }
    j -= 32;
    if (j < 88) {   // row kernel: tile t = (product p = t/2, columns (t&1)*256..): K slabs over [pf 0..127 | state 160..255] (p < 2) or pf only
        int t, ks;
        if (j < 56) { t = j / 14; ks = j - t * 14; } else { t = 4 + (j - 56) / 8; ks = (j - 56) & 7; }
        const int p = t >> 1, nb = t & 1;
        const sttode_mlp3_weights& mw = p == 0 ? w0.dec_x : p == 1 ? w0.dec_y : p == 2 ? w1.dec_y : w1.dec_x;
        const uint32_t blk = L.rowp_off + (uint32_t)j * 16384;
        J = PackJob{mw.w0, 256, nb * 256, ks < 8 ? ks * 16 : 160 + (ks - 8) * 16, 256, 256, 16, blk, blk + 8192};
        return J;
    }
    J.K = 0;            // last job: clear the 256-byte status tail
    return J;
}

__global__ void k_pack_all(sttode_decompose_weights w0, sttode_decompose_weights w1, const __grid_constant__ DecPackLayout L, unsigned char* __restrict__ pk) {
    const int j = blockIdx.y;
    const PackJob J = pack_job(j, w0, w1, L);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (J.K == 0) {
        if (idx < 64) reinterpret_cast<uint32_t*>(pk + L.total)[idx] = 0u;
        return;
    }
    if (idx >= (J.K >> 3) * J.nrows) return;
    const int kc = idx / J.nrows, r = idx - kc * J.nrows;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = r < J.nvalid ? __ldg(J.W + (size_t)(J.row0 + gru_perm_row(J.perm, r)) * J.ldw + J.col0 + kc * 8 + i) : 0.f;
    tc::store_split8(pk + J.hi, pk + J.lo, J.nrows, r, kc, v);
}
void launch_pack_all(const sttode_decompose_weights& w0, const sttode_decompose_weights& w1, const DecPackLayout& L, unsigned char* pk, cudaStream_t st) {
    k_pack_all<<<dim3(8, PACK_JOBS), 128, 0, st>>>(w0, w1, L, pk);      // largest image: 4 k-chunks x 256 rows = 1024 threads
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised tile kernel (576 threads):
//   warps 0..15 epilogue : TMEM -> registers -> bias/ReLU/gates -> fp16 hi/lo operand images in shared memory, outputs
//   warp  16    issuer   : tcgen05.mma + tcgen05.commit (warp-uniform bookkeeping, one elected lane issues)
//   warp  17    producer : streams the packed weight blocks L2 -> shared ring with cp.async.bulk
// The roles talk through mbarriers only.  All phase parities are closed-form functions of the loop indices.
// Epilogue warp w reads TMEM lanes 32*(w & 3)..; sub = w >> 2 (0..3) selects its share of the columns: four epilogue warps per
// scheduler keep the ALU / MUFU pipes fed while one of them waits on TMEM or shared memory.
constexpr int ISSUER_WARP = 16, PRODUCER_WARP = 17;
constexpr int TC_THREADS = 576;
constexpr int EPI_THREADS = 512;
constexpr int SM_A1 = 0;                         // [hi 32 KB | lo 32 KB]   A1 = [z (k 0..31) | state (k 32..127)]
constexpr int SM_H = 65536;                      // 2 x [hi 8 KB | lo 8 KB]: conv image of GRU step t in H[t & 1]; the last step leaves H[T & 1] to the next tile's latent image
constexpr int SM_RING = SM_H + 32768;
constexpr int SM_XY = SM_RING + TC_NST * TC_STAGE;
constexpr int COL_D2 = 0, COL_D1 = 256, COL_D3 = 384;     // TMEM columns, MLP phases (D1: 2 buffers of 128 = one 4-chunk hidden group each; D3 <= 128 columns aliases the second)
// GRU phase accumulators (the MLP accumulators are all read out by then).  A step runs as two N-halves, each [W_in x | r | z | W_hn h]:
//   half A (units 0..63): columns 0..255        half B (units 64..95): columns 256..383 (even steps) / 384..511 (odd steps)
// Half B alternates between two buffers and half A is read out first, so the MMAs of step t+1 run UNDER the gate math of step t: the x-slabs as soon
// as half A's accumulators are in registers, h-slabs 0..3 once units 0..63 of h(t+1) are staged; only h-slabs 4..5 wait for the whole state.
constexpr int COL_GA = 0, COL_GB = 256;
constexpr int B_WFULL = 0, B_WFREE = 4, B_D1FULL = 8, B_HFULL = 12, B_ACC = 16, B_XFULL = 17, B_D3FREE = 18, B_ACC2 = 19,      // B_ACC2: GRU half B complete
              B_GREAD = 20, B_HA = 21, B_HB = 22, B_CFULL = 23, B_L3RDY = 31, B_COUNT = 33;    // GRU: half-A accumulators read + next conv image staged / units 0..63 / units 64..95 of the new state staged
// B_CFULL + (ch & 7): hidden chunk ch converted in place to an A operand in TMEM (8 arrivals); B_L3RDY + s: half s of relu(D2 + b1) converted (16 arrivals)

// Weight ring.  Blocks travel in PAIRS (2 stages of TC_STAGE bytes, one full / one free mbarrier per slot).  Pairs [0, n3) of a tile rotate over THREE
// slots, every later pair alternates over slots 0 and 1: the third slot is the state half of A1 (k chunks 4..15 of the hi and of the lo image, 24 KB each),
// which is idle while the two K1 = 32 MLPs run before the GRU (phase order x0, y0, GRU, y1).  A pair is requested when the previous occupant of its slot
// has been consumed, and a 40 KB pair takes ~2 k cycles from L2 against 0.87 k of MMA work per chunk: with two slots every chunk waited for its weights.
// n3 = 0 is the plain two-slot ring (row kernel; pair kernel when the separate y buffer does not fit).  slot / use index are closed forms of the pair index.
// The issuing thread runs these between its waits and its MMAs, one dependent instruction every ~5 cycles: no divisions (a first version with p / 3 and
// (n3 - slot + 2) / 3 in here cost 4 % of the kernel).
struct RingCfg { int n3, c30, c31; };               // c3s = uses of slot s by the three-slot pairs = (n3 - s + 2) / 3
__host__ __device__ __forceinline__ RingCfg make_ring_cfg(int n3) { return RingCfg{n3, (n3 + 2) / 3, (n3 + 1) / 3}; }
__device__ __forceinline__ void ring_slot(const RingCfg& rc, int p, int& slot, int& use) {
    if (p < rc.n3) { use = (p * 0xAAAB) >> 17; slot = p - 3 * use; }                    // p / 3, p % 3 for p < 32768
    else { const int q = p - rc.n3; slot = q & 1; use = (slot ? rc.c31 : rc.c30) + (q >> 1); }
}
__device__ __forceinline__ int ring_stage_off(int slot, int half) { return slot < 2 ? SM_RING + (2 * slot + half) * TC_STAGE : SM_A1 + 8192 + half * 32768; }

// debug timeline: absolute clock64 stamps into the status tail (CTA 0 only, when status[1] == 2)
// compiled in only with -DSTTODE_TC_TIMELINE (scripts/tc_timing.py); the release build has no trace of it
// STTODE_TC_TIMELINE=1: phase / GRU-step stamps (TLOG); =2: issuer and epilogue stamps of the y0 MLP (TLOG2, TACC2 accumulates a duration)
#ifdef STTODE_TC_TIMELINE
__device__ long long* g_tl = nullptr;
#endif
#if defined(STTODE_TC_TIMELINE) && STTODE_TC_TIMELINE == 1
#define TLOG(cond, slot) do { if (blockIdx.x == 0 && (cond)) { long long* tl_ = *const_cast<long long* volatile*>(&g_tl); if (tl_) tl_[slot] = clock64(); } } while (0)
#else
#define TLOG(cond, slot) do { (void)(cond); } while (0)
#endif
// =3: whole-tile accounting incl. the prologue and the hand-over of the SM to the next CTA (TLOG3)
#if defined(STTODE_TC_TIMELINE) && STTODE_TC_TIMELINE == 3
#define TLOG3(cond, slot) do { if (blockIdx.x == 0 && (cond)) { long long* tl_ = *const_cast<long long* volatile*>(&g_tl); if (tl_) tl_[slot] = clock64(); } } while (0)
#else
#define TLOG3(cond, slot) do { (void)(cond); } while (0)
#endif
#if defined(STTODE_TC_TIMELINE) && STTODE_TC_TIMELINE == 2
#define TLOG2(cond, slot) do { if (blockIdx.x == 0 && (cond)) { long long* tl_ = *const_cast<long long* volatile*>(&g_tl); if (tl_) tl_[slot] = clock64(); } } while (0)
#define TACC2(cond, slot, t0) do { if (blockIdx.x == 0 && (cond)) { long long* tl_ = *const_cast<long long* volatile*>(&g_tl); if (tl_) tl_[slot] += clock64() - (t0); } } while (0)
#define TNOW() clock64()
#else
#define TLOG2(cond, slot) do { (void)(cond); } while (0)
#define TACC2(cond, slot, t0) do { (void)(cond); (void)(t0); } while (0)
#define TNOW() 0LL
#endif

__device__ __forceinline__ void tc_wait(uint64_t* bars, int bar, uint32_t parity) {
    if (!tc::mbar_wait(bars + bar, parity)) {      // a protocol bug must fail loudly, never hang the GPU
        printf("k_dec_pair_tc: mbarrier %d timed out (block %d thread %d)\n", bar, blockIdx.x, threadIdx.x);
        __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// epilogue warp: publish generic-proxy smem writes + finished TMEM reads, one arrival per warp
__device__ __forceinline__ void epi_signal(uint64_t* bars, int bar) {
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + bar);
}

// global block sequence: [mlp x0: 38 blocks] [mlp y0: 38 when y_first] [gru1: 8*Tp] [mlp y0: 38 otherwise] [mlp y1: 38] [mlp x1: 38 when recon]
// one MLP = 4 lead blocks, 32 chunk blocks, layer-3 pair
__device__ __forceinline__ void block_src(const DecPackLayout& lay, int Tp, int y_first, int i, uint32_t& off, uint32_t& bytes) {
    int m;
    const int gru0 = y_first ? 2 * TC_MLP_BLOCKS : TC_MLP_BLOCKS, ngru = TC_GRU_STEP_BLOCKS * Tp;
    if (i < gru0) { m = i < TC_MLP_BLOCKS ? 0 : 1; if (m) i -= TC_MLP_BLOCKS; }
    else if (i < gru0 + ngru) {                     // the same 8 blocks every step
        gru_block((i - gru0) & 7, off, bytes);
        off += lay.gru_off[1];
        return;
    } else {
        i -= gru0 + ngru;
        m = y_first ? 2 : 1;                        // MLPs after the GRU: (y0,) y1, x1
        while (i >= TC_MLP_BLOCKS) { i -= TC_MLP_BLOCKS; ++m; }
    }
    if (i < 2) { off = lay.mlp_off[m] + (uint32_t)i * lay.lead_blk[m]; bytes = lay.lead_blk[m]; }
    else if (i < TC_LEAD) { bytes = lay.lead_blk[m] + lay.slice_blk[m]; off = lay.mlp_off[m] + 2 * lay.lead_blk[m] + (uint32_t)(i - 2) * bytes; }
    else if (i < TC_LEAD + TC_NCH) {
        // chunk block (hi or lo) of chunk ch: W2 image (16 KB) + the layer-1 K slice that rides behind it, when there is one (slice (ch+1) & 3 of group ((ch+1) >> 2) + 1:
        // groups 1..3, K1 = 32 has two slices per group): blocks without a slice copy 16 KB only (they used to drag 4..8 dead KB each: 8 % of the tile's L2 traffic)
        const int ch = (i - TC_LEAD) >> 1, cp = ch + 1;
        const bool has_slice = (cp >> 2) + 1 < 4 && (cp & 3) < (lay.k1[m] == 32 ? 2 : 4);
        off = lay.mlp_off[m] + 4 * lay.lead_blk[m] + 2 * lay.slice_blk[m] + (uint32_t)(i - TC_LEAD) * lay.mlp_blk[m];
        bytes = 16384 + (has_slice ? lay.slice_blk[m] : 0);
    }
    else { off = lay.l3_off[m] + (uint32_t)(i - TC_LEAD - TC_NCH) * lay.l3_blk[m]; bytes = lay.l3_blk[m]; }
}

// ================================================================ producer warp ============================
// The ring is managed in PAIRS of stages (every phase starts at an even block index and uses blocks two at a time):
// one full / one free mbarrier per pair halves the number of barrier hand-offs on the issuer's critical path.
template <typename SrcFn>
__device__ void producer_loop(unsigned char* sm, uint64_t* bars, const unsigned char* __restrict__ packed, int total_blocks, int dbg_nocopy, const RingCfg rc,
                              SrcFn src_of) {
#pragma unroll 1
    for (int j = 0; j < (total_blocks >> 1); ++j) {
        int slot, use;
        ring_slot(rc, j, slot, use);
        uint32_t off0, bytes0, off1, bytes1;
        src_of(2 * j, off0, bytes0);                // address arithmetic first: nothing but the copies follows the wait
        src_of(2 * j + 1, off1, bytes1);
        if (use >= 1) tc_wait(bars, B_WFREE + slot, (uint32_t)((use - 1) & 1));      // the slot's previous occupant has been consumed
        if (elect_one()) {
            if (dbg_nocopy) mbar_arrive(bars + B_WFULL + slot);       // debug: timing without the weight stream (results invalid)
            else {
                tc::mbar_arrive_expect_tx(bars + B_WFULL + slot, bytes0 + bytes1);
                tc::bulk_g2s(sm + ring_stage_off(slot, 0), packed + off0, bytes0, bars + B_WFULL + slot);
                tc::bulk_g2s(sm + ring_stage_off(slot, 1), packed + off1, bytes1, bars + B_WFULL + slot);
            }
        }
        __syncwarp();
    }
}
static_assert(TC_NST == 4, "ring index arithmetic assumes 4 stages");
static_assert(2 * TC_GRU_A <= TC_STAGE && (TC_GRU_STEP_BLOCKS & 1) == 0, "a GRU slab pair fills at most one ring stage");

// ================================================================ issuer warp ==============================
__device__ __forceinline__ uint32_t ring_wait(unsigned char* sm, uint64_t* bars, int i) {      // block i resident -> stage address
    if (!(i & 1)) tc_wait(bars, B_WFULL + ((i >> 1) & 1), (uint32_t)((i >> 2) & 1));           // the pair's barrier covers block i+1 too
    return tc::smem_u32(sm + SM_RING + (i & 3) * TC_STAGE);
}

// n_d3 is a running count kept by the caller (uniform across the warp).
// Layer 1 runs per group of 128 hidden units (chunks 4g..4g+3) into one of two 128-column D1 buffers, and well ahead of layer 2: group 0 from the
// lead blocks, group g >= 1 in K slices that ride behind W2 of chunks 4(g-1)-1 .. (complete at least one chunk before chunk 4g is consumed), so the
// epilogue always has a finished D1 to work on while the tensor pipe runs layer 2.  D3 aliases the second D1 buffer (free once chunk 15 is staged).
template <int K1>
__device__ void issuer_mlp(unsigned char* sm, uint64_t* bars, uint32_t tmem, int n3p, int blk0, uint32_t& n_d3, const RingCfg rc) {
    const uint32_t a1 = tc::smem_u32(sm + SM_A1);
    const uint32_t idesc1 = tc::make_idesc_f16(128), idesc2 = tc::make_idesc_f16(256), idesc3 = tc::make_idesc_f16(n3p);
    constexpr int nsl = K1 == 32 ? 2 : 4, nks = (K1 >> 4) / nsl;          // K slices per group, K = 16 steps per slice
    // A operand of slice / half starting at pair column k0: A1 k chunk k0 / 8 (2048 B per chunk), lo image 32 KB behind
    auto a_hi = [&](int k0) { return tc::op_desc(a1 + (uint32_t)(k0 >> 3) * 2048u, 128); };
    auto a_lo = [&](int k0) { return tc::op_desc(a1 + 32768u + (uint32_t)(k0 >> 3) * 2048u, 128); };
    const uint32_t sm0 = tc::smem_u32(sm);
    auto st = [&](int slot, int half) { return sm0 + (uint32_t)ring_stage_off(slot, half); };      // stage address of a slot
    // Ring cursor: the 19 pairs of an MLP are consecutive and lie in ONE regime (three- or two-slot rotation), so slot and use index advance by a compare and an
    // add per pair.  The issuing thread runs one dependent instruction every ~5 cycles and is what paces the tensor pipe whenever it is not back-pressured:
    // everything it executes between a wait and its MMAs is kept to a minimum.
    int cslot, cuse;
    ring_slot(rc, blk0 >> 1, cslot, cuse);
    const int cns = (blk0 >> 1) < rc.n3 ? 3 : 2;
    auto ring_pair = [&]() {                                               // next pair resident -> its slot; the cursor moves on
        const int slot = cslot;
        tc_wait(bars, B_WFULL + slot, (uint32_t)(cuse & 1));
        tc::tc_fence_after();
        if (++cslot == cns) { cslot = 0; ++cuse; }
        return slot;
    };
    // one elected lane runs the whole phase (waits included): no reconvergence between the MMA groups
    const bool tl = K1 == 32 && blk0 != 0;          // y0 MLP (timeline build 2 only)
    if (elect_one()) {
        TLOG2(tl, 19);
        // lead: layer 1 of group 0 (D1 buffer 0: its previous readers finished phases ago), one K half per ring pair
#pragma unroll 1
        for (int kh = 0; kh < 2; ++kh) {
            const int sl_ = ring_pair();
            tc::mma_split3_d(tmem + COL_D1, a_hi(kh * (K1 >> 1)), a_lo(kh * (K1 >> 1)), 256, tc::op_desc(st(sl_, 0), 128), tc::op_desc(st(sl_, 1), 128), 256,
                             idesc1, K1 >> 5, kh > 0);
            if (kh == 1) {
                tc::mma_commit(bars + B_D1FULL);
                // first K slice of group 1 (behind the second K half in the same ring pair) -> D1 buffer 1 = D3 of the previous MLP: its read-out precedes
                // the epilogue's first staged chunk, so this wait ends about a thousand cycles before the first layer-2 MMA can go
                const uint32_t LB = 128u * (uint32_t)K1;
                if (n_d3 >= 1) { tc_wait(bars, B_D3FREE, (n_d3 - 1) & 1); tc::tc_fence_after(); }
                tc::mma_split3_d(tmem + COL_D1 + 128, a_hi(0), a_lo(0), 256, tc::op_desc(st(sl_, 0) + LB, 128), tc::op_desc(st(sl_, 1) + LB, 128), 256, idesc1, nks,
                                 false);
            }
            tc::mma_commit(bars + B_WFREE + sl_);
        }
#pragma unroll 1
        for (int ch = 0; ch < 16; ++ch) {
            if (ch == 0) TLOG2(tl, 20);
            const long long tq0 = TNOW();
            const int sl_ = ring_pair();
            TACC2(tl, 16, tq0);
            const long long tq1 = TNOW();
            const uint32_t sa = st(sl_, 0), sb = st(sl_, 1);
            tc_wait(bars, B_CFULL + (ch & 7), (uint32_t)((ch >> 3) & 1));   // relu(D1(ch) + rowpre) converted in place: per K = 16 step [hi 8 columns | lo 8 columns]
            tc::tc_fence_after();
            TACC2(tl, 17, tq1);
            TLOG2(tl, ch);
            const long long tq2 = TNOW();
            const uint32_t ac = tmem + COL_D1 + (uint32_t)((ch >> 2) & 1) * 128u + (uint32_t)(ch & 3) * 32u;
            tc::mma_split3_ts(tmem + COL_D2, ac, ac + 8, 16, tc::op_desc(sa, 256), tc::op_desc(sb, 256), 512, idesc2, 2, ch > 0);
            if (ch == 15) tc::mma_commit(bars + B_ACC);            // D2 complete
            // K slice (ch+1) & 3 of group ((ch+1) >> 2) + 1: its D1 buffer held group g-2, whose last chunk 4g-5 <= ch was converted before the wait above and read by the layer-2 MMAs issued before this one (the tensor pipe runs them in order)
            const int cp = ch + 1, g = (cp >> 2) + 1, sl = cp & 3;
            if (g < 4 && sl < nsl) {
                const int k0 = sl * (K1 / nsl);
                tc::mma_split3_d(tmem + COL_D1 + (g & 1) * 128, a_hi(k0), a_lo(k0), 256, tc::op_desc(sa + 16384, 128), tc::op_desc(sb + 16384, 128), 256, idesc1,
                                 nks, sl > 0);
                if (sl == nsl - 1) tc::mma_commit(bars + B_D1FULL + (g & 1));
            }
            tc::mma_commit(bars + B_WFREE + sl_);
            TACC2(tl, 18, tq2);
        }
        // layer 3: A = relu(D2 + b1) converted in place (column block 16 s = K step s: [hi 8 | lo 8]), two halves of 8 steps; D3 = D1 buffer 1, whose last
        // reader (layer 2 of chunks 12..15) precedes these MMAs in the pipe
        {
            const int sl_ = ring_pair();
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                tc_wait(bars, B_L3RDY + hf, n_d3 & 1);
                tc::tc_fence_after();
                TLOG2(tl, 21 + hf);
#pragma unroll 1
                for (int p = 4 * hf; p < 4 * hf + 4; ++p) {            // 32-wide K piece p: image [hi n3p x 32 | lo n3p x 32] inside block p >> 2
                    const uint32_t sp_ = st(sl_, p >> 2) + (uint32_t)(p & 3) * n3p * 128;
                    const uint32_t ac = tmem + COL_D2 + (uint32_t)p * 32u;
                    tc::mma_split3_ts(tmem + COL_D3, ac, ac + 8, 16, tc::op_desc(sp_, n3p), tc::op_desc(sp_ + n3p * 64, n3p), 2 * n3p, idesc3, 2, p > 0);
                }
            }
            tc::mma_commit(bars + B_WFREE + sl_);
            tc::mma_commit(bars + B_ACC);                              // D3 complete
            TLOG2(tl, 23);
        }
    }
    __syncwarp();
    ++n_d3;
}

// GRU issue schedule of step t (one elected lane; every wait is on the completion index t of its barrier: index 0 is the epilogue's prologue,
// index t >= 1 comes from the gate epilogue of step t-1):
//   B_GREAD : conv image t staged, half-A accumulators of step t-1 in registers        -> x-slabs of both halves
//   B_HA    : units 0..63 of h(t) staged (every MMA of step t-1 has completed)          -> h-slabs 0..3 (half B first: it frees the ring pair sooner)
//   B_HB    : units 64..95 staged, half-B accumulators of step t-1 read                  -> h-slabs 4, 5; commit half A, then half B
__device__ void issuer_gru(unsigned char* sm, uint64_t* bars, uint32_t tmem, int blk0, int T, const RingCfg rc) {
    const uint32_t sm0 = tc::smem_u32(sm), a1 = tc::smem_u32(sm + SM_A1), xh = tc::smem_u32(sm + SM_H);
    const uint32_t idesc192 = tc::make_idesc_f16(192), idesc128 = tc::make_idesc_f16(128), idesc96 = tc::make_idesc_f16(96),
                   idesc64 = tc::make_idesc_f16(64), idesc32 = tc::make_idesc_f16(32);
    int gslot[4], gpar[4];                          // pair k of step 0: slot, parity of its use index (pair k of step t: use index + 2 t)
#pragma unroll
    for (int k = 0; k < 4; ++k) { int u; ring_slot(rc, (blk0 >> 1) + k, gslot[k], u); gpar[k] = u & 1; }
    if (elect_one()) {
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const bool tl = t == 2;
            const uint32_t par = (uint32_t)(t & 1);
            const uint32_t colA = tmem + COL_GA, colB = tmem + COL_GB + (uint32_t)(t & 1) * 128u;
            // the step's four pairs sit in the two-slot regime, 4 pairs per step: slot and barrier parity of pair k are the same in every step
            auto ring_pair = [&](int k) {                          // pair k of this step resident -> its slot
                tc_wait(bars, B_WFULL + gslot[k], (uint32_t)gpar[k]);      // a slot is used twice per step: the parity of its use index does not change with t
                tc::tc_fence_after();
                return gslot[k];
            };
            auto st = [&](int slot, int half) { return sm0 + (uint32_t)ring_stage_off(slot, half); };
            // state slab j (units 16j..16j+15): A1 k chunks 4 + 2j, 5 + 2j; lo image 32 KB behind the hi image
            auto h_slab = [&](int j, uint64_t& ah, uint64_t& al) {
                const uint32_t a = a1 + 4 * 2048 + (uint32_t)j * 4096;
                ah = tc::op_desc(a, 128); al = tc::op_desc(a + 32768, 128);
            };
            // h-slab into half A (s = image inside the ring stage): r|z|nh = columns 64..255; the FIRST h-slab must not accumulate into nh
            auto mma_hA = [&](uint32_t s, uint64_t ah, uint64_t al, bool first) {
                if (first) {
                    tc::mma_split3_d(colA + 64, ah, al, 0, tc::op_desc(s, 192), tc::op_desc(s + 6144, 192), 0, idesc128, 1, true);
                    tc::mma_split3_d(colA + 192, ah, al, 0, tc::op_desc(s + 128 * 16, 192), tc::op_desc(s + 6144 + 128 * 16, 192), 0, idesc64, 1, false);
                } else tc::mma_split3_d(colA + 64, ah, al, 0, tc::op_desc(s, 192), tc::op_desc(s + 6144, 192), 0, idesc192, 1, true);
            };
            auto mma_hB = [&](uint32_t s, uint64_t ah, uint64_t al, bool first) {
                if (first) {
                    tc::mma_split3_d(colB + 32, ah, al, 0, tc::op_desc(s, 96), tc::op_desc(s + 3072, 96), 0, idesc64, 1, true);
                    tc::mma_split3_d(colB + 96, ah, al, 0, tc::op_desc(s + 64 * 16, 96), tc::op_desc(s + 3072 + 64 * 16, 96), 0, idesc32, 1, false);
                } else tc::mma_split3_d(colB + 32, ah, al, 0, tc::op_desc(s, 96), tc::op_desc(s + 3072, 96), 0, idesc96, 1, true);
            };
            // ---- x-slabs: conv image of step t (H[t & 1], k chunks 0..3) . W_ih^T -> [ni | r | z] of both halves
            TLOG(tl, 20);
            tc_wait(bars, B_GREAD, par);
            tc::tc_fence_after();
            TLOG(tl, 21);
            const int s0 = ring_pair(0);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                const uint32_t xa = xh + (uint32_t)(t & 1) * 16384 + kb * 4096;
                const uint64_t ah = tc::op_desc(xa, 128), al = tc::op_desc(xa + 8192, 128);
                const uint32_t sa = st(s0, 0) + kb * TC_GRU_A, sb = st(s0, 1) + kb * TC_GRU_B;
                tc::mma_split3_d(colA, ah, al, 0, tc::op_desc(sa, 192), tc::op_desc(sa + 6144, 192), 0, idesc192, 1, kb > 0);
                tc::mma_split3_d(colB, ah, al, 0, tc::op_desc(sb, 96), tc::op_desc(sb + 3072, 96), 0, idesc96, 1, kb > 0);
            }
            tc::mma_commit(bars + B_WFREE + s0);
            // ---- h-slabs 0..3
            tc_wait(bars, B_HA, par);
            tc::tc_fence_after();
            TLOG(tl, 22);
            const int s1 = ring_pair(1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t ah, al;
                h_slab(j, ah, al);
                mma_hB(st(s1, j >> 1) + (j & 1) * TC_GRU_B, ah, al, j == 0);
            }
            tc::mma_commit(bars + B_WFREE + s1);
            const int s2 = ring_pair(2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t ah, al;
                h_slab(j, ah, al);
                mma_hA(st(s2, j >> 1) + (j & 1) * TC_GRU_A, ah, al, j == 0);
            }
            tc::mma_commit(bars + B_WFREE + s2);
            // ---- h-slabs 4, 5: the only MMAs on the recurrence's critical path
            tc_wait(bars, B_HB, par);
            tc::tc_fence_after();
            TLOG(tl, 7);
            const int s3 = ring_pair(3);
#pragma unroll
            for (int j = 4; j < 6; ++j) {
                uint64_t ah, al;
                h_slab(j, ah, al);
                mma_hA(st(s3, 0) + (j & 1) * TC_GRU_A, ah, al, false);
            }
            tc::mma_commit(bars + B_ACC);
#pragma unroll
            for (int j = 4; j < 6; ++j) {
                uint64_t ah, al;
                h_slab(j, ah, al);
                mma_hB(st(s3, 1) + (j & 1) * TC_GRU_B, ah, al, false);
            }
            tc::mma_commit(bars + B_ACC2);
            tc::mma_commit(bars + B_WFREE + s3);
        }
        // final state staged and every gate accumulator read out (completion index T)
        tc_wait(bars, B_HA, (uint32_t)(T & 1));
        tc_wait(bars, B_HB, (uint32_t)(T & 1));
    }
    __syncwarp();
    tc::tc_fence_after();
}

// ================================================================ epilogue (256 threads) ===================
// fp16 operands overflow at 65504: any staged activation beyond FP16_SAFE marks the tile, and the FP32-FFMA kernel redoes it
constexpr float FP16_SAFE = 60000.f;
template <int N>
__device__ __forceinline__ void track_range(const float (&v)[N], int& ovf) {
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) mx = fmaxf(mx, fabsf(v[i]));
    ovf |= !(mx <= FP16_SAFE);                      // also true for NaN/inf
}

__device__ __forceinline__ void load_rowbias16(float (&rb)[16], const float* __restrict__ rowpre, int prow, int col) {
    if (prow >= 0) {
        const float4* p = reinterpret_cast<const float4*>(rowpre + rowpre_off(prow, col));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 r = __ldg(p + q * 128);            // next 4-column group of the same row: 512 floats on
            rb[4 * q] = r.x; rb[4 * q + 1] = r.y; rb[4 * q + 2] = r.z; rb[4 * q + 3] = r.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) rb[i] = 0.f;
    }
}

// 16 accumulator columns of one row -> relu -> fp16 hi/lo -> written back IN PLACE as the A operand of the next layer's MMA (tensor-memory A):
// the 16 fp32 columns of a K = 16 step become [hi: 8 columns of packed pairs | lo: 8 columns], so a thread only overwrites columns it has read itself
__device__ __forceinline__ void convert_in_place16(uint32_t taddr, const float (&v)[16], int& ovf) {
    track_range(v, ovf);
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float t8[8];
        uint32_t h4[4], l4[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = v[q * 8 + i];
        tc::split8_packed(t8, h4, l4);
#pragma unroll
        for (int i = 0; i < 4; ++i) { hw[q * 4 + i] = h4[i]; lw[q * 4 + i] = l4[i]; }
    }
    tc::tmem_st8(taddr, hw);
    tc::tmem_st8(taddr + 8, lw);
}
// epilogue warp: finished TMEM writes (and reads) published to the issuer, one arrival per warp
__device__ __forceinline__ void epi_signal_tmem(uint64_t* bars, int bar) {
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + bar);
}

// Epilogue warp groups: sub = warp >> 2.  wh = sub & 1 owns the chunks of index parity wh, half = sub >> 1 owns columns 16*half..16*half+15 (one K = 16
// step) of each 32-wide chunk, so each chunk is converted by 8 warps.  Nothing here waits for the tensor pipe except the layer-1 completion of a
// 4-chunk group: the conversions run up to 7 chunks ahead of layer 2 (8 chunk barriers).
__device__ void epi_mlp(unsigned char* sm, uint64_t* bars, uint32_t tmem, int n3, int n3p, bool first_mlp, const float* __restrict__ rowpre,
                        int prow, const float* __restrict__ b1, const float* __restrict__ b2, float* xy_s, int ldxy, bool accumulate,
                        uint32_t& n_acc, int& ovf) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, sub = warp >> 2, wh = sub & 1, half = sub >> 1;
    const int m = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    (void)sm;
    float rb[16];
    const bool tl2 = !first_mlp && !accumulate && threadIdx.x == 0;      // timeline of the y0 phase (timeline builds only)
    TLOG2(tl2, 24);
#pragma unroll 1
    for (int ch = wh; ch < 16; ch += 2) {
        load_rowbias16(rb, rowpre, prow, ch * 32 + half * 16);
        tc_wait(bars, B_D1FULL + ((ch >> 2) & 1), (uint32_t)((ch >> 3) & 1));      // group ch / 4 in D1 buffer (ch / 4) & 1; 2 completions per buffer per MLP: parity restarts at 0
        tc::tc_fence_after();
        if (ch == 0) TLOG2(tl2, 25);
        float v[16];
        const uint32_t col = tmem + lane_base + COL_D1 + ((ch >> 2) & 1) * 128 + (ch & 3) * 32 + half * 16;
        tc::tmem_ld16(col, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + rb[i], 0.f);
        convert_in_place16(col, v, ovf);
        epi_signal_tmem(bars, B_CFULL + (ch & 7));
        if (ch == 0) TLOG2(tl2, 26);
    }
    TLOG(tl2, 8);
    TLOG2(tl2, 27);
    tc_wait(bars, B_ACC, n_acc & 1);                // D2 complete
    ++n_acc;
    tc::tc_fence_after();
    TLOG(tl2, 9);
    TLOG2(tl2, 28);
    // relu(D2 + b1) in place; this thread converts K steps sub, sub + 4 (first half) and sub + 8, sub + 12 (second half) of its row
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int s = hf * 8 + j * 4 + sub;
            float v[16];
            float4 bq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) bq[q] = __ldg(reinterpret_cast<const float4*>(b1 + s * 16) + q);      // four 16-byte loads, in flight under the TMEM load
            const uint32_t col = tmem + lane_base + COL_D2 + s * 16;
            tc::tmem_ld16(col, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[4 * q] = fmaxf(v[4 * q] + bq[q].x, 0.f); v[4 * q + 1] = fmaxf(v[4 * q + 1] + bq[q].y, 0.f);
                v[4 * q + 2] = fmaxf(v[4 * q + 2] + bq[q].z, 0.f); v[4 * q + 3] = fmaxf(v[4 * q + 3] + bq[q].w, 0.f);
            }
            convert_in_place16(col, v, ovf);
        }
        epi_signal_tmem(bars, B_L3RDY + hf);
    }
    TLOG(tl2, 19);
    TLOG2(tl2, 29);
    tc_wait(bars, B_ACC, n_acc & 1);                // D3 complete
    ++n_acc;
    tc::tc_fence_after();
    TLOG(tl2, 23);
    TLOG2(tl2, 30);
    for (int c0 = 0; c0 < n3p; c0 += 16) {
        if (((c0 >> 4) & 3) != sub) continue;       // warp-uniform: the four groups take 16-wide column groups in turn
        float v[16];
        tc::tmem_ld16(tmem + lane_base + COL_D3 + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int o = c0 + i;
            if (o < n3) {
                const float r = v[i] + __ldg(b2 + o);
                xy_s[m * ldxy + o] = accumulate ? xy_s[m * ldxy + o] + r : r;
            }
        }
    }
    epi_signal(bars, B_D3FREE);
    epi_barrier();
    TLOG(tl2, 30);
}

constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / x for x in [1, 2^65] on the FMA pipe (bit-trick seed, relative error 5 %, three Newton steps in residual form: <= 6e-8, the class of rcp.approx).
// The gate math is bound by the 16-lane XU pipe (3 ex2 + 2 rcp per unit); this moves one of the five to the 128-lane pipes.
__device__ __forceinline__ float rcp_newton(float x) {
    float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
#pragma unroll
    for (int i = 0; i < 3; ++i) r = fmaf(r, fmaf(-x, r, 1.f), r);
    return r;
}
// GRU constants in shared memory, biases pre-scaled for the ex2-based gates of epi_gru: [-log2e (b_ir+b_hr | b_iz+b_hz) 192 | 2 log2e b_in 96 | 2 log2e b_hn 96 | conv_w 192 | conv_b 32]
__device__ __forceinline__ float gru_const(const sttode_decompose_weights& w, int i) {
    if (i < 192) return -LOG2E * (w.gru_b_ih[i] + w.gru_b_hh[i]);
    if (i < 288) return 2.f * LOG2E * w.gru_b_ih[i];
    if (i < 384) return 2.f * LOG2E * w.gru_b_hh[i - 96];
    if (i < 576) return w.conv_w[i - 384];
    return w.conv_b[i - 576];
}

// Conv1d+ReLU -> GRU over T steps; conv input d = xy_s; final state left in A1 k 32..127 (fp16 hi/lo images).
// gb: shared-memory constants of gru_const()
// last_step_work(): run once in the slot where the last step has no next conv image to build (the tensor pipe is busy with h-slabs 4, 5 then):
// the pair kernel stages the NEXT tile's latent there, into the conv buffer H[T & 1] that is free from that point on
template <typename LastStepWork>
__device__ void epi_gru(unsigned char* sm, uint64_t* bars, uint32_t tmem, const float* gb, const float* xy_s, int ldxy, int valid_rows, int T,
                        uint32_t& n_acc, int& ovf, LastStepWork last_step_work) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, sub = warp >> 2;
    const int m = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    unsigned char* a1_hi_p = sm + SM_A1;
    unsigned char* a1_lo_p = a1_hi_p + 32768;
    const float *brz = gb, *bin = gb + 192, *bhn = gb + 288, *cw = gb + 384, *cb = gb + 576;
    // h0 = 0: clear k chunks 4..15 of both images
    for (int i = tid; i < 2 * 12 * 128; i += EPI_THREADS) {
        const int img = i / (12 * 128), r = i - img * 12 * 128;
        *reinterpret_cast<uint4*>(sm + SM_A1 + img * 32768 + 4 * 2048 + r * 16) = make_uint4(0, 0, 0, 0);
    }
    float hreg[24];                                 // this thread's 24 hidden units of row m (chunks 2*sub, 2*sub+1, 8+sub), fp32 across the recurrence
#pragma unroll
    for (int i = 0; i < 24; ++i) hreg[i] = 0.f;
    TLOG(tid == 0 && ldxy != 2 * T, 5);            // pair kernel only (the row kernel passes ldxy = 2 T)
    // conv channels sub*8 .. +8 of row m at time t -> k chunk `sub` of the conv image of step t, H[t & 1]: the images alternate between the two
    // H buffers so that step t+1's image is computed while the tensor pipe works on step t (it does not depend on the state)
    auto stage_conv = [&](int t) {
        float d[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int tt = t + k - 1;
            const bool in = tt >= 0 && tt < T && m < valid_rows;
            d[k] = in ? xy_s[m * ldxy + tt * 2] : 0.f;
            d[3 + k] = in ? xy_s[m * ldxy + tt * 2 + 1] : 0.f;
        }
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int chn = sub * 8 + i;
            float a = cb[chn];
#pragma unroll
            for (int e = 0; e < 6; ++e) a = fmaf(cw[chn * 6 + e], d[e], a);
            v[i] = fmaxf(a, 0.f);
        }
        track_range(v, ovf);
        unsigned char* img = sm + SM_H + (t & 1) * 16384;
        tc::store_split8(img, img + 8192, 128, m, sub, v);
    };
    stage_conv(0);
    // completion index 0 of the three hand-off barriers: conv image 0 + zero state staged, nothing to read out yet
    epi_signal(bars, B_GREAD);
    epi_signal(bars, B_HA);
    epi_signal(bars, B_HB);
    // this thread's three 8-unit chunks of row m: chunks 2*sub, 2*sub+1 (half A, units 0..63) and 8+sub (half B, units 64..95)
    constexpr std::integral_constant<int, 0> Q0{};
    constexpr std::integral_constant<int, 1> Q1{};
    constexpr std::integral_constant<int, 2> Q2{};
    auto gates = [&](auto qc, int t) {
        constexpr int q = decltype(qc)::value;
        const int chunk = q < 2 ? sub * 2 + q : 8 + sub, u = chunk * 8;
        // accumulator columns of a half: [W_in x | r | z | W_hn h], each `w` wide
        const uint32_t base = tmem + lane_base + (q < 2 ? COL_GA + u : COL_GB + (uint32_t)(t & 1) * 128u + (u - 64)), w = q < 2 ? 64u : 32u;
        float gr[8], gz[8], gni[8], gnh[8];
        tc::tmem_ld8x4(base + w, base + 2 * w, base, base + 3 * w, gr, gz, gni, gnh);
        TLOG(t == 2 && tid == 0, 10 + 2 * q);
        if (q == 1) epi_signal(bars, B_GREAD);      // half A is in registers (q = 0 was read before) and conv image t+1 is staged: x-slabs of step t+1 may run
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // r = 1/(1+ea), z = 1/(1+eb) with one reciprocal of the product.  The biases are pre-scaled by -log2(e) (r, z) and 2 log2(e) (n), so a
            // pre-activation is ONE fma away from the ex2 argument; the clamp (pre-activation >= -30) keeps da * db finite, large positive x gives 2^-x -> 0.
            // ex2 / rcp in their .ftz forms: no operand here is denormal, and the non-ftz forms cost three extra instructions each.
            const float da = 1.f + ex2_ftz(fminf(fmaf(gr[i], -LOG2E, brz[u + i]), 30.f * LOG2E));
            const float db = 1.f + ex2_ftz(fminf(fmaf(gz[i], -LOG2E, brz[96 + u + i]), 30.f * LOG2E));
            const float inv = rcp_ftz(da * db);
            const float r = inv * db, zg = inv * da;
            // tanh(y) = 1 - 2 / (1 + e^(2y));  the exponent is clamped at 64 (tanh = 1 to fp32 long before) so that 1 + e^(2y) stays in rcp_newton's range; 0 gives -1
            const float ec = ex2_ftz(fminf(fmaf(r, fmaf(gnh[i], 2.f * LOG2E, bhn[u + i]), fmaf(gni[i], 2.f * LOG2E, bin[u + i])), 64.f));
            const float n = fmaf(-2.f, rcp_newton(1.f + ec), 1.f);
            const float hp = hreg[q * 8 + i];
            hreg[q * 8 + i] = fmaf(hp - n, zg, n);
        }
        TLOG(t == 2 && tid == 0, 11 + 2 * q);
    };
    auto stage_h = [&](auto qc) {
        constexpr int q = decltype(qc)::value;
        float t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = hreg[q * 8 + i];
        tc::store_split8(a1_hi_p, a1_lo_p, 128, m, 4 + (q < 2 ? sub * 2 + q : 8 + sub), t8);
    };
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        const bool tl = (t == 2 && tid == 0);
        TLOG(tl, 0);
        // H[(t+1) & 1] held the image of step t-1, whose MMAs completed before the half-A completion this warp waited for in the last iteration;
        // the image is published by the B_GREAD arrival below.  It is built while the tensor pipe runs the last two h-slabs of step t.
        if (t + 1 < T) stage_conv(t + 1);
        else last_step_work();
        TLOG(tl, 1);
        tc_wait(bars, B_ACC, n_acc & 1);            // half A of step t complete
        ++n_acc;
        tc::tc_fence_after();
        TLOG(tl, 2);
        gates(Q0, t);
        gates(Q1, t);
        // the old state image is read by the MMAs of BOTH halves: wait for half B too (complete long ago) before overwriting it
        tc_wait(bars, B_ACC2, (uint32_t)(t & 1));
        tc::tc_fence_after();
        TLOG(tl, 16);
        stage_h(Q0);
        stage_h(Q1);
        TLOG(tl, 17);
        epi_signal(bars, B_HA);                     // units 0..63 of h(t+1): h-slabs 0..3 of step t+1 run under the gate math of half B
        TLOG(tl, 3);
        gates(Q2, t);
        stage_h(Q2);
        epi_signal(bars, B_HB);                     // units 64..95 staged, half-B accumulators read
        TLOG(tl, 18);
    }
    TLOG(tid == 0 && ldxy != 2 * T, 6);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_dec_pair_tc(const float* __restrict__ traj_norm, const float* __restrict__ z, unsigned long long seed, unsigned long long pair_base,
              const unsigned long long* __restrict__ rng_state,
              const float* __restrict__ rowpre, const int* __restrict__ row_scene, const float* __restrict__ origin, int rows, int K,
              int rpt, int Tp, int Tf, sttode_decompose_weights w0, sttode_decompose_weights w1,
              const unsigned char* __restrict__ packed, const __grid_constant__ DecPackLayout lay, float* __restrict__ pred,
              float* __restrict__ recon, int* __restrict__ status, int* __restrict__ tile_flags, int ntiles, int* __restrict__ tile_counter, int y_first) {
    extern __shared__ __align__(128) unsigned char sm[];
    // y_first: phase order x0, y0, GRU, y1 (y0 = decoder_y0 needs only the latent): both K1 = 32 MLPs then run while the state half of A1 is idle and serves
    // as a third ring slot; y0's output must survive the GRU, so it gets its own buffer (set by the host when the extra shared memory fits).
    // Otherwise x0, GRU, y0, y1 with one buffer for x and y.
    const int Tm = Tp > Tf ? Tp : Tf;
    const int ldxy = y_first ? 2 * Tp + 1 : 2 * Tm + 1, ldy = y_first ? 2 * Tf + 1 : ldxy;
    float* xy_s = reinterpret_cast<float*>(sm + SM_XY);
    float* y_s = y_first ? xy_s + BM * ldxy : xy_s;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_XY + ((BM * (y_first ? ldxy + ldy : ldxy) * 4 + 15) / 16) * 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);
    float* gb = reinterpret_cast<float*>(bars + B_COUNT + 3);          // GRU/conv constants of block 1 (608 floats)
    float* cur_base = gb + 608;                                        // 2 x [BM][4]: current position (x, y) and scene origin (x, y) of every pair's agent (this tile / next tile)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, sub = (warp >> 2) & 3;
    const int m = wq * 32 + lane;
    const size_t P = (size_t)rows * K;
    if (rng_state) seed += *rng_state;
    const int total_blocks = TC_MLP_BLOCKS * 3 + TC_GRU_STEP_BLOCKS * Tp + (recon ? TC_MLP_BLOCKS : 0);
    const int blk_y0 = y_first ? TC_MLP_BLOCKS : TC_MLP_BLOCKS + TC_GRU_STEP_BLOCKS * Tp, blk_gru = y_first ? 2 * TC_MLP_BLOCKS : TC_MLP_BLOCKS,
              blk_y1 = 2 * TC_MLP_BLOCKS + TC_GRU_STEP_BLOCKS * Tp;
    const RingCfg rc = make_ring_cfg(y_first ? TC_MLP_BLOCKS : 0);      // pairs of x0 and y0 (2 x TC_MLP_BLOCKS / 2) rotate over three slots
    // Persistent CTA (one per SM): tensor memory, the GRU constants and the CTA itself are set up once (a tile costs ~160 k cycles, a CTA hand-over
    // + set-up ~8 k).  Tiles are claimed from a counter one tile ahead (tile_s[it & 1] = this tile, tile_s[(it + 1) & 1] = the next one): the SMs do
    // not run at the same speed (distance to the L2 slices), and a static split of equal tiles lost ~3 % to the slowest SM.
    int* tile_s = reinterpret_cast<int*>(cur_base + 2 * BM * 4);
    if (tid == PRODUCER_WARP * 32) tile_s[0] = atomicAdd(tile_counter, 1);
    __syncthreads();
    if (warp == ISSUER_WARP) tc::tmem_alloc(tmem_slot, 512);
    if (warp < ISSUER_WARP)
        for (int i = tid; i < 608; i += EPI_THREADS) gb[i] = gru_const(w1, i);
    int ovf_next = 0, it = 0;
    // latent z of pair tile `t` -> fp16 hi/lo image of 128 rows x 32 k (this thread: k chunk `sub` of row m), plus the current position / scene origin of
    // the pairs' agents (dependent global loads: row_scene -> origin) for the output phase
    auto stage_tile_inputs = [&](int t, unsigned char* img_hi, unsigned char* img_lo, float* cur_dst, int& ovf_t) {
        const size_t t_pair0 = ((size_t)rpt + t) * BM;
        const int t_np = (int)min((size_t)BM, P - t_pair0);
        const int t_prow = m < t_np ? (int)((t_pair0 + m) / K) : -1;
        float v[8];
        if (m < t_np) {
            const size_t pair = t_pair0 + m;
            float4 lo4, hi4;
            if (z) {
                lo4 = *reinterpret_cast<const float4*>(z + pair * 32 + sub * 8);
                hi4 = *reinterpret_cast<const float4*>(z + pair * 32 + sub * 8 + 4);
            } else {
                lo4 = latent_normal4(seed, pair_base + pair, sub * 2);
                hi4 = latent_normal4(seed, pair_base + pair, sub * 2 + 1);
            }
            v[0] = lo4.x; v[1] = lo4.y; v[2] = lo4.z; v[3] = lo4.w; v[4] = hi4.x; v[5] = hi4.y; v[6] = hi4.z; v[7] = hi4.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        track_range(v, ovf_t);
        tc::store_split8(img_hi, img_lo, 128, m, sub, v);
        if (sub == 0) {
            float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t_prow >= 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(traj_norm + (size_t)t_prow * 2 * Tp));       // the row the conv-input build reads after the x0 MLP
                const float2 tn = *reinterpret_cast<const float2*>(traj_norm + ((size_t)t_prow * Tp + Tp - 1) * 2);
                c4.x = tn.x; c4.y = tn.y;
                if (origin) { const int sc = row_scene[t_prow]; c4.z = origin[sc * 2]; c4.w = origin[sc * 2 + 1]; }
            }
            *reinterpret_cast<float2*>(cur_dst + m * 4) = make_float2(c4.x, c4.y);          // 8-byte aligned (it follows the mbarriers)
            *reinterpret_cast<float2*>(cur_dst + m * 4 + 2) = make_float2(c4.z, c4.w);
        }
    };
#pragma unroll 1
    for (;; ++it) {
    const int tile = tile_s[it & 1];
    if (tile >= ntiles) break;
    float* cur_s = cur_base + (it & 1) * BM * 4;
    unsigned char* znext = sm + SM_H + (Tp & 1) * 16384;           // [hi 8 KB | lo 8 KB]: the conv buffer the last GRU step leaves free
    // A tile is 128 CONSECUTIVE (agent, sample) pairs of the whole batch (pair p = agent row p / K, sample p % K): tiles cut on agent
    // boundaries would leave 128 - 6*20 = 8 of 128 MMA rows idle at K = 20.  `rpt` carries the index of this launch's first tile.
    const size_t pair0 = ((size_t)rpt + tile) * BM;
    const int np = (int)min((size_t)BM, P - pair0);
    const int prow = m < np ? (int)((pair0 + m) / K) : -1;
    int ovf = ovf_next;                             // set when a staged fp16 operand would leave the fp16 range (the latent was staged one tile ahead)
    ovf_next = 0;
#ifdef STTODE_TC_TIMELINE
    if (blockIdx.x == 0 && tid == 0) {              // the stamps are taken on CTA 0's SIXTH tile (every SM busy); published to the CTA by the prologue's barrier
        long long* tl = reinterpret_cast<long long*>(status + 2);
        const bool on = (status[1] & 3) == 2;
        g_tl = on && it == 5 ? tl : nullptr;
#if STTODE_TC_TIMELINE == 3
        if (on && (it == 5 || it == 6)) tl[it == 5 ? 0 : 8] = clock64();      // slot 0: start of the stamped tile, slot 8: start of the next one
        if (on && it < 56 && (it % 10 == 5 || it % 10 == 6)) tl[12 + 2 * (it / 10) + (it % 10 - 5)] = clock64();      // slots 12..23: tile periods at it = 5, 15, .., 55
#endif
    }
#endif

    // every phase parity below is a closed-form function of the loop indices of ONE tile: the barriers are re-created per tile.  All their users are
    // past the barrier that ends the previous tile, and the issuer has waited for the last asynchronous arrivals (tcgen05.commit on the ring barriers).
    if (tid == PRODUCER_WARP * 32) {                // the producer warp is idle here (ONE thread: inval / init of many barriers by the lanes of a warp at once lost phases on the GPU)
        tile_s[(it + 1) & 1] = atomicAdd(tile_counter, 1);      // published by the barrier below; its previous value was last read a whole tile ago
        for (int i = 0; i < B_COUNT; ++i) {
            const int cnt = (i == B_HFULL || i == B_HFULL + 1) ? 8 : (i >= B_CFULL && i < B_CFULL + 8) ? 8 : (i == B_XFULL || i == B_D3FREE || i == B_GREAD || i == B_HA || i == B_HB || i == B_L3RDY || i == B_L3RDY + 1) ? 16 : 1;     // arrivals: one per epilogue warp
            if (it) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bars + i)) : "memory");
            tc::mbar_init(bars + i, cnt);
        }
        tc::fence_mbar_init();
    }
    if (warp < ISSUER_WARP) {
        // latent z -> A1 k chunks 0..3, current positions -> cur_s: generated here for the CTA's first tile; every later tile finds them staged by the
        // previous tile's GRU phase (z images in the spare conv buffer: two 8 KB copies)
        if (it == 0) stage_tile_inputs(tile, sm + SM_A1, sm + SM_A1 + 32768, cur_s, ovf);
        else
            for (int i = tid; i < 1024; i += EPI_THREADS) {
                const int img = i >> 9, o = (i & 511) * 16;
                *reinterpret_cast<uint4*>(sm + SM_A1 + img * 32768 + o) = *reinterpret_cast<const uint4*>(znext + img * 8192 + o);
            }
        if (sub == 1 && prow >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(traj_norm + (size_t)prow * 2 * Tp));
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    TLOG3(tid == 0, 1);

    if (warp == PRODUCER_WARP) {                    // ---------------- producer
        producer_loop(sm, bars, packed, total_blocks, (status[1] & 4) != 0, rc, [&](int i, uint32_t& off, uint32_t& bytes) { block_src(lay, Tp, y_first, i, off, bytes); });
    } else if (warp == ISSUER_WARP) {               // ---------------- issuer
        uint32_t n_d3 = 0;
        issuer_mlp<32>(sm, bars, tmem, (int)lay.n3p[0], 0, n_d3, rc);
        if (y_first) issuer_mlp<32>(sm, bars, tmem, (int)lay.n3p[1], blk_y0, n_d3, rc);
        issuer_gru(sm, bars, tmem, blk_gru, Tp, rc);
        if (!y_first) issuer_mlp<32>(sm, bars, tmem, (int)lay.n3p[1], blk_y0, n_d3, rc);
        issuer_mlp<128>(sm, bars, tmem, (int)lay.n3p[2], blk_y1, n_d3, rc);
        if (recon) issuer_mlp<128>(sm, bars, tmem, (int)lay.n3p[3], blk_y1 + TC_MLP_BLOCKS, n_d3, rc);
        // the last commit on every slot's free barrier has no other waiter: it must have arrived before the barriers are re-created
        {
            const int np_ = total_blocks >> 1;
            for (int j = np_ - 2; j < np_; ++j) { int slot, use; ring_slot(rc, j, slot, use); tc_wait(bars, B_WFREE + slot, (uint32_t)(use & 1)); }
            if (rc.n3 >= 3) { const int j = rc.n3 - 1 - (rc.n3 % 3); int slot, use; ring_slot(rc, j, slot, use); tc_wait(bars, B_WFREE + slot, (uint32_t)(use & 1)); }   // last pair in slot 2
        }
    } else {                                        // ---------------- epilogue warps
        uint32_t n_acc = 0;
        TLOG(tid == 0, 24);
        TLOG3(tid == 0, 2);
        epi_mlp(sm, bars, tmem, 2 * Tp, (int)lay.n3p[0], true, rowpre + 0, prow, w0.dec_x.b1, w0.dec_x.b2, xy_s, ldxy, false, n_acc, ovf);
        TLOG(tid == 0, 25);
        TLOG3(tid == 0, 3);
        if (recon)
            for (int idx = tid; idx < np * 2 * Tp; idx += EPI_THREADS) {
                const int mm = idx / (2 * Tp), e = idx - mm * 2 * Tp;
                recon[(pair0 + mm) * 2 * Tp + e] = xy_s[mm * ldxy + e];
            }
        if (y_first) {                              // y0 right away (the issuer is already on its layer 1); the conv input is built after it
            TLOG(tid == 0, 26);
            epi_mlp(sm, bars, tmem, 2 * Tf, (int)lay.n3p[1], false, rowpre + rowpre_col(512), prow, w0.dec_y.b1, w0.dec_y.b2, y_s, ldy, false, n_acc, ovf);
            TLOG(tid == 0, 27);
            TLOG3(tid == 0, 5);
        } else if (recon) epi_barrier();            // the copy above and the build below map threads to elements differently
        {   // conv input of block 1: x_true - x_hat0.  Thread (m, sub) owns a quarter of row m's 2 T_p values: no index division, the
            // agent row is already known (prow), independent loads of a line the prologue prefetched (this loop cost 3.8 k cycles per tile)
            const int per = (2 * Tp + 3) >> 2, e0 = sub * per, e1 = min(2 * Tp, e0 + per);
            const float* tn = traj_norm + (size_t)(prow >= 0 ? prow : 0) * 2 * Tp;
#pragma unroll 4
            for (int e = e0; e < e1; ++e) xy_s[m * ldxy + e] = prow >= 0 ? __ldg(tn + e) - xy_s[m * ldxy + e] : 0.f;
        }
        epi_barrier();
        TLOG(tid == 0, 4);
        epi_gru(sm, bars, tmem, gb, xy_s, ldxy, BM, Tp, n_acc, ovf, [&] {
            const int nt = tile_s[(it + 1) & 1];
            if (nt < ntiles) stage_tile_inputs(nt, znext, znext + 8192, cur_base + ((it + 1) & 1) * BM * 4, ovf_next);
        });
        epi_barrier();                              // every thread is done reading the conv input before y0 overwrites xy_s
        TLOG3(tid == 0, 4);
        if (!y_first) {
            TLOG(tid == 0, 26);
            epi_mlp(sm, bars, tmem, 2 * Tf, (int)lay.n3p[1], false, rowpre + rowpre_col(512), prow, w0.dec_y.b1, w0.dec_y.b2, y_s, ldy, false, n_acc, ovf);
            TLOG(tid == 0, 27);
            TLOG3(tid == 0, 5);
        }
        epi_mlp(sm, bars, tmem, 2 * Tf, (int)lay.n3p[2], false, rowpre + rowpre_col(1024), prow, w1.dec_y.b1, w1.dec_y.b2, y_s, ldy, true, n_acc, ovf);
        TLOG(tid == 0, 28);
        TLOG3(tid == 0, 6);
        if (!(Tf & 1) && !(reinterpret_cast<uintptr_t>(pred) & 15)) {
            // the tile's predictions are ONE contiguous span of np * 8 T_f bytes: consecutive threads write consecutive 16-byte pieces, so every store instruction of a
            // warp covers 512 contiguous bytes (whole 128-byte lines: this also keeps a host-mapped `pred`, written over PCIe, efficient)
            const int hp = Tf >> 1;
            float4* dst = reinterpret_cast<float4*>(pred + pair0 * 2 * Tf);
#pragma unroll 2
            for (int idx = tid; idx < np * hp; idx += EPI_THREADS) {
                const int mm = idx / hp, j = idx - mm * hp;
                const float* src = y_s + mm * ldy + 4 * j;
                const float cx = cur_s[mm * 4], cy = cur_s[mm * 4 + 1];
                float4 o = make_float4(src[0] + cx, src[1] + cy, src[2] + cx, src[3] + cy);
                if (origin) { const float ox = cur_s[mm * 4 + 2], oy = cur_s[mm * 4 + 3]; o.x += ox; o.y += oy; o.z += ox; o.w += oy; }
                dst[idx] = o;
            }
        } else {
            for (int idx = tid; idx < np * 2 * Tf; idx += EPI_THREADS) {
                const int mm = idx / (2 * Tf), e = idx - mm * 2 * Tf, cc = e & 1;
                float v = y_s[mm * ldy + e] + cur_s[mm * 4 + cc];
                if (origin) v += cur_s[mm * 4 + 2 + cc];
                pred[(pair0 + mm) * 2 * Tf + e] = v;
            }
        }
        if (recon) {
            epi_barrier();
            epi_mlp(sm, bars, tmem, 2 * Tp, (int)lay.n3p[3], false, rowpre + rowpre_col(1536), prow, w1.dec_x.b1, w1.dec_x.b2, xy_s, ldxy, false, n_acc, ovf);
            for (int idx = tid; idx < np * 2 * Tp; idx += EPI_THREADS) {
                const int mm = idx / (2 * Tp), e = idx - mm * 2 * Tp;
                recon[(pair0 + mm) * 2 * Tp + e] += xy_s[mm * ldxy + e];
            }
        }
    }
    if (ovf) tile_flags[tile] = 1;
    TLOG(tid == 0, 29);
    TLOG3(tid == 0, 7);
    tc::tc_fence_before();
    __syncthreads();
    TLOG3(tid == 0, 10);
    }   // tile loop
    tc::tc_fence_after();
    if (warp == ISSUER_WARP) tc::tmem_dealloc(*tmem_slot, 512);
}


// ================================================================ per-agent-row kernel on tensor cores ======
// k_dec_row_tc: block-0 Conv1d+GRU on the past trajectory (sample independent) and the four layer-1 partial products
//   rowpre[r] = [ W0x[:, pf|st].[pf|st0]+b | W0y[:, pf|st].[pf|st0]+b | W1y[:, pf].pf+b | W1x[:, pf].pf+b ]   (512 each)
// for a tile of 128 agent rows.  Same roles / ring / GRU code as the pair kernel; the P phase runs 8 output tiles of
// [128 x 256] (double buffered in TMEM), K = 224 or 128 in 16-wide slabs streamed through the ring.
constexpr int SMR_PFHI = SM_H;                   // P phase: past_feature image hi (32 KB) overlays the conv-image / chunk buffers
constexpr int SMR_PFLO = SM_RING + TC_NST * TC_STAGE;     // past_feature image lo (32 KB)
constexpr int SMR_BARS = SMR_PFLO + 32768;

__device__ __forceinline__ void row_block_src(const DecPackLayout& lay, int Tp, int i, uint32_t& off, uint32_t& bytes) {
    if (i < TC_GRU_STEP_BLOCKS * Tp) {
        gru_block(i & 7, off, bytes);
        off += lay.gru_off[0];
        return;
    }
    off = lay.rowp_off + (uint32_t)(i - TC_GRU_STEP_BLOCKS * Tp) * 16384;
    bytes = 16384;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_dec_row_tc(const float* __restrict__ past_feature, const float* __restrict__ traj_norm, int rows, int Tp, sttode_decompose_weights w0,
             sttode_decompose_weights w1, const unsigned char* __restrict__ packed, const __grid_constant__ DecPackLayout lay,
             float* __restrict__ rowpre, int* __restrict__ tile_flags) {
    extern __shared__ __align__(128) unsigned char sm[];
    int ovf = 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SMR_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);
    float* gb = reinterpret_cast<float*>(bars + B_COUNT + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, sub = (warp >> 2) & 3;
    const int m = wq * 32 + lane;
    const int row0 = blockIdx.x * BM;
    const int nrow = min(BM, rows - row0);
    const int total_blocks = TC_GRU_STEP_BLOCKS * Tp + 88;

    if (tid == 0) {
        for (int i = 0; i < B_COUNT; ++i)
            tc::mbar_init(bars + i, (i == B_HFULL || i == B_HFULL + 1 || i == B_XFULL || i == B_D3FREE || i == B_GREAD || i == B_HA || i == B_HB) ? 16 : 1);
        tc::fence_mbar_init();
    }
    if (warp == ISSUER_WARP) tc::tmem_alloc(tmem_slot, 512);
    if (warp < ISSUER_WARP)
        for (int i = tid; i < 608; i += EPI_THREADS) gb[i] = gru_const(w0, i);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == PRODUCER_WARP) {                    // ---------------- producer
        producer_loop(sm, bars, packed, total_blocks, 0, make_ring_cfg(0), [&](int i, uint32_t& off, uint32_t& bytes) { row_block_src(lay, Tp, i, off, bytes); });
    } else if (warp == ISSUER_WARP) {               // ---------------- issuer
        issuer_gru(sm, bars, tmem, 0, Tp, make_ring_cfg(0));
        tc_wait(bars, B_XFULL, 0);                  // past_feature images staged
        tc::tc_fence_after();
        const uint32_t pf_hi = tc::smem_u32(sm + SMR_PFHI), pf_lo = tc::smem_u32(sm + SMR_PFLO);
        const uint32_t st_hi = tc::smem_u32(sm + SM_A1) + 4 * 2048, st_lo = st_hi + 32768;
        const uint32_t idesc = tc::make_idesc_f16(256);
        int blk = TC_GRU_STEP_BLOCKS * Tp;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {               // output tile j: product p = j/2, columns (j&1)*256 ..
            const int b = j & 1, nks = j < 4 ? 14 : 8;
            if (j >= 2) { tc_wait(bars, B_HFULL + b, (uint32_t)(((j >> 1) + 1) & 1)); tc::tc_fence_after(); }     // D[b] read out
#pragma unroll 1
            for (int ks = 0; ks < nks; ++ks, ++blk) {
                const uint32_t st = ring_wait(sm, bars, blk);
                tc::tc_fence_after();
                const uint32_t ah = ks < 8 ? pf_hi + ks * 4096 : st_hi + (ks - 8) * 4096;
                const uint32_t al = ks < 8 ? pf_lo + ks * 4096 : st_lo + (ks - 8) * 4096;
                if (elect_one()) {
                    tc::mma_split3_d(tmem + b * 256, tc::op_desc(ah, 128), tc::op_desc(al, 128), 0, tc::op_desc(st, 256), tc::op_desc(st + 8192, 256), 0,
                                     idesc, 1, ks > 0);
                    if (blk & 1) tc::mma_commit(bars + B_WFREE + ((blk >> 1) & 1));
                    if (ks == nks - 1) tc::mma_commit(bars + B_D1FULL + b);
                }
                __syncwarp();
            }
        }
    } else {                                        // ---------------- epilogue warps
        uint32_t n_acc = 0;
        epi_gru(sm, bars, tmem, gb, traj_norm + (size_t)row0 * 2 * Tp, 2 * Tp, nrow, Tp, n_acc, ovf, [] {});
        // past_feature tile -> fp16 hi/lo images (k chunks 4*sub .. 4*sub+3 of row m); safe: every GRU MMA has completed
        {
            const int r = row0 + m;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const int kc = sub * 4 + q;
                float v[8];
                if (m < nrow) {
                    const float4 a = *reinterpret_cast<const float4*>(past_feature + (size_t)r * 128 + kc * 8);
                    const float4 bb = *reinterpret_cast<const float4*>(past_feature + (size_t)r * 128 + kc * 8 + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bb.x; v[5] = bb.y; v[6] = bb.z; v[7] = bb.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                }
                track_range(v, ovf);
                tc::store_split8(sm + SMR_PFHI, sm + SMR_PFLO, 128, m, kc, v);
            }
        }
        epi_signal(bars, B_XFULL);
        const float* bias[4] = {w0.dec_x.b0, w0.dec_y.b0, w1.dec_y.b0, w1.dec_x.b0};
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            const int b = j & 1, p = j >> 1, nb = j & 1;
            tc_wait(bars, B_D1FULL + b, (uint32_t)((j >> 1) & 1));
            tc::tc_fence_after();
            const float4* bp = reinterpret_cast<const float4*>(bias[p] + nb * 256 + sub * 64);
            float4* dst = reinterpret_cast<float4*>(rowpre + rowpre_off(row0 + m, p * 512 + nb * 256 + sub * 64));      // + 128 per 4-column group; lanes = consecutive rows: 512 contiguous bytes per store
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                tc::tmem_ld32(tmem + lane_base + b * 256 + sub * 64 + c0, v);
                if (m < nrow) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bq = __ldg(bp + (c0 >> 2) + q);
                        dst[((c0 >> 2) + q) * 128] = make_float4(v[q * 4] + bq.x, v[q * 4 + 1] + bq.y, v[q * 4 + 2] + bq.z, v[q * 4 + 3] + bq.w);
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + B_HFULL + b);
        }
    }
    if (ovf) tile_flags[blockIdx.x] = 1;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ISSUER_WARP) tc::tmem_dealloc(tmem, 512);
}

size_t dec_row_tc_smem_bytes() { return (size_t)SMR_BARS + (B_COUNT + 3) * 8 + 608 * 4; }

size_t dec_pair_tc_smem_bytes(int Tp, int Tf, int y_first) {
    const int Tm = Tp > Tf ? Tp : Tf;
    const int ld = y_first ? (2 * Tp + 1) + (2 * Tf + 1) : 2 * Tm + 1;
    return (size_t)SM_XY + ((size_t)(BM * ld * 4 + 15) / 16) * 16 + (B_COUNT + 3) * 8 + 608 * 4 + 2 * BM * 4 * 4 + 16;
}
int dec_pair_tc_threads() { return TC_THREADS; }

}  // namespace sttode
