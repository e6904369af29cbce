// Tensor-core (tcgen05 / TMEM) versions of the named hot path: the MHGSA in-projection and the tensor-ODE right-hand side, sm_100a.
//
// Reference semantics (file:line relative to the reference root):
//   k_qkv_tc        [q|k|v] = x.W_in^T + b_in ; q *= 8^-1/2 ; per-head Oblique.proj of q and k      hyptransformerlib.py:115,191,214-218,252-253;
//                                                                                                    core/manifolds/oblique.py:15-16
//   k_ode_tail_tc   (v-proj when L = 1) -> out_proj -> tanh*sigmoid gate -> +x, LN1 -> FFN 64->1024->64 -> +s, LN2 -> base + h*f -> relu
//                                                                hyptransformerlib.py:304-305; hypertransformer.py:81-83,148-152; ode_demo.py:188,231
// Same algebra as k_proj / k_ode_tail (encoder.cu, FP32 FFMA); every dense contraction runs as tcgen05.mma kind::f16 with the
// 3xFP16 hi/lo split of tc.cuh (fp32-accurate: SURVEY finding 9 rules out single-pass 16/19-bit operands near the acos clamp),
// accumulators in TMEM.  The K = 8 Gram of the geodesic score stays on FFMA (k_geo_attn): a 128-row tcgen05 tile would be > 90 % padding.
//
// Tile = 128 rows.  Roles (320 threads): warps 0..7 epilogue (TMEM -> registers -> bias / activation / LayerNorm -> fp16 hi/lo
// operand images in shared memory, or the output rows), warp 8 issuer (one elected lane issues every tcgen05.mma / commit),
// warp 9 producer (streams the pre-split weight units L2 -> shared ring with cp.async.bulk).  Activations chain through ONE
// 128 x 64 operand image A: each GEMM's epilogue overwrites A with the next GEMM's input after the commit of the MMA that read it.
// The FFN hidden layer never leaves the SM: 16 chunks of 64 hidden units, layer 1 of chunk c+1 issued before layer 2 of chunk c
// (D1 and the staged chunk operand H are double buffered).
#include "kernels.cuh"
#include "tc.cuh"
#include <cstdio>

namespace sttode {

// ---- packed weights of one ODE block -------------------------------------------------------------------------------------------
// unit u = fp16 hi (8 KB) | lo (8 KB) image of a 64 x 64 weight slice (64 B-rows, K = 64), canonical no-swizzle K-major layout:
//   u0 W_v  u1 W_o  u2 W_info  u3 W_gate  u(4+2c) linear1 rows 64c..  u(5+2c) linear2 columns 64c..   (c < 16)
// followed by one 192-row image of in_proj_weight (hi 24 KB | lo 24 KB) for k_qkv_tc.
constexpr int ODE_UNIT = 16384, ODE_UNITS = 36, ODE_QKV_OFF = ODE_UNITS * ODE_UNIT, ODE_QKV_IMG = 192 * 8 * 16;
constexpr size_t ODE_PACK_BYTES = (size_t)ODE_QKV_OFF + 2 * ODE_QKV_IMG;

__global__ void k_pack_ode(sttode_ode_weights w, unsigned char* __restrict__ pk) {
    const int j = blockIdx.y, idx = blockIdx.x * blockDim.x + threadIdx.x;
    const float* W;
    int ldw, row0 = 0, col0 = 0, nrows = 64;
    unsigned char *hi, *lo;
    if (j < ODE_UNITS) {
        hi = pk + (size_t)j * ODE_UNIT; lo = hi + 8192;
        if (j == 0) { W = w.attn.in_proj_w; ldw = 64; row0 = 128; }
        else if (j == 1) { W = w.attn.out_proj_w; ldw = 64; }
        else if (j == 2) { W = w.info_w; ldw = 64; }
        else if (j == 3) { W = w.gate_w; ldw = 64; }
        else if (!(j & 1)) { W = w.lin1_w; ldw = 64; row0 = ((j - 4) >> 1) * 64; }
        else { W = w.lin2_w; ldw = STTODE_FF; col0 = ((j - 5) >> 1) * 64; }
    } else {
        W = w.attn.in_proj_w; ldw = 64; nrows = 192;
        hi = pk + ODE_QKV_OFF; lo = hi + ODE_QKV_IMG;
    }
    if (idx >= 8 * nrows) return;
    const int kc = idx / nrows, r = idx - kc * nrows;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(W + (size_t)(row0 + r) * ldw + col0 + kc * 8 + i);
    tc::store_split8(hi, lo, nrows, r, kc, v);
}
size_t ode_pack_bytes() { return ODE_PACK_BYTES; }
void launch_pack_ode(const sttode_ode_weights& w, unsigned char* pk, cudaStream_t st) { k_pack_ode<<<dim3(12, ODE_UNITS + 1), 128, 0, st>>>(w, pk); }

// ---- folded encoder front end ------------------------------------------------------------------------------------------------------
// In eval mode (PE dropout off) PastEncoder.forward up to the ODE block (model/STTODE.py:214-223) is a chain of Linear layers with no
// nonlinearity in between — input_fc (4 -> 64), pos_encoder.fc on [. | pe_t] (128 -> 64), input_fc2 over the flattened time axis
// (64 T -> 64), input_fc3 on [. | category] (67 -> 64) — i.e. ONE affine map of the 4 T input features:
//     x0 = sum_t G_t in_t + g0 + flag * w3f ,   G_t = W3a W2_t Wpe1 Wfc  [64 x 4],
//     g0 = W3a (sum_t W2_t (Wpe1 b_fc + Wpe2 pe_t + b_pe) + b_2) + b_3 ,   w3f = W3[:, 66].
// k_fold_prologue builds the table [(4 T + 2), 64] = [G (k-major) | g0 | w3f] once per weight version, accumulating in float64 (so the
// folded map is at least as accurate as the reference's fp32 chain); the tensor-core kernels then compute x0 in their input staging
// (1024 FMAs per thread at T = 8) and the raw trajectory is the only HBM input of the encoder.  209 kFLOP/row -> 4 kFLOP/row.
__global__ void __launch_bounds__(320) k_fold_prologue(sttode_encoder_weights w, int T, float* __restrict__ out) {
    __shared__ double P1[64][5], U[64][5], V[64], acc4[64];
    const int tid = threadIdx.x, n = tid / 5, c = tid - n * 5;
    {   // P1[:, 0..3] = Wpe1 Wfc ; P1[:, 4] = Wpe1 b_fc + b_pe
        double a = c == 4 ? (double)w.pefc_b[n] : 0.0;
        for (int k = 0; k < 64; ++k) a += (double)w.pefc_w[n * 128 + k] * (c < 4 ? (double)w.fc_w[k * 4 + c] : (double)w.fc_b[k]);
        P1[n][c] = a;
    }
    if (tid < 64) acc4[tid] = 0.0;
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        if (tid < 64) {
            double a = P1[tid][4];
            for (int k = 0; k < 64; ++k) a += (double)w.pefc_w[tid * 128 + 64 + k] * (double)w.pe[t * 64 + k];
            V[tid] = a;
        }
        __syncthreads();
        {
            double a = 0.0;
            for (int k = 0; k < 64; ++k) a += (double)w.fc2_w[(size_t)n * 64 * T + t * 64 + k] * (c < 4 ? P1[k][c] : V[k]);
            U[n][c] = a;
        }
        __syncthreads();
        if (c < 4) {
            double a = 0.0;
            for (int k = 0; k < 64; ++k) a += (double)w.fc3_w[n * 67 + k] * U[k][c];
            out[(size_t)(t * 4 + c) * 64 + n] = (float)a;
        } else {
            acc4[n] += U[n][4];
        }
        __syncthreads();
    }
    if (tid < 64) {
        double a = (double)w.fc3_b[tid];
        for (int k = 0; k < 64; ++k) a += (double)w.fc3_w[tid * 67 + k] * (acc4[k] + (double)w.fc2_b[k]);
        out[(size_t)(4 * T) * 64 + tid] = (float)a;
        out[(size_t)(4 * T + 1) * 64 + tid] = w.fc3_w[tid * 67 + 66];
    }
}
size_t fold_table_bytes(int T) { return (size_t)(4 * T + 2) * 64 * sizeof(float); }
void launch_fold_prologue(const sttode_encoder_weights& w, int T, float* out, cudaStream_t st) { k_fold_prologue<<<1, 320, 0, st>>>(w, T, out); }

// columns 16*cq .. 16*cq+15 of x0 for row r from the raw trajectory (scene-centred positions + velocities, first step duplicated:
// model/STTODE.py:417,429,582-596) or from explicit [x, y, vx, vy] features; tab_s = the folded table in shared memory
__device__ __forceinline__ void folded_x0(const FoldedIn& f, const float* tab_s, int r, bool live, int cq, bool write_norm, float (&v)[16]) {
    const int T = f.T;
    float flag = 0.f, ox = 0.f, oy = 0.f;
    if (live) {
        if (f.scene_ptr) {
            const int sidx = f.row_scene[r];
            flag = (r == f.scene_ptr[sidx + 1] - 1) ? 1.f : 0.f;
            ox = f.origin[sidx * 2]; oy = f.origin[sidx * 2 + 1];
        } else {
            flag = ((r % f.agents) == f.agents - 1) ? 1.f : 0.f;
        }
    }
    const float* g0 = tab_s + (size_t)(4 * T) * 64 + cq * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(flag, g0[64 + i], g0[i]);
    if (!live) return;
    const float* p = f.traj + (size_t)r * T * (f.layout == STTODE_FEAT_RT4 ? 4 : 2);
    float ppx = 0.f, ppy = 0.f;                     // previous normalised position
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        float in4[4];
        if (f.layout == STTODE_FEAT_RT4) {
            const float4 q = *reinterpret_cast<const float4*>(p + t * 4);
            in4[0] = q.x; in4[1] = q.y; in4[2] = q.z; in4[3] = q.w;
        } else {
            auto pos = [&](int tt, float& x_, float& y_) {
                if (f.layout == STTODE_TRAJ_RT2) { x_ = p[tt * 2] - ox; y_ = p[tt * 2 + 1] - oy; }
                else                             { x_ = p[tt] - ox;     y_ = p[T + tt] - oy; }
            };
            float px, py;
            pos(t, px, py);
            if (t == 0) {                           // velocity of the first step duplicates the second
                float qx, qy;
                pos(1, qx, qy);
                in4[2] = qx - px; in4[3] = qy - py;
            } else {
                in4[2] = px - ppx; in4[3] = py - ppy;
            }
            in4[0] = px; in4[1] = py;
            ppx = px; ppy = py;
            if (write_norm && f.traj_norm) *reinterpret_cast<float2*>(f.traj_norm + ((size_t)r * T + t) * 2) = make_float2(px, py);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4* row = reinterpret_cast<const float4*>(tab_s + (size_t)(t * 4 + c) * 64 + cq * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 m4 = row[q];
                v[4 * q] = fmaf(in4[c], m4.x, v[4 * q]); v[4 * q + 1] = fmaf(in4[c], m4.y, v[4 * q + 1]);
                v[4 * q + 2] = fmaf(in4[c], m4.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(in4[c], m4.w, v[4 * q + 3]);
            }
        }
    }
}
__device__ __forceinline__ void store_row16(float* p, const float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// ---- shared pieces ----------------------------------------------------------------------------------------------------------------
namespace otc {
constexpr int THREADS = 576, EPI_WARPS = 16, EPI_THREADS = 512, ISSUER = 16, PRODUCER = 17;
constexpr int STAGE = 2 * ODE_UNIT, NST = 3, NBLK = ODE_UNITS / 2;          // 18 ring blocks of two units
constexpr int SM_A = 0;                          // A operand image, 128 rows x K 64: hi 16 KB | lo 16 KB
constexpr int SM_H = 32768;                      // 2 x (hi 16 KB | lo 16 KB): staged FFN chunk operand
constexpr int SM_RING = SM_H + 65536;
constexpr int SM_BARS = SM_RING + NST * STAGE;   // 196608
constexpr int SM_RED = SM_BARS + 256;            // LayerNorm partial sums: 2 x [128 rows][4 column quarters] fp32
constexpr int SM_FOLD = SM_RED + 4096;           // folded front-end table [(4T+2), 64] fp32 (optional)
constexpr int COL_D0 = 0, COL_DG = 64, COL_D1 = 192, COL_D2 = 320;
enum { B_WFULL = 0, B_WFREE = 3, B_AFULL = 6, B_DFULL = 7, B_D1FULL = 8, B_HFULL = 10, B_HFREE = 12, B_D2FULL = 14, B_COUNT = 15 };
constexpr float FP16_SAFE = 60000.f;

__device__ __forceinline__ void wait(uint64_t* bars, int bar, uint32_t parity) {
    if (!tc::mbar_wait(bars + bar, parity)) {      // a protocol bug must fail loudly, never hang the GPU
        printf("k_ode_tail_tc: mbarrier %d timed out (block %d thread %d)\n", bar, blockIdx.x, threadIdx.x);
        __trap();
    }
}
__device__ __forceinline__ void arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 epilogue warps only
// epilogue warp: publish generic-proxy smem writes + finished TMEM reads, one arrival per warp
__device__ __forceinline__ void signal(uint64_t* bars, int bar) {
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) arrive(bars + bar);
}
template <int N>
__device__ __forceinline__ void track(const float (&v)[N], int& ovf) {
    // integer max over |v| bit patterns: monotonic for finite values, and NaN patterns sort above inf (fmaxf would silently drop a NaN)
    uint32_t mx = 0u;
#pragma unroll
    for (int i = 0; i < N; ++i) mx = max(mx, __float_as_uint(v[i]) & 0x7fffffffu);
    ovf |= mx > __float_as_uint(FP16_SAFE);
}
// 16 consecutive-k values of row m (k chunks kc0, kc0+1) -> the 128-row operand image at `img` (hi) / img + 16384 (lo)
__device__ __forceinline__ void stage16(unsigned char* img, int m, int kc0, const float (&v)[16], int& ovf) {
    track(v, ovf);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = v[q * 8 + i];
        tc::store_split8(img, img + 16384, 128, m, kc0 + q, t8);
    }
}
__device__ __forceinline__ void load_row16(float (&v)[16], const float* __restrict__ p) {      // 64 contiguous bytes through L2; p == nullptr -> zeros
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 r = p ? __ldcg(reinterpret_cast<const float4*>(p) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * q] = r.x; v[4 * q + 1] = r.y; v[4 * q + 2] = r.z; v[4 * q + 3] = r.w;
    }
}
__device__ __forceinline__ void add_bias16(float (&v)[16], const float* __restrict__ b) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(b) + q);
        v[4 * q] += r.x; v[4 * q + 1] += r.y; v[4 * q + 2] += r.z; v[4 * q + 3] += r.w;
    }
}
// LayerNorm over the 64 features of a row held 16 per thread by the four threads (m, cq = 0..3) — they sit in different warps, so the two
// row statistics travel through shared memory (biased variance about the mean, eps inside the sqrt: torch native_layer_norm).
__device__ __forceinline__ void layernorm_rows16(float (&s)[16], float* red, int m, int cq, const float* __restrict__ gamma, const float* __restrict__ beta) {
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) part += s[i];
    red[m * 4 + cq] = part;
    epi_barrier();
    const float4 a = *reinterpret_cast<const float4*>(red + m * 4);
    const float mean = (a.x + a.y + a.z + a.w) * (1.f / 64.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s[i] -= mean; var = fmaf(s[i], s[i], var); }
    red[512 + m * 4 + cq] = var;
    epi_barrier();
    const float4 b = *reinterpret_cast<const float4*>(red + 512 + m * 4);
    const float rstd = 1.f / sqrtf((b.x + b.y + b.z + b.w) * (1.f / 64.f) + 1e-5f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + cq * 16) + q), b4 = __ldg(reinterpret_cast<const float4*>(beta + cq * 16) + q);
        s[4 * q] = s[4 * q] * rstd * g4.x + b4.x; s[4 * q + 1] = s[4 * q + 1] * rstd * g4.y + b4.y;
        s[4 * q + 2] = s[4 * q + 2] * rstd * g4.z + b4.z; s[4 * q + 3] = s[4 * q + 3] * rstd * g4.w + b4.w;
    }
}
}  // namespace otc

// =====================================================================================================================================
// ODE right-hand side tail + solver update for a 128-row tile (see k_ode_tail for the algebra).  attn_o == nullptr: attention axis of
// length 1 (softmax over one key is 1), the attention output is the value projection of x itself.
// Epilogue thread (m = 32*(warp & 3) + lane, cq = warp >> 2) owns columns 16*cq .. 16*cq+15 of row m: four warps per scheduler hide the
// TMEM / shared-memory latencies of one another (two per scheduler left the tensor pipe 11 % busy).
// =====================================================================================================================================
__global__ void __launch_bounds__(otc::THREADS, 1)
k_ode_tail_tc(const float* __restrict__ x, int ldx, const float* __restrict__ attn_o, int ldo, const float* __restrict__ base, int ldb, float coef,
              int relu, float* __restrict__ out, int ldout, int rows, sttode_ode_weights w, const unsigned char* __restrict__ packed,
              int* __restrict__ tile_flags, const FoldedIn fold) {
    using namespace otc;
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    float* red = reinterpret_cast<float*>(sm + SM_RED);
    float* tab_s = reinterpret_cast<float*>(sm + SM_FOLD);
    if (fold.table)                                 // folded front end: x (= fold.x0_out, ld ldx) is produced by this kernel's input staging
        for (int i = threadIdx.x; i < (4 * fold.T + 2) * 16; i += THREADS) reinterpret_cast<float4*>(tab_s)[i] = __ldg(reinterpret_cast<const float4*>(fold.table) + i);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, cq = (warp >> 2) & 3;
    const int m = wq * 32 + lane, r = blockIdx.x * BM + m;
    const bool live = r < rows;
    int ovf = 0;

    if (tid == 0) {
        for (int i = 0; i < B_COUNT; ++i) tc::mbar_init(bars + i, (i == B_AFULL || i == B_HFULL || i == B_HFULL + 1) ? EPI_WARPS : 1);
        tc::fence_mbar_init();
    }
    if (warp == ISSUER) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const bool vproj = attn_o == nullptr;

    if (warp == PRODUCER) {                         // ---------------- producer: 18 blocks of [unit 2b | unit 2b+1] through a 3-stage ring
#pragma unroll 1
        for (int b = 0; b < NBLK; ++b) {
            const int s = b % NST;
            if (b >= NST) wait(bars, B_WFREE + s, (uint32_t)((b / NST - 1) & 1));
            if (elect_one()) {
                tc::mbar_arrive_expect_tx(bars + B_WFULL + s, STAGE);
                tc::bulk_g2s(sm + SM_RING + s * STAGE, packed + (size_t)b * STAGE, STAGE, bars + B_WFULL + s);
            }
            __syncwarp();
        }
    } else if (warp == ISSUER) {                    // ---------------- issuer: one elected lane runs the whole chain
        if (elect_one()) {
            const uint32_t a_addr = tc::smem_u32(sm + SM_A), h_addr = tc::smem_u32(sm + SM_H), ring = tc::smem_u32(sm + SM_RING);
            const uint64_t a_hi = tc::op_desc(a_addr, 128), a_lo = tc::op_desc(a_addr + 16384, 128);
            const uint32_t idesc = tc::make_idesc_f16(64);
            uint32_t n_a = 0;
            auto unit_mma = [&](uint32_t dcol, uint64_t ah, uint64_t al, uint32_t unit_addr, bool acc) {
                tc::mma_split3_d(tmem + dcol, ah, al, 256, tc::op_desc(unit_addr, 64), tc::op_desc(unit_addr + 8192, 64), 128, idesc, 4, acc);
            };
            // block 0: [W_v | W_o]
            wait(bars, B_WFULL + 0, 0);
            if (vproj) {
                wait(bars, B_AFULL, n_a++ & 1); tc::tc_fence_after();            // x staged
                unit_mma(COL_D0, a_hi, a_lo, ring, false);
                tc::mma_commit(bars + B_DFULL);
            }
            wait(bars, B_AFULL, n_a++ & 1); tc::tc_fence_after();                // attention output (or v) staged
            unit_mma(COL_D0, a_hi, a_lo, ring + ODE_UNIT, false);
            tc::mma_commit(bars + B_DFULL);
            tc::mma_commit(bars + B_WFREE + 0);
            // block 1: [W_info | W_gate]
            wait(bars, B_WFULL + 1, 0);
            wait(bars, B_AFULL, n_a++ & 1); tc::tc_fence_after();                // m = out_proj(o) staged
            unit_mma(COL_DG, a_hi, a_lo, ring + STAGE, false);
            unit_mma(COL_DG + 64, a_hi, a_lo, ring + STAGE + ODE_UNIT, false);
            tc::mma_commit(bars + B_DFULL);
            tc::mma_commit(bars + B_WFREE + 1);
            // FFN: chunk c lives in ring block 2 + c = [linear1 rows 64c.. | linear2 columns 64c..]
            wait(bars, B_AFULL, n_a++ & 1); tc::tc_fence_after();                // s = LN1(x + gate) staged
            wait(bars, B_WFULL + 2, 0); tc::tc_fence_after();
            unit_mma(COL_D1, a_hi, a_lo, ring + 2 * STAGE, false);
            tc::mma_commit(bars + B_D1FULL + 0);
#pragma unroll 1
            for (int c = 0; c < 16; ++c) {
                const int b = c & 1, blk = 2 + c, s = blk % NST;
                if (c + 1 < 16) {                   // layer 1 of chunk c+1 -> D1[(c+1)&1] (chunk c-1 was read out before HFULL(c-1), waited last iteration)
                    const int blk1 = blk + 1, s1 = blk1 % NST;
                    wait(bars, B_WFULL + s1, (uint32_t)((blk1 / NST) & 1)); tc::tc_fence_after();
                    unit_mma(COL_D1 + ((c + 1) & 1) * 64, a_hi, a_lo, ring + s1 * STAGE, false);
                    tc::mma_commit(bars + B_D1FULL + ((c + 1) & 1));
                }
                wait(bars, B_HFULL + b, (uint32_t)((c >> 1) & 1)); tc::tc_fence_after();      // relu(D1(c) + b1) staged as the A operand
                unit_mma(COL_D2, tc::op_desc(h_addr + b * 32768, 128), tc::op_desc(h_addr + b * 32768 + 16384, 128), ring + s * STAGE + ODE_UNIT, c > 0);
                tc::mma_commit(bars + B_WFREE + s);
                tc::mma_commit(bars + B_HFREE + b);
            }
            tc::mma_commit(bars + B_D2FULL);
        }
        __syncwarp();
    } else {                                        // ---------------- epilogue warps
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        unsigned char* A = sm + SM_A;
        uint32_t n_d = 0;
        float v[16];
        // input operand: x (value projection follows) or the attention output, columns 16*cq .. of row m
        if (fold.table) {                           // x0 from the raw trajectory; also this kernel's x / base (re-read below through L2)
            folded_x0(fold, tab_s, r, live, cq, cq == 0, v);
            if (live) store_row16(fold.x0_out + (size_t)r * ldx + cq * 16, v);
        }
        if (!fold.table || !vproj) {
            const float* src = vproj ? x + (size_t)r * ldx : attn_o + (size_t)r * ldo;
            load_row16(v, live ? src + cq * 16 : nullptr);
        }
        stage16(A, m, cq * 2, v, ovf);
        signal(bars, B_AFULL);
        if (vproj) {                                // v = x.W_v^T + b_v
            wait(bars, B_DFULL, n_d++ & 1); tc::tc_fence_after();
            tc::tmem_ld16(tmem + lane_base + COL_D0 + cq * 16, v);
            add_bias16(v, w.attn.in_proj_b + 128 + cq * 16);
            stage16(A, m, cq * 2, v, ovf);
            signal(bars, B_AFULL);
        }
        // m = o.W_o^T + b_o
        wait(bars, B_DFULL, n_d++ & 1); tc::tc_fence_after();
        tc::tmem_ld16(tmem + lane_base + COL_D0 + cq * 16, v);
        add_bias16(v, w.attn.out_proj_b + cq * 16);
        stage16(A, m, cq * 2, v, ovf);
        signal(bars, B_AFULL);
        // gate + residual + LN1; s (this thread's 16 features) stays in registers for the second residual
        float s[16];
        wait(bars, B_DFULL, n_d++ & 1); tc::tc_fence_after();
        {
            float ui[16], ug[16];
            tc::tmem_ld16(tmem + lane_base + COL_DG + cq * 16, ui);
            tc::tmem_ld16(tmem + lane_base + COL_DG + 64 + cq * 16, ug);
            add_bias16(ui, w.info_b + cq * 16);
            add_bias16(ug, w.gate_b + cq * 16);
            load_row16(s, live ? x + (size_t)r * ldx + cq * 16 : nullptr);          // through L2: x may have been written by this CTA (folded front end)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] += tanhf(ui[i]) * sigmoidf_(ug[i]);
        }
        layernorm_rows16(s, red, m, cq, w.norm1_g, w.norm1_b);
        stage16(A, m, cq * 2, s, ovf);
        signal(bars, B_AFULL);
        // FFN chunk epilogue: H[c&1] = relu(D1(c) + b1[64c..]), this thread: columns 16*cq .. of row m
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
            const int b = c & 1;
            float4 bq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) bq[q] = __ldg(reinterpret_cast<const float4*>(w.lin1_b + c * 64 + cq * 16) + q);      // before the wait: off the critical path
            wait(bars, B_D1FULL + b, (uint32_t)((c >> 1) & 1)); tc::tc_fence_after();
            tc::tmem_ld16(tmem + lane_base + COL_D1 + b * 64 + cq * 16, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[4 * q] = fmaxf(v[4 * q] + bq[q].x, 0.f); v[4 * q + 1] = fmaxf(v[4 * q + 1] + bq[q].y, 0.f);
                v[4 * q + 2] = fmaxf(v[4 * q + 2] + bq[q].z, 0.f); v[4 * q + 3] = fmaxf(v[4 * q + 3] + bq[q].w, 0.f);
            }
            if (c >= 2) wait(bars, B_HFREE + b, (uint32_t)(((c >> 1) - 1) & 1));       // layer 2 of chunk c-2 has read H[b]
            stage16(sm + SM_H + b * 32768, m, cq * 2, v, ovf);
            signal(bars, B_HFULL + b);
        }
        // second residual + LN2 + solver update: out = act(base + coef * f)
        wait(bars, B_D2FULL, 0); tc::tc_fence_after();
        tc::tmem_ld16(tmem + lane_base + COL_D2 + cq * 16, v);
        add_bias16(v, w.lin2_b + cq * 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] += v[i];
        layernorm_rows16(s, red, m, cq, w.norm2_g, w.norm2_b);
        if (live) {
            if (base) load_row16(v, base + (size_t)r * ldb + cq * 16);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float t = __fmul_rn(coef, s[i]);                                // dt * f, then + y0: two roundings like the reference
                if (base) t = __fadd_rn(v[i], t);
                v[i] = (relu && t < 0.f) ? 0.f : t;                              // torch.relu keeps NaN
            }
            store_row16(out + (size_t)r * ldout + cq * 16, v);
        }
    }
    if (ovf && tile_flags) tile_flags[blockIdx.x] = 1;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ISSUER) tc::tmem_dealloc(tmem, 512);
}
size_t ode_tail_tc_smem_bytes(int fold_T) { return (size_t)otc::SM_FOLD + (fold_T > 0 ? fold_table_bytes(fold_T) : 0); }
int ode_tail_tc_threads() { return otc::THREADS; }

// =====================================================================================================================================
// MHGSA in-projection for a 128-row tile: out[:, 0:192] = [ q^ | k^ | v ] (q scaled by head_dim^-1/2 then both q and k projected onto the
// Oblique manifold per head: p / |p|, no epsilon — a zero row yields NaN exactly as the reference does).  One N = 192 MMA group.
// =====================================================================================================================================
namespace qtc {
constexpr int SM_A = 0, SM_W = 32768, SM_BARS = SM_W + 2 * ODE_QKV_IMG, SM_FOLD = SM_BARS + 64;      // 32 KB + 48 KB (+ folded table)
enum { B_W = 0, B_A = 1, B_D = 2, B_COUNT = 3 };
}
__global__ void __launch_bounds__(otc::THREADS, 1)
k_qkv_tc(const float* __restrict__ x, int ldx, int rows, const float* __restrict__ in_b, float scale, const unsigned char* __restrict__ packed,
         float* __restrict__ out, int ldo_, int* __restrict__ tile_flags, const FoldedIn fold) {
    using namespace otc;
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + qtc::SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + qtc::B_COUNT);
    float* tab_s = reinterpret_cast<float*>(sm + qtc::SM_FOLD);
    if (fold.table)
        for (int i = threadIdx.x; i < (4 * fold.T + 2) * 16; i += THREADS) reinterpret_cast<float4*>(tab_s)[i] = __ldg(reinterpret_cast<const float4*>(fold.table) + i);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3, cq = (warp >> 2) & 3;
    const int m = wq * 32 + lane, r = blockIdx.x * BM + m;
    const bool live = r < rows;
    int ovf = 0;
    if (tid == 0) {
        tc::mbar_init(bars + qtc::B_W, 1);
        tc::mbar_init(bars + qtc::B_A, EPI_WARPS);
        tc::mbar_init(bars + qtc::B_D, 1);
        tc::fence_mbar_init();
    }
    if (warp == ISSUER) tc::tmem_alloc(tmem_slot, 256);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == PRODUCER) {
        if (elect_one()) {
            tc::mbar_arrive_expect_tx(bars + qtc::B_W, 2 * ODE_QKV_IMG);
            tc::bulk_g2s(sm + qtc::SM_W, packed + ODE_QKV_OFF, 2 * ODE_QKV_IMG, bars + qtc::B_W);
        }
        __syncwarp();
    } else if (warp == ISSUER) {
        if (elect_one()) {
            const uint32_t a_addr = tc::smem_u32(sm + qtc::SM_A), w_addr = tc::smem_u32(sm + qtc::SM_W);
            if (!tc::mbar_wait(bars + qtc::B_W, 0) || !tc::mbar_wait(bars + qtc::B_A, 0)) { printf("k_qkv_tc: barrier timed out\n"); __trap(); }
            tc::tc_fence_after();
            tc::mma_split3_d(tmem, tc::op_desc(a_addr, 128), tc::op_desc(a_addr + 16384, 128), 256, tc::op_desc(w_addr, 192),
                             tc::op_desc(w_addr + ODE_QKV_IMG, 192), 384, tc::make_idesc_f16(192), 4, false);
            tc::mma_commit(bars + qtc::B_D);
        }
        __syncwarp();
    } else {
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        float v[16];
        if (fold.table) {
            folded_x0(fold, tab_s, r, live, cq, cq == 0, v);
            if (live) store_row16(fold.x0_out + (size_t)r * ldx + cq * 16, v);
        } else {
            load_row16(v, live ? x + (size_t)r * ldx + cq * 16 : nullptr);
        }
        stage16(sm + qtc::SM_A, m, cq * 2, v, ovf);
        signal(bars, qtc::B_A);
        if (!tc::mbar_wait(bars + qtc::B_D, 0)) { printf("k_qkv_tc: MMA never completed\n"); __trap(); }
        tc::tc_fence_after();
        // twelve 16-column groups (2 heads each): groups 0..3 = q, 4..7 = k, 8..11 = v; the warps of column quarter cq take groups cq, cq+4, cq+8
#pragma unroll 1
        for (int g = cq; g < 12; g += 4) {
            tc::tmem_ld16(tmem + lane_base + g * 16, v);
            add_bias16(v, in_b + g * 16);
            if (g < 4) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= scale;
            }
            if (g < 8) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float n2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) n2 = fmaf(v[h * 8 + i], v[h * 8 + i], n2);
                    const float nrm = sqrtf(n2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[h * 8 + i] = v[h * 8 + i] / nrm;          // oblique.py:16, no epsilon
                }
            }
            if (live) store_row16(out + (size_t)r * ldo_ + g * 16, v);
        }
    }
    if (ovf && tile_flags) tile_flags[blockIdx.x] = 1;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == ISSUER) tc::tmem_dealloc(tmem, 256);
}
size_t qkv_tc_smem_bytes(int fold_T) { return (size_t)qtc::SM_FOLD + (fold_T > 0 ? fold_table_bytes(fold_T) : 0); }

}  // namespace sttode
