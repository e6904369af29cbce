// Past-encoder front end, fused geodesic attention, and the tensor-ODE right-hand side (sm_100a, FP32 FFMA).
//
// Reference semantics (file:line relative to the reference root):
//   k_scene_origin   set_data scene centring                    model/STTODE.py:417,429
//   k_prologue       PastEncoder.forward up to the ODE block     model/STTODE.py:214-223 (+ :582-596 velocities)
//   k_proj           in-projection + q scaling + Oblique.proj    hyptransformerlib.py:115,191,214-218,252-253
//   k_geo_attn       Oblique.dist -> neg -> softmax -> bmm       core/manifolds/oblique.py:36-45; hyptransformerlib.py:254-300
//   k_attn_weights   head-averaged weights                       hyptransformerlib.py:306-309
//   k_ode_tail       out_proj, tanh*sigmoid gate, LN1, FFN, LN2, solver axpy, relu
//                                                                hyptransformerlib.py:304-305; hypertransformer.py:81-83,148-152;
//                                                                ode_demo.py:188,231
#include "kernels.cuh"

namespace sttode {

// ------------------------------------------------------------------------------------------------
// scene origin: mean over the scene's agents of the last observed position.  One warp per scene.
// ------------------------------------------------------------------------------------------------
__global__ void k_scene_origin(const float* __restrict__ traj, int layout, const int* __restrict__ scene_ptr,
                               int n_scenes, int T, float* __restrict__ origin) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_scenes) return;
    const int r0 = scene_ptr[warp], r1 = scene_ptr[warp + 1];
    float sx = 0.f, sy = 0.f;
    for (int r = r0 + lane; r < r1; r += 32) {
        const float* p = traj + (size_t)r * T * 2;
        if (layout == STTODE_TRAJ_RT2) { sx += p[(T - 1) * 2]; sy += p[(T - 1) * 2 + 1]; }
        else                           { sx += p[T - 1];       sy += p[T + T - 1]; }
    }
    sx = warp_sum(sx); sy = warp_sum(sy);
    if (lane == 0) {
        const float n = (float)(r1 - r0);
        origin[warp * 2] = sx / n;
        origin[warp * 2 + 1] = sy / n;
    }
}

// ------------------------------------------------------------------------------------------------
// prologue: raw trajectory -> x0 = input_fc3([input_fc2(flatten_t(pos_fc([input_fc(in_t) | pe_t]))) | category])
// ------------------------------------------------------------------------------------------------
constexpr int LDS64 = 64 + PAD;

__global__ void __launch_bounds__(NTHREADS, 1)
k_prologue(const float* __restrict__ traj, int layout, const int* __restrict__ scene_ptr, const int* __restrict__ row_scene,
           const float* __restrict__ origin, int rows, int agents, int T, sttode_encoder_weights w,
           float* __restrict__ x0, int ldx0, float* __restrict__ traj_norm) {
    extern __shared__ __align__(16) float smem[];
    const int ldin = T * 4 + PAD;
    float* in_s = smem;                       // [BM][T*4+PAD]
    float* a_s = in_s + BM * ldin;            // [BM][68]
    float* e_s = a_s + BM * LDS64;            // [BM][68]
    float* wpe_s = e_s + BM * LDS64;          // [64][68]
    float* wbuf = wpe_s + 64 * LDS64;         // [2][64][68]
    float* c_s = wbuf + 2 * 64 * LDS64;       // [64]
    float* flag_s = c_s + 64;                 // [BM]

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * BM;

    // positions (scene-normalised) and the category flag
    if (layout == STTODE_FEAT_RT4) {       // explicit [x, y, vx, vy] features: copy the tile as it is
        for (int idx = tid; idx < BM * T; idx += NTHREADS) {
            const int m = idx / T, t = idx - m * T, r = row0 + m;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) v = *reinterpret_cast<const float4*>(traj + ((size_t)r * T + t) * 4);
            *reinterpret_cast<float4*>(in_s + m * ldin + t * 4) = v;
        }
    } else
    for (int idx = tid; idx < BM * T; idx += NTHREADS) {
        const int m = idx / T, t = idx - m * T, r = row0 + m;
        float px = 0.f, py = 0.f;
        if (r < rows) {
            const float* p = traj + (size_t)r * T * 2;
            if (layout == STTODE_TRAJ_RT2) { px = p[t * 2]; py = p[t * 2 + 1]; }
            else                           { px = p[t];     py = p[T + t]; }
            if (scene_ptr) {
                const int s = row_scene[r];
                px -= origin[s * 2]; py -= origin[s * 2 + 1];
            }
            if (traj_norm) { traj_norm[((size_t)r * T + t) * 2] = px; traj_norm[((size_t)r * T + t) * 2 + 1] = py; }
        }
        in_s[m * ldin + t * 4] = px;
        in_s[m * ldin + t * 4 + 1] = py;
    }
    if (tid < BM) {
        const int r = row0 + tid;
        float f = 0.f;
        if (r < rows) {
            if (scene_ptr) f = (r == scene_ptr[row_scene[r] + 1] - 1) ? 1.f : 0.f;
            else           f = ((r % agents) == agents - 1) ? 1.f : 0.f;
        }
        flag_s[tid] = f;
    }
    __syncthreads();
    // velocities from the normalised positions, first step duplicated (model/STTODE.py:588-589)
    if (layout != STTODE_FEAT_RT4)
    for (int idx = tid; idx < BM * T; idx += NTHREADS) {
        const int m = idx / T, t = idx - m * T;
        const int t1 = t == 0 ? 1 : t, t0 = t1 - 1;
        in_s[m * ldin + t * 4 + 2] = in_s[m * ldin + t1 * 4] - in_s[m * ldin + t0 * 4];
        in_s[m * ldin + t * 4 + 3] = in_s[m * ldin + t1 * 4 + 1] - in_s[m * ldin + t0 * 4 + 1];
    }
    // pos_encoder.fc[:, 0:64] stays resident for every t
    load_w_chunk_async(wpe_s, LDS64, w.pefc_w, 128, 64, 64);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    int wrow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wrow[j] = j * 16 + tx;
    float accf[8][4];
    zero_acc(accf);
    float fcw[4][4], fcb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        fcb[j] = __ldg(w.fc_b + wrow[j]);
#pragma unroll
        for (int c = 0; c < 4; ++c) fcw[j][c] = __ldg(w.fc_w + wrow[j] * 4 + c);
    }

    for (int t = 0; t < T; ++t) {
        if (tid < 64) {          // row-independent half of pos_encoder.fc: W[:, 64:128] . pe[t] + b
            float c = __ldg(w.pefc_b + tid);
            const float* wr = w.pefc_w + tid * 128 + 64;
            const float* pe = w.pe + t * 64;
            for (int k = 0; k < 64; ++k) c = fmaf(__ldg(pe + k), __ldg(wr + k), c);
            c_s[tid] = c;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = i * 16 + ty;
            const float4 v = *reinterpret_cast<const float4*>(in_s + m * ldin + t * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = fcb[j];
                a = fmaf(v.x, fcw[j][0], a); a = fmaf(v.y, fcw[j][1], a);
                a = fmaf(v.z, fcw[j][2], a); a = fmaf(v.w, fcw[j][3], a);
                a_s[m * LDS64 + wrow[j]] = a;
            }
        }
        __syncthreads();
        float acce[8][4];
        zero_acc(acce);
        mma_tile<4>(acce, a_s, LDS64, wpe_s, LDS64, wrow, 64, ty);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) e_s[(i * 16 + ty) * LDS64 + wrow[j]] = acce[i][j] + c_s[wrow[j]];
        // input_fc2 consumes the flattened [t][64] features: accumulate this t's 64-wide slice
        gemm_step<4, 64>(accf, e_s, LDS64, w.fc2_w + t * 64, 64 * T, 64, wbuf);
    }
    // f = input_fc2 output; x0 = input_fc3([f | 0 0 flag])
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a_s[(i * 16 + ty) * LDS64 + wrow[j]] = accf[i][j] + __ldg(w.fc2_b + wrow[j]);
    for (int idx = tid; idx < 64 * 64; idx += NTHREADS) {       // fc3 rows are 67 floats: not 16 B aligned
        const int n = idx >> 6, k = idx & 63;
        wpe_s[n * LDS64 + k] = __ldg(w.fc3_w + n * 67 + k);
    }
    __syncthreads();
    float acc[8][4];
    zero_acc(acc);
    mma_tile<4>(acc, a_s, LDS64, wpe_s, LDS64, wrow, 64, ty);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = i * 16 + ty, r = row0 + m;
        if (r >= rows) continue;
        const float fl = flag_s[m];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = wrow[j];
            x0[(size_t)r * ldx0 + n] = acc[i][j] + fl * __ldg(w.fc3_w + n * 67 + 66) + __ldg(w.fc3_b + n);
        }
    }
}

size_t prologue_smem_bytes(int T) {
    return sizeof(float) * ((size_t)BM * (T * 4 + PAD) + 2 * BM * LDS64 + 64 * LDS64 + 2 * 64 * LDS64 + 64 + BM);
}

// ------------------------------------------------------------------------------------------------
// projection: out[:, col0 + blk*64 + n] = post(X . W[blk*64+n, :]^T + b), post = scale and/or per-head
// manifold projection (Oblique row-normalise, or expmap0+project onto the Poincare ball).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_proj(const float* __restrict__ X, int ldx, int rows, const float* __restrict__ W, const float* __restrict__ b, int nblk,
       float* __restrict__ out, int ldo, unsigned scale_mask, unsigned norm_mask, float scale, int score_mode, float curv,
       float* __restrict__ norms, int ldn, const int* __restrict__ only_flagged) {
    if (only_flagged && !only_flagged[blockIdx.x]) return;      // redo pass after k_qkv_tc: only tiles whose operands left the fp16 range
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                      // [BM][68]
    float* wbuf = xs + BM * LDS64;         // [2][64][68]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * BM;
    for (int idx = tid; idx < BM * 16; idx += NTHREADS) {
        const int m = idx >> 4, c4 = (idx & 15) << 2, r = row0 + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) v = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + c4);
        *reinterpret_cast<float4*>(xs + m * LDS64 + c4) = v;
    }
    const float sqrt_c = sqrtf(curv);
    for (int blk = 0; blk < nblk; ++blk) {
        float acc[8][4];
        zero_acc(acc);
        gemm_step<4, 64>(acc, xs, LDS64, W + (size_t)blk * 64 * 64, 64, 64, wbuf);
        const bool do_scale = (scale_mask >> blk) & 1u, do_norm = (norm_mask >> blk) & 1u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row0 + i * 16 + ty;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = j * 16 + tx;
                float v = acc[i][j] + __ldg(b + blk * 64 + n);
                if (do_scale) v *= scale;
                if (do_norm) {   // a head is 8 consecutive columns = 8 consecutive lanes
                    float nrm = sqrtf(group8_sum(v * v));
                    if (norms && r < rows && (tx & 7) == 0) norms[(size_t)r * ldn + blk * 8 + (n >> 3)] = nrm;   // saved for the backward
                    if (score_mode == STTODE_SCORE_OBLIQUE_ACOS) {
                        v = v / nrm;                                   // oblique.py:16, no epsilon: 0/0 -> NaN as in the reference
                    } else {                                           // expmap0 (pmath.py:300-304) then project (:98-103)
                        nrm = fmaxf(nrm, 1e-5f);
                        float g = tanhf(clamp_nan(sqrt_c * nrm, -15.f, 15.f)) * v / (sqrt_c * nrm);
                        float gn = fmaxf(sqrtf(group8_sum(g * g)), 1e-5f);
                        const float maxnorm = (1.f - 1e-3f) / sqrt_c;
                        v = gn > maxnorm ? g / gn * maxnorm : g;
                    }
                }
                if (r < rows) out[(size_t)r * ldo + blk * 64 + n] = v;
            }
        }
    }
}

size_t proj_smem_bytes() { return sizeof(float) * ((size_t)BM * LDS64 + 2 * 64 * LDS64); }

// ------------------------------------------------------------------------------------------------
// fused geodesic attention for one (group, head, 128-row tile): Gram -> clamp -> acos -> neg -> softmax -> AV
// rowop_i / colop_j are the already-projected head vectors.  Scores lie in [-pi, 0] (or are <= 0 for the
// Poincare distance), so exp() needs no running-max rescale: softmax_j = exp(s_ij) / sum_j exp(s_ij).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float geodesic_score(const float (&a)[8], float a2, const float4& b0, const float4& b1,
                                                int score_mode, float lo, float hi, float curv, float sqrt_c) {
    float g = a[0] * b0.x;
    g = fmaf(a[1], b0.y, g); g = fmaf(a[2], b0.z, g); g = fmaf(a[3], b0.w, g);
    g = fmaf(a[4], b1.x, g); g = fmaf(a[5], b1.y, g); g = fmaf(a[6], b1.z, g); g = fmaf(a[7], b1.w, g);
    if (score_mode == STTODE_SCORE_OBLIQUE_ACOS) return -acosf(clamp_nan(g, lo, hi));
    // -dist_c(a, b) from the Gram value and the two squared norms (pmath.py:416-427,482-488)
    float b2 = b0.x * b0.x + b0.y * b0.y + b0.z * b0.z + b0.w * b0.w + b1.x * b1.x + b1.y * b1.y + b1.z * b1.z + b1.w * b1.w;
    float A = 1.f - 2.f * curv * g + curv * b2, B = 1.f - curv * a2;
    float n2 = fmaxf(A * A * a2 - 2.f * A * B * g + B * B * b2, 0.f);
    float den = 1.f - 2.f * curv * g + curv * curv * a2 * b2 + 1e-5f;
    float u = clamp_nan(sqrt_c * sqrtf(n2) / den, -1.f + 1e-5f, 1.f - 1e-5f);
    return -(2.f / sqrt_c) * 0.5f * (logf(1.f + u) - logf(1.f - u));
}

__global__ void __launch_bounds__(128)
k_geo_attn(const float* __restrict__ rowop, int ldr, const float* __restrict__ colop, int ldc, const float* __restrict__ vals,
           int ldv, float* __restrict__ out, int ldo, float* __restrict__ denom, int Lr, int Lc, int bsz,
           int score_mode, float lo, float hi, float curv) {
    __shared__ __align__(16) float cs[128][8];
    __shared__ __align__(16) float vs[128][8];
    const int head = blockIdx.x & 7, g = blockIdx.x >> 3;
    const int kb = g / bsz, n = g - kb * bsz;
    const int i = blockIdx.y * 128 + threadIdx.x;
    const size_t rbase = (size_t)kb * Lr * bsz + n, cbase = (size_t)kb * Lc * bsz + n;
    const float sqrt_c = sqrtf(curv);
    float a[8], o[8], a2 = 0.f, den = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) { a[d] = 0.f; o[d] = 0.f; }
    if (i < Lr) {
        const float* p = rowop + (rbase + (size_t)i * bsz) * ldr + head * 8;
        const float4 p0 = *reinterpret_cast<const float4*>(p), p1 = *reinterpret_cast<const float4*>(p + 4);
        a[0] = p0.x; a[1] = p0.y; a[2] = p0.z; a[3] = p0.w; a[4] = p1.x; a[5] = p1.y; a[6] = p1.z; a[7] = p1.w;
#pragma unroll
        for (int d = 0; d < 8; ++d) a2 = fmaf(a[d], a[d], a2);
    }
    for (int j0 = 0; j0 < Lc; j0 += 128) {
        const int j = j0 + threadIdx.x;
        if (j < Lc) {
            const float* pc = colop + (cbase + (size_t)j * bsz) * ldc + head * 8;
            const float* pv = vals + (cbase + (size_t)j * bsz) * ldv + head * 8;
            *reinterpret_cast<float4*>(&cs[threadIdx.x][0]) = *reinterpret_cast<const float4*>(pc);
            *reinterpret_cast<float4*>(&cs[threadIdx.x][4]) = *reinterpret_cast<const float4*>(pc + 4);
            *reinterpret_cast<float4*>(&vs[threadIdx.x][0]) = *reinterpret_cast<const float4*>(pv);
            *reinterpret_cast<float4*>(&vs[threadIdx.x][4]) = *reinterpret_cast<const float4*>(pv + 4);
        }
        __syncthreads();
        const int jn = min(128, Lc - j0);
        for (int jj = 0; jj < jn; ++jj) {
            const float4 b0 = *reinterpret_cast<const float4*>(&cs[jj][0]), b1 = *reinterpret_cast<const float4*>(&cs[jj][4]);
            const float p = expf(geodesic_score(a, a2, b0, b1, score_mode, lo, hi, curv, sqrt_c));
            const float4 v0 = *reinterpret_cast<const float4*>(&vs[jj][0]), v1 = *reinterpret_cast<const float4*>(&vs[jj][4]);
            den += p;
            o[0] = fmaf(p, v0.x, o[0]); o[1] = fmaf(p, v0.y, o[1]); o[2] = fmaf(p, v0.z, o[2]); o[3] = fmaf(p, v0.w, o[3]);
            o[4] = fmaf(p, v1.x, o[4]); o[5] = fmaf(p, v1.y, o[5]); o[6] = fmaf(p, v1.z, o[6]); o[7] = fmaf(p, v1.w, o[7]);
        }
        __syncthreads();
    }
    if (i < Lr) {
        float* po = out + (rbase + (size_t)i * bsz) * ldo + head * 8;
        *reinterpret_cast<float4*>(po) = make_float4(o[0] / den, o[1] / den, o[2] / den, o[3] / den);
        *reinterpret_cast<float4*>(po + 4) = make_float4(o[4] / den, o[5] / den, o[6] / den, o[7] / den);
        if (denom) denom[((size_t)g * 8 + head) * Lr + i] = den;
    }
}

// head-averaged attention weights [groups, Lr, Lc] (need_weights=True), from the saved denominators.
__global__ void __launch_bounds__(128)
k_attn_weights(const float* __restrict__ rowop, int ldr, const float* __restrict__ colop, int ldc,
               const float* __restrict__ denom, float* __restrict__ wout, int Lr, int Lc, int bsz, int score_mode,
               float lo, float hi, float curv) {
    const int g = blockIdx.x, i = blockIdx.y;
    const int kb = g / bsz, n = g - kb * bsz;
    const size_t rbase = (size_t)kb * Lr * bsz + n, cbase = (size_t)kb * Lc * bsz + n;
    const float sqrt_c = sqrtf(curv);
    __shared__ __align__(16) float arow[64];
    __shared__ float dinv[8];
    if (threadIdx.x < 64) arow[threadIdx.x] = rowop[(rbase + (size_t)i * bsz) * ldr + threadIdx.x];
    if (threadIdx.x < 8) dinv[threadIdx.x] = denom[((size_t)g * 8 + threadIdx.x) * Lr + i];
    __syncthreads();
    for (int j = threadIdx.x; j < Lc; j += 128) {
        const float* pc = colop + (cbase + (size_t)j * bsz) * ldc;
        float sum = 0.f;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            float a[8], a2 = 0.f;
#pragma unroll
            for (int d = 0; d < 8; ++d) { a[d] = arow[h * 8 + d]; a2 = fmaf(a[d], a[d], a2); }
            const float4 b0 = *reinterpret_cast<const float4*>(pc + h * 8), b1 = *reinterpret_cast<const float4*>(pc + h * 8 + 4);
            sum += expf(geodesic_score(a, a2, b0, b1, score_mode, lo, hi, curv, sqrt_c)) / dinv[h];
        }
        wout[((size_t)g * Lr + i) * Lc + j] = sum / 8.f;
    }
}

// ------------------------------------------------------------------------------------------------
// ODE right-hand side tail + solver update for a 128-row tile:
//   o   = attention output (pre out_proj), or x.W_v^T + b_v when the attention axis has length 1
//   m   = o.W_o^T + b_o ;  g = tanh(m.W_i^T + b_i) * sigmoid(m.W_g^T + b_g)
//   s   = LN1(x + g) ;  f = LN2(s + relu(s.W_1^T + b_1).W_2^T + b_2)
//   out = act(base + coef * f)          (Euler step: base = x, coef = h; relu on the last step)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
k_ode_tail(const float* __restrict__ x, int ldx, const float* __restrict__ attn_o, int ldo, const float* __restrict__ base,
           int ldb, float coef, int relu, float* __restrict__ out, int ldout, int rows, sttode_ode_weights w, int gate_only,
           const int* __restrict__ only_flagged) {
    if (only_flagged && !only_flagged[blockIdx.x]) return;      // redo pass after k_ode_tail_tc: only tiles whose operands left the fp16 range
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                       // [BM][68]
    float* As = xs + BM * LDS64;            // [BM][68]
    float* Bs = As + BM * LDS64;            // [BM][68]
    float* wbuf = Bs + BM * LDS64;          // [2][64][68]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * BM;
    int col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) col[j] = j * 16 + tx;

    for (int idx = tid; idx < BM * 16; idx += NTHREADS) {
        const int m = idx >> 4, c4 = (idx & 15) << 2, r = row0 + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = v;
        if (r < rows) {
            v = *reinterpret_cast<const float4*>(x + (size_t)r * ldx + c4);
            if (attn_o) u = *reinterpret_cast<const float4*>(attn_o + (size_t)r * ldo + c4);
        }
        *reinterpret_cast<float4*>(xs + m * LDS64 + c4) = v;
        if (attn_o) *reinterpret_cast<float4*>(As + m * LDS64 + c4) = u;
    }
    float acc[8][4], acc2[8][4];
    if (!attn_o) {      // softmax over one key is 1: the attention output is the value projection itself
        zero_acc(acc);
        gemm_step<4, 64>(acc, xs, LDS64, w.attn.in_proj_w + 128 * 64, 64, 64, wbuf);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) As[(i * 16 + ty) * LDS64 + col[j]] = acc[i][j] + __ldg(w.attn.in_proj_b + 128 + col[j]);
    }
    zero_acc(acc);
    gemm_step<4, 64>(acc, As, LDS64, w.attn.out_proj_w, 64, 64, wbuf);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[(i * 16 + ty) * LDS64 + col[j]] = acc[i][j] + __ldg(w.attn.out_proj_b + col[j]);
    zero_acc(acc);
    zero_acc(acc2);
    gemm_step<4, 64>(acc, Bs, LDS64, w.info_w, 64, 64, wbuf);
    gemm_step<4, 64>(acc2, Bs, LDS64, w.gate_w, 64, 64, wbuf);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = (i * 16 + ty) * LDS64 + col[j];
            const float g = tanhf(acc[i][j] + __ldg(w.info_b + col[j])) * sigmoidf_(acc2[i][j] + __ldg(w.gate_b + col[j]));
            As[o] = xs[o] + g;
            if (gate_only) {                 // Hypattention.forward on its own (hypertransformer.py:81-83)
                const int r = row0 + i * 16 + ty;
                if (r < rows) out[(size_t)r * ldout + col[j]] = g;
            }
        }
    if (gate_only) return;
    __syncthreads();
    layernorm64_rows(As, LDS64, BM, w.norm1_g, w.norm1_b, 1e-5f);
    // FFN 64 -> 1024 -> 64; the hidden layer never leaves the SM: 16 chunks of 64 hidden units
    zero_acc(acc2);
    for (int c = 0; c < STTODE_FF / 64; ++c) {
        zero_acc(acc);
        gemm_step<4, 64>(acc, As, LDS64, w.lin1_w + (size_t)c * 64 * 64, 64, 64, wbuf);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                Bs[(i * 16 + ty) * LDS64 + col[j]] = fmaxf(acc[i][j] + __ldg(w.lin1_b + c * 64 + col[j]), 0.f);
        gemm_step<4, 64>(acc2, Bs, LDS64, w.lin2_w + c * 64, STTODE_FF, 64, wbuf);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = (i * 16 + ty) * LDS64 + col[j];
            Bs[o] = As[o] + (acc2[i][j] + __ldg(w.lin2_b + col[j]));
        }
    __syncthreads();
    layernorm64_rows(Bs, LDS64, BM, w.norm2_g, w.norm2_b, 1e-5f);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = i * 16 + ty, r = row0 + m;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = __fmul_rn(coef, Bs[m * LDS64 + col[j]]);          // dt * f, then + y0: two roundings like the reference
            if (base) v = __fadd_rn(base[(size_t)r * ldb + col[j]], v);
            out[(size_t)r * ldout + col[j]] = (relu && v < 0.f) ? 0.f : v;          // torch.relu keeps NaN (fmaxf would drop it)
        }
    }
}

size_t ode_tail_smem_bytes() { return sizeof(float) * ((size_t)3 * BM * LDS64 + 2 * 64 * LDS64); }

// y[rows,N] = act(x[rows,K] . W[N,K]^T + b): small dense heads (K <= 256, K % 4 == 0, N % 64 == 0)
__global__ void __launch_bounds__(NTHREADS, 1)
k_linear(const float* __restrict__ X, int rows, int K, const float* __restrict__ W, const float* __restrict__ b, int N, int relu,
         float* __restrict__ Y) {
    extern __shared__ __align__(16) float smem[];
    const int lda = K + PAD;
    float* xs = smem;                      // [BM][K+4]
    float* wbuf = xs + BM * lda;           // [2][64][36]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * BM, kv = K >> 2;
    for (int idx = tid; idx < BM * kv; idx += NTHREADS) {
        const int m = idx / kv, c4 = (idx - m * kv) << 2, r = row0 + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) v = *reinterpret_cast<const float4*>(X + (size_t)r * K + c4);
        *reinterpret_cast<float4*>(xs + m * lda + c4) = v;
    }
    for (int nb = 0; nb < N / 64; ++nb) {
        float acc[8][4];
        zero_acc(acc);
        gemm_step<4, 32>(acc, xs, lda, W + (size_t)nb * 64 * K, K, K, wbuf);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row0 + i * 16 + ty;
            if (r >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = nb * 64 + j * 16 + tx;
                const float v = acc[i][j] + __ldg(b + n);
                Y[(size_t)r * N + n] = relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
}
size_t linear_smem_bytes(int K) { return sizeof(float) * ((size_t)BM * (K + PAD) + 2 * 64 * (32 + PAD)); }

// out = base + sum_i coef[i] * k_i  (explicit Runge-Kutta stage combinations), optional relu.
__global__ void k_lincomb(const float* __restrict__ base, int ldb, const float* __restrict__ k0, const float* __restrict__ k1,
                          const float* __restrict__ k2, const float* __restrict__ k3, float c0, float c1, float c2, float c3,
                          int relu, float* __restrict__ out, int ldout, int rows) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * 64) return;
    const int r = (int)(idx >> 6), n = (int)(idx & 63);
    float v = base[(size_t)r * ldb + n];
    if (k0) v += c0 * k0[idx];
    if (k1) v += c1 * k1[idx];
    if (k2) v += c2 * k2[idx];
    if (k3) v += c3 * k3[idx];
    out[(size_t)r * ldout + n] = (relu && v < 0.f) ? 0.f : v;
}

}  // namespace sttode
