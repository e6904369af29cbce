// Backward of the MHGSA + tensor-ODE right-hand side (training config), sm_100a, FP32.
//
// What is differentiated (file:line relative to the reference root, forward semantics in encoder.cu):
//   k_geo_attn_bwd   Oblique.proj / Oblique.dist / clamp / acos / softmax / bmm     core/manifolds/oblique.py:15-16,36-45;
//                                                                                  hyptransformerlib.py:252-300
//   k_gate_ln_fwd    tanh(info) * sigmoid(gate), residual, LayerNorm (recompute)   hypertransformer.py:81-83,148
//   k_gate_bwd       d tanh * sigmoid                                              hypertransformer.py:81-83
//   k_ln_bwd         LayerNorm(64) input gradient (+ x^ * dy for the gamma sum)    hypertransformer.py:148,152
//   k_gemm           every dense product of the chain (y = x W^T, dx = dy W, dW += dy^T x)
//   k_colsum         bias / LayerNorm parameter gradients (deterministic column sums)
// The reference gets these from torch autograd; the chain below is the hand derivation of the same graph.
// The training shapes are small (NBA: 32 x 11 rows, ETH: one scene), so this round's kernels favour a
// deterministic, launch-by-launch decomposition over fusion; the forward they re-materialise is bit-identical
// in structure to k_proj / k_geo_attn and within rounding of k_ode_tail.
#include "kernels.cuh"

namespace sttode {

// ------------------------------------------------------------------------------------------------
// C[M,N] = epi(sum_k A(m,k) B(k,n)),  A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]
// epi(v) = ((relu? max(v + bias[n], 0) : v + bias[n]) * (mask[m,n] > 0 ? 1 : 0 if mask)) + res[m,n] (+ C[m,n] if accumulate)
// ------------------------------------------------------------------------------------------------
constexpr int GTM = 128, GTN = 64, GK = 16;     // 128 x 64 output tile per CTA, 256 threads x (8 x 4) register tile, K in slabs of 16

__global__ void __launch_bounds__(256)
k_gemm(int M, int N, int K, const float* __restrict__ A, long sam, long sak, const float* __restrict__ B, long sbk, long sbn,
       float* __restrict__ C, long ldc, const float* __restrict__ bias, const float* __restrict__ res, long ldres,
       const float* __restrict__ mask, long ldmask, int relu, int accumulate, int kchunk, long slice_stride) {
    __shared__ __align__(16) float As[GK][GTM + 4];
    __shared__ __align__(16) float Bs[GK][GTN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GTM, n0 = blockIdx.x * GTN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_kfast = sak == 1, b_nfast = sbn == 1;
    // split-K: slice blockIdx.z owns k in [z * kchunk, (z+1) * kchunk) and writes its partial product to C + z * slice_stride
    const int kbeg = blockIdx.z * kchunk;
    K = min(K, kbeg + kchunk);
    C += (size_t)blockIdx.z * slice_stride;
    // element e of this thread's share of a slab: coordinates chosen so that consecutive threads walk the contiguous axis
    float ra[8], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * 256;
            int mm, kk;
            if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 127; kk = idx >> 7; }
            ra[e] = (m0 + mm < M && k0 + kk < K) ? A[(long)(m0 + mm) * sam + (long)(k0 + kk) * sak] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            int nn, kb;
            if (b_nfast) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
            rb[e] = (n0 + nn < N && k0 + kb < K) ? B[(long)(k0 + kb) * sbk + (long)(n0 + nn) * sbn] : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + e * 256;
            int mm, kk;
            if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 127; kk = idx >> 7; }
            As[kk][mm] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            int nn, kb;
            if (b_nfast) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
            Bs[kb][nn] = rb[e];
        }
    };
    if (kbeg < K) fetch(kbeg);
    for (int k0 = kbeg; k0 < K; k0 += GK) {
        stash();
        __syncthreads();
        if (k0 + GK < K) fetch(k0 + GK);            // global loads of the next slab overlap the FMAs of this one
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += __ldg(bias + n);
            if (relu) v = fmaxf(v, 0.f);
            if (mask && !(mask[(long)m * ldmask + n] > 0.f)) v = 0.f;
            if (res) v += res[(long)m * ldres + n];
            if (accumulate) v += C[(long)m * ldc + n];
            C[(long)m * ldc + n] = v;
        }
    }
}

// C[m,n] = epi(sum_z part[z][m][n]), z ascending, with k_gemm's epilogue (bias, ReLU, mask, residual, accumulate): second pass of a split-K product
// whose output is too small to fill the GPU in one pass (a few hundred rows x K = 1024 in the ODE-block backward)
__global__ void k_splitk_epilogue(const float* __restrict__ part, int slices, int M, int N, float* __restrict__ C, long ldc,
                                  const float* __restrict__ bias, const float* __restrict__ res, long ldres, const float* __restrict__ mask,
                                  long ldmask, int relu, int accumulate) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (long)m * N);
    float v = 0.f;
    for (int z = 0; z < slices; ++z) v += part[(long)z * M * N + i];
    if (bias) v += __ldg(bias + n);
    if (relu) v = fmaxf(v, 0.f);
    if (mask && !(mask[(long)m * ldmask + n] > 0.f)) v = 0.f;
    if (res) v += res[(long)m * ldres + n];
    if (accumulate) v += C[(long)m * ldc + n];
    C[(long)m * ldc + n] = v;
}

// out[i] += sum_s part[s * stride + i], s ascending (fixed order: the split-K weight gradients stay deterministic)
__global__ void k_sum_slices(const float* __restrict__ part, int slices, long stride, long count, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int z = 0; z < slices; ++z) s += part[(long)z * stride + i];
    out[i] += s;
}

// part[z * N + n] = sum over rows [z * rchunk, (z+1) * rchunk) of X[r, n]   (tall inputs: many CTAs, summed later in slice order)
__global__ void __launch_bounds__(256)
k_colsum_part(const float* __restrict__ X, long ldx, int rows, int N, int rchunk, float* __restrict__ part) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rchunk, r1 = min(rows, r0 + rchunk);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (n < N) {
        int r = r0 + ty;
        for (; r + 24 < r1; r += 32) {              // four independent loads in flight per thread
            s0 += X[(long)r * ldx + n];
            s1 += X[(long)(r + 8) * ldx + n];
            s2 += X[(long)(r + 16) * ldx + n];
            s3 += X[(long)(r + 24) * ldx + n];
        }
        for (; r < r1; r += 8) s0 += X[(long)r * ldx + n];
    }
    sh[ty][tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][tx];
        part[(long)blockIdx.y * N + n] = t;
    }
}

// out[n] += sum_r X[r, n]     (one CTA owns 32 columns: no atomics, fixed summation order)
__global__ void __launch_bounds__(256)
k_colsum(const float* __restrict__ X, long ldx, int rows, int N, float* __restrict__ out) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (n < N)
        for (int r = ty; r < rows; r += 8) s += X[(long)r * ldx + n];
    part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][tx];
        out[n] += t;
    }
}

// z1 = x + tanh(ui) * sigmoid(ug);  s1 = LayerNorm(z1).  One warp per row, two features per lane.
__global__ void __launch_bounds__(256)
k_gate_ln_fwd(const float* __restrict__ x, const float* __restrict__ ui, const float* __restrict__ ug,
              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ z1, float* __restrict__ s1,
              int rows) {
    const int r = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    const size_t o = (size_t)r * 64;
    const float v0 = x[o + lane] + tanhf(ui[o + lane]) * sigmoidf_(ug[o + lane]);
    const float v1 = x[o + lane + 32] + tanhf(ui[o + lane + 32]) * sigmoidf_(ug[o + lane + 32]);
    const float mean = warp_sum(v0 + v1) * (1.f / 64.f);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float rstd = 1.f / sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
    z1[o + lane] = v0;
    z1[o + lane + 32] = v1;
    s1[o + lane] = d0 * rstd * __ldg(gamma + lane) + __ldg(beta + lane);
    s1[o + lane + 32] = d1 * rstd * __ldg(gamma + lane + 32) + __ldg(beta + lane + 32);
}

// LayerNorm(64) backward for the input: z = pre-norm rows, ds = gradient of the normalised output.
//   dz = rstd * (g - mean(g) - x^ * mean(g * x^)),  g = ds * gamma;   xhat_ds = x^ * ds  (column-summed into d gamma)
__global__ void __launch_bounds__(256)
k_ln_bwd(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ ds, long ldds,
         float* __restrict__ dz, float* __restrict__ xhat_ds, int rows) {
    const int r = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    const size_t o = (size_t)r * 64;
    const float v0 = z[o + lane], v1 = z[o + lane + 32];
    const float mean = warp_sum(v0 + v1) * (1.f / 64.f);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float rstd = 1.f / sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
    const float h0 = d0 * rstd, h1 = d1 * rstd;
    const float e0 = ds[(long)r * ldds + lane], e1 = ds[(long)r * ldds + lane + 32];
    const float g0 = e0 * __ldg(gamma + lane), g1 = e1 * __ldg(gamma + lane + 32);
    const float m1 = warp_sum(g0 + g1) * (1.f / 64.f);
    const float m2 = warp_sum(g0 * h0 + g1 * h1) * (1.f / 64.f);
    dz[o + lane] = rstd * (g0 - m1 - h0 * m2);
    dz[o + lane + 32] = rstd * (g1 - m1 - h1 * m2);
    if (xhat_ds) {
        xhat_ds[o + lane] = h0 * e0;
        xhat_ds[o + lane + 32] = h1 * e1;
    }
}

// g = tanh(ui) * sigmoid(ug):  dui = dg * s * (1 - t^2),  dug = dg * t * s * (1 - s)
__global__ void k_gate_bwd(const float* __restrict__ ui, const float* __restrict__ ug, const float* __restrict__ dg, long lddg,
                           float* __restrict__ dui, float* __restrict__ dug, int rows) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * 64) return;
    const int r = (int)(idx >> 6), n = (int)(idx & 63);
    const float t = tanhf(ui[idx]), s = sigmoidf_(ug[idx]), d = dg[(long)r * lddg + n];
    dui[idx] = d * s * (1.f - t * t);
    dug[idx] = d * t * s * (1.f - s);
}

// out[r, 0:64] = alpha * a[r] * (gate[r] > 0 if gate) + (b ? b[r] : 0)       (row strides in floats)
__global__ void k_rows_axpb(const float* __restrict__ a, long lda, float alpha, const float* __restrict__ gate, long ldg,
                            const float* __restrict__ b, long ldb, float* __restrict__ out, long ldo, int rows) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * 64) return;
    const long r = (long)(idx >> 6), n = (long)(idx & 63);
    float v = alpha * a[r * lda + n];
    if (gate && !(gate[r * ldg + n] > 0.f)) v = 0.f;
    if (b) v += b[r * ldb + n];
    out[r * ldo + n] = v;
}

// ------------------------------------------------------------------------------------------------
// Geodesic attention backward for one (group, head).  Forward (k_geo_attn):
//   s_ij = -acos(clamp(<row_i, col_j>, lo, hi)),  P_ij = softmax_j(s_ij),  o_i = sum_j P_ij v_j
// with row = k^, col = q^ when L == S (the reference's un-transposed matrix) and row = q^, col = k^ otherwise.
// Given d o_i:   delta_i = <do_i, o_i>,  dS_ij = P_ij (<do_i, v_j> - delta_i),
//   dG_ij = dS_ij / sqrt(1 - g_ij^2) inside the clamp range (torch.clamp passes the gradient on [lo, hi]), 0 outside,
//   d row_i = sum_j dG_ij col_j,  d col_j = sum_i dG_ij row_i,  d v_j = sum_i P_ij do_i,
// then through p^ = p / |p| (oblique.py:16):  dp = (dp^ - p^ <p^, dp^>) / |p|   (times the q scaling for q).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8(float (&d)[8], const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&d)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(d[4], d[5], d[6], d[7]);
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
    float g = a[0] * b[0];
#pragma unroll
    for (int d = 1; d < 8; ++d) g = fmaf(a[d], b[d], g);
    return g;
}
// gradient w.r.t. the un-normalised head vector from the gradient w.r.t. the unit vector
__device__ __forceinline__ void unnormalise_grad(float (&g)[8], const float (&unit)[8], float norm, float scale) {
    const float t = dot8(unit, g), inv = scale / norm;
#pragma unroll
    for (int d = 0; d < 8; ++d) g[d] = (g[d] - unit[d] * t) * inv;
}

__global__ void __launch_bounds__(128)
k_geo_attn_bwd(const float* __restrict__ qh, int ldq, const float* __restrict__ kh, int ldk, const float* __restrict__ v, int ldv,
               const float* __restrict__ nq, const float* __restrict__ nk, int ldn, const float* __restrict__ dout, int ldo,
               float* __restrict__ dq, int lddq, float* __restrict__ dk, int lddk, float* __restrict__ dv, int lddv,
               int L, int S, int bsz, int swap, float lo, float hi, float qscale) {
    extern __shared__ __align__(16) float sm[];
    const int Lr = L, Lc = S;
    float* rowv = sm;                    // [Lr][8]
    float* dov = rowv + Lr * 8;          // [Lr][8]
    float* colv = dov + Lr * 8;          // [Lc][8]
    float* vv = colv + Lc * 8;           // [Lc][8]
    float* den = vv + Lc * 8;            // [Lr]
    float* delta = den + Lr;             // [Lr]
    const int head = blockIdx.x & 7, g = blockIdx.x >> 3;
    const int kb = g / bsz, n = g - kb * bsz;
    const size_t qbase = (size_t)kb * L * bsz + n, kbase = (size_t)kb * S * bsz + n;
    const float* rowsrc = swap ? kh : qh;  const int ldr = swap ? ldk : ldq;  const size_t rbase = swap ? kbase : qbase;
    const float* colsrc = swap ? qh : kh;  const int ldc = swap ? ldq : ldk;  const size_t cbase = swap ? qbase : kbase;
    for (int i = threadIdx.x; i < Lr; i += 128) {
        float t[8];
        ld8(t, rowsrc + (rbase + (size_t)i * bsz) * ldr + head * 8);  st8(rowv + i * 8, t);
        ld8(t, dout + (qbase + (size_t)i * bsz) * ldo + head * 8);    st8(dov + i * 8, t);
    }
    for (int j = threadIdx.x; j < Lc; j += 128) {
        float t[8];
        ld8(t, colsrc + (cbase + (size_t)j * bsz) * ldc + head * 8);  st8(colv + j * 8, t);
        ld8(t, v + (kbase + (size_t)j * bsz) * ldv + head * 8);       st8(vv + j * 8, t);
    }
    __syncthreads();
    // ---- row pass: softmax statistics, delta, d row
    for (int i = threadIdx.x; i < Lr; i += 128) {
        float a[8], d_o[8], o[8], dr[8];
        ld8(a, rowv + i * 8);
        ld8(d_o, dov + i * 8);
        float dn = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) { o[d] = 0.f; dr[d] = 0.f; }
        for (int j = 0; j < Lc; ++j) {
            float c[8], w[8];
            ld8(c, colv + j * 8);
            ld8(w, vv + j * 8);
            const float p = expf(-acosf(clamp_nan(dot8(a, c), lo, hi)));
            dn += p;
#pragma unroll
            for (int d = 0; d < 8; ++d) o[d] = fmaf(p, w[d], o[d]);
        }
        const float dl = dot8(d_o, o) / dn;
        for (int j = 0; j < Lc; ++j) {
            float c[8], w[8];
            ld8(c, colv + j * 8);
            ld8(w, vv + j * 8);
            const float gij = dot8(a, c);
            if (gij >= lo && gij <= hi) {
                const float p = expf(-acosf(gij)) / dn;
                const float dg = p * (dot8(d_o, w) - dl) / sqrtf(1.f - gij * gij);
#pragma unroll
                for (int d = 0; d < 8; ++d) dr[d] = fmaf(dg, c[d], dr[d]);
            }
        }
        den[i] = dn;
        delta[i] = dl;
        const size_t r = rbase + (size_t)i * bsz;
        unnormalise_grad(dr, a, (swap ? nk : nq)[r * ldn + head], swap ? 1.f : qscale);
        st8((swap ? dk : dq) + r * (swap ? lddk : lddq) + head * 8, dr);
    }
    __syncthreads();
    // ---- column pass: d col, d v
    for (int j = threadIdx.x; j < Lc; j += 128) {
        float c[8], w[8], dc[8], dw[8];
        ld8(c, colv + j * 8);
        ld8(w, vv + j * 8);
#pragma unroll
        for (int d = 0; d < 8; ++d) { dc[d] = 0.f; dw[d] = 0.f; }
        for (int i = 0; i < Lr; ++i) {
            float a[8], d_o[8];
            ld8(a, rowv + i * 8);
            ld8(d_o, dov + i * 8);
            const float gij = dot8(a, c);
            const float p = expf(-acosf(clamp_nan(gij, lo, hi))) / den[i];
#pragma unroll
            for (int d = 0; d < 8; ++d) dw[d] = fmaf(p, d_o[d], dw[d]);
            if (gij >= lo && gij <= hi) {
                const float dg = p * (dot8(d_o, w) - delta[i]) / sqrtf(1.f - gij * gij);
#pragma unroll
                for (int d = 0; d < 8; ++d) dc[d] = fmaf(dg, a[d], dc[d]);
            }
        }
        const size_t rk = kbase + (size_t)j * bsz;
        st8(dv + rk * lddv + head * 8, dw);
        const size_t r = cbase + (size_t)j * bsz;
        unnormalise_grad(dc, c, (swap ? nq : nk)[r * ldn + head], swap ? qscale : 1.f);
        st8((swap ? dq : dk) + r * (swap ? lddq : lddk) + head * 8, dc);
    }
}

// ------------------------------------------------------------------------------------------------
// Decoder backward helpers (DecomposeBlock: Conv1d(2->32,k3,pad1)+ReLU -> GRU(32->96) -> two 3-layer MLPs; model/STTODE.py:51-77).
// ------------------------------------------------------------------------------------------------
// col[(p*T + t), c*3 + k] = x[p, t + k - 1, c] (zero outside [0, T)), x = x_true - x_hat (x_hat nullable): the conv as a GEMM with
// conv_past.weight viewed as [32, 6] (c-major, k-minor = its memory order).  Also writes xin = x_true - x_hat when xin != null.
__global__ void k_conv_im2col(const float* __restrict__ x_true, const float* __restrict__ x_hat, int P, int T, float* __restrict__ col) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P * T * 6) return;
    const int e = (int)(idx % 6), c = e / 3, k = e - c * 3;
    const size_t pt = idx / 6;
    const int t = (int)(pt % T);
    const size_t p = pt / T;
    const int tt = t + k - 1;
    float v = 0.f;
    if (tt >= 0 && tt < T) {
        const size_t o = (p * T + tt) * 2 + c;
        v = x_true[o] - (x_hat ? x_hat[o] : 0.f);
    }
    col[idx] = v;
}
// dxin[p, t, c] = sum_k dcol[(p*T + t - k + 1), c*3 + k]  (transpose of im2col);  out = base (nullable) - dxin
// (block 1 sees x_true - x_hat0, so d x_hat0 = d recon - d xin1)
__global__ void k_conv_col2im_neg(const float* __restrict__ dcol, const float* __restrict__ base, int P, int T, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P * T * 2) return;
    const int c = (int)(idx & 1);
    const size_t pt = idx >> 1;
    const int t = (int)(pt % T);
    const size_t p = pt / T;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ts = t - k + 1;                   // source time whose tap k reads x[t]
        if (ts >= 0 && ts < T) s += dcol[((p * T + ts) * 6) + c * 3 + k];
    }
    out[idx] = (base ? base[idx] : 0.f) - s;
}
// One GRU step (PyTorch gate order r, z, n):  gi = W_ih x_t + b_ih (row stride ldgi), gh = W_hh h + b_hh.
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) n + z h.   Saves r | z | n and gh_n for the backward.
__global__ void k_gru_step_fwd(const float* __restrict__ gi, long ldgi, const float* __restrict__ gh, const float* __restrict__ h, int P,
                               float* __restrict__ rzn, float* __restrict__ ghn, float* __restrict__ hout) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P * 96) return;
    const size_t p = idx / 96;
    const int u = (int)(idx - p * 96);
    const float* gip = gi + p * ldgi;
    const float* ghp = gh + p * 288;
    const float r = sigmoidf_(gip[u] + ghp[u]), z = sigmoidf_(gip[96 + u] + ghp[96 + u]);
    const float gn = ghp[192 + u];
    const float n = tanhf(gip[192 + u] + r * gn);
    rzn[p * 288 + u] = r;
    rzn[p * 288 + 96 + u] = z;
    rzn[p * 288 + 192 + u] = n;
    ghn[idx] = gn;
    hout[idx] = (1.f - z) * n + z * h[idx];
}
// Backward of one step.  dh = gradient of h'.  Writes d gi (row stride lddgi), d gh [P,288] and the direct carry dh * z into dh_prev.
__global__ void k_gru_step_bwd(const float* __restrict__ dh, const float* __restrict__ rzn, const float* __restrict__ ghn,
                               const float* __restrict__ hprev, int P, float* __restrict__ dgi, long lddgi, float* __restrict__ dgh,
                               float* __restrict__ dh_prev) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P * 96) return;
    const size_t p = idx / 96;
    const int u = (int)(idx - p * 96);
    const float r = rzn[p * 288 + u], z = rzn[p * 288 + 96 + u], n = rzn[p * 288 + 192 + u];
    const float d = dh[idx], gn = ghn[idx];
    const float dn = d * (1.f - z) * (1.f - n * n);         // through tanh
    const float dz = d * (hprev[idx] - n) * z * (1.f - z);
    const float dr = dn * gn * r * (1.f - r);
    float* gp = dgi + p * lddgi;
    gp[u] = dr; gp[96 + u] = dz; gp[192 + u] = dn;
    float* hp = dgh + p * 288;
    hp[u] = dr; hp[96 + u] = dz; hp[192 + u] = dn * r;
    dh_prev[idx] = d * z;
}
// ---- GRU through time, one launch per direction ---------------------------------------------------------------------------------
// The recurrence is row-local (h_t of a row depends only on that row's h_{t-1}), so a CTA owns GRU_R rows for ALL T steps with W_hh
// resident in shared memory: one launch replaces the T x (GEMM + gate kernel) launches of the step-by-step form.
// Thread (u = tid % 96, rg = tid / 96) owns hidden unit u (forward) / state column k = u (backward) of rows rg*8 .. rg*8+7.
constexpr int GRU_R = 32, GRU_THREADS = 384, GRU_LDH = 100, GRU_LDG = 292;

// forward: gi [P, T, 288] (row stride ldgi, input projections incl. b_ih), h [T+1, P, 96] (h[0] given), saves rzn [T, P, 288], ghn [T, P, 96]
__global__ void __launch_bounds__(GRU_THREADS, 1)
k_gru_seq_fwd(const float* __restrict__ gi, long ldgi, const float* __restrict__ w_hh, const float* __restrict__ b_hh, int P, int T,
              float* __restrict__ h, float* __restrict__ rzn, float* __restrict__ ghn) {
    extern __shared__ __align__(16) float smem[];
    float* Wt = smem;                               // [96][288]: Wt[k][n] = W_hh[n][k]
    float* hs = Wt + 96 * 288;                      // [GRU_R][GRU_LDH]
    const int tid = threadIdx.x, u = tid % 96, rg = tid / 96, row0 = blockIdx.x * GRU_R;
    for (int i = tid; i < 288 * 96; i += GRU_THREADS) { const int k = i / 288, n = i - k * 288; Wt[i] = __ldg(w_hh + n * 96 + k); }      // conflict-free stores; the strided reads hit L1/L2
    for (int i = tid; i < GRU_R * 96; i += GRU_THREADS) {
        const int rr = i / 96, c = i - rr * 96, p = row0 + rr;
        hs[rr * GRU_LDH + c] = p < P ? h[(size_t)p * 96 + c] : 0.f;
    }
    const float br = __ldg(b_hh + u), bz = __ldg(b_hh + 96 + u), bn = __ldg(b_hh + 192 + u);
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        float acc[8][3], gin[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i) {               // the step's input projections: issued now, consumed after the dot products (latency hidden)
            const int p = row0 + rg * 8 + i;
            const float* gip = gi + (size_t)(p < P ? p : 0) * ldgi + (size_t)t * 288;
            gin[i][0] = p < P ? __ldg(gip + u) : 0.f; gin[i][1] = p < P ? __ldg(gip + 96 + u) : 0.f; gin[i][2] = p < P ? __ldg(gip + 192 + u) : 0.f;
            acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
        }
#pragma unroll 2
        for (int k = 0; k < 96; k += 4) {
            float w[3][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int g = 0; g < 3; ++g) w[g][j] = Wt[(k + j) * 288 + g * 96 + u];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 h4 = *reinterpret_cast<const float4*>(hs + (rg * 8 + i) * GRU_LDH + k);
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    acc[i][g] = fmaf(h4.x, w[g][0], acc[i][g]); acc[i][g] = fmaf(h4.y, w[g][1], acc[i][g]);
                    acc[i][g] = fmaf(h4.z, w[g][2], acc[i][g]); acc[i][g] = fmaf(h4.w, w[g][3], acc[i][g]);
                }
            }
        }
        __syncthreads();                            // every dot product has read the old state
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = rg * 8 + i, p = row0 + rr;
            if (p >= P) continue;
            const float r = sigmoidf_(gin[i][0] + acc[i][0] + br), z = sigmoidf_(gin[i][1] + acc[i][1] + bz);
            const float gn = acc[i][2] + bn;
            const float n = tanhf(gin[i][2] + r * gn);
            const float hn = (1.f - z) * n + z * hs[rr * GRU_LDH + u];
            float* rz = rzn + ((size_t)t * P + p) * 288;
            rz[u] = r; rz[96 + u] = z; rz[192 + u] = n;
            ghn[((size_t)t * P + p) * 96 + u] = gn;
            h[((size_t)(t + 1) * P + p) * 96 + u] = hn;
            hs[rr * GRU_LDH + u] = hn;
        }
        __syncthreads();
    }
}

// backward (BPTT): dh_T [P, 96] (row stride lddh) -> d gi [P, T, 288] (row stride lddgi), d gh [T, P, 288]; dh_0 is discarded (h_0 = 0 is a constant)
__global__ void __launch_bounds__(GRU_THREADS, 1)
k_gru_seq_bwd(const float* __restrict__ dh_T, long lddh, const float* __restrict__ w_hh, const float* __restrict__ rzn, const float* __restrict__ ghn,
              const float* __restrict__ h, int P, int T, float* __restrict__ dgi, long lddgi, float* __restrict__ dgh) {
    extern __shared__ __align__(16) float smem[];
    float* W = smem;                                // [288][96] as stored
    float* gs = W + 288 * 96;                       // [GRU_R][GRU_LDG]: d gh of the current step
    float* ds = gs + GRU_R * GRU_LDG;               // [GRU_R][GRU_LDH]: running dh
    const int tid = threadIdx.x, u = tid % 96, rg = tid / 96, row0 = blockIdx.x * GRU_R;
    for (int i = tid; i < 288 * 96; i += GRU_THREADS) W[i] = __ldg(w_hh + i);
    for (int i = tid; i < GRU_R * 96; i += GRU_THREADS) {
        const int rr = i / 96, c = i - rr * 96, p = row0 + rr;
        ds[rr * GRU_LDH + c] = p < P ? dh_T[(size_t)p * lddh + c] : 0.f;
    }
    __syncthreads();
    float sv[8][5];                                 // saved (r, z, n, W_hn h + b, h_prev) of the step being differentiated, fetched one step ahead
    auto fetch = [&](int t) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = row0 + rg * 8 + i;
            if (p < P) {
                const float* rz = rzn + ((size_t)t * P + p) * 288;
                sv[i][0] = __ldg(rz + u); sv[i][1] = __ldg(rz + 96 + u); sv[i][2] = __ldg(rz + 192 + u);
                sv[i][3] = __ldg(ghn + ((size_t)t * P + p) * 96 + u);
                sv[i][4] = __ldg(h + ((size_t)t * P + p) * 96 + u);
            }
        }
    };
    fetch(T - 1);
    for (int t = T - 1; t >= 0; --t) {
        float carry[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = rg * 8 + i, p = row0 + rr;
            float dr = 0.f, dz = 0.f, dn = 0.f, dnr = 0.f;
            carry[i] = 0.f;
            if (p < P) {
                const float r = sv[i][0], z = sv[i][1], n = sv[i][2];
                const float d = ds[rr * GRU_LDH + u], gn = sv[i][3];
                dn = d * (1.f - z) * (1.f - n * n);         // through tanh
                dz = d * (sv[i][4] - n) * z * (1.f - z);
                dr = dn * gn * r * (1.f - r);
                dnr = dn * r;
                carry[i] = d * z;
                float* gp = dgi + (size_t)p * lddgi + (size_t)t * 288;
                gp[u] = dr; gp[96 + u] = dz; gp[192 + u] = dn;
                float* hp = dgh + ((size_t)t * P + p) * 288;
                hp[u] = dr; hp[96 + u] = dz; hp[192 + u] = dnr;
            }
            gs[rr * GRU_LDG + u] = dr; gs[rr * GRU_LDG + 96 + u] = dz; gs[rr * GRU_LDG + 192 + u] = dnr;
        }
        __syncthreads();
        if (t == 0) break;                          // dh_0 is not needed
        fetch(t - 1);                               // in flight during the matvec below
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 2
        for (int n = 0; n < 288; n += 4) {
            const float w0 = W[n * 96 + u], w1 = W[(n + 1) * 96 + u], w2 = W[(n + 2) * 96 + u], w3 = W[(n + 3) * 96 + u];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 g4 = *reinterpret_cast<const float4*>(gs + (rg * 8 + i) * GRU_LDG + n);
                acc[i] = fmaf(g4.x, w0, acc[i]); acc[i] = fmaf(g4.y, w1, acc[i]); acc[i] = fmaf(g4.z, w2, acc[i]); acc[i] = fmaf(g4.w, w3, acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ds[(rg * 8 + i) * GRU_LDH + u] = carry[i] + acc[i];      // own (row, column) element: no other thread reads ds before the sync
        __syncthreads();
    }
}
size_t gru_seq_fwd_smem_bytes() { return sizeof(float) * ((size_t)96 * 288 + GRU_R * GRU_LDH); }
size_t gru_seq_bwd_smem_bytes() { return sizeof(float) * ((size_t)288 * 96 + GRU_R * GRU_LDG + GRU_R * GRU_LDH); }
int gru_seq_rows() { return GRU_R; }
int gru_seq_threads() { return GRU_THREADS; }

// dst[r, 0:n] = src[r, 0:n]   (row strides in floats)
__global__ void k_rows_copy(const float* __restrict__ src, long lds, float* __restrict__ dst, long ldd, int rows, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * n) return;
    const size_t r = idx / n;
    const int c = (int)(idx - r * n);
    dst[r * ldd + c] = src[r * lds + c];
}

// out[r, 0:n] = a[r, 0:n] + b[r, 0:n]
__global__ void k_rows_add(const float* __restrict__ a, long lda, const float* __restrict__ b, long ldb, float* __restrict__ out, long ldo, int rows,
                           int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * n) return;
    const size_t r = idx / n;
    const int c = (int)(idx - r * n);
    out[r * ldo + c] = a[r * lda + c] + b[r * ldb + c];
}

size_t geo_attn_bwd_smem_bytes(int L, int S) { return sizeof(float) * ((size_t)L * 18 + (size_t)S * 16); }

}  // namespace sttode
