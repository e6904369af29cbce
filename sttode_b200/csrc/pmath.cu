// Poincare-ball math library kernels (hyptorch/pmath.py, rows P1-P8 of SURVEY.md §8a), sm_100a.
// Row-wise ops: one warp per row, lane-strided coalesced loads, warp-shuffle reductions; every clamp /
// epsilon constant of the reference is reproduced (1e-5 norm floor, 1e-3 ball margin, 1e-5 artanh clamp,
// 1e-5 Mobius denominator, tanh clamp 15).  dist_matrix uses the Gram-epilogue closed form: nothing of
// size b*n*d is materialised (the reference builds [b,n,d], pmath.py:416-427).
#include "kernels.cuh"

namespace sttode {

enum { P_PROJECT, P_LAMBDA, P_MOBADD, P_DIST, P_DIST0, P_EXPMAP0, P_LOGMAP0, P_EXPMAP, P_LOGMAP };

__device__ __forceinline__ float p_tanh(float x) { return tanhf(clamp_nan(x, -15.f, 15.f)); }           // pmath.py:11-12
__device__ __forceinline__ float p_artanh(float x) {                                                     // pmath.py:16-22
    x = clamp_nan(x, -1.f + 1e-5f, 1.f - 1e-5f);
    return (logf(1.f + x) - logf(1.f - x)) * 0.5f;
}

// mobius_add coefficients: x (+) y = (A x + B y) / D   (pmath.py:171-177)
struct MobCoef { float A, B, D; };
__device__ __forceinline__ MobCoef mob_coef(float x2, float y2, float xy, float c) {
    MobCoef m;
    m.A = 1.f + 2.f * c * xy + c * y2;
    m.B = 1.f - c * x2;
    m.D = (1.f + 2.f * c * xy + c * c * x2 * y2) + 1e-5f;
    return m;
}

template <int OP>
__global__ void k_pmath_rows(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int rows,
                             int d, float c) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * d;
    const float* yr = y ? y + (size_t)row * d : nullptr;
    const float sqrt_c = sqrtf(c);
    float x2 = 0.f, y2 = 0.f, xy = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float a = xr[k];
        x2 = fmaf(a, a, x2);
        if (yr) { const float b = yr[k]; y2 = fmaf(b, b, y2); xy = fmaf(a, b, xy); }
    }
    x2 = warp_sum(x2); y2 = warp_sum(y2); xy = warp_sum(xy);
    float* o = out + (size_t)row * d;
    if (OP == P_PROJECT) {
        const float norm = fmaxf(sqrtf(x2), 1e-5f), maxnorm = (1.f - 1e-3f) / sqrt_c;
        for (int k = lane; k < d; k += 32) o[k] = norm > maxnorm ? xr[k] / norm * maxnorm : xr[k];
    } else if (OP == P_LAMBDA) {
        if (lane == 0) out[row] = 2.f / (1.f - c * x2);
    } else if (OP == P_MOBADD) {
        const MobCoef m = mob_coef(x2, y2, xy, c);
        for (int k = lane; k < d; k += 32) o[k] = (m.A * xr[k] + m.B * yr[k]) / m.D;
    } else if (OP == P_DIST) {               // (-x) (+) y : xy -> -xy
        const MobCoef m = mob_coef(x2, y2, -xy, c);
        float n2 = 0.f;
        for (int k = lane; k < d; k += 32) { const float v = (m.B * yr[k] - m.A * xr[k]) / m.D; n2 = fmaf(v, v, n2); }
        n2 = warp_sum(n2);
        if (lane == 0) out[row] = p_artanh(sqrt_c * sqrtf(n2)) * 2.f / sqrt_c;
    } else if (OP == P_DIST0) {
        if (lane == 0) out[row] = p_artanh(sqrt_c * sqrtf(x2)) * 2.f / sqrt_c;
    } else if (OP == P_EXPMAP0) {
        const float un = fmaxf(sqrtf(x2), 1e-5f), th = p_tanh(sqrt_c * un);
        for (int k = lane; k < d; k += 32) o[k] = th * xr[k] / (sqrt_c * un);
    } else if (OP == P_LOGMAP0) {
        const float yn = fmaxf(sqrtf(x2), 1e-5f), at = p_artanh(sqrt_c * yn);
        for (int k = lane; k < d; k += 32) o[k] = xr[k] / yn / sqrt_c * at;
    } else if (OP == P_EXPMAP) {             // x (+) tanh(sqrt_c/2 * lambda_x * |u|) u / (sqrt_c |u|),  y == u
        const float un = fmaxf(sqrtf(y2), 1e-5f), lam = 2.f / (1.f - c * x2);
        const float s = p_tanh(sqrt_c / 2.f * lam * un) / (sqrt_c * un);       // second = s * u
        const MobCoef m = mob_coef(x2, s * s * y2, s * xy, c);
        for (int k = lane; k < d; k += 32) o[k] = (m.A * xr[k] + m.B * (s * yr[k])) / m.D;
    } else if (OP == P_LOGMAP) {
        const MobCoef m = mob_coef(x2, y2, -xy, c);
        float n2 = 0.f;
        for (int k = lane; k < d; k += 32) { const float v = (m.B * yr[k] - m.A * xr[k]) / m.D; n2 = fmaf(v, v, n2); }
        const float sn = sqrtf(warp_sum(n2)), lam = 2.f / (1.f - c * x2);
        const float f = 2.f / sqrt_c / lam * p_artanh(sqrt_c * sn);
        for (int k = lane; k < d; k += 32) o[k] = f * ((m.B * yr[k] - m.A * xr[k]) / m.D) / sn;
    }
}

// mobius_matvec: one warp per row; mx kept in shared memory (dout <= 1024).
__global__ void k_pmath_matvec(const float* __restrict__ mat, const float* __restrict__ x, float* __restrict__ out, int rows,
                               int d, int dout, float c) {
    extern __shared__ float sm[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + wid;
    if (row >= rows) return;
    float* mx = sm + wid * dout;
    const float* xr = x + (size_t)row * d;
    float x2 = 0.f;
    for (int k = lane; k < d; k += 32) x2 = fmaf(xr[k], xr[k], x2);
    x2 = warp_sum(x2);
    float m2 = 0.f;
    int allzero = 1;
    for (int o = lane; o < dout; o += 32) {
        float acc = 0.f;
        for (int k = 0; k < d; ++k) acc = fmaf(xr[k], __ldg(mat + (size_t)o * d + k), acc);
        mx[o] = acc;
        m2 = fmaf(acc, acc, m2);
        if (acc != 0.f) allzero = 0;
    }
    m2 = warp_sum(m2);
    allzero = __all_sync(0xffffffffu, allzero);
    const float sqrt_c = sqrtf(c), xn = fmaxf(sqrtf(x2), 1e-5f), mn = sqrtf(m2);
    const float f = p_tanh(mn / xn * p_artanh(sqrt_c * xn)) / (mn * sqrt_c);
    // ||res|| for the trailing project(): res = f * mx
    const float rn = fmaxf(fabsf(f) * mn, 1e-5f), maxnorm = (1.f - 1e-3f) / sqrt_c;
    __syncwarp();
    for (int o = lane; o < dout; o += 32) {
        float r = allzero ? 0.f : f * mx[o];
        if (!allzero && rn > maxnorm) r = r / rn * maxnorm;
        out[(size_t)row * dout + o] = r;
    }
}

// dist_matrix: out[i][j] = 2/sqrt_c * artanh(sqrt_c * ||(-x_i) (+) y_j||), 16x16 output tile per CTA,
// d streamed through shared memory in 32-wide slabs; Gram + both squared norms accumulated together.
__global__ void __launch_bounds__(256)
k_pmath_dist_matrix(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int b, int n, int d, float c) {
    __shared__ float xs[16][33], ys[16][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    float g = 0.f, x2 = 0.f, y2 = 0.f;
    for (int k0 = 0; k0 < d; k0 += 32) {
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int r = e >> 5, k = e & 31;
            const int xi = blockIdx.y * 16 + r, yj = blockIdx.x * 16 + r;
            xs[r][k] = (xi < b && k0 + k < d) ? x[(size_t)xi * d + k0 + k] : 0.f;
            ys[r][k] = (yj < n && k0 + k < d) ? y[(size_t)yj * d + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float a = xs[ty][k], bb = ys[tx][k];
            g = fmaf(a, bb, g); x2 = fmaf(a, a, x2); y2 = fmaf(bb, bb, y2);
        }
        __syncthreads();
    }
    if (i < b && j < n) {
        const float sqrt_c = sqrtf(c);
        const float A = 1.f - 2.f * c * g + c * y2, B = 1.f - c * x2;
        const float n2 = fmaxf(A * A * x2 - 2.f * A * B * g + B * B * y2, 0.f);
        const float den = (1.f - 2.f * c * g + c * c * x2 * y2) + 1e-5f;
        out[(size_t)i * n + j] = 2.f / sqrt_c * p_artanh(sqrt_c * (sqrtf(n2) / den));
    }
}

enum { E_TANH, E_ARTANH, E_ARSINH };
template <int OP>
__global__ void k_pmath_elem(const float* __restrict__ x, float* __restrict__ out, size_t n, float clampv) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        float r;
        if (OP == E_TANH) r = tanhf(clamp_nan(v, -clampv, clampv));
        else if (OP == E_ARTANH) r = p_artanh(v);
        else { const float u = v + sqrtf(1.f + v * v); r = logf(u < 1e-5f ? 1e-5f : u); }   // pmath.py:55, NaN propagates like clamp_min_
        out[i] = r;
    }
}

// ---- Oblique manifold stand-alone ops (core/manifolds/oblique.py:15-16, 36-45) ---------------------
__global__ void k_oblique_proj(const float* __restrict__ x, float* __restrict__ out, int rows, int d) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * d;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) s = fmaf(xr[k], xr[k], s);
    const float nrm = sqrtf(warp_sum(s));
    for (int k = lane; k < d; k += 32) out[(size_t)row * d + k] = xr[k] / nrm;     // no epsilon, as the reference
}

// out[bt][j][i] = acos(clamp(<p2[bt][j], p1[bt][i]>, lo, hi)); p1 [batch,n1,d], p2 [batch,n2,d]
__global__ void __launch_bounds__(256)
k_oblique_dist(const float* __restrict__ p1, const float* __restrict__ p2, float* __restrict__ out, int n1, int n2, int d,
               float lo, float hi) {
    __shared__ float as[16][33], bs[16][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, bt = blockIdx.z;
    const int j = blockIdx.y * 16 + ty, i = blockIdx.x * 16 + tx;
    const float* A = p2 + (size_t)bt * n2 * d;
    const float* B = p1 + (size_t)bt * n1 * d;
    float g = 0.f;
    for (int k0 = 0; k0 < d; k0 += 32) {
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int r = e >> 5, k = e & 31;
            const int aj = blockIdx.y * 16 + r, bi = blockIdx.x * 16 + r;
            as[r][k] = (aj < n2 && k0 + k < d) ? A[(size_t)aj * d + k0 + k] : 0.f;
            bs[r][k] = (bi < n1 && k0 + k < d) ? B[(size_t)bi * d + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) g = fmaf(as[ty][k], bs[tx][k], g);
        __syncthreads();
    }
    if (j < n2 && i < n1) out[((size_t)bt * n2 + j) * n1 + i] = acosf(clamp_nan(g, lo, hi));
}

void launch_oblique_proj(const float* x, float* out, int rows, int d, cudaStream_t st) {
    k_oblique_proj<<<(rows * 32 + 255) / 256, 256, 0, st>>>(x, out, rows, d);
}
void launch_oblique_dist(const float* p1, const float* p2, float* out, int batch, int n1, int n2, int d, float lo, float hi, cudaStream_t st) {
    dim3 grid((n1 + 15) / 16, (n2 + 15) / 16, batch);
    k_oblique_dist<<<grid, 256, 0, st>>>(p1, p2, out, n1, n2, d, lo, hi);
}

// ---- host launchers (called from capi.cu) --------------------------------------------------------
void launch_pmath_rows(int op, const float* x, const float* y, float* out, int rows, int d, float c, cudaStream_t st) {
    const int threads = 256, blocks = (rows * 32 + threads - 1) / threads;
    switch (op) {
        case P_PROJECT: k_pmath_rows<P_PROJECT><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_LAMBDA: k_pmath_rows<P_LAMBDA><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_MOBADD: k_pmath_rows<P_MOBADD><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_DIST: k_pmath_rows<P_DIST><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_DIST0: k_pmath_rows<P_DIST0><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_EXPMAP0: k_pmath_rows<P_EXPMAP0><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_LOGMAP0: k_pmath_rows<P_LOGMAP0><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_EXPMAP: k_pmath_rows<P_EXPMAP><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
        case P_LOGMAP: k_pmath_rows<P_LOGMAP><<<blocks, threads, 0, st>>>(x, y, out, rows, d, c); break;
    }
}
void launch_pmath_matvec(const float* m, const float* x, float* out, int rows, int d, int dout, float c, cudaStream_t st) {
    const int warps = 4;
    k_pmath_matvec<<<(rows + warps - 1) / warps, warps * 32, warps * dout * sizeof(float), st>>>(m, x, out, rows, d, dout, c);
}
void launch_pmath_dist_matrix(const float* x, const float* y, float* out, int b, int n, int d, float c, cudaStream_t st) {
    dim3 grid((n + 15) / 16, (b + 15) / 16);
    k_pmath_dist_matrix<<<grid, 256, 0, st>>>(x, y, out, b, n, d, c);
}
void launch_pmath_elem(int op, const float* x, float* out, size_t n, float clampv, cudaStream_t st) {
    const int threads = 256;
    size_t want = (n + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
    if (op == E_TANH) k_pmath_elem<E_TANH><<<blocks, threads, 0, st>>>(x, out, n, clampv);
    else if (op == E_ARTANH) k_pmath_elem<E_ARTANH><<<blocks, threads, 0, st>>>(x, out, n, clampv);
    else k_pmath_elem<E_ARSINH><<<blocks, threads, 0, st>>>(x, out, n, clampv);
}

// ------------------------------------------------------------------------------------------------
// min-over-K ADE / FDE (utils/metrics.py:7-26): one warp per agent row, lanes over samples.
// ------------------------------------------------------------------------------------------------
// One warp per agent row, lanes over samples; a block of 32 warps reduces its rows in shared memory (fixed order) and issues ONE
// atomic triple — 3 double atomics per row on the same three addresses serialised the whole launch (0.27 ms at 53 760 rows, round 1).
__global__ void __launch_bounds__(1024) k_min_ade_fde(const float* __restrict__ pred, const float* __restrict__ gt, int gt_layout, int rows, int K,
                                                     int Tf, double* __restrict__ sums) {
    __shared__ float s_ade[32], s_fde[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 32 + warp;
    float best_ade = INFINITY, best_fde = INFINITY;
    if (row < rows) {
        for (int k = lane; k < K; k += 32) {
            const float2* p = reinterpret_cast<const float2*>(pred + ((size_t)row * K + k) * Tf * 2);
            float s = 0.f, last = 0.f;
            for (int t = 0; t < Tf; ++t) {
                float gx, gy;
                if (gt_layout == STTODE_TRAJ_RT2) { gx = gt[((size_t)row * Tf + t) * 2]; gy = gt[((size_t)row * Tf + t) * 2 + 1]; }
                else                              { gx = gt[(size_t)row * 2 * Tf + t];   gy = gt[(size_t)row * 2 * Tf + Tf + t]; }
                const float2 q = __ldg(p + t);
                const float dx = q.x - gx, dy = q.y - gy;
                last = sqrtf(dx * dx + dy * dy);
                s += last;
            }
            best_ade = fminf(best_ade, s / (float)Tf);
            best_fde = fminf(best_fde, last);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            best_ade = fminf(best_ade, __shfl_xor_sync(0xffffffffu, best_ade, o));
            best_fde = fminf(best_fde, __shfl_xor_sync(0xffffffffu, best_fde, o));
        }
    }
    if (lane == 0) { s_ade[warp] = row < rows ? best_ade : 0.f; s_fde[warp] = row < rows ? best_fde : 0.f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, f = 0.0;
        const int n = min(32, rows - blockIdx.x * 32);
        for (int i = 0; i < n; ++i) { a += (double)s_ade[i]; f += (double)s_fde[i]; }
        atomicAdd(sums, a);
        atomicAdd(sums + 1, f);
        atomicAdd(sums + 2, (double)n);
    }
}

}  // namespace sttode
