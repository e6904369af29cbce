// Data-parallel gradient exchange fused with the optimizer over NVLink / NVSwitch peer memory (sm_100a, one process per GPU, one node).
//
// Every rank owns an arena [gradient buffer | flag words] allocated with cudaMalloc and exported through CUDA IPC; every rank maps all
// peers' arenas.  One training step then needs no NCCL call and no host synchronisation:
//     k_peer_barrier   every rank's gradients are complete and visible (release / acquire flags in peer memory, system scope)
//     k_allreduce_adam each rank reads all `world` gradient buffers over NVLink, sums them IN RANK ORDER (bit-identical on every rank),
//                      scales by 1/world and applies the Adam update to its replica of the weights in the same pass
//     k_peer_barrier   every peer has finished reading this rank's gradients (they may be zeroed for the next step)
// All three are plain kernels on the caller's stream: the whole step — forward, backward, exchange, update — is capturable as ONE CUDA graph at
// any world size (a captured NCCL collective hung an 8-rank run in round 1).  A 6.6 MB buffer is latency-bound (tens of microseconds), so the
// one-shot form (every rank reads everything) beats reduce-scatter + all-gather, which would need a third barrier.
// Spins are bounded: a peer that never arrives raises the arena's error word instead of hanging the GPU.
#include "kernels.cuh"

namespace sttode {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// flags_of[j] = device address of rank j's flag words (each rank: [0..world) arrival epochs, [PEER_EPOCH] local epoch, [PEER_ERR] error, [PEER_STEP] Adam step)
__global__ void k_peer_barrier(unsigned* const* __restrict__ flags_of, int world, int rank, int bump_step) {
    unsigned* mine = flags_of[rank];
    __shared__ unsigned epoch;
    if (threadIdx.x == 0) epoch = mine[PEER_EPOCH] + 1u;
    __syncthreads();
    const unsigned e = epoch;
    const int j = threadIdx.x;
    if (j < world && !mine[PEER_ERR]) {
        __threadfence_system();                     // everything this GPU wrote before the barrier is visible to the peers
        st_release_sys(flags_of[j] + rank, e);      // my arrival, in peer j's flag words
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine + j) - e) < 0) {
            if (clock64() - t0 > 6000000000LL) { mine[PEER_ERR] = 1u; break; }       // ~3 s: a peer never arrived; fail loudly on the host side, never hang
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mine[PEER_EPOCH] = e;
        if (bump_step) mine[PEER_STEP] += 1u;
    }
}

// grads_of[r] = rank r's gradient buffer.  lr_dev: device scalar (a scheduler may change it between replays of a captured step).
__global__ void __launch_bounds__(256)
k_allreduce_adam(const float* const* __restrict__ grads_of, int world, int rank, unsigned* const* __restrict__ flags_of, float* __restrict__ param,
                 float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, float* __restrict__ grad_avg, size_t n4, const float* __restrict__ lr_dev,
                 float beta1, float beta2, float eps) {
    const unsigned step = flags_of[rank][PEER_STEP] + 1u;    // the closing barrier bumps the stored count
    const float lr = *lr_dev;
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2), inv_world = 1.f / (float)world;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 g = __ldcg(reinterpret_cast<const float4*>(grads_of[0]) + i);
        for (int r = 1; r < world; ++r) {           // fixed rank order: every rank computes the same bits
            const float4 a = __ldcg(reinterpret_cast<const float4*>(grads_of[r]) + i);
            g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
        }
        g.x *= inv_world; g.y *= inv_world; g.z *= inv_world; g.w *= inv_world;
        float4 m = reinterpret_cast<float4*>(exp_avg)[i], v = reinterpret_cast<float4*>(exp_avg_sq)[i], p = reinterpret_cast<float4*>(param)[i];
        const float gg[4] = {g.x, g.y, g.z, g.w};
        float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w}, pp[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {               // torch.optim.Adam (no weight decay, no amsgrad): same operation order as its single-tensor path
            mm[c] = mm[c] + (gg[c] - mm[c]) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
            vv[c] = vv[c] * beta2 + (1.f - beta2) * gg[c] * gg[c];      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
            const float denom = sqrtf(vv[c]) * inv_sqrt_bc2 + eps;
            pp[c] -= step_size * (mm[c] / denom);
        }
        reinterpret_cast<float4*>(exp_avg)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(exp_avg_sq)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        reinterpret_cast<float4*>(param)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        if (grad_avg) reinterpret_cast<float4*>(grad_avg)[i] = g;
    }
}

void launch_peer_barrier(unsigned* const* flags_of, int world, int rank, int bump_step, cudaStream_t st) {
    k_peer_barrier<<<1, 32, 0, st>>>(flags_of, world, rank, bump_step);
}
void launch_allreduce_adam(const float* const* grads_of, int world, int rank, unsigned* const* flags_of, float* param, float* exp_avg, float* exp_avg_sq,
                           float* grad_avg, size_t n4, const float* lr_dev, float beta1, float beta2, float eps, cudaStream_t st) {
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    k_allreduce_adam<<<blocks, 256, 0, st>>>(grads_of, world, rank, flags_of, param, exp_avg, exp_avg_sq, grad_avg, n4, lr_dev, beta1, beta2, eps);
}

}  // namespace sttode
