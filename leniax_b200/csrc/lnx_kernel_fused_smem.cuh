// The two earlier fused single-channel kernels (state in shared memory, one world per SM, statistics warp): 256 compute
// threads (lnx_world128_fused) and 512 (lnx_world128_r16).  Kept for A/B runs and as cross-checks of the TMEM kernel.
#pragma once
#include "lnx_resident_common.cuh"

namespace lnx {

// ---------------------------------------------------------------------------------------------------------------------
// fused kernel: C = K = 1
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FUSED_SMEM = 65536 * 3 + KPQ_F4 * 16 + NPART_FUSED * NT * 4 + CTRL_BYTES + SCRATCH_BYTES + TW_BYTES;

template <int GF, int SF, bool NP>
__global__ void __launch_bounds__(NTHREADS, 1) lnx_world128_fused(const RunArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float4* A4 = reinterpret_cast<float4*>(smem + 65536);
    float4* Kt = reinterpret_cast<float4*>(smem + 131072);
    float4* Kpq = reinterpret_cast<float4*>(smem + 196608);
    float* part = reinterpret_cast<float*>(smem + 196608 + KPQ_F4 * 16);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + 196608 + KPQ_F4 * 16 + NPART_FUSED * NT * 4);
    float2* scratch = reinterpret_cast<float2*>(smem + 196608 + KPQ_F4 * 16 + NPART_FUSED * NT * 4 + CTRL_BYTES);
    float4* twtab = reinterpret_cast<float4*>(smem + 196608 + KPQ_F4 * 16 + NPART_FUSED * NT * 4 + CTRL_BYTES + SCRATCH_BYTES);

    const int tid = threadIdx.x;
    const int n_worlds = P.n_sols * P.n_init;
    const bool early = (P.flags & LNX_RUN_EARLY_STOP) != 0;
    Regs R;
    init_twiddle_table(tid, twtab, c_tw128);  // made visible by the __syncthreads of the first world fetch
    int loaded_sol = -1;

    for (;;) {
        if (tid == NT) {
            ctrl->world = atomicAdd(P.queue, 1);
            ctrl->shift0 = ctrl->shift1 = 0;
            ctrl->stop = 0;
            ctrl->done = 0;
        }
        __syncthreads();
        const int world = ctrl->world;
        if (world >= n_worlds) break;
        const int sol = world / P.n_init, init = world - sol * P.n_init;

        if (tid < NT) {
            // ------------------------------------------------ compute threads ------------------------------------------
            gather_state(A4, P.cells0 + (size_t)world * (WS * WS), tid);
            if (sol != loaded_sol) {
                const float4* src = P.table + (size_t)sol * KTAB_F4;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) Kt[i * NT + tid] = __ldg(src + i * NT + tid);
                if (tid < KPQ_F4) Kpq[tid] = __ldg(src + KT_F4 + tid);
                loaded_sol = sol;
            }
            FusedConsts fc;
            {
                const float m = __ldg(P.gf_params + (size_t)sol * 2), s = __ldg(P.gf_params + (size_t)sol * 2 + 1);
                fc = fused_consts(GF, m, s, __ldg(P.weights + sol), P.mean, __ldg(P.dt + sol));
            }
            bar_sync(BAR_COMPUTE, NT);  // Kt / Kpq visible to every compute thread

            for (int t = 0; t < P.max_iter; ++t) {
                load_state_regs(R, A4, tid);
                __syncwarp();  // previous step's phase5 reads of this group's region are complete
                phase1(tid, R, W);
                __syncwarp();
                phase2_load(tid, R, W);
                __syncwarp();
                phase2_compute_store(tid, R, W, twtab);
                bar_sync(BAR_COMPUTE, NT);
                phase3_load_fft(tid, R, W);
                if (tid < 32) {  // warp 0 owns the packed DC|Nyquist column
                    phase3_col0_stash(tid, R, scratch);
                    __syncwarp();
                    phase3_col0_compute(tid, scratch, Kpq);
                    __syncwarp();
                }
                phase3_multiply(tid, R, Kt);
                if (tid < 32) phase3_col0_fetch(tid, R, scratch);
                phase3_ifft_store(tid, R, W);
                bar_sync(BAR_COMPUTE, NT);
                phase4_load(tid, R, W);
                __syncwarp();
                phase4_compute_store(tid, R, W, twtab);
                __syncwarp();
                phase5_load(tid, R, W);
                phase5_ifft(R);
                if (t > 0) {
                    while (ld_acquire_smem(&ctrl->done) < t) {}  // statistics of step t-1 are final: shift carry + stop flag
                    if (ctrl->stop) break;
                }
                cells_fused<GF, SF, NP>(tid, R.v, A4, fc, ctrl->shift0, ctrl->shift1, part);
                __threadfence_block();
                bar_arrive(BAR_PARTIALS, NTHREADS);
            }
            if (P.final_cells) scatter_state(P.final_cells + (size_t)world * (WS * WS), A4, tid);
        } else {
            // ------------------------------------------------ statistics warp ------------------------------------------
            const int lane = tid - NT;
            const float invR2 = 1.0f / (P.R * P.R), invR = 1.0f / P.R, inv_dt = 1.0f / P.stats_dt;
            StatsCarry S;
            S.reset();
            for (int t = 0; t < P.max_iter; ++t) {
                bar_sync(BAR_PARTIALS, NTHREADS);
                const float sc = stats_step(P, part, NPART_FUSED, lane, t, sol, init, S, ctrl, invR2, invR, inv_dt);
                const int stop = (early && sc == 0.f && t + 1 >= 128) ? 1 : 0;
                if (lane == 0) {
                    ctrl->shift0 = S.shift[0];
                    ctrl->shift1 = S.shift[1];
                    ctrl->stop = stop;
                    st_release_smem(&ctrl->done, t + 1);
                }
                if (stop) break;
            }
            if (lane == 0) P.n_alive[world] = S.n_alive;
        }
        __syncthreads();  // world done: ctrl / part / A4 can be reused
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// fused kernel, R16 variant: 512 compute threads (one real row quarter each) + statistics warp
// ---------------------------------------------------------------------------------------------------------------------
constexpr int R16_THREADS = r16::NT + 32;
constexpr int R16_NPART = PT_FIXED + 1;
constexpr int R16_OFF_A = 65536, R16_OFF_KT = 131072, R16_OFF_KPQ = 196608;
constexpr int R16_OFF_PART = R16_OFF_KPQ + R16_KPQ_F4 * 16;
constexpr int R16_OFF_CTRL = R16_OFF_PART + R16_NPART * r16::NT * 4;
constexpr int R16_OFF_SCRATCH = R16_OFF_CTRL + CTRL_BYTES;
constexpr int R16_OFF_TW = R16_OFF_SCRATCH + SCRATCH_BYTES;
constexpr int R16_SMEM = R16_OFF_TW + r16::TW_TABLE_F4 * 16;

__device__ __forceinline__ float stats_step_r16(const RunArgs& P, const float* part, int lane, int t, int sol, int init, StatsCarry& S,
                                                Ctrl* ctrl, float invR2, float invR, float inv_dt) {
#pragma unroll 1
    for (int k = 0; k < R16_NPART; ++k) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) a += part[k * r16::NT + lane + 32 * i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) ctrl->tot[k] = a;
    }
    __syncwarp();
    float sc = 0.f;
    if (lane == 0) sc = stats_finalize(ctrl->tot, 1, t, invR2, invR, inv_dt, S, ctrl->row);
    sc = __shfl_sync(0xffffffffu, sc, 0);
    const size_t plane = (size_t)P.n_sols * P.max_iter * P.n_init;
    const size_t idx = ((size_t)sol * P.max_iter + t) * P.n_init + init;
    if (lane < ST_COUNT)
        P.stats[lane * plane + idx] = ctrl->row[lane];
    else if (lane == ST_COUNT)
        P.channel_mass[idx] = ctrl->row[lane];
    return sc;
}

template <int GF, int SF, bool NP>
__global__ void __launch_bounds__(R16_THREADS, 1) lnx_world128_r16(const RunArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float4* A4 = reinterpret_cast<float4*>(smem + R16_OFF_A);
    float4* Kt = reinterpret_cast<float4*>(smem + R16_OFF_KT);
    float4* Kpq = reinterpret_cast<float4*>(smem + R16_OFF_KPQ);
    float* part = reinterpret_cast<float*>(smem + R16_OFF_PART);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + R16_OFF_CTRL);
    float2* scratch = reinterpret_cast<float2*>(smem + R16_OFF_SCRATCH);
    float4* twtab = reinterpret_cast<float4*>(smem + R16_OFF_TW);

    const int u = threadIdx.x;
    const int n_worlds = P.n_sols * P.n_init;
    const bool early = (P.flags & LNX_RUN_EARLY_STOP) != 0;
    r16::init_twiddle_table(u, twtab, c_tw128);
    int loaded_sol = -1;

    for (;;) {
        if (u == r16::NT) {
            ctrl->world = atomicAdd(P.queue, 1);
            ctrl->shift0 = ctrl->shift1 = 0;
            ctrl->stop = 0;
            ctrl->done = 0;
        }
        __syncthreads();
        const int world = ctrl->world;
        if (world >= n_worlds) break;
        const int sol = world / P.n_init, init = world - sol * P.n_init;

        if (u < r16::NT) {
            const int l = r16::t_l(u);
            {  // gather the initial state into the thread-private layout
                const float* img = P.cells0 + (size_t)world * (WS * WS) + r16::cell_row(u) * WS + l;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4)
                    A4[i4 * r16::NT + u] = make_float4(__ldg(img + 16 * i4), __ldg(img + 16 * i4 + 4), __ldg(img + 16 * i4 + 8), __ldg(img + 16 * i4 + 12));
            }
            if (sol != loaded_sol) {
                const float4* src = P.table + (size_t)sol * KTAB_F4 + KTAB32_F4;
#pragma unroll
                for (int i = 0; i < 8; ++i) Kt[i * r16::NT + u] = __ldg(src + i * r16::NT + u);
                if (u < R16_KPQ_F4) Kpq[u] = __ldg(src + R16_KT_F4 + u);
                loaded_sol = sol;
            }
            const FusedConsts fc = fused_consts(GF, __ldg(P.gf_params + (size_t)sol * 2), __ldg(P.gf_params + (size_t)sol * 2 + 1),
                                                __ldg(P.weights + sol), P.mean, __ldg(P.dt + sol));
            bar_sync(BAR_COMPUTE, r16::NT);

            for (int t = 0; t < P.max_iter; ++t) {
                float x[32];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 c = A4[i4 * r16::NT + u];
                    x[4 * i4 + 0] = c.x;
                    x[4 * i4 + 1] = c.y;
                    x[4 * i4 + 2] = c.z;
                    x[4 * i4 + 3] = c.w;
                }
                __syncwarp();  // the previous step's P5' reads of this warp's region are complete
                r16::phase1(u, x, W);
                __syncwarp();
                {
                    r16::P2State s2;
                    r16::phase2_compute(u, s2, W, twtab);
                    __syncwarp();
                    r16::phase2_store(u, s2, W);
                }
                bar_sync(BAR_COMPUTE, r16::NT);
                {
                    r16::Regs R;
                    r16::phase3_load_fft(u, R, W);
                    if (u < 32) {
                        r16::phase3_col0_stash(u, R, scratch);
                        __syncwarp();
                        r16::phase3_col0_compute(u, scratch, Kpq);
                        __syncwarp();
                    }
                    r16::phase3_multiply(u, R, Kt);
                    if (u < 32) r16::phase3_col0_fetch(u, R, scratch);
                    r16::phase3_ifft_store(u, R, W);
                }
                bar_sync(BAR_COMPUTE, r16::NT);
                {
                    r16::P4State s4;
                    r16::phase4_load_ifft(u, s4, W, twtab);
                    float2 pA[4], pB[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        pA[j].x = __shfl_xor_sync(0xffffffffu, s4.cA[4 + j].x, 1);
                        pA[j].y = __shfl_xor_sync(0xffffffffu, s4.cA[4 + j].y, 1);
                        pB[j].x = __shfl_xor_sync(0xffffffffu, s4.cB[4 + j].x, 1);
                        pB[j].y = __shfl_xor_sync(0xffffffffu, s4.cB[4 + j].y, 1);
                    }
                    __syncwarp();
                    r16::phase4_finish_store(u, s4, pA, pB, W, twtab);
                }
                __syncwarp();
                r16::phase5(u, x, W);
                if (t > 0) {
                    while (ld_acquire_smem(&ctrl->done) < t) {}
                    if (ctrl->stop) break;
                }
                r16::cells_fused<GF, SF, NP>(u, x, A4, fc, ctrl->shift0, ctrl->shift1, part);
                __threadfence_block();
                bar_arrive(BAR_PARTIALS, R16_THREADS);
            }
            if (P.final_cells) {
                float* img = P.final_cells + (size_t)world * (WS * WS) + r16::cell_row(u) * WS + l;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 c = A4[i4 * r16::NT + u];
                    img[16 * i4] = c.x;
                    img[16 * i4 + 4] = c.y;
                    img[16 * i4 + 8] = c.z;
                    img[16 * i4 + 12] = c.w;
                }
            }
        } else {
            const int lane = u - r16::NT;
            const float invR2 = 1.0f / (P.R * P.R), invR = 1.0f / P.R, inv_dt = 1.0f / P.stats_dt;
            StatsCarry S;
            S.reset();
            for (int t = 0; t < P.max_iter; ++t) {
                bar_sync(BAR_PARTIALS, R16_THREADS);
                const float sc = stats_step_r16(P, part, lane, t, sol, init, S, ctrl, invR2, invR, inv_dt);
                const int stop = (early && sc == 0.f && t + 1 >= 128) ? 1 : 0;
                if (lane == 0) {
                    ctrl->shift0 = S.shift[0];
                    ctrl->shift1 = S.shift[1];
                    ctrl->stop = stop;
                    st_release_smem(&ctrl->done, t + 1);
                }
                if (stop) break;
            }
            if (lane == 0) P.n_alive[world] = S.n_alive;
        }
        __syncthreads();
    }
}

}  // namespace lnx
