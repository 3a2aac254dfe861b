// leniax_b200 CUDA kernels (sm_100a) + C ABI (include/leniax_b200.h).
//
// Two persistent kernels, one CTA per world, 256 compute threads + 1 statistics warp:
//   world128_fused   : 1 channel, 1 kernel, stats only (run_scan_mem_optimized hot path, BASELINE config B).
//                      State, work buffer and kernel spectrum stay in shared memory for the whole run;
//                      only the statistics rows go to HBM.
//   world128_generic : any C <= 8, K <= 32, all growth/state functions, optional full trajectory output
//                      (run_scan, core.update); channel states / spectra / field accumulators live in a per-CTA
//                      global scratch that stays L2 resident.
// There is no CPU fallback anywhere in this file.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/leniax_b200.h"
#include "lnx_step.cuh"
#include "lnx_step_r16.cuh"
#include "lnx_stats_batch.cuh"
#include "lnx_tmem.cuh"
#include "lnx_tiled.cuh"
#include "lnx_conv.cuh"

namespace lnx {

constexpr int NTHREADS = NT + 32;   // 256 compute threads + 1 statistics warp
constexpr int BAR_COMPUTE = 1;      // named barrier: the 256 compute threads
constexpr int BAR_PARTIALS = 2;     // compute arrive  -> statistics warp sync   (partials of step t are in smem)
constexpr int KT_F4 = 16 * NT;      // float4 per kernel table (complex multipliers)
constexpr int KPQ_F4 = 32 * KPQ_LANES;  // float4 per packed-column table (Kp, Kq)
constexpr int SCRATCH_BYTES = 2 * 4 * 32 * 8;  // packed-column exchange of warp 0
constexpr int TW_BYTES = TW_TABLE_F4 * 16;     // run-time twiddle table of P2/P4
constexpr int KTAB32_F4 = KT_F4 + KPQ_F4;                 // T32 layout (generic kernel, fused T32 variant)
constexpr int R16_KT_F4 = 8 * r16::NT, R16_KPQ_F4 = 16 * 8;
constexpr int KTAB16_F4 = R16_KT_F4 + R16_KPQ_F4;        // R16 layout (fused R16 variant)
constexpr int KTAB_F4 = KTAB32_F4 + KTAB16_F4;           // per (solution, kernel): both layouts back to back
constexpr int NPART_FUSED = PT_FIXED + 1;
constexpr int NPART_MAX = PT_FIXED + MAX_C;

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// release/acquire flag in shared memory: the statistics warp publishes "step t is final" without forcing the compute
// warps through a CTA-wide barrier (a bar.sync with all 288 threads re-aligned the 8 compute warps once more per step)
__device__ __forceinline__ int ld_acquire_smem(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_smem(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

__constant__ float2 c_tw128[128];

struct Ctrl {
    int world;
    int shift0, shift1;
    int stop;
    int done;                      // number of steps whose statistics (carry + stop flag) are final
    float tot[NPART_MAX];          // CTA-wide sums of the step (statistics warp)
    float row[ST_COUNT + MAX_C];   // finished statistics row
};
constexpr int CTRL_BYTES = 256;
static_assert(sizeof(Ctrl) <= CTRL_BYTES, "Ctrl does not fit its shared-memory slot");

struct RunArgs {
    const float* cells0;
    const float4* table;
    const float* gf_params;
    const float* weights;
    const float* dt;
    float* stats;
    float* channel_mass;
    float* n_alive;
    float* final_cells;
    float* cells_out;
    float* field_out;
    float* potential_out;
    float4* scratch;
    int* queue;
    int n_sols, n_init, max_iter;
    int C, K;
    int state_fn, mean;
    float R, stats_dt;
    unsigned flags;
    int c_in[MAX_K];
    int gf_id[MAX_K];
};

// ---------------------------------------------------------------------------------------------------------------------
// kernel-spectrum table builder: K_fft [n_sols][nb_slots][128][128] complex64 -> per-thread multipliers
// ---------------------------------------------------------------------------------------------------------------------
struct PrepArgs {
    const float2* K_fft;
    float4* table;
    int K, nb_slots;
    int slot[MAX_K];
};
__global__ void __launch_bounds__(NT) lnx_prepare_kernel(PrepArgs P) {
    const int sol = blockIdx.x / P.K, k = blockIdx.x % P.K, tid = threadIdx.x;
    const float2* Kf = P.K_fft + ((size_t)sol * P.nb_slots + P.slot[k]) * (WS * WS);
    float4* tab = P.table + ((size_t)sol * P.K + k) * KTAB_F4;
    const float scale = 1.0f / (2.0f * WS * WS);
    const int col = t_col(tid);
    for (int i = 0; i < 16; ++i) {
        float2 v[2];
        for (int e = 0; e < 2; ++e) {
            const int m = p3_slot_m(tid, 2 * i + e);
            v[e] = col == 0 ? make_float2(0.f, 0.f) : Kf[m * WS + col];
        }
        tab[i * NT + tid] = make_float4(v[0].x * scale, v[0].y * scale, v[1].x * scale, v[1].y * scale);
    }
    if (tid < KPQ_LANES) {
        for (int s = 0; s < 32; ++s) {
            const int m = p3_slot_m(tid, s);
            const float2 k0 = Kf[m * WS], k64 = Kf[m * WS + 64];
            const float h = 0.5f * scale;
            tab[KT_F4 + s * KPQ_LANES + tid] = make_float4((k0.x + k64.x) * h, (k0.y + k64.y) * h, (k0.x - k64.x) * h, (k0.y - k64.y) * h);
        }
    }
    // R16 layout: thread u = (col, m2), 16 slots
    float4* tab16 = tab + KTAB32_F4;
    for (int u = tid; u < r16::NT; u += NT) {
        const int c16 = r16::t_col(u);
        for (int i = 0; i < 8; ++i) {
            float2 v[2];
            for (int e = 0; e < 2; ++e) {
                const int m = r16::p3_slot_m(u, 2 * i + e);
                v[e] = c16 == 0 ? make_float2(0.f, 0.f) : Kf[m * WS + c16];
            }
            tab16[i * r16::NT + u] = make_float4(v[0].x * scale, v[0].y * scale, v[1].x * scale, v[1].y * scale);
        }
        if (u < 8) {
            for (int pos = 0; pos < 16; ++pos) {
                const int m = r16::p3_slot_m(u, pos);
                const float2 k0 = Kf[m * WS], k64 = Kf[m * WS + 64];
                const float h = 0.5f * scale;
                tab16[R16_KT_F4 + pos * 8 + u] = make_float4((k0.x + k64.x) * h, (k0.y + k64.y) * h, (k0.x - k64.x) * h, (k0.y - k64.y) * h);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// plain 2-D FFT of real 128x128 images -> full complex spectrum (used to build K = fftn(fftshift(kernel)) like
// leniax/kernels.py:145-149 without cuFFT).  One CTA per image, phases P1..P3a of the resident pipeline.
// ---------------------------------------------------------------------------------------------------------------------
template <bool B0, int S>
__device__ __forceinline__ void rfft2_col0(const float2* v, float2* out, int tid) {
    if constexpr (S < 32) {
        const int m = p3_slot_m(tid, S);
        const float2 g = v[S], gp = v[col0_partner(B0, S)];
        // v = 2 (F0 + i F64):  F0 = (G + conj G')/4, F64 = -i (G - conj G')/4
        out[m * WS] = make_float2((g.x + gp.x) * 0.25f, (g.y - gp.y) * 0.25f);
        out[m * WS + 64] = make_float2((g.y + gp.y) * 0.25f, (gp.x - g.x) * 0.25f);
        rfft2_col0<B0, S + 1>(v, out, tid);
    }
}
__global__ void __launch_bounds__(NT) lnx_rfft2_kernel(const float* __restrict__ images, float2* __restrict__ spectra) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    const int tid = threadIdx.x, l = t_sub(tid) & 3;
    const float* img = images + (size_t)blockIdx.x * (WS * WS);
    float2* out = spectra + (size_t)blockIdx.x * (WS * WS);
    float4* twtab = reinterpret_cast<float4*>(smem + 65536);
    Regs R;
    init_twiddle_table(tid, twtab, c_tw128);
#pragma unroll 8
    for (int j = 0; j < 32; ++j) R.v[j] = make_float2(img[cell_row(tid, 0) * WS + 4 * j + l], img[cell_row(tid, 1) * WS + 4 * j + l]);
    phase1(tid, R, W);
    __syncthreads();
    phase2_load(tid, R, W);
    __syncthreads();
    phase2_compute_store(tid, R, W, twtab);
    __syncthreads();
    phase3_load_fft(tid, R, W);
    const int col = t_col(tid);
    if (col != 0) {
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            const int m = p3_slot_m(tid, s);
            const float2 v = make_float2(R.v[s].x * 0.5f, R.v[s].y * 0.5f);
            out[m * WS + col] = v;
            out[((WS - m) & (WS - 1)) * WS + (WS - col)] = make_float2(v.x, -v.y);
        }
    } else if (tid == 0) {
        rfft2_col0<true, 0>(R.v, out, tid);
    } else {
        rfft2_col0<false, 0>(R.v, out, tid);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// FP32 FMA-throughput probe: the roofline denominator for the resident kernels (MEASURED_PEAKS.json has no FP32 entry)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) lnx_fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456f) out[0] = r;  // never true in practice; keeps the loop alive
}

// ---------------------------------------------------------------------------------------------------------------------
// shared helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_state_regs(Regs& R, const float4* A4, int tid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 c0 = A4[i * NT + tid], c1 = A4[(8 + i) * NT + tid];
        R.v[4 * i + 0] = make_float2(c0.x, c1.x);
        R.v[4 * i + 1] = make_float2(c0.y, c1.y);
        R.v[4 * i + 2] = make_float2(c0.z, c1.z);
        R.v[4 * i + 3] = make_float2(c0.w, c1.w);
    }
}
// gather one channel image [128][128] (row major, global) into the thread-private state layout
__device__ __forceinline__ void gather_state(float4* A4, const float* img, int tid) {
    const int l = t_sub(tid) & 3;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const float* row = img + cell_row(tid, i >> 3) * WS + 16 * (i & 7) + l;
        A4[i * NT + tid] = make_float4(__ldg(row), __ldg(row + 4), __ldg(row + 8), __ldg(row + 12));
    }
}
__device__ __forceinline__ void scatter_state(float* img, const float4* A4, int tid) {
    const int l = t_sub(tid) & 3;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        float* row = img + cell_row(tid, i >> 3) * WS + 16 * (i & 7) + l;
        const float4 c = A4[i * NT + tid];
        row[0] = c.x;
        row[4] = c.y;
        row[8] = c.z;
        row[12] = c.w;
    }
}

// statistics warp: reduce the partials of one step (rolled loop: this code is fetched every step, keep it small),
// lane 0 finalises, lanes 0..10+C store the row
__device__ __forceinline__ float stats_step(const RunArgs& P, const float* part, int npart, int lane, int t, int sol, int init,
                                            StatsCarry& S, Ctrl* ctrl, float invR2, float invR, float inv_dt) {
#pragma unroll 1
    for (int k = 0; k < npart; ++k) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a += part[k * NT + lane + 32 * i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) ctrl->tot[k] = a;
    }
    __syncwarp();
    float sc = 0.f;
    if (lane == 0) sc = stats_finalize(ctrl->tot, P.C, t, invR2, invR, inv_dt, S, ctrl->row);
    sc = __shfl_sync(0xffffffffu, sc, 0);
    const size_t plane = (size_t)P.n_sols * P.max_iter * P.n_init;
    const size_t idx = ((size_t)sol * P.max_iter + t) * P.n_init + init;
    if (lane < ST_COUNT)
        P.stats[lane * plane + idx] = ctrl->row[lane];
    else if (lane < ST_COUNT + P.C)
        P.channel_mass[idx * P.C + (lane - ST_COUNT)] = ctrl->row[lane];
    return sc;
}

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
// fused kernel, TMEM variant (default): C = K = 1, two worlds (CTAs) per SM
//
// Same five phases as lnx_world128_fused, but the two thread-private arrays (state, kernel multipliers) live in tensor
// memory instead of shared memory (lnx_tmem.cuh), the new state stays in registers from the cell phase to phase 1 of the
// next step, and there is no statistics warp: the partial sums of step t are reduced by warps 1..7 at the start of phase 3
// of step t+1 (behind the barrier that is there anyway), warp 1 advances the shift carry, and warp 7 turns 32 steps of
// totals into statistics rows at once (lnx_stats_batch.cuh).  256 threads x 128 registers + 83 KB of shared memory per
// CTA => two CTAs per SM, i.e. four warps per scheduler from two INDEPENDENT worlds, whose barriers do not align.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TM_COLS = 256;  // per CTA: state in columns [0,128), kernel multipliers in [128,256)
constexpr int TM_OFF_PART = 65536;
constexpr int TM_OFF_RING = TM_OFF_PART + NPART_FUSED * NT * 4;
constexpr int TM_OFF_KPQ = TM_OFF_RING + RING_ROWS * RING_STRIDE_1 * 4;
constexpr int TM_OFF_SCRATCH = TM_OFF_KPQ + KPQ_F4 * 16;
constexpr int TM_OFF_TW = TM_OFF_SCRATCH + SCRATCH_BYTES;
constexpr int TM_OFF_XT = TM_OFF_TW + TW_BYTES;
constexpr int TM_OFF_CTRL = TM_OFF_XT + XT_F4 * 16;
struct TmCtrl {
    // written by thread 0 at world start / by warp 7 at batch boundaries; kept in their own 16 bytes so that a vectorised read of
    // them never touches the words warp 1 updates every step (compute-sanitizer racecheck flagged exactly that overlap)
    int world;
    int stop;
    uint32_t tmem_base;
    int pad0;
    alignas(16) int shift0;  // total_shift_idx used by the next cell phase (advanced by warp 1)
    int shift1;
    int pad1[2];
    alignas(16) BatchCarry carry;  // statistics carry between batches (warp 7)
};
constexpr int TM_SMEM = TM_OFF_CTRL + 160;
static_assert(sizeof(TmCtrl) <= 160, "TmCtrl does not fit its shared-memory slot");

struct TmemStore {  // Store concept of cells_fused_rs
    uint32_t addr;
    __device__ __forceinline__ void load(int i, float* d) const { tm::ld8(addr + 8 * i, d); }
    __device__ __forceinline__ void wait_load(float* d) const { tm::wait_ld8(d); }
    __device__ __forceinline__ void store(int i, const float* s) const { tm::st8(addr + 8 * i, s); }
};

__device__ __forceinline__ float tm_reduce_one(const float* part, int k, int lane) {  // same order as stats_step
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += part[k * NT + lane + 32 * i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}
// CTA-wide totals of one step -> ring row; warp 1 also advances the shift carry (statistics.py:117-119)
__device__ __forceinline__ void tm_reduce_partials(const float* part, float* row, TmCtrl* ctrl, float4* xt, int warp, int lane) {
    if (warp == 1) {
        const float m = tm_reduce_one(part, PT_M00_C0, lane), r = tm_reduce_one(part, PT_MX_R, lane), c = tm_reduce_one(part, PT_MX_C, lane);
        const float m00 = 0.f + m;
        const float im = sdiv(1.0f, m00 + EPS);
        const float c0 = r * im, c1 = c * im;
        const int shift1 = (ctrl->shift1 + trunc_to_int(c1)) & (WS - 1);  // every lane computes the same value
        __syncwarp();
        if (lane == 0) {
            row[RING_M00] = m;
            row[PT_MX_R] = r;
            row[PT_MX_C] = c;
            row[RING_C0] = c0;
            row[RING_C1] = c1;
            ctrl->shift0 = (ctrl->shift0 + trunc_to_int(c0)) & (WS - 1);
            ctrl->shift1 = shift1;
        }
        xt_build(lane, shift1, xt);  // column coordinates of the next cell phase
    } else if (warp >= 2) {
        // CNT_A, G00, CNT_G, CNT_P, GX_R, GX_C on warps 2..7; MX2_R, MX2_C as second item of warps 2, 3
        const int ka = warp < 6 ? warp - 2 : warp + 2;
        const float a = tm_reduce_one(part, ka, lane);
        if (lane == 0) row[ka] = a;
        if (warp < 4) {
            const float b = tm_reduce_one(part, warp + 4, lane);
            if (lane == 0) row[warp + 4] = b;
        }
    }
}
static_assert(PT_CNT_A == 0 && PT_G00 == 1 && PT_CNT_G == 2 && PT_CNT_P == 3 && PT_MX2_R == 6 && PT_MX2_C == 7 && PT_GX_R == 8 &&
                  PT_GX_C == 9 && PT_MX_R == 4 && PT_MX_C == 5 && PT_M00_C0 == 10,
              "tm_reduce_partials assumes this order of the partial sums");

__device__ __forceinline__ void phase3_multiply_tm(Regs& R, uint32_t kt_addr, float (&k)[2][8]) {  // k[0] already in flight
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        tm::wait_ld8(k[c & 1]);
        if (c + 1 < 8) tm::ld8(kt_addr + 8 * (c + 1), k[(c + 1) & 1]);
#pragma unroll
        for (int e = 0; e < 4; ++e) R.v[4 * c + e] = cmul(R.v[4 * c + e], make_float2(k[c & 1][2 * e], k[c & 1][2 * e + 1]));
    }
}

template <int GF, int SF, bool NP>
__global__ void __launch_bounds__(NT, 2) lnx_world128_tm(const RunArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float* part = reinterpret_cast<float*>(smem + TM_OFF_PART);
    float* ring = reinterpret_cast<float*>(smem + TM_OFF_RING);
    float4* Kpq = reinterpret_cast<float4*>(smem + TM_OFF_KPQ);
    float2* scratch = reinterpret_cast<float2*>(smem + TM_OFF_SCRATCH);
    float4* twtab = reinterpret_cast<float4*>(smem + TM_OFF_TW);
    float4* xt = reinterpret_cast<float4*>(smem + TM_OFF_XT);
    TmCtrl* ctrl = reinterpret_cast<TmCtrl*>(smem + TM_OFF_CTRL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_worlds = P.n_sols * P.n_init;
    const bool early = (P.flags & LNX_RUN_EARLY_STOP) != 0;
    const float invR2 = 1.0f / (P.R * P.R), invR = 1.0f / P.R, inv_dt = 1.0f / P.stats_dt;
    const size_t plane = (size_t)P.n_sols * P.max_iter * P.n_init;

    if (warp == 0) tm::alloc(&ctrl->tmem_base, TM_COLS);
    init_twiddle_table(tid, twtab, c_tw128);
    tm::fence_before_sync();
    __syncthreads();
    tm::fence_after_sync();
    const uint32_t tbase = ctrl->tmem_base;
    const TmemStore st{tm::warp_addr(tbase, warp, (warp >> 2) * 64)};
    const uint32_t kt_addr = tm::warp_addr(tbase, warp, 128 + (warp >> 2) * 64);
    int loaded_sol = -1;
    Regs R;

    for (;;) {
        if (tid == 0) {
            ctrl->world = atomicAdd(P.queue, 1);
            ctrl->shift0 = ctrl->shift1 = 0;
            ctrl->stop = 0;
            ctrl->carry.reset();
        }
        __syncthreads();
        const int world = ctrl->world;
        if (world >= n_worlds) break;
        const int sol = world / P.n_init, init = world - sol * P.n_init;
        {  // initial state: global -> registers (phase-1 layout) and tensor memory
            const int l = t_sub(tid) & 3;
            const float* r0 = P.cells0 + (size_t)world * (WS * WS) + cell_row(tid, 0) * WS + l;
            const float* r1 = P.cells0 + (size_t)world * (WS * WS) + cell_row(tid, 1) * WS + l;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float n[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    n[2 * e] = __ldg(r0 + 16 * i + 4 * e);
                    n[2 * e + 1] = __ldg(r1 + 16 * i + 4 * e);
                    R.v[4 * i + e] = make_float2(n[2 * e], n[2 * e + 1]);
                }
                st.store(i, n);
            }
        }
        if (sol != loaded_sol) {
            const float4* src = P.table + (size_t)sol * KTAB_F4;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) tm::st4(kt_addr + 4 * i, __ldg(src + i * NT + tid));
            if (tid < KPQ_F4) Kpq[tid] = __ldg(src + KT_F4 + tid);  // made visible by the first barrier of step 0
            loaded_sol = sol;
        }
        tm::wait_st();
        if (warp == 1) xt_build(lane, 0, xt);  // coordinates for step 0 (made visible by the barriers of step 0)
        const FusedConsts fc = fused_consts(GF, __ldg(P.gf_params + (size_t)sol * 2), __ldg(P.gf_params + (size_t)sol * 2 + 1),
                                            __ldg(P.weights + sol), P.mean, __ldg(P.dt + sol));
        const size_t idx_world = (size_t)sol * P.max_iter * P.n_init + init;  // statistics index of step 0

        int t = 0;
        for (; t < P.max_iter; ++t) {
            __syncwarp();  // previous step's phase-5 reads of this group's region are complete
            phase1(tid, R, W);
            __syncwarp();
            phase2_load(tid, R, W);
            __syncwarp();
            phase2_compute_store(tid, R, W, twtab);
            __syncthreads();
            if (t > 0) {
                if (ctrl->stop) break;  // written by warp 7 before this barrier, next written after the following one
                tm_reduce_partials(part, ring + ((t - 1) & (RING_ROWS - 1)) * RING_STRIDE_1, ctrl, xt, warp, lane);
            }
            float kbuf[2][8];
            tm::ld8(kt_addr, kbuf[0]);  // first multiplier chunk: lands during the column transforms
            phase3_load_fft(tid, R, W);
            if (tid < 32) phase3_col0_stash(tid, R, scratch);  // warp 0 owns the packed DC|Nyquist column (threads 0..3)
            phase3_multiply_tm(R, kt_addr, kbuf);              // (their plain products are overwritten by the fetch below)
            if (tid < 32) {  // the stash has landed behind the multiply; G' = G Kp + conj(G[-m]) Kq through the scratch
                __syncwarp();
                phase3_col0_compute(tid, scratch, Kpq);
                __syncwarp();
                phase3_col0_fetch(tid, R, scratch);
            }
            phase3_ifft_store(tid, R, W);
            __syncthreads();
            if (warp == 7 && t > 0 && (t & (RING_ROWS - 1)) == 0) {  // rows t-32 .. t-1 are complete
                BatchCarry S = ctrl->carry;
                stats_finalize_batch<1, RING_STRIDE_1>(ring, RING_ROWS, lane, 1, P.stats, P.channel_mass, plane,
                                                       idx_world + (size_t)S.rows * P.n_init, P.n_init, invR2, invR, inv_dt, S);
                __syncwarp();
                if (lane == 0) {
                    ctrl->carry = S;
                    if (early && S.should_continue == 0.f && S.rows >= 128) ctrl->stop = 1;
                }
            }
            phase4_load(tid, R, W);
            __syncwarp();
            phase4_compute_store(tid, R, W, twtab);
            __syncwarp();
            phase5_load(tid, R, W);
            phase5_ifft(R);
            tm::wait_st();  // the previous step's state stores (long complete by now)
            cells_fused_rs<GF, SF, NP>(tid, R.v, st, fc, ctrl->shift0, xt, part);
        }
        // the partial sums of the last completed cell phase (step t-1) are not reduced yet; t >= 1 here
        tm::wait_st();
        __syncthreads();
        tm_reduce_partials(part, ring + ((t - 1) & (RING_ROWS - 1)) * RING_STRIDE_1, ctrl, xt, warp, lane);
        __syncthreads();
        if (warp == 7) {  // flush the pending rows S.rows .. t-1 (1..32 of them)
            BatchCarry S = ctrl->carry;
            stats_finalize_batch<1, RING_STRIDE_1>(ring, t - S.rows, lane, 1, P.stats, P.channel_mass, plane,
                                                   idx_world + (size_t)S.rows * P.n_init, P.n_init, invR2, invR, inv_dt, S);
            if (lane == 0) P.n_alive[world] = S.n_alive;
        }
        if (P.final_cells) {
            const int l = t_sub(tid) & 3;
            float* r0 = P.final_cells + (size_t)world * (WS * WS) + cell_row(tid, 0) * WS + l;
            float* r1 = P.final_cells + (size_t)world * (WS * WS) + cell_row(tid, 1) * WS + l;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float n[8];
                st.load(i, n);
                st.wait_load(n);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    r0[16 * i + 4 * e] = n[2 * e];
                    r1[16 * i + 4 * e] = n[2 * e + 1];
                }
            }
        }
        __syncthreads();  // world done: ctrl / part / ring can be reused
    }
    __syncthreads();
    if (warp == 0) tm::dealloc(tbase, TM_COLS);
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

// ---------------------------------------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------------------------------------
struct GenericConsts {
    GfConst gf[MAX_K];
    float w[MAX_C * MAX_K];
    float inv_wsum[MAX_C];
    float dt;
};
constexpr int GENERIC_SMEM = 65536 + NPART_MAX * NT * 4 + CTRL_BYTES + SCRATCH_BYTES + TW_BYTES + (int)sizeof(GenericConsts);
constexpr int PLANE_F4 = 16 * NT;  // float4 per thread-private image

__global__ void __launch_bounds__(NTHREADS, 1) lnx_world128_generic(const RunArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float* part = reinterpret_cast<float*>(smem + 65536);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + 65536 + NPART_MAX * NT * 4);
    float2* scratch = reinterpret_cast<float2*>(smem + 65536 + NPART_MAX * NT * 4 + CTRL_BYTES);
    float4* twtab = reinterpret_cast<float4*>(smem + 65536 + NPART_MAX * NT * 4 + CTRL_BYTES + SCRATCH_BYTES);
    GenericConsts* gc = reinterpret_cast<GenericConsts*>(smem + 65536 + NPART_MAX * NT * 4 + CTRL_BYTES + SCRATCH_BYTES + TW_BYTES);

    const int tid = threadIdx.x;
    const int C = P.C, K = P.K;
    const int n_worlds = P.n_sols * P.n_init;
    const int npart = PT_FIXED + C;
    const bool early = (P.flags & LNX_RUN_EARLY_STOP) != 0;
    float4* Ast = P.scratch + (size_t)blockIdx.x * (3 * C) * PLANE_F4;  // [C] states
    float4* Sp = Ast + (size_t)C * PLANE_F4;                            // [C] forward spectra (P3 layout)
    float4* Fa = Sp + (size_t)C * PLANE_F4;                             // [C] field accumulators
    Regs R;
    init_twiddle_table(tid, twtab, c_tw128);

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
        if (tid < K) gc->gf[tid] = gf_prepare(P.gf_id[tid], P.gf_params[((size_t)sol * K + tid) * 2], P.gf_params[((size_t)sol * K + tid) * 2 + 1]);
        if (tid < C * K) gc->w[tid] = P.weights[(size_t)sol * C * K + tid];
        if (tid < C) {
            float sum = 0.f;
            for (int k = 0; k < K; ++k) sum += P.weights[((size_t)sol * C + tid) * K + k];
            gc->inv_wsum[tid] = P.mean ? 1.0f / sum : 1.0f;
        }
        if (tid == 0) gc->dt = P.dt[sol];
        __syncthreads();

        if (tid < NT) {
            const float dt = gc->dt;
            const int l = t_sub(tid) & 3;
            for (int c = 0; c < C; ++c) gather_state(Ast + (size_t)c * PLANE_F4, P.cells0 + ((size_t)world * C + c) * (WS * WS), tid);
            const float4* tab = P.table + (size_t)sol * K * KTAB_F4;
            for (int t = 0; t < P.max_iter; ++t) {
                const size_t tstep = ((size_t)sol * P.max_iter + t) * P.n_init + init;  // index of this world-step in trajectories
                // ---- forward transforms of every channel ----
                for (int c = 0; c < C; ++c) {
                    load_state_regs(R, Ast + (size_t)c * PLANE_F4, tid);
                    __syncwarp();
                    phase1(tid, R, W);
                    __syncwarp();
                    phase2_load(tid, R, W);
                    __syncwarp();
                    phase2_compute_store(tid, R, W, twtab);
                    bar_sync(BAR_COMPUTE, NT);
                    phase3_load_fft(tid, R, W);
                    float4* sp = Sp + (size_t)c * PLANE_F4;
#pragma unroll
                    for (int i = 0; i < 16; ++i) sp[i * NT + tid] = make_float4(R.v[2 * i].x, R.v[2 * i].y, R.v[2 * i + 1].x, R.v[2 * i + 1].y);
                    bar_sync(BAR_COMPUTE, NT);
                }
                // ---- one inverse transform per kernel, growth, accumulate into the target channels ----
                unsigned touched = 0;
                float cnt_p = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float4* sp = Sp + (size_t)P.c_in[k] * PLANE_F4;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 s4 = sp[i * NT + tid];
                        R.v[2 * i] = make_float2(s4.x, s4.y);
                        R.v[2 * i + 1] = make_float2(s4.z, s4.w);
                    }
                    if (tid < 32) {
                        phase3_col0_stash(tid, R, scratch);
                        __syncwarp();
                        phase3_col0_compute(tid, scratch, tab + (size_t)k * KTAB_F4 + KT_F4);
                        __syncwarp();
                    }
                    phase3_multiply(tid, R, tab + (size_t)k * KTAB_F4);
                    if (tid < 32) phase3_col0_fetch(tid, R, scratch);
                    phase3_ifft_store(tid, R, W);
                    bar_sync(BAR_COMPUTE, NT);
                    phase4_load(tid, R, W);
                    __syncwarp();
                    phase4_compute_store(tid, R, W, twtab);
                    __syncwarp();
                    phase5_load(tid, R, W);
                    bar_sync(BAR_COMPUTE, NT);  // W is free for the next kernel's spectrum
                    phase5_ifft(R);
                    if (P.potential_out) {
                        float* img = P.potential_out + (tstep * K + k) * (WS * WS);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {  // fully unrolled: R.v must keep compile-time indices (registers)
                            img[cell_row(tid, 0) * WS + 4 * j + l] = R.v[j].x;
                            img[cell_row(tid, 1) * WS + 4 * j + l] = R.v[j].y;
                        }
                    }
                    growth_vec_dyn<true, 32>(P.gf_id[k], R.v, gc->gf[k], cnt_p);
                    for (int c = 0; c < C; ++c) {
                        const float w = gc->w[c * K + k];
                        if (w == 0.f) continue;
                        float4* fa = Fa + (size_t)c * PLANE_F4;
                        const bool first = !(touched & (1u << c));
                        touched |= 1u << c;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
                            if (!first) {
                                f0 = fa[i * NT + tid];
                                f1 = fa[(8 + i) * NT + tid];
                            }
                            f0.x += w * R.v[4 * i + 0].x; f0.y += w * R.v[4 * i + 1].x; f0.z += w * R.v[4 * i + 2].x; f0.w += w * R.v[4 * i + 3].x;
                            f1.x += w * R.v[4 * i + 0].y; f1.y += w * R.v[4 * i + 1].y; f1.z += w * R.v[4 * i + 2].y; f1.w += w * R.v[4 * i + 3].y;
                            fa[i * NT + tid] = f0;
                            fa[(8 + i) * NT + tid] = f1;
                        }
                    }
                }
                if (t > 0) {
                    while (ld_acquire_smem(&ctrl->done) < t) {}
                    if (ctrl->stop) break;
                }
                // ---- state update + statistics partials ----
                const int sh0 = ctrl->shift0, sh1 = ctrl->shift1;
                const float xr0 = rolled_coord(cell_row(tid, 0), sh0), xr1 = rolled_coord(cell_row(tid, 1), sh0);
                const float cbase = (float)(((l - sh1) & (WS - 1)) - WS / 2);
                float mx_r = 0.f, mx2_r = 0.f, gx_r = 0.f, mxc = 0.f, mx2c = 0.f, gxc = 0.f, g00 = 0.f, cnt_a = 0.f, cnt_g = 0.f;
                for (int c = 0; c < C; ++c) {
                    float4* st = Ast + (size_t)c * PLANE_F4;
                    const float4* fa = Fa + (size_t)c * PLANE_F4;
                    const float inv = gc->inv_wsum[c];
                    const bool has = (touched >> c) & 1u;
                    float* cimg = P.cells_out ? P.cells_out + (tstep * C + c) * (WS * WS) : nullptr;
                    float* fimg = P.field_out ? P.field_out + (tstep * C + c) * (WS * WS) : nullptr;
                    CellAcc A;
                    A.clear();
#pragma unroll 2
                    for (int i = 0; i < 8; ++i) {
                        const float4 c0 = st[i * NT + tid], c1 = st[(8 + i) * NT + tid];
                        float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
                        if (has) {
                            f0 = fa[i * NT + tid];
                            f1 = fa[(8 + i) * NT + tid];
                        }
                        const float a0[4] = {c0.x, c0.y, c0.z, c0.w}, a1[4] = {c1.x, c1.y, c1.z, c1.w};
                        const float q0[4] = {f0.x * inv, f0.y * inv, f0.z * inv, f0.w * inv}, q1[4] = {f1.x * inv, f1.y * inv, f1.z * inv, f1.w * inv};
                        float n0[4], n1[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = 4 * i + e;
                            acc_cells(A, col_coord(cbase, j), a0[e], a1[e], q0[e], q1[e]);
                            n0[e] = state_update_dyn<true>(P.state_fn, a0[e], q0[e], dt);
                            n1[e] = state_update_dyn<true>(P.state_fn, a1[e], q1[e], dt);
                            if (cimg) {
                                cimg[cell_row(tid, 0) * WS + 4 * j + l] = a0[e];
                                cimg[cell_row(tid, 1) * WS + 4 * j + l] = a1[e];
                            }
                            if (fimg) {
                                fimg[cell_row(tid, 0) * WS + 4 * j + l] = q0[e];
                                fimg[cell_row(tid, 1) * WS + 4 * j + l] = q1[e];
                            }
                        }
                        st[i * NT + tid] = make_float4(n0[0], n0[1], n0[2], n0[3]);
                        st[(8 + i) * NT + tid] = make_float4(n1[0], n1[1], n1[2], n1[3]);
                    }
                    part[(PT_M00_C0 + c) * NT + tid] = A.sa0 + A.sa1;
                    mx_r += xr0 * A.sa0 + xr1 * A.sa1;
                    mx2_r += (xr0 * xr0) * A.sa0 + (xr1 * xr1) * A.sa1;
                    gx_r += xr0 * A.sg0 + xr1 * A.sg1;
                    mxc += A.mxc;
                    mx2c += A.mx2c;
                    gxc += A.gxc;
                    cnt_a += A.cnt_a;
                    cnt_g += A.cnt_g;
                    g00 += A.sg0 + A.sg1;
                }
                part[PT_CNT_A * NT + tid] = cnt_a;
                part[PT_G00 * NT + tid] = g00;
                part[PT_CNT_G * NT + tid] = cnt_g;
                part[PT_CNT_P * NT + tid] = cnt_p;
                part[PT_MX_R * NT + tid] = mx_r;
                part[PT_MX_C * NT + tid] = mxc;
                part[PT_MX2_R * NT + tid] = mx2_r;
                part[PT_MX2_C * NT + tid] = mx2c;
                part[PT_GX_R * NT + tid] = gx_r;
                part[PT_GX_C * NT + tid] = gxc;
                __threadfence_block();
                bar_arrive(BAR_PARTIALS, NTHREADS);
            }
            if (P.final_cells)
                for (int c = 0; c < C; ++c) scatter_state(P.final_cells + ((size_t)world * C + c) * (WS * WS), Ast + (size_t)c * PLANE_F4, tid);
        } else {
            const int lane = tid - NT;
            const float invR2 = 1.0f / (P.R * P.R), invR = 1.0f / P.R, inv_dt = 1.0f / P.stats_dt;
            StatsCarry S;
            S.reset();
            for (int t = 0; t < P.max_iter; ++t) {
                bar_sync(BAR_PARTIALS, NTHREADS);
                const float sc = stats_step(P, part, npart, lane, t, sol, init, S, ctrl, invR2, invR, inv_dt);
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

// ---------------------------------------------------------------------------------------------------------------------
// generic kernel, TMEM variant (default for C <= 4): any K <= 32, all growth / state functions, optional trajectory
//
// One CTA (256 threads) per world, one world per SM.  Per thread and channel, 64 floats of FIELD ACCUMULATOR live in tensor
// memory (the array that is read-modify-written once per kernel); the channel states sit in a per-CTA global scratch that
// stays L2 resident (read at the forward transform and at the update, written at the update); the spectrum of the current
// input channel is kept in shared memory for all the kernels that read it (kernels are sorted by input channel,
// leniax/kernels.py:90), and the multipliers of the NEXT kernel are prefetched into shared memory with cp.async while the
// current kernel's inverse transform runs.  Statistics as in lnx_world128_tm (no statistics warp, batched finaliser).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int G2_MAX_C = 4;
constexpr int G2_NPART = PT_FIXED + G2_MAX_C;
constexpr int G2_TM_COLS = 512;
constexpr int G2_OFF_SP = 65536;
constexpr int G2_OFF_KT = 131072;
constexpr int G2_OFF_PART = 196608;
constexpr int G2_OFF_RING = G2_OFF_PART + G2_NPART * NT * 4;
constexpr int G2_OFF_SCRATCH = G2_OFF_RING + RING_ROWS * RING_STRIDE_C * 4;
constexpr int G2_OFF_TW = G2_OFF_SCRATCH + SCRATCH_BYTES;
constexpr int G2_OFF_XT = G2_OFF_TW + TW_BYTES;
constexpr int G2_OFF_GC = G2_OFF_XT + XT_F4 * 16;
constexpr int G2_OFF_CTRL = G2_OFF_GC + (((int)sizeof(GenericConsts) + 15) / 16) * 16;
constexpr int G2_SMEM = G2_OFF_CTRL + 160;
static_assert(G2_SMEM <= 227 * 1024, "generic TMEM kernel: shared memory over the per-CTA limit");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// totals of one step -> ring row (C channels); warp 1 advances the shift carry and rebuilds the coordinate table
__device__ __forceinline__ void g2_reduce_partials(const float* part, float* row, TmCtrl* ctrl, float4* xt, int C, int warp, int lane) {
    if (warp == 1) {
        float m00 = 0.f;
        for (int c = 0; c < C; ++c) {
            const float m = tm_reduce_one(part, PT_M00_C0 + c, lane);
            if (lane == 0) row[RING_M00 + c] = m;
            m00 += m;
        }
        const float r = tm_reduce_one(part, PT_MX_R, lane), cc = tm_reduce_one(part, PT_MX_C, lane);
        const float im = sdiv(1.0f, m00 + EPS);
        const float c0 = r * im, c1 = cc * im;
        const int shift1 = (ctrl->shift1 + trunc_to_int(c1)) & (WS - 1);
        __syncwarp();
        if (lane == 0) {
            row[PT_MX_R] = r;
            row[PT_MX_C] = cc;
            row[RING_C0] = c0;
            row[RING_C1] = c1;
            ctrl->shift0 = (ctrl->shift0 + trunc_to_int(c0)) & (WS - 1);
            ctrl->shift1 = shift1;
        }
        xt_build(lane, shift1, xt);
    } else if (warp >= 2) {
        const int ka = warp < 6 ? warp - 2 : warp + 2;
        const float a = tm_reduce_one(part, ka, lane);
        if (lane == 0) row[ka] = a;
        if (warp < 4) {
            const float b = tm_reduce_one(part, warp + 4, lane);
            if (lane == 0) row[warp + 4] = b;
        }
    }
}

__global__ void __launch_bounds__(NT, 1) lnx_world128_gen_tm(const RunArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float4* SpBuf = reinterpret_cast<float4*>(smem + G2_OFF_SP);
    float4* KtBuf = reinterpret_cast<float4*>(smem + G2_OFF_KT);
    float* part = reinterpret_cast<float*>(smem + G2_OFF_PART);
    float* ring = reinterpret_cast<float*>(smem + G2_OFF_RING);
    float2* scratch = reinterpret_cast<float2*>(smem + G2_OFF_SCRATCH);
    float4* twtab = reinterpret_cast<float4*>(smem + G2_OFF_TW);
    float4* xt = reinterpret_cast<float4*>(smem + G2_OFF_XT);
    GenericConsts* gc = reinterpret_cast<GenericConsts*>(smem + G2_OFF_GC);
    TmCtrl* ctrl = reinterpret_cast<TmCtrl*>(smem + G2_OFF_CTRL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = P.C, K = P.K;
    const int n_worlds = P.n_sols * P.n_init;
    const bool early = (P.flags & LNX_RUN_EARLY_STOP) != 0;
    const float invR2 = 1.0f / (P.R * P.R), invR = 1.0f / P.R, inv_dt = 1.0f / P.stats_dt;
    const size_t plane = (size_t)P.n_sols * P.max_iter * P.n_init;
    const int l = t_sub(tid) & 3;
    float4* Ast = P.scratch + (size_t)blockIdx.x * C * PLANE_F4;  // [C] states: float4 [16][256], chunk i = float4 2i, 2i+1

    if (warp == 0) tm::alloc(&ctrl->tmem_base, G2_TM_COLS);
    init_twiddle_table(tid, twtab, c_tw128);
    tm::fence_before_sync();
    __syncthreads();
    tm::fence_after_sync();
    const uint32_t tbase = ctrl->tmem_base;
    const uint32_t acc0 = tm::warp_addr(tbase, warp, (warp >> 2) * 256);  // accumulator of channel c at acc0 + 64 c
    Regs R;

    for (;;) {
        if (tid == 0) {
            ctrl->world = atomicAdd(P.queue, 1);
            ctrl->shift0 = ctrl->shift1 = 0;
            ctrl->stop = 0;
            ctrl->carry.reset();
        }
        __syncthreads();
        const int world = ctrl->world;
        if (world >= n_worlds) break;
        const int sol = world / P.n_init, init = world - sol * P.n_init;
        if (tid < K) gc->gf[tid] = gf_prepare(P.gf_id[tid], P.gf_params[((size_t)sol * K + tid) * 2], P.gf_params[((size_t)sol * K + tid) * 2 + 1]);
        if (tid < C * K) gc->w[tid] = P.weights[(size_t)sol * C * K + tid];
        if (tid < C) {
            float sum = 0.f;
            for (int k = 0; k < K; ++k) sum += P.weights[((size_t)sol * C + tid) * K + k];
            gc->inv_wsum[tid] = P.mean ? 1.0f / sum : 1.0f;
        }
        if (tid == 0) gc->dt = P.dt[sol];
        if (warp == 1) xt_build(lane, 0, xt);
        for (int c = 0; c < C; ++c) {  // initial state -> scratch, chunk layout ((row p, row p+64) pairs)
            const float* r0 = P.cells0 + ((size_t)world * C + c) * (WS * WS) + cell_row(tid, 0) * WS + l;
            const float* r1 = P.cells0 + ((size_t)world * C + c) * (WS * WS) + cell_row(tid, 1) * WS + l;
            float4* st = Ast + (size_t)c * PLANE_F4;
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
                st[(2 * i) * NT + tid] = make_float4(__ldg(r0 + 16 * i), __ldg(r1 + 16 * i), __ldg(r0 + 16 * i + 4), __ldg(r1 + 16 * i + 4));
                st[(2 * i + 1) * NT + tid] = make_float4(__ldg(r0 + 16 * i + 8), __ldg(r1 + 16 * i + 8), __ldg(r0 + 16 * i + 12), __ldg(r1 + 16 * i + 12));
            }
        }
        const float4* tab = P.table + (size_t)sol * K * KTAB_F4;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) cp_async16(KtBuf + i * NT + tid, tab + i * NT + tid);  // multipliers of kernel 0
        cp_async_commit();
        __syncthreads();  // consts, coordinate table
        const float dt = gc->dt;
        const size_t idx_world = (size_t)sol * P.max_iter * P.n_init + init;

        int t = 0;
        bool stopped = false;
        for (; t < P.max_iter; ++t) {
            const size_t tstep = ((size_t)sol * P.max_iter + t) * P.n_init + init;  // world-step slot of the trajectory outputs
            unsigned touched = 0;
            float cnt_p = 0.f;
            for (int k = 0; k < K; ++k) {
                const int cin = P.c_in[k];
                if (k == 0 || cin != P.c_in[k - 1]) {
                    // ---- forward transform of input channel `cin` ----
                    const float4* st = Ast + (size_t)cin * PLANE_F4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 lo = st[(2 * i) * NT + tid], hi = st[(2 * i + 1) * NT + tid];
                        R.v[4 * i + 0] = make_float2(lo.x, lo.y);
                        R.v[4 * i + 1] = make_float2(lo.z, lo.w);
                        R.v[4 * i + 2] = make_float2(hi.x, hi.y);
                        R.v[4 * i + 3] = make_float2(hi.z, hi.w);
                    }
                    __syncwarp();
                    phase1(tid, R, W);
                    __syncwarp();
                    phase2_load(tid, R, W);
                    __syncwarp();
                    phase2_compute_store(tid, R, W, twtab);
                    __syncthreads();
                    if (k == 0 && t > 0) {
                        if (ctrl->stop) {
                            stopped = true;
                            break;
                        }
                        g2_reduce_partials(part, ring + ((t - 1) & (RING_ROWS - 1)) * RING_STRIDE_C, ctrl, xt, C, warp, lane);
                    }
                    phase3_load_fft(tid, R, W);
                    if (k + 1 < K && P.c_in[k + 1] == cin) {  // other kernels read this spectrum too
#pragma unroll
                        for (int i = 0; i < 16; ++i) SpBuf[i * NT + tid] = make_float4(R.v[2 * i].x, R.v[2 * i].y, R.v[2 * i + 1].x, R.v[2 * i + 1].y);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 s4 = SpBuf[i * NT + tid];
                        R.v[2 * i] = make_float2(s4.x, s4.y);
                        R.v[2 * i + 1] = make_float2(s4.z, s4.w);
                    }
                }
                // ---- multiply by kernel k, inverse transform ----
                cp_async_wait_all();  // this thread's multipliers of kernel k are in KtBuf
                if (tid < 32) {
                    phase3_col0_stash(tid, R, scratch);
                    __syncwarp();
                    phase3_col0_compute(tid, scratch, tab + (size_t)k * KTAB_F4 + KT_F4);
                    __syncwarp();
                }
                phase3_multiply(tid, R, KtBuf);
                if (tid < 32) phase3_col0_fetch(tid, R, scratch);
                {  // prefetch the next kernel's multipliers (thread-private slots: no barrier needed)
                    const int kn = k + 1 < K ? k + 1 : 0;
                    const float4* src = tab + (size_t)kn * KTAB_F4;
#pragma unroll 4
                    for (int i = 0; i < 16; ++i) cp_async16(KtBuf + i * NT + tid, src + i * NT + tid);
                    cp_async_commit();
                }
                phase3_ifft_store(tid, R, W);
                __syncthreads();
                if (k == 0 && warp == 7 && t > 0 && (t & (RING_ROWS - 1)) == 0) {  // rows t-32 .. t-1 are complete
                    BatchCarry S = ctrl->carry;
                    stats_finalize_batch<G2_MAX_C, RING_STRIDE_C>(ring, RING_ROWS, lane, C, P.stats, P.channel_mass, plane,
                                                                  idx_world + (size_t)S.rows * P.n_init, P.n_init, invR2, invR, inv_dt, S);
                    __syncwarp();
                    if (lane == 0) {
                        ctrl->carry = S;
                        if (early && S.should_continue == 0.f && S.rows >= 128) ctrl->stop = 1;
                    }
                }
                phase4_load(tid, R, W);
                __syncwarp();
                phase4_compute_store(tid, R, W, twtab);
                __syncwarp();
                phase5_load(tid, R, W);
                __syncthreads();  // W is free for the next transform
                phase5_ifft(R);
                if (P.potential_out) {
                    float* img = P.potential_out + (tstep * K + k) * (WS * WS);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        img[cell_row(tid, 0) * WS + 4 * j + l] = R.v[j].x;
                        img[cell_row(tid, 1) * WS + 4 * j + l] = R.v[j].y;
                    }
                }
                growth_vec_dyn<true, 32>(P.gf_id[k], R.v, gc->gf[k], cnt_p);
                for (int c = 0; c < C; ++c) {  // field accumulators (tensor memory), core.py:202-242
                    const float w = gc->w[c * K + k];
                    if (w == 0.f) continue;
                    const bool first = !(touched & (1u << c));
                    touched |= 1u << c;
                    const uint32_t aa = acc0 + 64 * c;
                    const float2 w2 = pk_bc(w);
                    float a[8][8];  // the whole accumulator at once: one wait instead of eight exposed TMEM latencies
                    if (!first) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) tm::ld8(aa + 8 * i, a[i]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) tm::wait_ld8(a[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int e = 0; e < 8; ++e) a[i][e] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 r = pk_fma(R.v[4 * i + e], w2, make_float2(a[i][2 * e], a[i][2 * e + 1]));
                            a[i][2 * e] = r.x;
                            a[i][2 * e + 1] = r.y;
                        }
                        tm::st8(aa + 8 * i, a[i]);
                    }
                    tm::wait_st();
                }
            }
            if (stopped) break;
            // ---- state update + statistics partials ----
            const int sh0 = ctrl->shift0;
            const float xr0 = rolled_coord(cell_row(tid, 0), sh0), xr1 = rolled_coord(cell_row(tid, 1), sh0);
            float mx_r = 0.f, mx2_r = 0.f, gx_r = 0.f, mxc = 0.f, mx2c = 0.f, gxc = 0.f, g00 = 0.f;
            int cnt_a = 0, cnt_g = 0;
            for (int c = 0; c < C; ++c) {
                float4* st = Ast + (size_t)c * PLANE_F4;
                const float2 inv2 = pk_bc(gc->inv_wsum[c]);
                const bool has = (touched >> c) & 1u;
                const uint32_t aa = acc0 + 64 * c;
                float* cimg = P.cells_out ? P.cells_out + (tstep * C + c) * (WS * WS) : nullptr;
                float* fimg = P.field_out ? P.field_out + (tstep * C + c) * (WS * WS) : nullptr;
                float2 sa = make_float2(0.f, 0.f), sg = sa, mx = sa, mx2 = sa, gx = sa;
                float4 sv[16];   // the whole state and accumulator of the channel first: all L2 / TMEM loads in flight together
                float fv[8][8];
#pragma unroll
                for (int i = 0; i < 16; ++i) sv[i] = st[i * NT + tid];
                if (has) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) tm::ld8(aa + 8 * i, fv[i]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) tm::wait_ld8(fv[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int e = 0; e < 8; ++e) fv[i][e] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 lo = sv[2 * i], hi = sv[2 * i + 1];
                    const float a[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    const float* f = fv[i];
                    const float4 x4 = xt[l * XT_STRIDE + i], q4 = xt[(4 + l) * XT_STRIDE + i];
                    const float xc[4] = {x4.x, x4.y, x4.z, x4.w}, xc2[4] = {q4.x, q4.y, q4.z, q4.w};
                    float n[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * i + e;
                        const float2 A = make_float2(a[2 * e], a[2 * e + 1]);
                        const float2 F = pk_mul(make_float2(f[2 * e], f[2 * e + 1]), inv2);
                        sa = pk_add(sa, A);
                        mx = pk_fma(A, pk_bc(xc[e]), mx);
                        mx2 = pk_fma(A, pk_bc(xc2[e]), mx2);
                        cnt_a += gt_bits(A.x, EPS) + gt_bits(A.y, EPS);
                        const float2 G = make_float2(fmaxf(F.x, 0.f), fmaxf(F.y, 0.f));
                        sg = pk_add(sg, G);
                        gx = pk_fma(G, pk_bc(xc[e]), gx);
                        cnt_g += gt_bits(F.x, EPS) + gt_bits(F.y, EPS);
                        n[2 * e] = state_update_dyn<true>(P.state_fn, A.x, F.x, dt);
                        n[2 * e + 1] = state_update_dyn<true>(P.state_fn, A.y, F.y, dt);
                        if (cimg) {
                            cimg[cell_row(tid, 0) * WS + 4 * j + l] = A.x;
                            cimg[cell_row(tid, 1) * WS + 4 * j + l] = A.y;
                        }
                        if (fimg) {
                            fimg[cell_row(tid, 0) * WS + 4 * j + l] = F.x;
                            fimg[cell_row(tid, 1) * WS + 4 * j + l] = F.y;
                        }
                    }
                    st[(2 * i) * NT + tid] = make_float4(n[0], n[1], n[2], n[3]);
                    st[(2 * i + 1) * NT + tid] = make_float4(n[4], n[5], n[6], n[7]);
                }
                part[(PT_M00_C0 + c) * NT + tid] = sa.x + sa.y;
                mx_r += xr0 * sa.x + xr1 * sa.y;
                mx2_r += (xr0 * xr0) * sa.x + (xr1 * xr1) * sa.y;
                gx_r += xr0 * sg.x + xr1 * sg.y;
                mxc += mx.x + mx.y;
                mx2c += mx2.x + mx2.y;
                gxc += gx.x + gx.y;
                g00 += sg.x + sg.y;
            }
            // counts: at most 64 C hits per thread and step, C <= 4 < 511 / 64
            part[PT_CNT_A * NT + tid] = count_from_bits(cnt_a);
            part[PT_G00 * NT + tid] = g00;
            part[PT_CNT_G * NT + tid] = count_from_bits(cnt_g);
            part[PT_CNT_P * NT + tid] = cnt_p;
            part[PT_MX_R * NT + tid] = mx_r;
            part[PT_MX_C * NT + tid] = mxc;
            part[PT_MX2_R * NT + tid] = mx2_r;
            part[PT_MX2_C * NT + tid] = mx2c;
            part[PT_GX_R * NT + tid] = gx_r;
            part[PT_GX_C * NT + tid] = gxc;
        }
        // the partial sums of the last completed update (step t-1) are not reduced yet; t >= 1 here
        cp_async_wait_all();
        __syncthreads();
        g2_reduce_partials(part, ring + ((t - 1) & (RING_ROWS - 1)) * RING_STRIDE_C, ctrl, xt, C, warp, lane);
        __syncthreads();
        if (warp == 7) {
            BatchCarry S = ctrl->carry;
            stats_finalize_batch<G2_MAX_C, RING_STRIDE_C>(ring, t - S.rows, lane, C, P.stats, P.channel_mass, plane,
                                                          idx_world + (size_t)S.rows * P.n_init, P.n_init, invR2, invR, inv_dt, S);
            if (lane == 0) P.n_alive[world] = S.n_alive;
        }
        if (P.final_cells) {
            for (int c = 0; c < C; ++c) {
                float* r0 = P.final_cells + ((size_t)world * C + c) * (WS * WS) + cell_row(tid, 0) * WS + l;
                float* r1 = P.final_cells + ((size_t)world * C + c) * (WS * WS) + cell_row(tid, 1) * WS + l;
                const float4* st = Ast + (size_t)c * PLANE_F4;
#pragma unroll 2
                for (int i = 0; i < 8; ++i) {
                    const float4 lo = st[(2 * i) * NT + tid], hi = st[(2 * i + 1) * NT + tid];
                    r0[16 * i] = lo.x;
                    r1[16 * i] = lo.y;
                    r0[16 * i + 4] = lo.z;
                    r1[16 * i + 4] = lo.w;
                    r0[16 * i + 8] = hi.x;
                    r1[16 * i + 8] = hi.y;
                    r0[16 * i + 12] = hi.z;
                    r1[16 * i + 12] = hi.w;
                }
            }
        }
        __syncthreads();  // world done
    }
    __syncthreads();
    if (warp == 0) tm::dealloc(tbase, G2_TM_COLS);
}

}  // namespace lnx

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace lnx;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define LNX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return fail(LNX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct lnx_plan {
    lnx_desc d;
    int device;
    int sm_count;
    bool tiled;            // false: resident 128x128 kernels, true: multi-pass tiled engine
    lnx::tiled::Geom g;    // tiled engine geometry
};


// ---------------------------------------------------------------------------------------------------------------------
// tiled engine: host side
// ---------------------------------------------------------------------------------------------------------------------
namespace th {
using namespace lnx::tiled;

static int ilog2i(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}
static bool make_geom(int nd, const int32_t* dims, Geom* g, const char** why) {
    memset(g, 0, sizeof(*g));
    if (nd != 2 && nd != 3) {
        *why = "only 2-D and 3-D worlds are supported";
        return false;
    }
    for (int d = 0; d < nd; ++d) {
        const int n = dims[d];
        if (n < 8 || n > NMAX || (n & (n - 1))) {
            *why = "every world dimension must be a power of two in [8, 4096]";
            return false;
        }
        g->dims[d] = n;
    }
    g->nd = nd;
    g->L = dims[0];
    g->A1 = nd == 3 ? dims[1] : 1;
    g->A2 = dims[nd - 1];
    g->logL = ilog2i(g->L);
    g->logA1 = ilog2i(g->A1);
    g->logA2 = ilog2i(g->A2);
    g->half = g->A2 / 2 + 1;
    g->rows = g->L * g->A1;
    g->cells = (long long)g->rows * g->A2;
    g->spec = (long long)g->rows * g->half;
    if (nd == 3) {
        g->slab_rows = g->A1;
    } else {
        int r = 4096 / g->A2;
        if (r < 2) r = 2;
        if (r > g->rows) r = g->rows;
        g->slab_rows = r;
    }
    g->n_slabs = g->rows / g->slab_rows;
    int tc = 8192 / g->L;
    if (tc < 4) tc = 4;
    if (tc > 64) tc = 64;
    g->tc = tc;
    return true;
}
static size_t tw_bytes(int logn) { return ((size_t)1 << (logn - 1)) * sizeof(float2); }  // shared-memory twiddle table of one pass
static int log_inner(const Geom& g) { return g.logA2 > g.logA1 ? g.logA2 : g.logA1; }
static size_t smem_a(const Geom& g) {
    return ((size_t)(g.slab_rows / 2) * g.A2 + (g.nd == 3 ? (size_t)g.A1 * g.half : 0)) * sizeof(float2) + tw_bytes(log_inner(g));
}
static size_t smem_b(const Geom& g, bool two_buf = true) { return (two_buf ? 2 : 1) * (size_t)g.L * g.tc * sizeof(float2) + tw_bytes(g.logL); }
static size_t smem_c(const Geom& g, int C) {
    return ((size_t)g.slab_rows * g.half + (size_t)(g.slab_rows / 2) * g.A2) * sizeof(float2) + (size_t)C * g.slab_rows * g.A2 * sizeof(float) +
           tw_bytes(log_inner(g));
}
constexpr size_t SMEM_LIMIT = 220 * 1024;  // dynamic part; pass C also has ~1 KB of static shared memory

static float2* g_tw[64] = {nullptr};  // library-owned twiddle master table per device: (cos, sin)(2 pi k / NMAX)
static int ensure_tiled_init(int dev) {
    if (g_tw[dev]) return LNX_OK;
    static float2 host[NMAX / 2];
    for (int k = 0; k < NMAX / 2; ++k) {
        const double a = 2.0 * 3.14159265358979323846 * k / NMAX;
        host[k] = make_float2((float)cos(a), (float)sin(a));
    }
    float2* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(host));
    if (e == cudaSuccess) e = cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pass_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pass_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    if (e != cudaSuccess) {
        fail(LNX_ERR_CUDA, "tiled engine setup failed: %s", cudaGetErrorString(e));
        return -1;
    }
    g_tw[dev] = d;
    return LNX_OK;
}
struct Workspace {   // carve-up of the caller's scratch for one lnx_run_scan call
    float* state;
    float2* spec;
    float2* pot;
    float* partials;
    WorldCarry* carry;
    size_t bytes;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static Workspace carve(const Geom& g, int C, int K, long long worlds, unsigned char* base) {
    Workspace w;
    size_t off = 256;
    w.state = reinterpret_cast<float*>(base + off);
    off += align256((size_t)worlds * C * g.cells * sizeof(float));
    w.spec = reinterpret_cast<float2*>(base + off);
    off += align256((size_t)worlds * C * g.spec * sizeof(float2));
    w.pot = reinterpret_cast<float2*>(base + off);
    off += align256((size_t)worlds * K * g.spec * sizeof(float2));
    w.partials = reinterpret_cast<float*>(base + off);
    off += align256((size_t)worlds * g.n_slabs * NP_T * sizeof(float));
    w.carry = reinterpret_cast<WorldCarry*>(base + off);
    off += align256((size_t)worlds * sizeof(WorldCarry));
    w.bytes = off;
    return w;
}
}  // namespace th

// one-time per-device setup: architecture check (no fallback), twiddle constants, dynamic shared memory opt-in
static int ensure_device_init(int* dev_out, int* sms_out) {
    static bool done[64] = {false};
    static int sms[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LNX_ERR_NO_DEVICE, "no CUDA device: %s (leniax_b200 has no CPU fallback)", cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= 64) return fail(LNX_ERR_INVALID, "device index %d out of range", dev);
    if (!done[dev]) {
        cudaDeviceProp prop;
        LNX_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10)
            return fail(LNX_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", dev, prop.major,
                        prop.minor);
        float2 tw[128];
        for (int k = 0; k < 128; ++k) tw[k] = make_float2(Tw128::c[k], Tw128::s[k]);
        LNX_CUDA(cudaMemcpyToSymbol(c_tw128, tw, sizeof(tw)));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_fused<GF_POLY_QUAD4, SF_V1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_fused<GF_POLY_QUAD4, SF_V1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_tm<GF_POLY_QUAD4, SF_V1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_tm<GF_POLY_QUAD4, SF_V1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_tm<GF_POLY_QUAD4, SF_V1, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_tm<GF_POLY_QUAD4, SF_V1, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, GENERIC_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_gen_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_r16<GF_POLY_QUAD4, SF_V1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, R16_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_world128_r16<GF_POLY_QUAD4, SF_V1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R16_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(lnx_rfft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + TW_BYTES));
        sms[dev] = prop.multiProcessorCount;
        done[dev] = true;
    }
    if (dev_out) *dev_out = dev;
    if (sms_out) *sms_out = sms[dev];
    return LNX_OK;
}

extern "C" {

int lnx_version(void) { return LNX_VERSION; }
const char* lnx_last_error(void) { return g_err; }

int lnx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int lnx_plan_create(const lnx_desc* d, lnx_plan** out) {
    if (!d || !out) return fail(LNX_ERR_INVALID, "lnx_plan_create: null argument");
    *out = nullptr;
    const bool resident = d->nb_dims == 2 && d->dims[0] == WS && d->dims[1] == WS && !(d->flags & LNX_PLAN_FORCE_TILED);
    lnx::tiled::Geom geom;
    memset(&geom, 0, sizeof(geom));
    if (!resident) {
        const char* why = "";
        if (!th::make_geom(d->nb_dims, d->dims, &geom, &why))
            return fail(LNX_ERR_UNSUPPORTED, "unsupported world shape (nb_dims=%d, dims=%d x %d x %d): %s", d->nb_dims, d->dims[0], d->dims[1],
                        d->nb_dims > 2 ? d->dims[2] : 1, why);
        if (th::smem_a(geom) > th::SMEM_LIMIT || th::smem_b(geom) > th::SMEM_LIMIT || th::smem_c(geom, d->nb_channels) > th::SMEM_LIMIT)
            return fail(LNX_ERR_UNSUPPORTED, "world too large for the tiled engine's shared-memory slabs (A %zu, B %zu, C %zu bytes)",
                        th::smem_a(geom), th::smem_b(geom), th::smem_c(geom, d->nb_channels));
    }
    if (d->nb_channels < 1 || d->nb_channels > MAX_C) return fail(LNX_ERR_INVALID, "nb_channels must be in [1, %d]", MAX_C);
    if (d->nb_kernels < 1 || d->nb_kernels > MAX_K) return fail(LNX_ERR_INVALID, "nb_kernels must be in [1, %d]", MAX_K);
    if (d->nb_slots < d->nb_kernels) return fail(LNX_ERR_INVALID, "nb_slots < nb_kernels");
    for (int k = 0; k < d->nb_kernels; ++k) {
        if (d->c_in[k] < 0 || d->c_in[k] >= d->nb_channels) return fail(LNX_ERR_INVALID, "c_in[%d] out of range", k);
        if (d->slot[k] < 0 || d->slot[k] >= d->nb_slots) return fail(LNX_ERR_INVALID, "slot[%d] out of range", k);
        if (d->gf_id[k] < 0 || d->gf_id[k] >= GF_COUNT) return fail(LNX_ERR_INVALID, "gf_id[%d]: unknown growth function", k);
    }
    if (d->state_fn < 0 || d->state_fn >= SF_COUNT) return fail(LNX_ERR_INVALID, "unknown state function %d", d->state_fn);
    if (!(d->R > 0.f) || !(d->stats_dt > 0.f)) return fail(LNX_ERR_INVALID, "R and stats_dt must be positive");

    int dev = 0, sms = 0;
    const int rc = ensure_device_init(&dev, &sms);
    if (rc != LNX_OK) return rc;
    lnx_plan* p = new (std::nothrow) lnx_plan;
    if (!p) return fail(LNX_ERR_INVALID, "out of host memory");
    p->d = *d;
    p->device = dev;
    p->sm_count = sms;
    p->tiled = !resident;
    p->g = geom;
    if (p->tiled && th::ensure_tiled_init(dev) != LNX_OK) {
        delete p;
        return LNX_ERR_CUDA;  // message set by ensure_tiled_init
    }
    *out = p;
    return LNX_OK;
}

int lnx_plan_destroy(lnx_plan* p) {
    if (!p) return LNX_OK;
    delete p;
    return LNX_OK;
}

size_t lnx_workspace_bytes(const lnx_plan* p);

size_t lnx_kernel_table_bytes(const lnx_plan* p) {
    if (!p) return 0;
    if (p->tiled) return (size_t)p->d.nb_kernels * p->g.spec * sizeof(float2);
    return (size_t)p->d.nb_kernels * KTAB_F4 * sizeof(float4);
}

size_t lnx_workspace_bytes_for(const lnx_plan* p, int32_t n_sols, int32_t n_init) {
    if (!p) return 0;
    if (!p->tiled) return lnx_workspace_bytes(p);
    return th::carve(p->g, p->d.nb_channels, p->d.nb_kernels, (long long)n_sols * n_init, nullptr).bytes;
}

size_t lnx_workspace_bytes(const lnx_plan* p) {
    if (!p) return 0;
    if (p->tiled) return th::carve(p->g, p->d.nb_channels, p->d.nb_kernels, 1, nullptr).bytes;
    // 256 B header (world queue counter) + per-CTA scratch of the generic kernel: [3][C] thread-private images
    return 256 + (size_t)p->sm_count * 3 * p->d.nb_channels * PLANE_F4 * sizeof(float4);
}

int lnx_kernels_prepare(const lnx_plan* p, int32_t n_sols, const void* K_fft, void* table, void* stream) {
    if (!p || !K_fft || !table || n_sols < 1) return fail(LNX_ERR_INVALID, "lnx_kernels_prepare: bad argument");
    if (p->tiled) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        int* slots_dev = nullptr;
        LNX_CUDA(cudaMallocAsync(&slots_dev, sizeof(int) * MAX_K, st));
        LNX_CUDA(cudaMemcpyAsync(slots_dev, p->d.slot, sizeof(int) * p->d.nb_kernels, cudaMemcpyHostToDevice, st));
        const dim3 grid((unsigned)((p->g.spec + 255) / 256 > 1024 ? 1024 : (p->g.spec + 255) / 256), p->d.nb_kernels, n_sols);
        lnx::tiled::gather_ktab_kernel<<<grid, 256, 0, st>>>(static_cast<const float2*>(K_fft), static_cast<float2*>(table), p->g,
                                                            p->d.nb_kernels, p->d.nb_slots, slots_dev, 1.0f / (float)p->g.cells);
        LNX_CUDA(cudaGetLastError());
        LNX_CUDA(cudaFreeAsync(slots_dev, st));
        return LNX_OK;
    }
    PrepArgs a;
    a.K_fft = static_cast<const float2*>(K_fft);
    a.table = static_cast<float4*>(table);
    a.K = p->d.nb_kernels;
    a.nb_slots = p->d.nb_slots;
    for (int k = 0; k < a.K; ++k) a.slot[k] = p->d.slot[k];
    lnx_prepare_kernel<<<n_sols * a.K, NT, 0, static_cast<cudaStream_t>(stream)>>>(a);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_rfft2(const lnx_plan* p, int32_t n_images, const float* images, void* spectra, void* stream) {
    (void)p;  // the plan is optional here
    if (!images || !spectra || n_images < 1) return fail(LNX_ERR_INVALID, "lnx_rfft2: bad argument");
    const int rc = ensure_device_init(nullptr, nullptr);
    if (rc != LNX_OK) return rc;
    lnx_rfft2_kernel<<<n_images, NT, 65536 + TW_BYTES, static_cast<cudaStream_t>(stream)>>>(images, static_cast<float2*>(spectra));
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_rfftn(int32_t nb_dims, const int32_t* dims, int32_t n_images, const float* images, void* spectra, void* stream) {
    if (!dims || !images || !spectra || n_images < 1) return fail(LNX_ERR_INVALID, "lnx_rfftn: bad argument");
    if (nb_dims == 2 && dims[0] == WS && dims[1] == WS) return lnx_rfft2(nullptr, n_images, images, spectra, stream);
    int dev = 0;
    int rc = ensure_device_init(&dev, nullptr);
    if (rc != LNX_OK) return rc;
    lnx::tiled::Geom g;
    const char* why = "";
    if (!th::make_geom(nb_dims, dims, &g, &why)) return fail(LNX_ERR_UNSUPPORTED, "lnx_rfftn: %s", why);
    if (th::smem_a(g) > th::SMEM_LIMIT || th::smem_b(g) > th::SMEM_LIMIT) return fail(LNX_ERR_UNSUPPORTED, "lnx_rfftn: world too large");
    if (th::ensure_tiled_init(dev) != LNX_OK) return LNX_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float2 *sa = nullptr, *sb = nullptr;
    const size_t bytes = (size_t)n_images * g.spec * sizeof(float2);
    LNX_CUDA(cudaMallocAsync(&sa, bytes, st));
    LNX_CUDA(cudaMallocAsync(&sb, bytes, st));
    lnx::tiled::PassAArgs a;
    a.state = images;
    a.spec = sa;
    a.tw = th::g_tw[dev];
    a.g = g;
    a.C = 1;
    lnx::tiled::pass_a_kernel<<<dim3(g.n_slabs, 1, n_images), lnx::tiled::TPB, th::smem_a(g), st>>>(a);
    lnx::tiled::PassBArgs b;
    memset(&b, 0, sizeof(b));
    b.spec = sa;
    b.fwd_out = sb;
    b.tw = th::g_tw[dev];
    b.g = g;
    b.C = 1;
    b.K = 0;
    b.n_init = 1;
    const long long M = g.spec / g.L;
    lnx::tiled::pass_b_kernel<<<dim3((unsigned)((M + g.tc - 1) / g.tc), 1, n_images), lnx::tiled::TPB, th::smem_b(g), st>>>(b);
    lnx::tiled::expand_hermitian_kernel<<<dim3(1024, 1, n_images), 256, 0, st>>>(sb, static_cast<float2*>(spectra), g);
    LNX_CUDA(cudaGetLastError());
    LNX_CUDA(cudaFreeAsync(sa, st));
    LNX_CUDA(cudaFreeAsync(sb, st));
    return LNX_OK;
}

int lnx_compute_stats(const lnx_plan* p, int32_t n_worlds, const float* cells, const float* field, const float* potential,
                      int32_t* total_shift_idx, float* mass_centroid, float* mass_angle, float* stats, float* channel_mass, void* stream) {
    using namespace lnx::tiled;
    if (!p || n_worlds < 1 || !cells || !field || !potential || !total_shift_idx || !mass_centroid || !mass_angle || !stats || !channel_mass)
        return fail(LNX_ERR_INVALID, "lnx_compute_stats: bad argument");
    if (n_worlds > 65535) return fail(LNX_ERR_INVALID, "lnx_compute_stats: at most 65535 worlds per call");
    Geom g;
    const char* why = "";
    if (!th::make_geom(p->d.nb_dims, p->d.dims, &g, &why)) return fail(LNX_ERR_UNSUPPORTED, "lnx_compute_stats: %s", why);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partials = nullptr;
    WorldCarry* carry = nullptr;
    float* n_alive = nullptr;
    LNX_CUDA(cudaMallocAsync(&partials, (size_t)n_worlds * g.n_slabs * NP_T * sizeof(float), st));
    LNX_CUDA(cudaMallocAsync(&carry, (size_t)n_worlds * sizeof(WorldCarry), st));
    LNX_CUDA(cudaMallocAsync(&n_alive, (size_t)n_worlds * sizeof(float), st));
    const int nd = g.nd, tb = 128, nb = (n_worlds + tb - 1) / tb;
    carry_pack_kernel<<<nb, tb, 0, st>>>(carry, total_shift_idx, mass_centroid, mass_angle, n_worlds, nd, false, nullptr, nullptr, nullptr);
    StatsPartialArgs a;
    a.cells = cells;
    a.field = field;
    a.potential = potential;
    a.carry = carry;
    a.partials = partials;
    a.g = g;
    a.C = p->d.nb_channels;
    a.K = p->d.nb_kernels;
    stats_partials_kernel<<<dim3(g.n_slabs, 1, n_worlds), TPB, 0, st>>>(a);
    PassDArgs d;
    d.partials = partials;
    d.carry = carry;
    d.stats = stats;
    d.channel_mass = channel_mass;
    d.n_alive = n_alive;
    d.g = g;
    d.C = p->d.nb_channels;
    d.n_sols = 1;
    d.n_init = n_worlds;
    d.max_iter = 1;
    d.t = 0;
    d.R = p->d.R;
    d.stats_dt = p->d.stats_dt;
    pass_d_kernel<<<n_worlds, 128, 0, st>>>(d);
    carry_pack_kernel<<<nb, tb, 0, st>>>(carry, nullptr, nullptr, nullptr, n_worlds, nd, true, total_shift_idx, mass_centroid, mass_angle);
    LNX_CUDA(cudaGetLastError());
    LNX_CUDA(cudaFreeAsync(partials, st));
    LNX_CUDA(cudaFreeAsync(carry, st));
    LNX_CUDA(cudaFreeAsync(n_alive, st));
    return LNX_OK;
}

int lnx_measure_fp32_peak(int32_t iters, double* tflops, double* ms, void* stream) {
    if (iters < 1 || !tflops) return fail(LNX_ERR_INVALID, "lnx_measure_fp32_peak: bad argument");
    int dev = 0, sms = 0;
    const int rc = ensure_device_init(&dev, &sms);
    if (rc != LNX_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* d_out = nullptr;
    LNX_CUDA(cudaMalloc(&d_out, 64));
    cudaEvent_t e0, e1;
    LNX_CUDA(cudaEventCreate(&e0));
    LNX_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 4, block = 512;
    lnx_fp32_peak_kernel<<<grid, block, 0, st>>>(d_out, iters / 8 + 1, 0.999f, 0.001f);  // warm-up
    LNX_CUDA(cudaEventRecord(e0, st));
    lnx_fp32_peak_kernel<<<grid, block, 0, st>>>(d_out, iters, 0.999f, 0.001f);
    LNX_CUDA(cudaEventRecord(e1, st));
    LNX_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    LNX_CUDA(cudaEventElapsedTime(&t, e0, e1));
    const double flops = 2.0 * 128.0 * (double)iters * (double)grid * (double)block;
    *tflops = flops / ((double)t * 1e-3) / 1e12;
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return LNX_OK;
}

static int run_scan_tiled(const lnx_plan* p, int32_t n_sols, int32_t n_init, int32_t max_run_iter, const float* cells0, const void* table,
                          const float* gf_params, const float* weights, const float* dt, float* stats, float* channel_mass, float* n_alive,
                          float* final_cells, float* cells_out, float* field_out, float* potential_out, void* workspace,
                          size_t workspace_bytes, void* stream) {
    using namespace lnx::tiled;
    const Geom& g = p->g;
    const int C = p->d.nb_channels, K = p->d.nb_kernels;
    const long long worlds = (long long)n_sols * n_init;
    if (worlds > 65535) return fail(LNX_ERR_INVALID, "tiled engine: at most 65535 worlds per call (got %lld)", worlds);
    const th::Workspace ws = th::carve(g, C, K, worlds, static_cast<unsigned char*>(workspace));
    if (!workspace || workspace_bytes < ws.bytes)
        return fail(LNX_ERR_INVALID, "lnx_run_scan: workspace too small (%zu < %zu); use lnx_workspace_bytes_for()", workspace_bytes, ws.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* tw = th::g_tw[p->device];
    float* state = final_cells ? final_cells : ws.state;  // working state (updated in place every step)
    LNX_CUDA(cudaMemcpyAsync(state, cells0, (size_t)worlds * C * g.cells * sizeof(float), cudaMemcpyDeviceToDevice, st));
    LNX_CUDA(cudaMemsetAsync(ws.carry, 0, (size_t)worlds * sizeof(WorldCarry), st));
    PassAArgs a;
    a.state = state;
    a.spec = ws.spec;
    a.tw = tw;
    a.g = g;
    a.C = C;
    PassBArgs b;
    memset(&b, 0, sizeof(b));
    b.spec = ws.spec;
    b.pot_spec = ws.pot;
    b.ktab = static_cast<const float2*>(table);
    b.fwd_out = nullptr;
    b.tw = tw;
    b.g = g;
    b.C = C;
    b.K = K;
    b.n_init = n_init;
    PassCArgs c;
    memset(&c, 0, sizeof(c));
    c.state = state;
    c.pot_spec = ws.pot;
    c.gf_params = gf_params;
    c.weights = weights;
    c.dt = dt;
    c.carry = ws.carry;
    c.partials = ws.partials;
    c.cells_out = cells_out;
    c.field_out = field_out;
    c.potential_out = potential_out;
    c.tw = tw;
    c.g = g;
    c.C = C;
    c.K = K;
    c.n_init = n_init;
    c.max_iter = max_run_iter;
    c.state_fn = p->d.state_fn;
    c.mean = p->d.weighted_average;
    PassDArgs d;
    d.partials = ws.partials;
    d.carry = ws.carry;
    d.stats = stats;
    d.channel_mass = channel_mass;
    d.n_alive = n_alive;
    d.g = g;
    d.C = C;
    d.n_sols = n_sols;
    d.n_init = n_init;
    d.max_iter = max_run_iter;
    d.R = p->d.R;
    d.stats_dt = p->d.stats_dt;
    int per_channel[MAX_C] = {0};
    for (int k = 0; k < K; ++k) {
        b.c_in[k] = p->d.c_in[k];
        c.gf_id[k] = p->d.gf_id[k];
        if (++per_channel[p->d.c_in[k]] > 1) b.two_buf = 1;
    }
    const long long M = g.spec / g.L;
    const dim3 grid_a(g.n_slabs, C, (unsigned)worlds), grid_b((unsigned)((M + g.tc - 1) / g.tc), C, (unsigned)worlds),
        grid_c(g.n_slabs, 1, (unsigned)worlds);
    for (int t = 0; t < max_run_iter; ++t) {
        c.t = t;
        d.t = t;
        pass_a_kernel<<<grid_a, TPB, th::smem_a(g), st>>>(a);
        pass_b_kernel<<<grid_b, TPB, th::smem_b(g, b.two_buf != 0), st>>>(b);
        pass_c_kernel<<<grid_c, TPB, th::smem_c(g, C), st>>>(c);
        pass_d_kernel<<<(unsigned)worlds, 128, 0, st>>>(d);
    }
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_update_conv(const lnx_desc* d, int32_t n_worlds, int32_t kh, int32_t kw, const float* state, const float* kernels,
                    const float* gf_params, const float* weights, float dt, float* state_out, float* field_out, float* potential_out,
                    void* stream) {
    if (!d || !state || !kernels || !gf_params || !weights || !state_out || !field_out || !potential_out)
        return fail(LNX_ERR_INVALID, "lnx_update_conv: null argument");
    if (d->nb_dims != 2) return fail(LNX_ERR_UNSUPPORTED, "lnx_update_conv: the direct-convolution potential is 2-D only (core.py:136: strides (1, 1))");
    if (n_worlds < 1 || n_worlds > 65535 || kh < 1 || kw < 1 || d->dims[0] < 1 || d->dims[1] < 1)
        return fail(LNX_ERR_INVALID, "lnx_update_conv: bad sizes");
    if (d->nb_channels < 1 || d->nb_channels > MAX_C || d->nb_kernels < 1 || d->nb_kernels > MAX_K)
        return fail(LNX_ERR_INVALID, "lnx_update_conv: C must be in [1, %d] and K in [1, %d]", MAX_C, MAX_K);
    if ((long long)n_worlds * d->nb_kernels > 65535) return fail(LNX_ERR_INVALID, "lnx_update_conv: n_worlds * K must be <= 65535");
    const int rc = ensure_device_init(nullptr, nullptr);
    if (rc != LNX_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    lnx::conv::ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.state = state;
    a.kernels = kernels;
    a.potential = potential_out;
    a.C = d->nb_channels;
    a.K = d->nb_kernels;
    a.H = d->dims[0];
    a.W = d->dims[1];
    a.kh = kh;
    a.kw = kw;
    lnx::conv::FieldArgs f;
    memset(&f, 0, sizeof(f));
    for (int k = 0; k < a.K; ++k) {
        if (d->c_in[k] < 0 || d->c_in[k] >= a.C || d->slot[k] < 0 || d->slot[k] >= d->nb_slots || d->gf_id[k] < 0 || d->gf_id[k] >= GF_COUNT)
            return fail(LNX_ERR_INVALID, "lnx_update_conv: kernel %d: bad c_in / slot / gf_id", k);
        a.slot[k] = d->slot[k];
        a.c_in[k] = d->c_in[k];
        f.gf_id[k] = d->gf_id[k];
    }
    lnx::conv::potential_kernel<<<dim3((a.W + 31) / 32, (a.H + 7) / 8, n_worlds * a.K), dim3(32, 8), 0, st>>>(a);
    f.state = state;
    f.potential = potential_out;
    f.gf_params = gf_params;
    f.weights = weights;
    f.state_out = state_out;
    f.field_out = field_out;
    f.cells = (long long)a.H * a.W;
    f.C = a.C;
    f.K = a.K;
    f.state_fn = d->state_fn;
    f.mean = d->weighted_average;
    f.dt = dt;
    lnx::conv::field_update_kernel<<<dim3((unsigned)((f.cells + 255) / 256), n_worlds), 256, 0, st>>>(f);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

static bool use_fused(const lnx_plan* p, bool trajectory) {
    const lnx_desc& d = p->d;
    return d.nb_channels == 1 && d.nb_kernels == 1 && !trajectory && d.gf_id[0] == GF_POLY_QUAD4 && d.state_fn == SF_V1;
}

const char* lnx_run_scan_variant(const lnx_plan* p, int32_t with_trajectory) {
    if (!p) return "";
    if (p->tiled) return "tiled";
    return use_fused(p, with_trajectory != 0) ? "fused" : "generic";
}

int lnx_run_scan(const lnx_plan* p, int32_t n_sols, int32_t n_init, int32_t max_run_iter, uint32_t run_flags, const float* cells0,
                 const void* table, const float* gf_params, const float* weights, const float* dt, float* stats, float* channel_mass,
                 float* n_alive, float* final_cells, float* cells_out, float* field_out, float* potential_out, void* workspace,
                 size_t workspace_bytes, void* stream) {
    if (!p) return fail(LNX_ERR_INVALID, "lnx_run_scan: null plan");
    if (n_sols < 1 || n_init < 1) return fail(LNX_ERR_INVALID, "lnx_run_scan: n_sols and n_init must be >= 1");
    if (max_run_iter < 1) return fail(LNX_ERR_INVALID, "max_run_iter must be positive, value given: %d", max_run_iter);  // runner.py:51
    if (!cells0 || !table || !gf_params || !weights || !dt || !stats || !channel_mass || !n_alive)
        return fail(LNX_ERR_INVALID, "lnx_run_scan: null required pointer");
    if (p->tiled)
        return run_scan_tiled(p, n_sols, n_init, max_run_iter, cells0, table, gf_params, weights, dt, stats, channel_mass, n_alive, final_cells,
                              cells_out, field_out, potential_out, workspace, workspace_bytes, stream);
    const bool trajectory = cells_out || field_out || potential_out;
    const bool fused = use_fused(p, trajectory);
    if (!workspace || workspace_bytes < (fused ? (size_t)256 : lnx_workspace_bytes(p)))
        return fail(LNX_ERR_INVALID, "lnx_run_scan: workspace too small (%zu < %zu)", workspace_bytes, lnx_workspace_bytes(p));

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RunArgs a;
    memset(&a, 0, sizeof(a));
    a.cells0 = cells0;
    a.table = static_cast<const float4*>(table);
    a.gf_params = gf_params;
    a.weights = weights;
    a.dt = dt;
    a.stats = stats;
    a.channel_mass = channel_mass;
    a.n_alive = n_alive;
    a.final_cells = final_cells;
    a.cells_out = cells_out;
    a.field_out = field_out;
    a.potential_out = potential_out;
    a.scratch = reinterpret_cast<float4*>(static_cast<unsigned char*>(workspace) + 256);
    a.queue = static_cast<int*>(workspace);
    a.n_sols = n_sols;
    a.n_init = n_init;
    a.max_iter = max_run_iter;
    a.C = p->d.nb_channels;
    a.K = p->d.nb_kernels;
    a.state_fn = p->d.state_fn;
    a.mean = p->d.weighted_average;
    a.R = p->d.R;
    a.stats_dt = p->d.stats_dt;
    a.flags = run_flags;
    for (int k = 0; k < a.K; ++k) {
        a.c_in[k] = p->d.c_in[k];
        a.gf_id[k] = p->d.gf_id[k];
    }
    LNX_CUDA(cudaMemsetAsync(a.queue, 0, sizeof(int), st));
    const long long n_worlds = (long long)n_sols * n_init;
    const int grid = (int)(n_worlds < p->sm_count ? n_worlds : p->sm_count);
    if (fused) {
        // NaN can only be born from s == 0 or a zero weight (0 * inf); the fast variant assumes neither.  The host
        // cannot see device-side parameters without a sync, so the NaN-propagating variant is the default and the
        // caller opts into the fast one with flag bit 8 (set by the Python layer after checking the parameters).
        const bool t32 = (run_flags & LNX_RUN_FUSED_R16) == 0;
        if (!(run_flags & (LNX_RUN_FUSED_R16 | LNX_RUN_FUSED_SMEM))) {  // default: TMEM-resident state, two worlds per SM
            const int grid2 = (int)(n_worlds < 2 * p->sm_count ? n_worlds : 2 * p->sm_count);
            if (run_flags & LNX_RUN_ASSUME_FINITE)
                lnx_world128_tm<GF_POLY_QUAD4, SF_V1, false><<<grid2, NT, TM_SMEM, st>>>(a);
            else
                lnx_world128_tm<GF_POLY_QUAD4, SF_V1, true><<<grid2, NT, TM_SMEM, st>>>(a);
        } else if (t32) {
            if (run_flags & LNX_RUN_ASSUME_FINITE)
                lnx_world128_fused<GF_POLY_QUAD4, SF_V1, false><<<grid, NTHREADS, FUSED_SMEM, st>>>(a);
            else
                lnx_world128_fused<GF_POLY_QUAD4, SF_V1, true><<<grid, NTHREADS, FUSED_SMEM, st>>>(a);
        } else {
            if (run_flags & LNX_RUN_ASSUME_FINITE)
                lnx_world128_r16<GF_POLY_QUAD4, SF_V1, false><<<grid, R16_THREADS, R16_SMEM, st>>>(a);
            else
                lnx_world128_r16<GF_POLY_QUAD4, SF_V1, true><<<grid, R16_THREADS, R16_SMEM, st>>>(a);
        }
    } else if (a.C <= G2_MAX_C && !(run_flags & LNX_RUN_FUSED_SMEM)) {
        lnx_world128_gen_tm<<<grid, NT, G2_SMEM, st>>>(a);
    } else {
        lnx_world128_generic<<<grid, NTHREADS, GENERIC_SMEM, st>>>(a);
    }
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

}  // extern "C"
