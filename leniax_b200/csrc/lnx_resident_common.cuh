// Shared pieces of the resident 128x128 kernels: launch arguments, control block of the shared-memory kernels, kernel-table
// builder, stand-alone rfft2, FP32 peak probe, state gather / scatter and the per-step statistics warp of the older kernels.
#pragma once
#include "../../include/leniax_b200.h"
#include "lnx_step.cuh"
#include "lnx_step_r16.cuh"
#include "lnx_stats_batch.cuh"
#include "lnx_tmem.cuh"

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

}  // namespace lnx
