// Shared pieces of the resident 128x128 kernels: launch arguments, control block of the shared-memory kernel, state gather /
// scatter and the per-step statistics warp of the older generic kernel.
#pragma once
#include "../../include/leniax_b200.h"
#include "lnx_step.cuh"
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
constexpr int KREAL_F4 = 8 * NT;         // float4 per kernel: the real parts of the 32 multipliers of every thread (lnx_world128_gen2)
constexpr int KTAB_REAL_F4 = KT_F4 + KPQ_F4;          // offsets inside the table of one (solution, kernel): complex multipliers,
constexpr int KTAB_FLAG_F4 = KTAB_REAL_F4 + KREAL_F4; // packed DC|Nyquist column, real multipliers, flag word (1: the spectrum is real)
constexpr int KTAB_F4 = KTAB_FLAG_F4 + 4;
constexpr int PLANE_F4 = 16 * NT;  // float4 per thread-private image (L2 scratch of the multi-channel kernels)
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

static __constant__ float2 c_tw128[128];  // one copy per translation unit, filled by its setup function

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
    // schedule of lnx_world128_gen2 (host-built from the plan's c_in / c_out, lnx_kernel_gen2.cuh)
    signed char c_out[MAX_K];      // output channel of kernel k, < 0: none
    signed char acc_slot[MAX_K];   // tensor-memory accumulator slot (0 / 1) kernel k adds into
    unsigned char upd_mask[MAX_K]; // channels whose state is updated after kernel k
    signed char chan_slot[MAX_C];  // slot holding channel c's field at its update, < 0: no kernel feeds it (field 0)
    unsigned acc_first;            // bit k: kernel k is the first to touch its slot in this step (plain store instead of add)
};

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
