// lnx_world128_tm: the default fused kernel of the batched single-channel search (BASELINE config B).
#pragma once
#include "lnx_resident_common.cuh"

namespace lnx {

// ---------------------------------------------------------------------------------------------------------------------
// fused kernel, TMEM variant (default): C = K = 1, two worlds (CTAs) per SM
//
// The five phases of lnx_world128.cuh with the two thread-private arrays (state, kernel multipliers) in tensor memory
// instead of shared memory (lnx_tmem.cuh; round 1 started with them in shared memory: one world per SM), the new state stays in registers from the cell phase to phase 1 of the
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

// real spectrum (every circle_2d kernel: an even function): 32 real multipliers per thread, one packed multiply each
__device__ __forceinline__ void phase3_multiply_tm_real(Regs& R, uint32_t kt_addr, float (&k)[2][8]) {  // k[0] already in flight
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        tm::wait_ld8(k[c & 1]);
        if (c + 1 < 4) tm::ld8(kt_addr + 8 * (c + 1), k[(c + 1) & 1]);
#pragma unroll
        for (int e = 0; e < 8; ++e) R.v[8 * c + e] = pk_mul(R.v[8 * c + e], pk_bc(k[c & 1][e]));
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
    bool kreal = false;  // the loaded solution's spectrum is real (flag written by the table builder): multipliers in columns [0, 32) of its half
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
            kreal = __float_as_int(__ldg(reinterpret_cast<const float*>(src + KTAB_FLAG_F4))) != 0;
            if (kreal) {
#pragma unroll 4
                for (int i = 0; i < 8; ++i) tm::st4(kt_addr + 4 * i, __ldg(src + KTAB_REAL_F4 + i * NT + tid));
            } else {
#pragma unroll 4
                for (int i = 0; i < 16; ++i) tm::st4(kt_addr + 4 * i, __ldg(src + i * NT + tid));
            }
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
            if (kreal)                                         // (the plain products of the packed column are overwritten by the fetch below)
                phase3_multiply_tm_real(R, kt_addr, kbuf);
            else
                phase3_multiply_tm(R, kt_addr, kbuf);
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

}  // namespace lnx
