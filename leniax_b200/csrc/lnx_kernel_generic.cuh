// Generic resident kernels (any number of channels / kernels, every growth and state function, optional trajectory):
// lnx_world128_gen_tm (default, C <= 4, field accumulators in tensor memory) and the older lnx_world128_generic.
#pragma once
#include "lnx_kernel_tm.cuh"  // TmCtrl, tm_reduce_one

namespace lnx {

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
