// lnx_world128_gen2: worlds with several channels / kernels, TWO worlds (CTAs) per SM (BASELINE config C: 3 channels, 6 kernels).
//
// lnx_world128_gen_tm (lnx_kernel_generic.cuh) keeps everything of one world on one SM - 192 KB of shared memory, all 512 tensor
// memory columns, 254 registers - and is latency-bound at two warps per scheduler (ncu: issue 0.44, FP32 pipe 38 %).  This kernel
// halves the on-chip footprint of a world so that two CTAs share an SM (four warps per scheduler from two INDEPENDENT worlds, as in
// lnx_world128_tm):
//   * registers: <= 128 per thread (every tensor-memory / L2 access in chunks of 8 floats, double-buffered);
//   * tensor memory: 256 columns = TWO field accumulators per thread.  The host orders the channel updates so that no more than two
//     accumulators are ever live: a channel is updated as soon as its last kernel has been added and its own spectrum has been
//     taken (schedule in RunArgs, built by gen2_schedule; 3c6k needs exactly two); kernel graphs that need more run in gen_tm;
//   * shared memory: 113 KB = the 64 KB exchange buffer + 32 KB of REAL multipliers of the current kernel (circle kernels are even
//     functions: their spectra are real; the table builder records per kernel whether the imaginary parts vanish, and a kernel
//     with a complex spectrum takes its multipliers straight from L2 instead) + partial sums pair-reduced by one shuffle;
//   * the channel states and the spectrum shared by the kernels of one input channel live in a per-CTA scratch that stays
//     L2-resident (64 KB each; read / written as thread-private 128-bit slots, no barrier).
// Statistics as in lnx_world128_tm: reduced behind the first barrier of the next step, batched finaliser every 32 steps.
// Same arithmetic as gen_tm (growth_vec_dyn, state_update_dyn, weight -> accumulate -> * 1 / sum(W)), hence the same results.
#pragma once
#include "lnx_kernel_tm.cuh"  // TmCtrl

namespace lnx {

constexpr int G3_MAX_C = 4;
constexpr int G3_MAX_K = 16;
constexpr int G3_NPART = PT_FIXED + G3_MAX_C;
constexpr int G3_PART_N = NT / 2;  // partial sums are added pairwise (one shuffle) before they go to shared memory
constexpr int G3_TM_COLS = 256;
constexpr int G3_RING_STRIDE = 16;  // floats per ring row: 12 channel-independent totals + G3_MAX_C channel masses
static_assert(RING_M00 + G3_MAX_C <= G3_RING_STRIDE, "gen2 ring row too short");

struct Gen2Consts {
    GfConst gf[G3_MAX_K];
    float w[G3_MAX_K];        // W[c_out[k]][k]
    float inv_wsum[G3_MAX_C];
    float dt;
    int kreal[G3_MAX_K];      // 1: the spectrum of kernel k is real (multipliers staged in shared memory)
};
constexpr int G3_OFF_KT = 65536;
constexpr int G3_OFF_PART = G3_OFF_KT + KREAL_F4 * 16;
constexpr int G3_OFF_KPQ = G3_OFF_PART + G3_NPART * G3_PART_N * 4;  // packed DC|Nyquist column multipliers of the current kernel
constexpr int G3_OFF_RING = G3_OFF_KPQ + KPQ_F4 * 16;
constexpr int G3_OFF_SCRATCH = G3_OFF_RING + RING_ROWS * G3_RING_STRIDE * 4;
constexpr int G3_OFF_TW = G3_OFF_SCRATCH + SCRATCH_BYTES;
constexpr int G3_OFF_XT = G3_OFF_TW + TW_BYTES;
constexpr int G3_OFF_GC = G3_OFF_XT + XT_F4 * 16;
constexpr int G3_OFF_CTRL = G3_OFF_GC + (((int)sizeof(Gen2Consts) + 15) / 16) * 16;
constexpr int G3_SMEM = G3_OFF_CTRL + 160;
static_assert(2 * (G3_SMEM + 1024) <= 228 * 1024, "gen2: two CTAs do not fit the shared memory of an SM");

__device__ __forceinline__ float g3_reduce_one(float* part, int k, int lane) {  // reads AND clears the array (next step adds into it)
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < G3_PART_N / 32; ++i) {
        a += part[k * G3_PART_N + lane + 32 * i];
        part[k * G3_PART_N + lane + 32 * i] = 0.f;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}
__device__ __forceinline__ void g3_reduce_partials(float* part, float* row, TmCtrl* ctrl, float4* xt, int C, int warp, int lane) {
    if (warp == 1) {
        float m00 = 0.f;
        for (int c = 0; c < C; ++c) {
            const float m = g3_reduce_one(part, PT_M00_C0 + c, lane);
            if (lane == 0) row[RING_M00 + c] = m;
            m00 += m;
        }
        const float r = g3_reduce_one(part, PT_MX_R, lane), cc = g3_reduce_one(part, PT_MX_C, lane);
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
        const float a = g3_reduce_one(part, ka, lane);
        if (lane == 0) row[ka] = a;
        if (warp < 4) {
            const float b = g3_reduce_one(part, warp + 4, lane);
            if (lane == 0) row[warp + 4] = b;
        }
    }
}
// v (this thread) + v (its neighbour) -> part[k][tid / 2] += ..., by the even lanes
__device__ __forceinline__ void g3_part_add(float* part, int k, int tid, float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (!(tid & 1)) part[k * G3_PART_N + (tid >> 1)] += v;
}

__device__ __forceinline__ void g3_mul_real(Regs& R, const float4* Kr, int tid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 k = Kr[i * NT + tid];
        R.v[4 * i + 0] = pk_mul(R.v[4 * i + 0], pk_bc(k.x));
        R.v[4 * i + 1] = pk_mul(R.v[4 * i + 1], pk_bc(k.y));
        R.v[4 * i + 2] = pk_mul(R.v[4 * i + 2], pk_bc(k.z));
        R.v[4 * i + 3] = pk_mul(R.v[4 * i + 3], pk_bc(k.w));
    }
}
__device__ __forceinline__ void g3_mul_complex_global(Regs& R, const float4* __restrict__ Kt, int tid) {  // multipliers from L2, 4 loads in flight
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float4 k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) k[i] = __ldg(Kt + (4 * b + i) * NT + tid);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int I = 4 * b + i;
            R.v[2 * I] = cmul(R.v[2 * I], make_float2(k[i].x, k[i].y));
            R.v[2 * I + 1] = cmul(R.v[2 * I + 1], make_float2(k[i].z, k[i].w));
        }
    }
}

// GF / SF >= 0: every kernel uses this growth function / the plan this state function (compile-time arithmetic, smaller loop);
// -1: selected per kernel / per launch at run time.  The loop must stay small: two CTAs stream it through one instruction cache.
// NP: propagate NaN through the clamps exactly like jnp.maximum / jnp.clip (needed when s == 0 or a zero weight row can occur); the
// caller passes LNX_RUN_ASSUME_FINITE when it cannot, and the clamps become single min / max / saturate instructions.
template <int GF, bool NP>
__device__ __forceinline__ void g3_growth(float2* v, const GfConst& g, float& cnt_p) {
    int bits = 0;  // potential > eps counts on the integer pipe (at most 64 hits: count_from_bits is exact up to 511)
    if constexpr (GF == GF_POLY_QUAD4) {  // the arithmetic of growth<GF_POLY_QUAD4, NP, false> on (row p, row p + 64) pairs
        const float2 nm = pk_bc(-g.m), nk0 = pk_bc(-g.k0);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            bits += gt_bits(v[j].x, EPS) + gt_bits(v[j].y, EPS);  // statistics.py:70
            const float2 t = pk_add(v[j], nm);
            float2 o = pk_fma(pk_mul(t, t), nk0, pk_bc(1.0f));
            if constexpr (NP)
                o = make_float2((o.x < 0.f) ? 0.f : o.x, (o.y < 0.f) ? 0.f : o.y);
            else
                o = make_float2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f));
            const float2 o2 = pk_mul(o, o);
            v[j] = pk_fma(pk_mul(o2, o2), pk_bc(2.0f), pk_bc(-1.0f));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            bits += gt_bits(v[j].x, EPS) + gt_bits(v[j].y, EPS);
            v[j].x = growth<GF, NP, true>(v[j].x, g);
            v[j].y = growth<GF, NP, true>(v[j].y, g);
        }
    }
    cnt_p += count_from_bits(bits);
}

template <int GF, int SF, bool NP>
__global__ void __launch_bounds__(NT, 2) lnx_world128_gen2(const RunArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float4* KtBuf = reinterpret_cast<float4*>(smem + G3_OFF_KT);
    float* part = reinterpret_cast<float*>(smem + G3_OFF_PART);
    float4* KpqBuf = reinterpret_cast<float4*>(smem + G3_OFF_KPQ);
    float* ring = reinterpret_cast<float*>(smem + G3_OFF_RING);
    float2* scratch = reinterpret_cast<float2*>(smem + G3_OFF_SCRATCH);
    float4* twtab = reinterpret_cast<float4*>(smem + G3_OFF_TW);
    float4* xt = reinterpret_cast<float4*>(smem + G3_OFF_XT);
    Gen2Consts* gc = reinterpret_cast<Gen2Consts*>(smem + G3_OFF_GC);
    TmCtrl* ctrl = reinterpret_cast<TmCtrl*>(smem + G3_OFF_CTRL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = P.C, K = P.K;
    const int n_worlds = P.n_sols * P.n_init;
    const bool early = (P.flags & LNX_RUN_EARLY_STOP) != 0;
    const float invR2 = 1.0f / (P.R * P.R), invR = 1.0f / P.R, inv_dt = 1.0f / P.stats_dt;
    const size_t plane = (size_t)P.n_sols * P.max_iter * P.n_init;
    const int l = t_sub(tid) & 3;
    float4* Ast = P.scratch + (size_t)blockIdx.x * (C + 1) * PLANE_F4;  // [C] states, chunk i = float4 2i, 2i+1; then one spectrum
    float4* Sp = Ast + (size_t)C * PLANE_F4;

    if (warp == 0) tm::alloc(&ctrl->tmem_base, G3_TM_COLS);
    init_twiddle_table(tid, twtab, c_tw128);
    tm::fence_before_sync();
    __syncthreads();
    tm::fence_after_sync();
    const uint32_t acc0 = tm::warp_addr(ctrl->tmem_base, warp, (warp >> 2) * 128);  // slot s at acc0 + 64 s
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
        const float4* tab = P.table + (size_t)sol * K * KTAB_F4;
        if (tid < K) {
            gc->gf[tid] = gf_prepare(P.gf_id[tid], P.gf_params[((size_t)sol * K + tid) * 2], P.gf_params[((size_t)sol * K + tid) * 2 + 1]);
            gc->w[tid] = P.c_out[tid] >= 0 ? P.weights[((size_t)sol * C + P.c_out[tid]) * K + tid] : 0.f;
            gc->kreal[tid] = __float_as_int(__ldg(reinterpret_cast<const float*>(tab + (size_t)tid * KTAB_F4 + KTAB_FLAG_F4))) != 0;
        }
        if (tid < C) {
            float sum = 0.f;
            for (int k = 0; k < K; ++k) sum += P.weights[((size_t)sol * C + tid) * K + k];
            gc->inv_wsum[tid] = P.mean ? 1.0f / sum : 1.0f;
        }
        if (tid == 0) gc->dt = P.dt[sol];
        if (warp == 1) xt_build(lane, 0, xt);
        for (int i = tid; i < G3_NPART * G3_PART_N; i += NT) part[i] = 0.f;
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
        __syncthreads();  // consts, coordinate table, cleared partial sums
        if (gc->kreal[0]) {  // multipliers of kernel 0
#pragma unroll 4
            for (int i = 0; i < 8; ++i) cp_async16(KtBuf + i * NT + tid, tab + KTAB_REAL_F4 + i * NT + tid);
        }
        if (tid < 32) {  // ... and of its packed column: staged by the warp that uses them (no CTA barrier needed)
#pragma unroll
            for (int i = 0; i < KPQ_F4 / 32; ++i) cp_async16(KpqBuf + i * 32 + tid, tab + KT_F4 + i * 32 + tid);
        }
        cp_async_commit();
        const float dt = gc->dt;
        const size_t idx_world = (size_t)sol * P.max_iter * P.n_init + init;

        int t = 0;
        bool stopped = false;
        for (; t < P.max_iter; ++t) {
            float cnt_p = 0.f;  // (a count: exact in fp32)
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
                        g3_reduce_partials(part, ring + ((t - 1) & (RING_ROWS - 1)) * G3_RING_STRIDE, ctrl, xt, C, warp, lane);
                    }
                    phase3_load_fft(tid, R, W);
                    if (k + 1 < K && P.c_in[k + 1] == cin) {  // other kernels read this spectrum too: thread-private slots of the L2 scratch
#pragma unroll
                        for (int i = 0; i < 16; ++i) Sp[i * NT + tid] = make_float4(R.v[2 * i].x, R.v[2 * i].y, R.v[2 * i + 1].x, R.v[2 * i + 1].y);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 s4 = Sp[i * NT + tid];
                        R.v[2 * i] = make_float2(s4.x, s4.y);
                        R.v[2 * i + 1] = make_float2(s4.z, s4.w);
                    }
                    __syncthreads();  // every warp has read the previous potential out of W (phase5_load): phase3_ifft_store may overwrite it
                }
                // ---- multiply by kernel k, inverse transform ----
                const float4* ktab = tab + (size_t)k * KTAB_F4;
                cp_async_wait_all();  // this thread's multipliers of kernel k are in KtBuf (if staged), warp 0's packed column in KpqBuf
                if (tid < 32) {
                    phase3_col0_stash(tid, R, scratch);
                    __syncwarp();
                    phase3_col0_compute(tid, scratch, KpqBuf);
                    __syncwarp();
                }
                if (gc->kreal[k])
                    g3_mul_real(R, KtBuf, tid);
                else
                    g3_mul_complex_global(R, ktab, tid);
                if (tid < 32) phase3_col0_fetch(tid, R, scratch);
                {  // stage the next kernel's multipliers (thread-private slots: no barrier needed)
                    const int kn = k + 1 < K ? k + 1 : 0;
                    if (gc->kreal[kn]) {
                        const float4* src = tab + (size_t)kn * KTAB_F4 + KTAB_REAL_F4;
#pragma unroll 4
                        for (int i = 0; i < 8; ++i) cp_async16(KtBuf + i * NT + tid, src + i * NT + tid);
                    }
                    if (tid < 32) {
#pragma unroll
                        for (int i = 0; i < KPQ_F4 / 32; ++i) cp_async16(KpqBuf + i * 32 + tid, tab + (size_t)kn * KTAB_F4 + KT_F4 + i * 32 + tid);
                    }
                    cp_async_commit();
                }
                phase3_ifft_store(tid, R, W);
                __syncthreads();
                if (k == 0 && warp == 7 && t > 0 && (t & (RING_ROWS - 1)) == 0) {  // rows t-32 .. t-1 are complete
                    BatchCarry S = ctrl->carry;
                    stats_finalize_batch<G3_MAX_C, G3_RING_STRIDE>(ring, RING_ROWS, lane, C, P.stats, P.channel_mass, plane,
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
                // (no CTA barrier here: P4 -> P5 and the next P1 -> P2 exchange inside the half-warp's own region of W; the next write
                // to OTHER regions is phase3_ifft_store, behind the P2 -> P3 barrier or the one in front of it below)
                phase5_ifft(R);
                if constexpr (GF >= 0)
                    g3_growth<GF, NP>(R.v, gc->gf[k], cnt_p);
                else
                    growth_vec_dyn<true, 32>(P.gf_id[k], R.v, gc->gf[k], cnt_p);
                if (P.c_out[k] >= 0) {  // field accumulator (tensor memory), core.py:202-242
                    const uint32_t aa = acc0 + 64 * P.acc_slot[k];
                    const float2 w2 = pk_bc(gc->w[k]);
                    if ((P.acc_first >> k) & 1u) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float a[8];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 r = pk_mul(R.v[4 * i + e], w2);  // (= gen_tm's fused multiply-add into a zero accumulator)
                                a[2 * e] = r.x;
                                a[2 * e + 1] = r.y;
                            }
                            tm::st8(aa + 8 * i, a);
                        }
                    } else {
                        float buf[2][8];
                        tm::wait_st();  // the stores of the previous kernels (issued one transform ago: long complete)
                        tm::ld8(aa, buf[0]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float* a = buf[i & 1];
                            tm::wait_ld8(a);
                            if (i + 1 < 8) tm::ld8(aa + 8 * (i + 1), buf[(i + 1) & 1]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 r = pk_fma(R.v[4 * i + e], w2, make_float2(a[2 * e], a[2 * e + 1]));
                                a[2 * e] = r.x;
                                a[2 * e + 1] = r.y;
                            }
                            tm::st8(aa + 8 * i, a);
                        }
                    }
                }
                // ---- state update + statistics partials of the channels that are complete now ----
                for (unsigned um = P.upd_mask[k]; um; um &= um - 1) {
                    const int c = __ffs(um) - 1;
                    const int sh0 = ctrl->shift0;
                    const float xr0 = rolled_coord(cell_row(tid, 0), sh0), xr1 = rolled_coord(cell_row(tid, 1), sh0);
                    float4* st = Ast + (size_t)c * PLANE_F4;
                    const float2 inv2 = pk_bc(gc->inv_wsum[c]);
                    const int slot = P.chan_slot[c];
                    const uint32_t aa = acc0 + 64 * (slot < 0 ? 0 : slot);
                    float2 sa = make_float2(0.f, 0.f), sg = sa, mx = sa, mx2 = sa, gx = sa;
                    int cnt_a = 0, cnt_g = 0;
                    // rolled loop over the 8 chunks, two per iteration (the code of this phase is fetched every step by both CTAs of the
                    // SM): chunk i + 2 is requested from L2 into the registers of chunk i as soon as that chunk has been consumed, the
                    // tensor-memory load runs one chunk ahead
                    float4 sv[2][2] = {{st[0 * NT + tid], st[1 * NT + tid]}, {st[2 * NT + tid], st[3 * NT + tid]}};
                    float fb[2][8];
                    if (slot >= 0) {
                        tm::wait_st();  // this kernel's accumulator stores
                        tm::ld8(aa, fb[0]);
                    }
#pragma unroll 1
                    for (int i = 0; i < 8; i += 2) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int ci = i + h;
                            float* f = fb[h];
                            if (slot >= 0) {
                                tm::wait_ld8(f);
                                if (ci + 1 < 8) tm::ld8(aa + 8 * (ci + 1), fb[h ^ 1]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = 0.f;
                            }
                            const float a[8] = {sv[h][0].x, sv[h][0].y, sv[h][0].z, sv[h][0].w, sv[h][1].x, sv[h][1].y, sv[h][1].z, sv[h][1].w};
                            const float4 x4 = xt[l * XT_STRIDE + ci], q4 = xt[(4 + l) * XT_STRIDE + ci];
                            const float xc[4] = {x4.x, x4.y, x4.z, x4.w}, xc2[4] = {q4.x, q4.y, q4.z, q4.w};
                            float n[8];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
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
                                if constexpr (SF == SF_V1 && !NP) {  // clip(a + dt f, 0, 1) as one saturating FMA (no NaN possible here)
                                    n[2 * e] = saturate01(A.x + dt * F.x);
                                    n[2 * e + 1] = saturate01(A.y + dt * F.y);
                                } else if constexpr (SF >= 0) {
                                    n[2 * e] = state_update<SF, NP>(A.x, F.x, dt);
                                    n[2 * e + 1] = state_update<SF, NP>(A.y, F.y, dt);
                                } else {
                                    n[2 * e] = state_update_dyn<true>(P.state_fn, A.x, F.x, dt);
                                    n[2 * e + 1] = state_update_dyn<true>(P.state_fn, A.y, F.y, dt);
                                }
                            }
                            st[(2 * ci) * NT + tid] = make_float4(n[0], n[1], n[2], n[3]);
                            st[(2 * ci + 1) * NT + tid] = make_float4(n[4], n[5], n[6], n[7]);
                            if (ci + 2 < 8) {
                                sv[h][0] = st[(2 * ci + 4) * NT + tid];
                                sv[h][1] = st[(2 * ci + 5) * NT + tid];
                            }
                        }
                    }
                    g3_part_add(part, PT_M00_C0 + c, tid, sa.x + sa.y);
                    g3_part_add(part, PT_MX_R, tid, xr0 * sa.x + xr1 * sa.y);
                    g3_part_add(part, PT_MX2_R, tid, (xr0 * xr0) * sa.x + (xr1 * xr1) * sa.y);
                    g3_part_add(part, PT_GX_R, tid, xr0 * sg.x + xr1 * sg.y);
                    g3_part_add(part, PT_MX_C, tid, mx.x + mx.y);
                    g3_part_add(part, PT_MX2_C, tid, mx2.x + mx2.y);
                    g3_part_add(part, PT_GX_C, tid, gx.x + gx.y);
                    g3_part_add(part, PT_G00, tid, sg.x + sg.y);
                    g3_part_add(part, PT_CNT_A, tid, count_from_bits(cnt_a));
                    g3_part_add(part, PT_CNT_G, tid, count_from_bits(cnt_g));
                }
            }
            if (stopped) break;
            g3_part_add(part, PT_CNT_P, tid, cnt_p);
        }
        // the partial sums of the last completed update (step t-1) are not reduced yet; t >= 1 here
        cp_async_wait_all();
        __syncthreads();
        g3_reduce_partials(part, ring + ((t - 1) & (RING_ROWS - 1)) * G3_RING_STRIDE, ctrl, xt, C, warp, lane);
        __syncthreads();
        if (warp == 7) {
            BatchCarry S = ctrl->carry;
            stats_finalize_batch<G3_MAX_C, G3_RING_STRIDE>(ring, t - S.rows, lane, C, P.stats, P.channel_mass, plane,
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
    tm::wait_st();
    __syncthreads();
    if (warp == 0) tm::dealloc(ctrl->tmem_base, G3_TM_COLS);
}

// Host side: order of the channel updates and the tensor-memory slot of every accumulator.  false = this kernel graph needs more than
// two live accumulators (or is too large): the caller launches lnx_world128_gen_tm instead.
inline bool gen2_schedule(int C, int K, const int* c_in, const int* c_out, RunArgs& a) {
    if (C > G3_MAX_C || K > G3_MAX_K) return false;
    int first_in[MAX_C], last_out[MAX_C], upd_at[MAX_C];
    for (int c = 0; c < C; ++c) first_in[c] = last_out[c] = -1;
    for (int k = 0; k < K; ++k) {
        if (c_out[k] == LNX_COUT_ANY || c_out[k] >= C) return false;  // undeclared weight pattern
        if (k > 0 && c_in[k] < c_in[k - 1]) return false;             // kernels must arrive sorted by input channel (kernels.py:90)
        if (first_in[c_in[k]] < 0) first_in[c_in[k]] = k;
        if (c_out[k] >= 0) last_out[c_out[k]] = k;
    }
    for (int k = 0; k < K; ++k) a.upd_mask[k] = 0;
    for (int c = 0; c < C; ++c) {
        // after its last kernel has been added AND its own spectrum has been taken (every reader uses the saved spectrum)
        int u = last_out[c] > first_in[c] ? last_out[c] : first_in[c];
        if (u < 0) u = K - 1;  // neither read nor written: field 0, any time
        upd_at[c] = u;
        a.upd_mask[u] |= (unsigned char)(1u << c);
        a.chan_slot[c] = -1;
    }
    int owner[2] = {-1, -1};  // channel whose accumulator occupies the slot
    a.acc_first = 0;
    for (int k = 0; k < K; ++k) {
        const int c = c_out[k];
        a.c_out[k] = (signed char)(c >= 0 ? c : -1);
        a.acc_slot[k] = 0;
        if (c >= 0) {
            int s = owner[0] == c ? 0 : (owner[1] == c ? 1 : -1);
            if (s < 0) {
                s = owner[0] < 0 ? 0 : (owner[1] < 0 ? 1 : -1);
                if (s < 0) return false;  // a third accumulator would be live
                owner[s] = c;
                a.chan_slot[c] = (signed char)s;
                a.acc_first |= 1u << k;
            }
            a.acc_slot[k] = (signed char)s;
        }
        for (int cc = 0; cc < C; ++cc)
            if (upd_at[cc] == k && a.chan_slot[cc] >= 0 && owner[a.chan_slot[cc]] == cc) owner[a.chan_slot[cc]] = -1;
    }
    return true;
}

}  // namespace lnx
