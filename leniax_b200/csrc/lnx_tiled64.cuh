// Thread-per-line engine for 64 x 64 x 64 worlds (BASELINE config E): the three passes of lnx_tiled.cuh with every 64-point
// transform done by ONE thread in registers (fft_dif<64> / ifft_dit<64>, compile-time twiddles and indices), so a Lenia step
// has no shared-memory butterfly pass and no CTA-wide barrier at all:
//
//   plane_fwd   one warp per plane (64 x 64 reals): lane j transforms the packed row pair (2j, 2j+1), untangles the two real
//               rows in its registers, the plane is exchanged once through shared memory, lane k transforms spectral column k
//               along axis 1 (columns 0 and 32 are real sequences: packed into one complex column, lane 0)
//   lead        one thread per (axis-1, axis-2) spectral column: 64 coalesced loads along the leading axis, forward transform,
//               multiply by K (pre-scaled by 1 / cells), inverse transform, 64 coalesced stores
//   plane_inv   mirror of plane_fwd; the potential plane goes back through shared memory so that growth / weighted mix / update /
//               statistics partials run on coalesced 128-bit accesses of the state (one channel, one kernel)
//
// HBM layouts, the kernel table, the per-plane statistics partials and pass D are those of lnx_tiled.cuh (natural-order half
// spectrum [rows][33]), so the two engines are interchangeable and the generic one remains the fallback for every other shape.
// The per-lane phase functions are __host__ __device__: tests/emul/lnx_t64_emul.cu runs them lane by lane on the CPU.
// Reference: leniax/core.py:52-102 (n-D FFT potential), :163-319, leniax/statistics.py:36-126.
#pragma once
#include "lnx_tiled.cuh"

namespace lnx {
namespace t64 {

using tiled::MAXD;
using tiled::NP_T;
using tiled::PassAArgs;
using tiled::PassBArgs;
using tiled::PassCArgs;
using tiled::WorldCarry;

constexpr int N = 64, HALF = 33;
constexpr int PLS = 33;  // row stride (complex) of a spectrum plane in shared memory = the global layout; odd: column accesses are conflict-free
constexpr int SRS = 68;  // row stride (floats) of a real plane in shared memory: 128-bit accesses by lanes owning one row each are conflict-free
constexpr int PLANE_SPEC = N * HALF;   // complex values in one plane of the half spectrum
constexpr int PLANE_CELLS = N * N;
constexpr int COLS = N * HALF;         // spectral columns along the leading axis per (world, channel)
constexpr int SMEM_FLOATS = N * SRS;   // one buffer serves both views (64 * 33 * 2 = 4224 <= 4352 floats)
constexpr int LEAD_TPB = 64;

#ifdef __CUDA_ARCH__
#define LNX_T64_LDG(p) __ldg(p)
#else
#define LNX_T64_LDG(p) (*(p))
#endif

LNX_HDC int br6(int x) { return ((x & 1) << 5) | ((x & 2) << 3) | ((x & 4) << 1) | ((x & 8) >> 1) | ((x & 16) >> 3) | ((x & 32) >> 5); }
LNX_HDC int rho(int r) { return (r >> 1) + 32 * (r & 1); }  // physical row of logical row r in shared memory: even rows first

// v[j] = Z[br6(j)] with Z the transform of x + i y (x, y real).  Afterwards v[br6(k)] = X[k], v[br6(64 - k)] = Y[k] for k = 1..31;
// v[0] = (X[0], Y[0]) and v[1] = (X[32], Y[32]) need no work (those four values are real).
template <int K>
LNX_HD void untangle(float2* v) {
    if constexpr (K < 32) {
        constexpr int a = br6(K), b = br6(64 - K);
        const float2 zk = v[a], zc = v[b];
        v[a] = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
        v[b] = make_float2(0.5f * (zk.y + zc.y), 0.5f * (zc.x - zk.x));
        untangle<K + 1>(v);
    }
}
// inverse of untangle without the 1/2: Z[k] = X[k] + i Y[k], Z[64 - k] = conj(X[k]) + i conj(Y[k])
template <int K>
LNX_HD void retangle(float2* v) {
    if constexpr (K < 32) {
        constexpr int a = br6(K), b = br6(64 - K);
        const float2 X = v[a], Y = v[b];
        v[a] = make_float2(X.x - Y.y, X.y + Y.x);
        v[b] = make_float2(X.x + Y.y, Y.x - X.y);
        retangle<K + 1>(v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// plane_fwd phases (lane = 0..31 of the warp that owns the plane)
// ---------------------------------------------------------------------------------------------------------------------
// coalesced copy of the real plane into shared memory, even rows first
LNX_HD void fwd_load(int lane, const float* __restrict__ src, float* sm) {
    const int hi = lane >> 4, n0 = (lane & 15) * 4;
#pragma unroll 8
    for (int it = 0; it < 32; ++it) {
        const float4 x = LNX_T64_LDG(reinterpret_cast<const float4*>(src + it * 128 + lane * 4));  // row 2 it + hi
        *reinterpret_cast<float4*>(sm + (it + 32 * hi) * SRS + n0) = x;
    }
}
// rows (2 lane, 2 lane + 1) as one complex line: transform, untangle
LNX_HD void fwd_rows(int lane, const float* sm, float2* v) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 ra = *reinterpret_cast<const float4*>(sm + lane * SRS + 4 * q);
        const float4 rb = *reinterpret_cast<const float4*>(sm + (32 + lane) * SRS + 4 * q);
        v[4 * q + 0] = make_float2(ra.x, rb.x);
        v[4 * q + 1] = make_float2(ra.y, rb.y);
        v[4 * q + 2] = make_float2(ra.z, rb.z);
        v[4 * q + 3] = make_float2(ra.w, rb.w);
    }
    fft_dif<64>(v);
    untangle<1>(v);
}
// spectrum plane [physical row][k], k = 1..31 plain, column 0 = (X[0], X[32]) of the row (both real)
LNX_HD void fwd_rows_store(int lane, float2* pl, const float2* v) {
    float2* ra = pl + lane * PLS;
    float2* rb = pl + (32 + lane) * PLS;
    ra[0] = make_float2(v[0].x, v[1].x);
    rb[0] = make_float2(v[0].y, v[1].y);
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        ra[k] = v[br6(k)];
        rb[k] = v[br6(64 - k)];
    }
}
// lane k: column k along axis 1 (lane 0: the packed column), results to the global half spectrum of the plane [m1][33]
LNX_HD void fwd_cols(int lane, const float2* pl, float2* __restrict__ dst) {
    float2 u[64];
#pragma unroll
    for (int r = 0; r < 64; ++r) u[r] = pl[rho(r) * PLS + lane];
    fft_dif<64>(u);
    if (lane != 0) {
#pragma unroll
        for (int m = 0; m < 64; ++m) dst[m * HALF + lane] = u[br6(m)];
    } else {
        untangle<1>(u);  // u[br6(m)] = column 0, u[br6(64 - m)] = column 32 at axis-1 frequency m = 1..31; conjugates above 32
        dst[0] = make_float2(u[0].x, 0.f);
        dst[32] = make_float2(u[0].y, 0.f);
        dst[32 * HALF] = make_float2(u[1].x, 0.f);
        dst[32 * HALF + 32] = make_float2(u[1].y, 0.f);
#pragma unroll
        for (int m = 1; m < 32; ++m) {
            const float2 f0 = u[br6(m)], f32 = u[br6(64 - m)];
            dst[m * HALF] = f0;
            dst[(64 - m) * HALF] = make_float2(f0.x, -f0.y);
            dst[m * HALF + 32] = f32;
            dst[(64 - m) * HALF + 32] = make_float2(f32.x, -f32.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// lead: one thread per spectral column (a1, k) of one (world, channel): src / dst / kt point at element [l = 0][col]
// ---------------------------------------------------------------------------------------------------------------------
LNX_HD void lead_load_fwd(const float2* __restrict__ src, float2* v) {
#pragma unroll
    for (int l = 0; l < 64; ++l) v[l] = LNX_T64_LDG(src + (size_t)l * COLS);
    fft_dif<64>(v);
}
LNX_HD void lead_mul_inv_store(const float2* __restrict__ kt, float2* __restrict__ dst, float2* v) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = cmul(v[j], LNX_T64_LDG(kt + (size_t)br6(j) * COLS));
    ifft_dit<64>(v);
#pragma unroll
    for (int l = 0; l < 64; ++l) dst[(size_t)l * COLS] = v[l];
}
LNX_HD void lead_store_fwd(float2* __restrict__ dst, const float2* v) {  // forward-only mode (kernel-spectrum builder): natural order out
#pragma unroll
    for (int m = 0; m < 64; ++m) dst[(size_t)m * COLS] = v[br6(m)];
}

// ---------------------------------------------------------------------------------------------------------------------
// plane_inv phases
// ---------------------------------------------------------------------------------------------------------------------
LNX_HD void inv_load(int lane, const float2* __restrict__ src, float2* pl) {  // flat copy: the shared plane has the global layout
#pragma unroll 11
    for (int it = 0; it < PLANE_SPEC / 32; ++it) pl[it * 32 + lane] = LNX_T64_LDG(src + it * 32 + lane);
}
// lane k: inverse transform of column k along axis 1, in place (rows re-ordered even-first for the row phase); lane 0 packs
// columns 0 and 32 (their inverse transforms are real sequences) into one complex column
LNX_HD void inv_cols(int lane, float2* pl) {
    float2 u[64];
    if (lane != 0) {
#pragma unroll
        for (int j = 0; j < 64; ++j) u[j] = pl[br6(j) * PLS + lane];
    } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const float2 f0 = pl[br6(j) * PLS], f32 = pl[br6(j) * PLS + 32];
            u[j] = make_float2(f0.x - f32.y, f0.y + f32.x);
        }
    }
    ifft_dit<64>(u);
#pragma unroll
    for (int r = 0; r < 64; ++r) pl[rho(r) * PLS + lane] = u[r];
}
// lane j: spectra of rows 2j and 2j+1 -> one complex line in the order ifft_dit expects
LNX_HD void inv_rows_load(int lane, const float2* pl, float2* v) {
    const float2* ra = pl + lane * PLS;
    const float2* rb = pl + (32 + lane) * PLS;
    const float2 pa = ra[0], pb = rb[0];
    v[0] = make_float2(pa.x, pb.x);
    v[1] = make_float2(pa.y, pb.y);
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        v[br6(k)] = ra[k];
        v[br6(64 - k)] = rb[k];
    }
}
// after this v[n] = (potential[row 2 lane][n], potential[row 2 lane + 1][n])
LNX_HD void inv_rows(float2* v) {
    retangle<1>(v);
    ifft_dit<64>(v);
}

struct CellParams {   // what the cell phases of one plane need
    int gf_id, state_fn, mean;
    GfConst gc;
    float wk, wsum, dt;
    int sh0, sh1, sh2;   // total_shift_idx of the world
    int l;               // plane index (leading axis)
};
// the two potential rows of this lane -> shared memory (even rows first) [and the trajectory output]
LNX_HD void inv_pot_store(int lane, const float2* v, float* ps, float* pot_plane) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 pa = make_float4(v[4 * q].x, v[4 * q + 1].x, v[4 * q + 2].x, v[4 * q + 3].x);
        const float4 pb = make_float4(v[4 * q].y, v[4 * q + 1].y, v[4 * q + 2].y, v[4 * q + 3].y);
        *reinterpret_cast<float4*>(ps + lane * SRS + 4 * q) = pa;
        *reinterpret_cast<float4*>(ps + (32 + lane) * SRS + 4 * q) = pb;
        if (pot_plane) {
            *reinterpret_cast<float4*>(pot_plane + (2 * lane) * N + 4 * q) = pa;
            *reinterpret_cast<float4*>(pot_plane + (2 * lane + 1) * N + 4 * q) = pb;
        }
    }
}
// coalesced growth / mix / update of the plane + this lane's statistics partials (acc[NP_T], layout of tiled::pass_d_kernel).
// GF / SF >= 0: growth and state function fixed at compile time (reciprocal forms of the divisions, like the resident fused
// kernel); GF < 0: selected per cell from cp (true divisions, like the generic tiled pass C).  Eight row pairs per batch: the
// sixteen 128-bit loads of a batch are in flight together.
// keep_state: the new cells also replace the potentials in `ps` (same row placement as fwd_load: the fused kernel transforms them
// for the next step).
template <int GF, int SF>
LNX_HD void inv_update(int lane, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                       const CellParams& cp, float* acc, bool keep_state) {
    constexpr int B = 8;
    const int hi = lane >> 4, n0 = (lane & 15) * 4;
    float colA[4] = {0.f, 0.f, 0.f, 0.f}, colG[4] = {0.f, 0.f, 0.f, 0.f};
    float mx1 = 0.f, mx21 = 0.f, gx1 = 0.f, cnt_a = 0.f, cnt_g = 0.f, cnt_p = 0.f;
    const float inv_wsum = cp.mean ? 1.0f / cp.wsum : 1.0f;
#pragma unroll 1
    for (int it0 = 0; it0 < 32; it0 += B) {
        float4 avs[B], pvs[B];
#pragma unroll
        for (int b = 0; b < B; ++b) avs[b] = *reinterpret_cast<const float4*>(st + (it0 + b) * 128 + lane * 4);
#pragma unroll
        for (int b = 0; b < B; ++b) pvs[b] = *reinterpret_cast<const float4*>(ps + (it0 + b + 32 * hi) * SRS + n0);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int it = it0 + b, r = 2 * it + hi, i = it * 128 + lane * 4;
            const float a4[4] = {avs[b].x, avs[b].y, avs[b].z, avs[b].w};
            const float p4[4] = {pvs[b].x, pvs[b].y, pvs[b].z, pvs[b].w};
            float f4[4], n4[4];
            const float x1 = (float)(((r - cp.sh1) & (N - 1)) - N / 2);
            float rowa = 0.f, rowg = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                cnt_p += p4[e] > EPS ? 1.f : 0.f;
                float f;
                if constexpr (GF >= 0) {
                    f = (cp.wk * growth<GF, true, GF != GF_POLY_QUAD4>(p4[e], cp.gc)) * inv_wsum;
                } else {
                    f = 0.f + cp.wk * growth_dyn<true>(cp.gf_id, p4[e], cp.gc);
                    if (cp.mean) f = f / cp.wsum;
                }
                f4[e] = f;
                const float a = a4[e];
                if constexpr (GF >= 0)
                    n4[e] = state_update<SF, true>(a, f, cp.dt);
                else
                    n4[e] = state_update_dyn<true>(cp.state_fn, a, f, cp.dt);
                const float gp = fmaxf(f, 0.f);
                colA[e] += a;
                colG[e] += gp;
                rowa += a;
                rowg += gp;
                cnt_a += a > EPS ? 1.f : 0.f;
                cnt_g += gp > EPS ? 1.f : 0.f;
            }
            mx1 += rowa * x1;
            mx21 += rowa * x1 * x1;
            gx1 += rowg * x1;
            *reinterpret_cast<float4*>(st + i) = make_float4(n4[0], n4[1], n4[2], n4[3]);
            if (keep_state) *reinterpret_cast<float4*>(ps + (it + 32 * hi) * SRS + n0) = make_float4(n4[0], n4[1], n4[2], n4[3]);
            if (cells_out) *reinterpret_cast<float4*>(cells_out + i) = avs[b];
            if (field_out) *reinterpret_cast<float4*>(field_out + i) = make_float4(f4[0], f4[1], f4[2], f4[3]);
        }
    }
    float m00 = 0.f, g00 = 0.f, mx2 = 0.f, mx22 = 0.f, gx2 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float x2 = (float)(((n0 + e - cp.sh2) & (N - 1)) - N / 2);
        m00 += colA[e];
        g00 += colG[e];
        mx2 += colA[e] * x2;
        mx22 += colA[e] * x2 * x2;
        gx2 += colG[e] * x2;
    }
    const float x0 = (float)(((cp.l - cp.sh0) & (N - 1)) - N / 2);
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;
    acc[0] = cnt_a;
    acc[1] = g00;
    acc[2] = cnt_g;
    acc[3] = cnt_p;
    acc[4] = m00 * x0;
    acc[5] = mx1;
    acc[6] = mx2;
    acc[4 + MAXD] = m00 * x0 * x0;
    acc[5 + MAXD] = mx21;
    acc[6 + MAXD] = mx22;
    acc[4 + 2 * MAXD] = g00 * x0;
    acc[5 + 2 * MAXD] = gx1;
    acc[6 + 2 * MAXD] = gx2;
    acc[4 + 3 * MAXD] = m00;
}
// the common growth functions with the v1 update get compile-time variants, everything else the per-cell selection
LNX_HD void inv_update_dispatch(int lane, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                                const CellParams& cp, float* acc, bool keep_state = false) {
    if (cp.state_fn == SF_V1 && cp.gf_id == GF_POLY_QUAD4)
        inv_update<GF_POLY_QUAD4, SF_V1>(lane, ps, st, cells_out, field_out, cp, acc, keep_state);
    else if (cp.state_fn == SF_V1 && cp.gf_id == GF_GAUSSIAN)
        inv_update<GF_GAUSSIAN, SF_V1>(lane, ps, st, cells_out, field_out, cp, acc, keep_state);
    else
        inv_update<-1, -1>(lane, ps, st, cells_out, field_out, cp, acc, keep_state);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------
// grid (64 planes, C, worlds), one warp
__global__ void __launch_bounds__(32) plane_fwd_kernel(PassAArgs P) {
    __shared__ __align__(16) float sm[SMEM_FLOATS];
    const int lane = threadIdx.x, l = blockIdx.x, c = blockIdx.y, w = blockIdx.z + P.world0;
    const size_t plane = ((size_t)w * P.C + c) * N + l;
    fwd_load(lane, P.state + plane * PLANE_CELLS, sm);
    __syncwarp();
    float2 v[64];
    fwd_rows(lane, sm, v);
    __syncwarp();  // every lane has its rows in registers: the buffer becomes the spectrum plane
    fwd_rows_store(lane, reinterpret_cast<float2*>(sm), v);
    __syncwarp();
    fwd_cols(lane, reinterpret_cast<const float2*>(sm), P.spec + plane * PLANE_SPEC);
}

// grid (33, C, worlds), 64 threads: thread = one spectral column
template <int MINB, int TPB = LEAD_TPB>
__global__ void __launch_bounds__(TPB, MINB) lead_kernel(PassBArgs P) {
    const int col = blockIdx.x * TPB + threadIdx.x, c = blockIdx.y, w = blockIdx.z + P.world0;
    const float2* src = P.spec + ((size_t)w * P.C + c) * ((size_t)N * COLS) + col;
    float2 v[64];
    if (P.fwd_out) {
        lead_load_fwd(src, v);
        lead_store_fwd(P.fwd_out + ((size_t)w * P.C + c) * ((size_t)N * COLS) + col, v);
        return;
    }
    const int sol = w / P.n_init;
    for (int k = 0; k < P.K; ++k) {
        if (P.c_in[k] != c) continue;
        lead_load_fwd(src, v);  // (a channel feeding several kernels re-reads its spectrum from L2: registers hold one line)
        lead_mul_inv_store(P.ktab + ((size_t)sol * P.K + k) * ((size_t)N * COLS) + col, P.pot_spec + ((size_t)w * P.K + k) * ((size_t)N * COLS) + col, v);
    }
}

// grid (64 planes, 1, worlds), one warp; one channel, one kernel.  next_spec != nullptr: fused step kernel — the updated plane is
// still in shared memory, so it is transformed for the NEXT step right away (plane_fwd without its launch and its state read); the
// time loop is then lead + this kernel + pass D, after one plane_fwd launch for the first step.
__global__ void __launch_bounds__(32, 12) plane_inv_kernel(PassCArgs P, float2* next_spec) {  // 12: three warps per scheduler (<= 168 registers)
    __shared__ __align__(16) float sm[SMEM_FLOATS];
    const int lane = threadIdx.x, l = blockIdx.x, w = blockIdx.z + P.world0;
    const int sol = w / P.n_init, init = w - sol * P.n_init;
    const size_t plane = (size_t)w * N + l;
    float2* pl = reinterpret_cast<float2*>(sm);
#pragma unroll
    for (int j = 0; j < 4; ++j)  // the state plane is needed after the two transform phases: have it in L2 by then
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.state + plane * PLANE_CELLS + (j * 32 + lane) * 32));
    inv_load(lane, P.pot_spec + plane * PLANE_SPEC, pl);
    __syncwarp();
    inv_cols(lane, pl);
    __syncwarp();
    float2 v[64];
    inv_rows_load(lane, pl, v);
    __syncwarp();  // every lane has its spectra in registers: the buffer becomes the field plane
    inv_rows(v);
    const WorldCarry cr = P.carry[w];
    CellParams cp;
    cp.gf_id = P.gf_id[0];
    cp.state_fn = P.state_fn;
    cp.mean = P.mean;
    cp.gc = gf_prepare(cp.gf_id, P.gf_params[(size_t)sol * 2], P.gf_params[(size_t)sol * 2 + 1]);
    cp.wk = P.weights[sol];
    cp.wsum = cp.wk;
    cp.dt = P.dt[sol];
    cp.sh0 = cr.shift[0];
    cp.sh1 = cr.shift[1];
    cp.sh2 = cr.shift[2];
    cp.l = l;
    const size_t traj = ((size_t)sol * P.max_iter + P.t) * P.n_init + init;
    const size_t toff = traj * ((size_t)N * PLANE_CELLS) + (size_t)l * PLANE_CELLS;
    inv_pot_store(lane, v, sm, P.potential_out ? P.potential_out + toff : nullptr);
    __syncwarp();
    float acc[NP_T];
    const bool fuse = next_spec != nullptr && P.t + 1 < P.max_iter;
    inv_update_dispatch(lane, sm, P.state + plane * PLANE_CELLS, P.cells_out ? P.cells_out + toff : nullptr,
                        P.field_out ? P.field_out + toff : nullptr, cp, acc, fuse);
#pragma unroll
    for (int i = 0; i < NP_T; ++i) {
        float x = acc[i];
        if (i < 5 + 3 * MAXD) {  // entries of channels >= 1 are zero
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        }
        acc[i] = x;
    }
    if (lane == 0) {
        float* p = P.partials + plane * NP_T;
#pragma unroll
        for (int i = 0; i < NP_T; ++i) p[i] = acc[i];
    }
    // (the partial sums are written first: they are not live across the transform below)
    if (fuse) {
        __syncwarp();
        fwd_rows(lane, sm, v);
        __syncwarp();
        fwd_rows_store(lane, pl, v);
        __syncwarp();
        fwd_cols(lane, pl, next_spec + plane * PLANE_SPEC);
    }
}
#endif  // __CUDACC__

}  // namespace t64
}  // namespace lnx
