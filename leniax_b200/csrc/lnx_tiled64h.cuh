// Half-line engine for 64 x 64 x 64 worlds (BASELINE config E, round 2): the step kernels of lnx_tiled64.cuh with every 64-point
// line split over TWO threads, so that a thread carries 32 complex values (64 registers) instead of 64 (128 registers) and an SM
// holds 20 warps instead of 12 — the thread-per-line kernels were latency-bound at three warps per scheduler.
//
// The split is the first radix-2 stage of the transform itself and needs no exchange on the input side: both threads of a line
// read all 64 inputs (from shared memory, or as one broadcast global load) and keep
//     h = 0:  u[n] =  x[n] + x[n + 32]              -> 32-point transform = the EVEN outputs
//     h = 1:  u[n] = (x[n] - x[n + 32]) W64^(+-n)   -> 32-point transform = the ODD outputs
// (decimation in frequency for the forward AND the inverse direction).  Real rows (axis 2) need no split at all: one thread per
// row, as a 32-point complex transform of (x[2n], x[2n + 1]) plus the twiddled untangle.
//
//   plane_step   64 threads per plane: inverse axis 1 (thread = column c, half h) -> inverse axis 2 (thread = row) -> growth /
//                update / statistics partials on coalesced 128-bit accesses -> forward axis 2 -> forward axis 1 of the NEXT step
//   lead_h       thread pair (lanes c, c + 16) per spectral column: forward, multiply by K, inverse; the only exchange is the last
//                inverse stage (e[n] +- W^-n o[n]) as 64 warp shuffles
//
// HBM layouts, kernel table, statistics partials and pass D are those of lnx_tiled64.cuh / lnx_tiled.cuh: the first step's forward
// planes and the forward-only mode (lnx_rfftn) still use t64::plane_fwd_kernel / t64::lead_kernel.
// The per-thread phase functions are __host__ __device__: tests/emul/lnx_t64_emul.cu runs them thread by thread on the CPU.
// Reference: leniax/core.py:52-102 (n-D FFT potential), :163-319, leniax/statistics.py:36-126.
#pragma once
#include "lnx_tiled64.cuh"

namespace lnx {
namespace t64h {

using t64::CellParams;
using t64::COLS;
using t64::HALF;
using t64::N;
using t64::PLANE_CELLS;
using t64::PLANE_SPEC;
using t64::PLS;
using t64::SMEM_FLOATS;
using t64::SRS;
using tiled::MAXD;
using tiled::NP_T;
using tiled::PassBArgs;
using tiled::PassCArgs;
using tiled::WorldCarry;

constexpr int TPB = 64;          // threads per plane / per 32 spectral columns
constexpr int LEAD_COLS = 32;    // spectral columns per CTA of lead_h (two warps of 16 thread pairs)
constexpr int NRED = 5 + 3 * MAXD;  // statistics partials of a one-channel world (the entries of channels >= 1 stay zero)

LNX_HDC int br5(int x) { return bitrev(x, 5); }

// u[idx(J)] *= W64^J (forward) or its conjugate (inverse), J = 0..31; BR: the values sit in bit-reversed slots (input of ifft_dit<32>)
template <int J, bool INV, bool BR>
LNX_HD void half_twiddle(float2* u) {
    if constexpr (J < 32) {
        constexpr int i = BR ? bitrev(J, 5) : J;
        u[i] = mul_tw<J, 64, INV>(u[i]);
        half_twiddle<J + 1, INV, BR>(u);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// plane_step, inverse half (tid = 0..63; column phases: c = tid & 31, h = tid >> 5; row phases: row = tid)
// ---------------------------------------------------------------------------------------------------------------------
// flat copy of the potential-spectrum plane into shared memory (the shared plane has the global layout [m1][33])
LNX_HD void inv_load(int tid, const float2* __restrict__ src, float2* pl) {
#pragma unroll 11
    for (int it = 0; it < PLANE_SPEC / TPB; ++it) pl[it * TPB + tid] = LNX_MUT_LD(src + it * TPB + tid);
}
// columns 0 and 32 of a row are spectra of real sequences along axis 1: they travel as ONE complex column f0 + i f32 (slot 0)
LNX_HD void inv_pack(int tid, float2* pl) {
    const float2 f0 = pl[tid * PLS], f32 = pl[tid * PLS + 32];
    pl[tid * PLS] = make_float2(f0.x - f32.y, f0.y + f32.x);
}
// inverse transform along axis 1 of column c, outputs of parity h: u[n] = x[2 n + h]
LNX_HD void inv_col(int c, int h, const float2* pl, float2* u) {
    const float2 sg = pk_bc(h ? -1.f : 1.f);
#pragma unroll
    for (int m = 0; m < 32; ++m) u[br5(m)] = pk_fma(pl[(m + 32) * PLS + c], sg, pl[m * PLS + c]);
    if (h) half_twiddle<0, true, true>(u);
    ifft_dit<32>(u);
}
LNX_HD void inv_col_store(int c, int h, float2* pl, const float2* u) {
#pragma unroll
    for (int n = 0; n < 32; ++n) pl[(2 * n + h) * PLS + c] = u[n];
}
// half spectrum X[0..32] of one real row -> Z[k] = (X[k] + conj X[32-k]) + i W64^-k (X[k] - conj X[32-k]) in the slots ifft_dit<32>
// expects; its (un-normalised) inverse transform is (x[2n], x[2n+1])
template <int K>
LNX_HD void inv_row_tangle(const float2* row, float2* u) {
    if constexpr (K < 16) {
        const float2 x = row[K], xc = row[32 - K];
        const float2 A = pk_add(x, make_float2(xc.x, -xc.y));
        const float2 B = pk_add(x, make_float2(-xc.x, xc.y));
        constexpr float c = Tw128::c[2 * K], s = Tw128::s[2 * K];
        const float2 Q = cmul(B, make_float2(-s, c));  // i W64^-K B
        u[br5(K)] = pk_add(A, Q);
        u[br5(32 - K)] = pk_add(make_float2(A.x, -A.y), make_float2(-Q.x, Q.y));  // conj(A - Q)
        inv_row_tangle<K + 1>(row, u);
    }
}
LNX_HD void inv_row_load(int r, const float2* pl, float2* u) {
    const float2* row = pl + r * PLS;
    const float2 p = row[0], q = row[16];  // p = (X[0], X[32]), both real
    u[0] = make_float2(p.x + p.y, p.x - p.y);
    u[br5(16)] = make_float2(2.f * q.x, -2.f * q.y);
    inv_row_tangle<1>(row, u);
}
// potentials of row r -> shared memory, natural placement
LNX_HD void inv_pot_store(int r, const float2* u, float* ps) {
#pragma unroll
    for (int q = 0; q < 16; ++q)
        *reinterpret_cast<float4*>(ps + r * SRS + 4 * q) = make_float4(u[2 * q].x, u[2 * q].y, u[2 * q + 1].x, u[2 * q + 1].y);
}

// coalesced growth / mix / update of the plane + this thread's statistics partials (acc[NP_T], layout of tiled::pass_d_kernel);
// thread t: columns 4 (t & 15) .. + 3 of the rows 4 it + (t >> 4).  Same arithmetic per cell as t64::inv_update.
template <int GF, int SF>
LNX_HD void update(int t, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                   float* __restrict__ pot_out, const CellParams& cp, float* acc, bool keep_state) {
    constexpr int B = 4;
    const int rs = t >> 4, n0 = (t & 15) * 4;
    float colA[4] = {0.f, 0.f, 0.f, 0.f}, colG[4] = {0.f, 0.f, 0.f, 0.f};
    float mx1 = 0.f, mx21 = 0.f, gx1 = 0.f, cnt_a = 0.f, cnt_g = 0.f, cnt_p = 0.f;
    const float inv_wsum = cp.mean ? 1.0f / cp.wsum : 1.0f;
#pragma unroll 1
    for (int it0 = 0; it0 < 16; it0 += B) {
        float4 avs[B], pvs[B];
#pragma unroll
        for (int b = 0; b < B; ++b) avs[b] = LNX_MUT_LD(reinterpret_cast<const float4*>(st + (it0 + b) * 256 + t * 4));
#pragma unroll
        for (int b = 0; b < B; ++b) pvs[b] = *reinterpret_cast<const float4*>(ps + (4 * (it0 + b) + rs) * SRS + n0);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int it = it0 + b, r = 4 * it + rs, i = it * 256 + t * 4;
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
            if (keep_state) *reinterpret_cast<float4*>(ps + r * SRS + n0) = make_float4(n4[0], n4[1], n4[2], n4[3]);
            if (cells_out) *reinterpret_cast<float4*>(cells_out + i) = avs[b];
            if (field_out) *reinterpret_cast<float4*>(field_out + i) = make_float4(f4[0], f4[1], f4[2], f4[3]);
            if (pot_out) *reinterpret_cast<float4*>(pot_out + i) = pvs[b];
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
// The same cell phase on PAIRS of neighbouring cells with the packed FP32 instructions, for compile-time growth functions with the v1
// update (what lnx_step.cuh's cells_fused_rs does for the resident kernel): field = c2 o^4 - c folded into one FMA (poly_quad4),
// threshold counts on the integer pipe (gt_bits), clip as one saturating FMA when NaN cannot occur (!NP).  Per cell 13 instructions
// instead of 34.  Column sums per thread (its four columns never change), row moments as packed partial sums.
template <int GF, bool NP>
LNX_HD void update_pk(int t, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                      float* __restrict__ pot_out, const FusedConsts& K, int sh0, int sh1, int sh2, int l, float* acc, bool keep_state) {
    constexpr int B = 4;
    const int rs = t >> 4, n0 = (t & 15) * 4;
    const float2 z2 = make_float2(0.f, 0.f);
    float2 cA0 = z2, cA1 = z2, cG0 = z2, cG1 = z2, m1 = z2, m21 = z2, g1 = z2;
    int cnt_a = 0, cnt_g = 0, cnt_p = 0;  // at most 64 hits each
#pragma unroll 1
    for (int it0 = 0; it0 < 16; it0 += B) {
        float4 avs[B], pvs[B];
#pragma unroll
        for (int b = 0; b < B; ++b) avs[b] = LNX_MUT_LD(reinterpret_cast<const float4*>(st + (it0 + b) * 256 + t * 4));
#pragma unroll
        for (int b = 0; b < B; ++b) pvs[b] = *reinterpret_cast<const float4*>(ps + (4 * (it0 + b) + rs) * SRS + n0);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int it = it0 + b, r = 4 * it + rs, i = it * 256 + t * 4;
            const float x1 = (float)(((r - sh1) & (N - 1)) - N / 2);
            const float2 A0 = make_float2(avs[b].x, avs[b].y), A1 = make_float2(avs[b].z, avs[b].w);
            const float2 P0 = make_float2(pvs[b].x, pvs[b].y), P1 = make_float2(pvs[b].z, pvs[b].w);
            cnt_p += gt_bits(P0.x, EPS) + gt_bits(P0.y, EPS);  // statistics.py:70
            cnt_p += gt_bits(P1.x, EPS) + gt_bits(P1.y, EPS);
            const float2 F0 = field_fused_pk<GF, NP>(P0, K), F1 = field_fused_pk<GF, NP>(P1, K);
            cA0 = pk_add(cA0, A0);
            cA1 = pk_add(cA1, A1);
            const float2 S = pk_add(A0, A1);
            m1 = pk_fma(S, pk_bc(x1), m1);
            m21 = pk_fma(S, pk_bc(x1 * x1), m21);
            cnt_a += gt_bits(A0.x, EPS) + gt_bits(A0.y, EPS);
            cnt_a += gt_bits(A1.x, EPS) + gt_bits(A1.y, EPS);
            const float2 G0 = make_float2(fmaxf(F0.x, 0.f), fmaxf(F0.y, 0.f)), G1 = make_float2(fmaxf(F1.x, 0.f), fmaxf(F1.y, 0.f));  // statistics.py:65
            cG0 = pk_add(cG0, G0);
            cG1 = pk_add(cG1, G1);
            g1 = pk_fma(pk_add(G0, G1), pk_bc(x1), g1);
            cnt_g += gt_bits(F0.x, EPS) + gt_bits(F0.y, EPS);  // max(f, 0) > eps <=> f > eps
            cnt_g += gt_bits(F1.x, EPS) + gt_bits(F1.y, EPS);
            float4 nw;
            if constexpr (!NP) {  // clip(a + dt f, 0, 1) as one saturating FMA (no NaN possible here)
                nw = make_float4(saturate01(A0.x + K.dt * F0.x), saturate01(A0.y + K.dt * F0.y), saturate01(A1.x + K.dt * F1.x),
                                 saturate01(A1.y + K.dt * F1.y));
            } else {
                nw = make_float4(state_update<SF_V1, NP>(A0.x, F0.x, K.dt), state_update<SF_V1, NP>(A0.y, F0.y, K.dt),
                                 state_update<SF_V1, NP>(A1.x, F1.x, K.dt), state_update<SF_V1, NP>(A1.y, F1.y, K.dt));
            }
            *reinterpret_cast<float4*>(st + i) = nw;
            if (keep_state) *reinterpret_cast<float4*>(ps + r * SRS + n0) = nw;
            if (cells_out) *reinterpret_cast<float4*>(cells_out + i) = avs[b];
            if (field_out) *reinterpret_cast<float4*>(field_out + i) = make_float4(F0.x, F0.y, F1.x, F1.y);
            if (pot_out) *reinterpret_cast<float4*>(pot_out + i) = pvs[b];
        }
    }
    const float colA[4] = {cA0.x, cA0.y, cA1.x, cA1.y}, colG[4] = {cG0.x, cG0.y, cG1.x, cG1.y};
    float m00 = 0.f, g00 = 0.f, mx2 = 0.f, mx22 = 0.f, gx2 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float x2 = (float)(((n0 + e - sh2) & (N - 1)) - N / 2);
        m00 += colA[e];
        g00 += colG[e];
        mx2 += colA[e] * x2;
        mx22 += colA[e] * x2 * x2;
        gx2 += colG[e] * x2;
    }
    const float x0 = (float)(((l - sh0) & (N - 1)) - N / 2);
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;
    acc[0] = count_from_bits(cnt_a);
    acc[1] = g00;
    acc[2] = count_from_bits(cnt_g);
    acc[3] = count_from_bits(cnt_p);
    acc[4] = m00 * x0;
    acc[5] = m1.x + m1.y;
    acc[6] = mx2;
    acc[4 + MAXD] = m00 * x0 * x0;
    acc[5 + MAXD] = m21.x + m21.y;
    acc[6 + MAXD] = mx22;
    acc[4 + 2 * MAXD] = g00 * x0;
    acc[5 + 2 * MAXD] = g1.x + g1.y;
    acc[6 + 2 * MAXD] = gx2;
    acc[4 + 3 * MAXD] = m00;
}
// cell phase of one thread in one of the compiled forms: MODE_DYN = growth / state function selected per cell (every combination the
// reference has); the others are compile-time growth functions with the v1 update, NaN-propagating (NP) or with single-instruction
// clamps (caller vouches that NaN cannot occur: LNX_RUN_ASSUME_FINITE)
enum Mode { MODE_DYN = 0, MODE_PQ4_NP, MODE_PQ4, MODE_GAUSS_NP, MODE_GAUSS, MODE_COUNT };
LNX_HD int select_mode(int gf_id, int state_fn, bool finite) {
    if (state_fn != SF_V1) return MODE_DYN;
    if (gf_id == GF_POLY_QUAD4) return finite ? MODE_PQ4 : MODE_PQ4_NP;
    if (gf_id == GF_GAUSSIAN) return finite ? MODE_GAUSS : MODE_GAUSS_NP;
    return MODE_DYN;
}
template <int MODE>
LNX_HD void update_mode(int t, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                        float* __restrict__ pot_out, const CellParams& cp, float* acc, bool keep_state) {
    if constexpr (MODE == MODE_DYN) {
        update<-1, -1>(t, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state);
    } else {
        FusedConsts K;
        K.gf = cp.gc;
        K.c = cp.mean ? cp.wk * (1.0f / cp.wsum) : cp.wk;
        K.c2 = 2.0f * K.c;
        K.dt = cp.dt;
        constexpr int GF = (MODE == MODE_PQ4 || MODE == MODE_PQ4_NP) ? GF_POLY_QUAD4 : GF_GAUSSIAN;
        constexpr bool NP = MODE == MODE_PQ4_NP || MODE == MODE_GAUSS_NP;
        update_pk<GF, NP>(t, ps, st, cells_out, field_out, pot_out, K, cp.sh0, cp.sh1, cp.sh2, cp.l, acc, keep_state);
    }
}
LNX_HD void update_dispatch(int mode, int t, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                            float* __restrict__ pot_out, const CellParams& cp, float* acc, bool keep_state) {  // emulator entry
    switch (mode) {
        case MODE_PQ4_NP: update_mode<MODE_PQ4_NP>(t, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        case MODE_PQ4: update_mode<MODE_PQ4>(t, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        case MODE_GAUSS_NP: update_mode<MODE_GAUSS_NP>(t, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        case MODE_GAUSS: update_mode<MODE_GAUSS>(t, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        default: update_mode<MODE_DYN>(t, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// plane_step, forward half (the updated cells are still in shared memory)
// ---------------------------------------------------------------------------------------------------------------------
LNX_HD void fwd_row_load(int r, const float* ps, float2* u) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(ps + r * SRS + 4 * q);
        u[2 * q] = make_float2(x.x, x.y);
        u[2 * q + 1] = make_float2(x.z, x.w);
    }
    fft_dif<32>(u);  // u[j] = Z[br5(j)], Z the transform of (x[2n], x[2n+1])
}
// X[k] = (Z[k] + conj Z[32-k]) / 2 - i W64^k (Z[k] - conj Z[32-k]) / 2, k = 1..31, straight into the spectrum plane
template <int K>
LNX_HD void fwd_row_untangle(const float2* u, float2* row) {
    if constexpr (K < 16) {
        const float2 z = u[br5(K)], zc = u[br5(32 - K)];
        const float2 A = pk_add(z, make_float2(zc.x, -zc.y));
        const float2 B = pk_add(z, make_float2(-zc.x, zc.y));
        constexpr float c = Tw128::c[2 * K], s = Tw128::s[2 * K];
        const float2 Q = cmul(B, make_float2(-0.5f * s, -0.5f * c));  // -i/2 W64^K B
        row[K] = pk_fma(A, pk_bc(0.5f), Q);
        row[32 - K] = pk_fma(A, make_float2(0.5f, -0.5f), make_float2(-Q.x, Q.y));  // conj(A/2 - Q)
        fwd_row_untangle<K + 1>(u, row);
    }
}
LNX_HD void fwd_row_store(int r, float2* pl, const float2* u) {
    float2* row = pl + r * PLS;
    const float2 z0 = u[0], z16 = u[br5(16)];
    row[0] = make_float2(z0.x + z0.y, z0.x - z0.y);  // (X[0], X[32]): the packed column
    row[16] = make_float2(z16.x, -z16.y);
    fwd_row_untangle<1>(u, row);
}
// forward transform along axis 1 of column c, outputs of parity h: u[j] = Z[2 br5(j) + h]
LNX_HD void fwd_col(int c, int h, const float2* pl, float2* u) {
    const float2 sg = pk_bc(h ? -1.f : 1.f);
#pragma unroll
    for (int n = 0; n < 32; ++n) u[n] = pk_fma(pl[(n + 32) * PLS + c], sg, pl[n * PLS + c]);
    if (h) half_twiddle<0, false, false>(u);
    fft_dif<32>(u);
}
// columns 1..31 go to the global half spectrum [m1][33]; the packed column (c = 0) to zs[64] for fwd_packed_store
LNX_HD void fwd_col_store(int c, int h, const float2* u, float2* __restrict__ dst, float2* zs) {
    if (c != 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[(2 * br5(j) + h) * HALF + c] = u[j];
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) zs[2 * br5(j) + h] = u[j];
    }
}
// thread m: columns 0 and 32 at axis-1 frequency m from the transform of the packed column
LNX_HD void fwd_packed_store(int m, const float2* zs, float2* __restrict__ dst) {
    const float2 z = zs[m], zc = zs[(N - m) & (N - 1)];
    dst[m * HALF] = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
    dst[m * HALF + 32] = make_float2(0.5f * (z.y + zc.y), 0.5f * (zc.x - z.x));
}

// ---------------------------------------------------------------------------------------------------------------------
// lead_h: thread pair per spectral column; src / kt / dst point at element [l = 0][col]
// ---------------------------------------------------------------------------------------------------------------------
LNX_HD void lead_fwd(int h, const float2* __restrict__ src, float2* u) {
    const float2 sg = pk_bc(h ? -1.f : 1.f);
#pragma unroll
    for (int n = 0; n < 32; ++n) u[n] = pk_fma(LNX_MUT_LD(src + (size_t)(n + 32) * COLS), sg, LNX_MUT_LD(src + (size_t)n * COLS));
    if (h) half_twiddle<0, false, false>(u);
    fft_dif<32>(u);  // u[j] = X[2 br5(j) + h]
}
// multiply by K, inverse 32-point transform of this parity; afterwards u[n] = e[n] (h = 0) or W64^-n o[n] (h = 1)
LNX_HD void lead_mul_inv(int h, const float2* __restrict__ kt, float2* u) {
#pragma unroll
    for (int j = 0; j < 32; ++j) u[j] = cmul(u[j], LNX_T64_LDG(kt + (size_t)(2 * br5(j) + h) * COLS));
    ifft_dit<32>(u);
    if (h) half_twiddle<0, true, false>(u);
}
// last inverse stage: h = 0 keeps y[n] = e[n] + o'[n], h = 1 keeps y[n + 32] = e[n] - o'[n]  (own = this thread's value, other = the partner's)
LNX_HD float2 lead_combine(int h, float2 own, float2 other) { return pk_fma(own, pk_bc(h ? -1.f : 1.f), other); }

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------
// grid (66, 1, worlds), 64 threads: lane = 16 h + cc, warp wq: column 32 blockIdx.x + 16 wq + cc
// 8 CTAs per SM = 16 warps at 128 registers (compiled for 10 CTAs / 96 registers the kernels spill: lead_h 202 us against 103 us,
// profiles/r2_t64h_ncu_v1.txt)
__device__ __forceinline__ void lead_h_body(const PassBArgs& P, const int tile, const int w) {
    const int lane = threadIdx.x & 31, h = lane >> 4;
    const int col = tile * LEAD_COLS + (threadIdx.x >> 5) * 16 + (lane & 15);
    const int sol = w / P.n_init;
    const size_t img = (size_t)N * COLS;
    float2 u[32];
    lead_fwd(h, P.spec + (size_t)w * img + col, u);
    lead_mul_inv(h, P.ktab + (size_t)sol * img + col, u);
    float2* dst = P.pot_spec + (size_t)w * img + (size_t)(32 * h) * COLS + col;
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, u[n].x, 16);
        o.y = __shfl_xor_sync(0xffffffffu, u[n].y, 16);
        dst[(size_t)n * COLS] = lead_combine(h, u[n], o);
    }
}
__global__ void __launch_bounds__(TPB, 8) lead_h_kernel(PassBArgs P) { lead_h_body(P, blockIdx.x, blockIdx.z + P.world0); }

// plane l of world w at step t; next_spec != nullptr: the updated plane is transformed for the NEXT step right away (as
// t64::plane_inv_kernel does).  sm / zs / red: the CTA's shared memory (SMEM_FLOATS floats, N complex, NRED floats)
template <int MODE>
__device__ __forceinline__ void plane_step_body(const PassCArgs& P, float2* next_spec, const int l, const int w, const int t, float* sm,
                                                float2* zs, float* red) {
    const int tid = threadIdx.x, c = tid & 31, h = tid >> 5;
    const int sol = w / P.n_init, init = w - sol * P.n_init;
    const size_t plane = (size_t)w * N + l;
    float2* pl = reinterpret_cast<float2*>(sm);
    float* stp = P.state + plane * PLANE_CELLS;
#pragma unroll
    for (int j = 0; j < 2; ++j)  // the state plane is needed after the two transform phases: have it in L2 by then
        asm volatile("prefetch.global.L2 [%0];" ::"l"(stp + (j * 64 + tid) * 32));
    inv_load(tid, P.pot_spec + plane * PLANE_SPEC, pl);
    __syncthreads();
    inv_pack(tid, pl);
    __syncthreads();
    float2 u[32];
    inv_col(c, h, pl, u);
    __syncthreads();  // every thread has read its column: the results replace it
    inv_col_store(c, h, pl, u);
    __syncthreads();
    inv_row_load(tid, pl, u);
    ifft_dit<32>(u);
    __syncthreads();  // every thread has its spectrum in registers: the buffer becomes the potential plane
    inv_pot_store(tid, u, sm);
    CellParams cp;
    cp.gf_id = P.gf_id[0];
    cp.state_fn = P.state_fn;
    cp.mean = P.mean;
    cp.gc = gf_prepare(cp.gf_id, P.gf_params[(size_t)sol * 2], P.gf_params[(size_t)sol * 2 + 1]);
    cp.wk = P.weights[sol];
    cp.wsum = cp.wk;
    cp.dt = P.dt[sol];
    cp.sh0 = LNX_MUT_LD(&P.carry[w].shift[0]);
    cp.sh1 = LNX_MUT_LD(&P.carry[w].shift[1]);
    cp.sh2 = LNX_MUT_LD(&P.carry[w].shift[2]);
    cp.l = l;
    const size_t traj = ((size_t)sol * P.max_iter + t) * P.n_init + init;
    const size_t toff = traj * ((size_t)N * PLANE_CELLS) + (size_t)l * PLANE_CELLS;
    __syncthreads();
    float acc[NP_T];
    const bool fuse = next_spec != nullptr && t + 1 < P.max_iter;
    update_mode<MODE>(tid, sm, stp, P.cells_out ? P.cells_out + toff : nullptr, P.field_out ? P.field_out + toff : nullptr,
                      P.potential_out ? P.potential_out + toff : nullptr, cp, acc, fuse);
#pragma unroll
    for (int i = 0; i < NRED; ++i) {
        float x = acc[i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        acc[i] = x;
    }
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < NRED; ++i) red[i] = acc[i];
    }
    __syncthreads();  // also: every cell of the plane is updated in shared memory
    if (tid == 0) {
        float* p = P.partials + plane * NP_T;
#pragma unroll
        for (int i = 0; i < NP_T; ++i) p[i] = i < NRED ? acc[i] + red[i] : 0.f;
    }
    if (!fuse) return;
    fwd_row_load(tid, sm, u);
    __syncthreads();  // every thread has its row in registers: the buffer becomes the spectrum plane
    fwd_row_store(tid, pl, u);
    __syncthreads();
    fwd_col(c, h, pl, u);
    float2* dst = next_spec + plane * PLANE_SPEC;
    fwd_col_store(c, h, u, dst, zs);
    __syncthreads();
    fwd_packed_store(tid, zs, dst);
}
// grid (64 planes, 1, worlds), 64 threads; one channel, one kernel
template <int MODE>
__global__ void __launch_bounds__(TPB, 8) plane_step_kernel(PassCArgs P, float2* next_spec) {
    __shared__ __align__(16) float sm[SMEM_FLOATS];
    __shared__ float2 zs[N];
    __shared__ float red[NRED];
    plane_step_body<MODE>(P, next_spec, blockIdx.x, blockIdx.z + P.world0, P.t, sm, zs, red);
}

// ---------------------------------------------------------------------------------------------------------------------
// The WHOLE scan as one persistent kernel.  The work of (world w, step t) is 66 lead tiles + the statistics finaliser of step t - 1 +
// 64 planes; CTAs take these items from one global queue (atomicAdd) and wait on per-world completion counters:
//     lead tile (w, t), finaliser (w, t - 1)   need all 64 planes of (w, t - 1)
//     plane (w, t)                              needs all 66 lead tiles of (w, t) and the finaliser of (w, t - 1)
// Queue order: windows of `window` worlds run ALL their steps before the next window starts; inside a window, step by step, first the
// lead tiles + finalisers of every world, then the planes of every world.  (a) A dependency is always EARLIER in the queue than its
// dependents, so whoever holds it is running or done: the spin-waits cannot deadlock, whatever the number of resident CTAs.  (b) It is
// about 1 500 items earlier - more than the CTAs in flight - so the waits are almost never taken.  (c) The working set of a window
// (3.2 MB per world: state, spectrum, potential spectrum) stays in the 126 MB L2: HBM sees the initial states and the statistics rows.
// (d) No launch gaps or wave tails, and HBM-latency-bound lead tiles share an SM with shared-memory-bound planes.
// Everything another CTA may have written is read with LNX_MUT_LD (L2); the per-solution tables stay on the read-only path.
// ---------------------------------------------------------------------------------------------------------------------
struct ScanArgs {
    int* queue;     // [1] next item, zeroed by the host
    int* cnt;       // [3][worlds] completed lead tiles / planes / finalisers per world, zeroed by the host
    int worlds;     // worlds of this launch (PassBArgs::world0 etc. = 0)
    int steps;
    int window;     // worlds per L2-resident window
};
constexpr int SCAN_TILES = COLS / LEAD_COLS;            // 66 lead tiles per world and step
constexpr int SCAN_ITEMS = SCAN_TILES + 1 + N;          // + finaliser + 64 planes
__device__ __forceinline__ void scan_wait(const int* p, const int need) {
    while (*reinterpret_cast<const volatile int*>(p) < need) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 700
        __nanosleep(64);
#endif
    }
}
__device__ __forceinline__ void scan_wait2(const int* p, const int need_p, const int* q, const int need_q) {  // both loads in flight together
    for (;;) {
        const int a = *reinterpret_cast<const volatile int*>(p), b = *reinterpret_cast<const volatile int*>(q);
        if (a >= need_p && b >= need_q) return;
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 700
        __nanosleep(64);
#endif
    }
}
template <int MODE>
__global__ void __launch_bounds__(TPB, 8) scan_kernel(PassBArgs B, PassCArgs C, tiled::PassDArgs D, float2* spec, ScanArgs M) {
    __shared__ __align__(16) float sm[SMEM_FLOATS];
    __shared__ float2 zs[N];
    __shared__ float red[NRED];
    __shared__ int s_item;
    const int tid = threadIdx.x;
    int* cntL = M.cnt;
    int* cntP = M.cnt + M.worlds;
    int* cntD = M.cnt + 2 * M.worlds;
    const int total = M.worlds * M.steps * SCAN_ITEMS;
    const int win_items = M.window * M.steps * SCAN_ITEMS;
    // the queue is read one item ahead (the atomic's round trip through L2 runs under the current item's work); a CTA then holds two
    // items, the later one not started: the no-deadlock argument is unchanged (the holder of the earliest unfinished item is running it)
    int next = 0;
    if (tid == 0) next = atomicAdd(M.queue, 1);
    for (;;) {
        if (tid == 0) {
            s_item = next;
            if (next < total) next = atomicAdd(M.queue, 1);
        }
        __syncthreads();
        const int q = s_item;
        if (q >= total) break;
        const int win = q / win_items, w0 = win * M.window;
        const int W = M.worlds - w0 < M.window ? M.worlds - w0 : M.window;
        int r = q - win * win_items;
        const int t = r / (W * SCAN_ITEMS);
        r -= t * (W * SCAN_ITEMS);
        int kind, wl, j;  // kind 0: lead tile j, 1: plane j, 2: finaliser of step t - 1
        if (r < W * (SCAN_TILES + 1)) {
            wl = r / (SCAN_TILES + 1);
            j = r - wl * (SCAN_TILES + 1);
            kind = j < SCAN_TILES ? 0 : 2;
        } else {
            r -= W * (SCAN_TILES + 1);
            wl = r / N;
            j = r - wl * N;
            kind = 1;
        }
        const int w = w0 + wl;
        if (tid == 0) {
            if (kind == 1) {
                scan_wait2(cntL + w, SCAN_TILES * (t + 1), cntD + w, t + 1);
            } else if (t > 0) {
                scan_wait(cntP + w, N * t);
            }
            __threadfence();
        }
        __syncthreads();
        if (kind == 0)
            lead_h_body(B, j, w);
        else if (kind == 1)
            plane_step_body<MODE>(C, spec, j, w, t, sm, zs, red);
        else if (t > 0)
            tiled::pass_d_body(D, w, t - 1);
        __syncthreads();
        if (tid == 0) {  // (the barrier orders the CTA's stores before this fence: the pattern of a grid-wide barrier)
            __threadfence();
            atomicAdd(kind == 0 ? cntL + w : (kind == 1 ? cntP + w : cntD + w), 1);
        }
    }
}
template <int MODE>
inline cudaError_t launch_scan_mode(const PassBArgs& b, const PassCArgs& c, const tiled::PassDArgs& d, float2* spec, const ScanArgs& m, int sms,
                                    cudaStream_t s) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<MODE>, TPB, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)per_sm * sms;
    const long long total = (long long)m.worlds * m.steps * SCAN_ITEMS;
    if (grid > total) grid = total;
    scan_kernel<MODE><<<(unsigned)grid, TPB, 0, s>>>(b, c, d, spec, m);
    return cudaGetLastError();
}
// all steps of `m.worlds` worlds (world0 = 0 in the argument structs) in ONE launch; the caller has run plane_fwd for step 0 and
// finalises the last step's statistics (pass D with t = steps - 1) afterwards
inline cudaError_t launch_scan(const PassBArgs& b, const PassCArgs& c, const tiled::PassDArgs& d, float2* spec, const ScanArgs& m, bool finite,
                               int sms, cudaStream_t s) {
    switch (select_mode(c.gf_id[0], c.state_fn, finite)) {
        case MODE_PQ4: return launch_scan_mode<MODE_PQ4>(b, c, d, spec, m, sms, s);
        case MODE_PQ4_NP: return launch_scan_mode<MODE_PQ4_NP>(b, c, d, spec, m, sms, s);
        case MODE_GAUSS: return launch_scan_mode<MODE_GAUSS>(b, c, d, spec, m, sms, s);
        case MODE_GAUSS_NP: return launch_scan_mode<MODE_GAUSS_NP>(b, c, d, spec, m, sms, s);
        default: return launch_scan_mode<MODE_DYN>(b, c, d, spec, m, sms, s);
    }
}

// lead_h + plane_step of one step for `nb` worlds (finite: LNX_RUN_ASSUME_FINITE)
inline void launch_step(const PassBArgs& b, const PassCArgs& c, float2* next_spec, unsigned nb, bool finite, cudaStream_t s) {
    lead_h_kernel<<<dim3(COLS / LEAD_COLS, 1, nb), TPB, 0, s>>>(b);
    const dim3 gp(N, 1, nb);
    switch (select_mode(c.gf_id[0], c.state_fn, finite)) {
        case MODE_PQ4: (plane_step_kernel<MODE_PQ4>)<<<gp, TPB, 0, s>>>(c, next_spec); break;
        case MODE_PQ4_NP: (plane_step_kernel<MODE_PQ4_NP>)<<<gp, TPB, 0, s>>>(c, next_spec); break;
        case MODE_GAUSS: (plane_step_kernel<MODE_GAUSS>)<<<gp, TPB, 0, s>>>(c, next_spec); break;
        case MODE_GAUSS_NP: (plane_step_kernel<MODE_GAUSS_NP>)<<<gp, TPB, 0, s>>>(c, next_spec); break;
        default: (plane_step_kernel<MODE_DYN>)<<<gp, TPB, 0, s>>>(c, next_spec); break;
    }
}
#endif  // __CUDACC__

}  // namespace t64h
}  // namespace lnx
