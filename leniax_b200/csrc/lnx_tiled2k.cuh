// Four-step engine for one-channel one-kernel 2048 x 2048 worlds (BASELINE config D): every 2048-point transform is done by
// ONE WARP as 2048 = 32 x 64 — lane n1 transforms the 64 elements n1 + 32 n2 in its registers (fft_dif<64>), multiplies by the
// twiddles W^(n1 k2), the warp transposes once through shared memory, lane t transforms the two 32-point lines k2 = t, t + 32
// (fft_dif<32>).  No radix pass over shared memory, no CTA-wide barrier, every global access a full coalesced line:
//
//   rows_fwd   warp per row pair (2p, 2p+1): packed complex transform, untangle, TRANSPOSED half spectrum T[k][row] (16 B per k)
//   lead       warp per spectral column k = one contiguous 16 KB line of T: transform along the leading axis, multiply by the
//              kernel table (stored in the warp's register order), inverse transform, in the same layout
//   rows_inv   warp per row pair: gather the 16-byte pieces of its two rows, retangle, inverse transform, growth / update /
//              statistics partials of the row pair (= one slab of lnx_tiled.cuh, so pass D is shared)
//
// n = n1 + 32 n2, k = k2 + 64 k1:  W_2048^(nk) = W_2048^(n1 k2) W_32^(n1 k1) W_64^(n2 k2).
// The per-lane phase functions are __host__ __device__: tests/emul/lnx_t64_emul.cu runs them lane by lane on the CPU.
// Reference: leniax/core.py:52-102, :163-319, leniax/statistics.py:36-126.
#pragma once
#include "lnx_tiled64h.cuh"

namespace lnx {
namespace t2k {

using t64::br6;
using tiled::MAXD;
using tiled::NP_T;
using tiled::PassAArgs;
using tiled::PassBArgs;
using tiled::PassCArgs;
using tiled::PassDArgs;
using tiled::WorldCarry;

constexpr int N = 2048, HALF = 1025;
constexpr int EXS = 33;                  // row stride (complex) of the 64 x 32 exchange buffer: odd => both views conflict-free
constexpr int SMEM_C2 = 64 * EXS;        // complex values of shared memory per warp (also holds 2048 complex / 4096 floats)
constexpr size_t SPEC = (size_t)HALF * N;  // complex values of one transposed half spectrum

LNX_HDC int br5(int x) { return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4); }
// leading-axis frequency held by lane t in register q (= h * 32 + j) after the second stage
LNX_HDC int freq_of(int q, int t) { return (t + 32 * (q >> 5)) + 64 * br5(q & 31); }

// v[j] *= W^(lane * k2) (INV: conjugate), k2 = br6(j); tab[i] = (cos, sin)(2 pi i / 2048).  Sixteen exact table values per lane,
// every twiddle one product of two of them.
template <bool INV>
LNX_HD void twiddle64(int lane, float2* v, const float2* __restrict__ tab) {
    float2 ta[8], tc[8];
#pragma unroll
    for (int a = 1; a < 8; ++a) {
        ta[a] = LNX_T64_LDG(tab + ((lane * a) & (N - 1)));
        tc[a] = LNX_T64_LDG(tab + ((lane * 8 * a) & (N - 1)));
    }
#pragma unroll
    for (int j = 1; j < 64; ++j) {
        const int k2 = br6(j), a = k2 & 7, c = k2 >> 3;
        float2 w;
        if (c == 0)
            w = ta[a];
        else if (a == 0)
            w = tc[c];
        else
            w = make_float2(ta[a].x * tc[c].x - ta[a].y * tc[c].y, ta[a].y * tc[c].x + ta[a].x * tc[c].y);
        v[j] = INV ? rot_inv(v[j], w.x, w.y) : rot_fwd(v[j], w.x, w.y);
    }
}

// ---- forward: v[n2] = x[lane + 32 n2]  ->  u[h * 32 + j] = X[(lane + 32 h) + 64 br5(j)] ----
LNX_HD void fs_fwd_a(int lane, float2* v, const float2* __restrict__ tab) {
    fft_dif<64>(v);
    twiddle64<false>(lane, v, tab);
}
LNX_HD void fs_fwd_store(int lane, const float2* v, float2* ex) {
#pragma unroll
    for (int j = 0; j < 64; ++j) ex[br6(j) * EXS + lane] = v[j];
}
LNX_HD void fs_fwd_b(int lane, const float2* ex, float2* u) {
#pragma unroll
    for (int q = 0; q < 64; ++q) u[q] = ex[(lane + 32 * (q >> 5)) * EXS + (q & 31)];
    fft_dif<32>(u);
    fft_dif<32>(u + 32);
}
// ---- inverse (un-normalised): u as above  ->  v[n2] = x[lane + 32 n2] ----
LNX_HD void fs_inv_a(float2* u) {
    ifft_dit<32>(u);
    ifft_dit<32>(u + 32);
}
LNX_HD void fs_inv_store(int lane, const float2* u, float2* ex) {
#pragma unroll
    for (int q = 0; q < 64; ++q) ex[(lane + 32 * (q >> 5)) * EXS + (q & 31)] = u[q];
}
LNX_HD void fs_inv_b(int lane, const float2* ex, float2* v, const float2* __restrict__ tab) {
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = ex[br6(j) * EXS + lane];
    twiddle64<true>(lane, v, tab);
    ifft_dit<64>(v);
}

// ---------------------------------------------------------------------------------------------------------------------
// rows_fwd phases
// ---------------------------------------------------------------------------------------------------------------------
LNX_HD void rf_load(int lane, const float* __restrict__ rows, float2* v) {  // rows: two consecutive rows of 2048 reals
#pragma unroll
    for (int n2 = 0; n2 < 64; ++n2) v[n2] = make_float2(LNX_T64_LDG(rows + lane + 32 * n2), LNX_T64_LDG(rows + N + lane + 32 * n2));
}
LNX_HD void rf_nat_store(int lane, const float2* u, float2* nat) {  // spectrum of the packed line in natural order
#pragma unroll
    for (int q = 0; q < 64; ++q) nat[freq_of(q, lane)] = u[q];
}
// untangle the two real rows and write T[k][2p], T[k][2p + 1] (one 16-byte store per k); dst = T + 2p
LNX_HD void rf_untangle_store(int lane, const float2* nat, float2* __restrict__ dst) {
#pragma unroll 8
    for (int i = 0; i <= 32; ++i) {
        const int k = lane + 32 * i;
        if (i == 32 && lane != 0) break;  // k = 1024: lane 0 only
        const float2 zk = nat[k], zc = nat[(N - k) & (N - 1)];
        const float4 ab = make_float4(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y), 0.5f * (zk.y + zc.y), 0.5f * (zc.x - zk.x));
        *reinterpret_cast<float4*>(dst + (size_t)k * N) = ab;
    }
}

// The same two jobs done by a CTA of ROWS_WARPS warps (one row pair each) together, so that the 16-byte pieces of the transposed
// layout combine into full 128-byte lines: thread -> (warp buffer w = tid & 7, k = (tid >> 3) + 32 i); the eight pieces of one k are
// adjacent in T.  nat_all: ROWS_WARPS buffers of NATS complex values (NATS = 2 mod 16: the eight buffers sit in distinct banks).
constexpr int ROWS_WARPS = 8;
constexpr int NATS = SMEM_C2 + 2;
LNX_HD void rf8_untangle_store(int tid, const float2* nat_all, float2* __restrict__ dst) {  // dst = T + first row of the CTA
    const int w = tid & 7, kq = tid >> 3;
    const float2* nat = nat_all + w * NATS;
#pragma unroll 8
    for (int i = 0; i <= 32; ++i) {
        const int k = kq + 32 * i;
        if (i == 32 && kq != 0) break;  // k = 1024
        const float2 zk = nat[k], zc = nat[(N - k) & (N - 1)];
        const float4 ab = make_float4(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y), 0.5f * (zk.y + zc.y), 0.5f * (zc.x - zk.x));
        *reinterpret_cast<float4*>(dst + (size_t)k * N + 2 * w) = ab;
    }
}
LNX_HD void ri8_gather(int tid, const float2* __restrict__ src, float2* nat_all) {  // src = P + first row of the CTA
    const int w = tid & 7, kq = tid >> 3;
    float2* nat = nat_all + w * NATS;
    float4 ab[33];
#pragma unroll
    for (int i = 0; i <= 32; ++i) {
        const int k = kq + 32 * i;
        if (i < 32 || kq == 0) ab[i] = LNX_T64_LDG(reinterpret_cast<const float4*>(src + (size_t)k * N + 2 * w));
    }
#pragma unroll
    for (int i = 0; i <= 32; ++i) {
        const int k = kq + 32 * i;
        if (i == 32 && kq != 0) break;
        nat[k] = make_float2(ab[i].x - ab[i].w, ab[i].y + ab[i].z);
        if (k != 0 && k != N / 2) nat[N - k] = make_float2(ab[i].x + ab[i].w, ab[i].z - ab[i].y);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// lead phases: src / dst = line k of T (2048 complex), kt = line k of the kernel table in register order
// ---------------------------------------------------------------------------------------------------------------------
LNX_HD void ld_load(int lane, const float2* __restrict__ src, float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 64; ++n2) v[n2] = LNX_T64_LDG(src + lane + 32 * n2);
}
LNX_HD void ld_mul(int lane, float2* u, const float2* __restrict__ kt) {
#pragma unroll
    for (int q = 0; q < 64; ++q) u[q] = cmul(u[q], LNX_T64_LDG(kt + q * 32 + lane));
}
LNX_HD void ld_store(int lane, float2* __restrict__ dst, const float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 64; ++n2) dst[lane + 32 * n2] = v[n2];
}

// ---------------------------------------------------------------------------------------------------------------------
// rows_inv phases
// ---------------------------------------------------------------------------------------------------------------------
// gather the spectra of rows 2p, 2p+1 (src = P + 2p, 16 bytes per k) and retangle into the natural-order packed spectrum
LNX_HD void ri_gather(int lane, const float2* __restrict__ src, float2* nat) {
#pragma unroll 8
    for (int i = 0; i <= 32; ++i) {
        const int k = lane + 32 * i;
        if (i == 32 && lane != 0) break;
        const float4 ab = LNX_T64_LDG(reinterpret_cast<const float4*>(src + (size_t)k * N));
        nat[k] = make_float2(ab.x - ab.w, ab.y + ab.z);
        if (k != 0 && k != N / 2) nat[N - k] = make_float2(ab.x + ab.w, ab.z - ab.y);
    }
}
LNX_HD void ri_nat_load(int lane, const float2* nat, float2* u) {
#pragma unroll
    for (int q = 0; q < 64; ++q) u[q] = nat[freq_of(q, lane)];
}
// potentials of the row pair -> shared memory as two rows of 2048 floats [and the trajectory output]
LNX_HD void ri_pot_store(int lane, const float2* v, float* ps) {
#pragma unroll
    for (int n2 = 0; n2 < 64; ++n2) {
        ps[lane + 32 * n2] = v[n2].x;
        ps[N + lane + 32 * n2] = v[n2].y;
    }
}
struct CellParams2 {
    int gf_id, state_fn, mean;
    GfConst gc;
    float wk, wsum, dt;
    int sh0, sh1;
    int row0;   // first row of the pair
};
// coalesced growth / mix / update of the two rows (4096 consecutive cells) + this lane's statistics partials
// keep_state: the new cells also replace the potentials in `ps` (the fused step kernel transforms them for the next step)
template <int GF, int SF, int ROWS = 2>
LNX_HD void ri_update(int lane, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                      float* __restrict__ pot_out, const CellParams2& cp, float* acc, bool keep_state) {
    constexpr int B = 8;
    float m00 = 0.f, g00 = 0.f, mx0 = 0.f, mx20 = 0.f, gx0 = 0.f, mx1 = 0.f, mx21 = 0.f, gx1 = 0.f, cnt_a = 0.f, cnt_g = 0.f, cnt_p = 0.f;
    const float inv_wsum = cp.mean ? 1.0f / cp.wsum : 1.0f;
    const float x0a = (float)(((cp.row0 - cp.sh0) & (N - 1)) - N / 2), x0b = (float)(((cp.row0 + 1 - cp.sh0) & (N - 1)) - N / 2);
#pragma unroll 1
    for (int it0 = 0; it0 < 16 * ROWS; it0 += B) {  // ROWS = 1: one row per warp (the real-row kernels)
        float4 avs[B], pvs[B];
#pragma unroll
        for (int b = 0; b < B; ++b) avs[b] = *reinterpret_cast<const float4*>(st + (it0 + b) * 128 + lane * 4);
#pragma unroll
        for (int b = 0; b < B; ++b) pvs[b] = *reinterpret_cast<const float4*>(ps + (it0 + b) * 128 + lane * 4);
        const float x0 = it0 < 16 ? x0a : x0b;
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int i = (it0 + b) * 128 + lane * 4, n0 = i & (N - 1);
            const float a4[4] = {avs[b].x, avs[b].y, avs[b].z, avs[b].w};
            const float p4[4] = {pvs[b].x, pvs[b].y, pvs[b].z, pvs[b].w};
            float f4[4], n4[4];
            float sa = 0.f, sg = 0.f;
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
                const float x1 = (float)(((n0 + e - cp.sh1) & (N - 1)) - N / 2);
                sa += a;
                sg += gp;
                const float ax = a * x1;
                mx1 += ax;
                mx21 += ax * x1;
                gx1 += gp * x1;
                cnt_a += a > EPS ? 1.f : 0.f;
                cnt_g += gp > EPS ? 1.f : 0.f;
            }
            m00 += sa;
            g00 += sg;
            mx0 += sa * x0;
            mx20 += sa * x0 * x0;
            gx0 += sg * x0;
            *reinterpret_cast<float4*>(st + i) = make_float4(n4[0], n4[1], n4[2], n4[3]);
            if (keep_state) *reinterpret_cast<float4*>(ps + i) = make_float4(n4[0], n4[1], n4[2], n4[3]);
            if (cells_out) *reinterpret_cast<float4*>(cells_out + i) = avs[b];
            if (field_out) *reinterpret_cast<float4*>(field_out + i) = make_float4(f4[0], f4[1], f4[2], f4[3]);
            if (pot_out) *reinterpret_cast<float4*>(pot_out + i) = pvs[b];
        }
    }
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;
    acc[0] = cnt_a;
    acc[1] = g00;
    acc[2] = cnt_g;
    acc[3] = cnt_p;
    acc[4] = mx0;
    acc[5] = mx1;
    acc[4 + MAXD] = mx20;
    acc[5 + MAXD] = mx21;
    acc[4 + 2 * MAXD] = gx0;
    acc[5 + 2 * MAXD] = gx1;
    acc[4 + 3 * MAXD] = m00;
}
// The same cell phase on PAIRS of neighbouring cells with the packed FP32 instructions (compile-time growth function, v1 update;
// lnx_tiled64h.cuh's update_pk for this layout): 15 instructions per cell instead of 34 - the cell phase was more than half of the
// dependency chain of the rows warp.  Column coordinates as packed pairs (a quad wraps at most once per row), row sums folded into
// the row moments at the end of each row.
template <int GF, bool NP, int ROWS = 2>
LNX_HD void ri_update_pk(int lane, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                         float* __restrict__ pot_out, const FusedConsts& K, int sh0, int sh1, int row0, float* acc, bool keep_state) {
    constexpr int B = 8;
    const float2 z2 = make_float2(0.f, 0.f);
    float2 mA = z2, gA = z2, mx1 = z2, mx21 = z2, gx1 = z2;
    float m00 = 0.f, g00 = 0.f, mx0 = 0.f, mx20 = 0.f, gx0 = 0.f;
    int cnt_a = 0, cnt_g = 0, cnt_p = 0;  // at most 128 hits each
    const float x0a = (float)(((row0 - sh0) & (N - 1)) - N / 2), x0b = (float)(((row0 + 1 - sh0) & (N - 1)) - N / 2);
#pragma unroll 1
    for (int it0 = 0; it0 < 16 * ROWS; it0 += B) {
        float4 avs[B], pvs[B];
#pragma unroll
        for (int b = 0; b < B; ++b) avs[b] = *reinterpret_cast<const float4*>(st + (it0 + b) * 128 + lane * 4);
#pragma unroll
        for (int b = 0; b < B; ++b) pvs[b] = *reinterpret_cast<const float4*>(ps + (it0 + b) * 128 + lane * 4);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int i = (it0 + b) * 128 + lane * 4, n0 = i & (N - 1);
            const float xb = (float)(((n0 - sh1) & (N - 1)) - N / 2);
            float2 X01 = make_float2(xb, xb + 1.f), X23 = make_float2(xb + 2.f, xb + 3.f);
            if (xb > (float)(N / 2 - 4)) {  // the quad crosses the seam of the rolled frame (one quad per row at most)
                X01.y = X01.y >= (float)(N / 2) ? X01.y - (float)N : X01.y;
                X23.x = X23.x >= (float)(N / 2) ? X23.x - (float)N : X23.x;
                X23.y = X23.y >= (float)(N / 2) ? X23.y - (float)N : X23.y;
            }
            const float2 A0 = make_float2(avs[b].x, avs[b].y), A1 = make_float2(avs[b].z, avs[b].w);
            const float2 P0 = make_float2(pvs[b].x, pvs[b].y), P1 = make_float2(pvs[b].z, pvs[b].w);
            cnt_p += gt_bits(P0.x, EPS) + gt_bits(P0.y, EPS);  // statistics.py:70
            cnt_p += gt_bits(P1.x, EPS) + gt_bits(P1.y, EPS);
            const float2 F0 = field_fused_pk<GF, NP>(P0, K), F1 = field_fused_pk<GF, NP>(P1, K);
            mA = pk_add(mA, pk_add(A0, A1));
            const float2 AX0 = pk_mul(A0, X01), AX1 = pk_mul(A1, X23);
            mx1 = pk_add(mx1, pk_add(AX0, AX1));
            mx21 = pk_fma(AX0, X01, mx21);
            mx21 = pk_fma(AX1, X23, mx21);
            cnt_a += gt_bits(A0.x, EPS) + gt_bits(A0.y, EPS);
            cnt_a += gt_bits(A1.x, EPS) + gt_bits(A1.y, EPS);
            const float2 G0 = make_float2(fmaxf(F0.x, 0.f), fmaxf(F0.y, 0.f)), G1 = make_float2(fmaxf(F1.x, 0.f), fmaxf(F1.y, 0.f));  // statistics.py:65
            gA = pk_add(gA, pk_add(G0, G1));
            gx1 = pk_fma(G0, X01, gx1);
            gx1 = pk_fma(G1, X23, gx1);
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
            if (keep_state) *reinterpret_cast<float4*>(ps + i) = nw;
            if (cells_out) *reinterpret_cast<float4*>(cells_out + i) = avs[b];
            if (field_out) *reinterpret_cast<float4*>(field_out + i) = make_float4(F0.x, F0.y, F1.x, F1.y);
            if (pot_out) *reinterpret_cast<float4*>(pot_out + i) = pvs[b];
        }
        if (((it0 + B) & 15) == 0) {  // end of a row: its sums enter the row-coordinate moments
            const float x0 = it0 < 16 ? x0a : x0b, sa = mA.x + mA.y, sg = gA.x + gA.y;
            m00 += sa;
            g00 += sg;
            mx0 += sa * x0;
            mx20 += sa * x0 * x0;
            gx0 += sg * x0;
            mA = z2;
            gA = z2;
        }
    }
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;
    acc[0] = count_from_bits(cnt_a);
    acc[1] = g00;
    acc[2] = count_from_bits(cnt_g);
    acc[3] = count_from_bits(cnt_p);
    acc[4] = mx0;
    acc[5] = mx1.x + mx1.y;
    acc[4 + MAXD] = mx20;
    acc[5 + MAXD] = mx21.x + mx21.y;
    acc[4 + 2 * MAXD] = gx0;
    acc[5 + 2 * MAXD] = gx1.x + gx1.y;
    acc[4 + 3 * MAXD] = m00;
}
// the compiled forms of the cell phase (t64h::Mode): per-cell selection, or a compile-time growth function with the v1 update
template <int MODE, int ROWS = 2>
LNX_HD void ri_update_mode(int lane, float* ps, float* __restrict__ st, float* __restrict__ cells_out, float* __restrict__ field_out,
                           float* __restrict__ pot_out, const CellParams2& cp, float* acc, bool keep_state) {
    if constexpr (MODE == t64h::MODE_DYN) {
        ri_update<-1, -1, ROWS>(lane, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state);
    } else {
        FusedConsts K;
        K.gf = cp.gc;
        K.c = cp.mean ? cp.wk * (1.0f / cp.wsum) : cp.wk;
        K.c2 = 2.0f * K.c;
        K.dt = cp.dt;
        constexpr int GF = (MODE == t64h::MODE_PQ4 || MODE == t64h::MODE_PQ4_NP) ? GF_POLY_QUAD4 : GF_GAUSSIAN;
        constexpr bool NP = MODE == t64h::MODE_PQ4_NP || MODE == t64h::MODE_GAUSS_NP;
        ri_update_pk<GF, NP, ROWS>(lane, ps, st, cells_out, field_out, pot_out, K, cp.sh0, cp.sh1, cp.row0, acc, keep_state);
    }
}
template <int ROWS = 2>
LNX_HD void ri_update_by_mode(int mode, int lane, float* ps, float* __restrict__ st, float* __restrict__ cells_out,
                              float* __restrict__ field_out, float* __restrict__ pot_out, const CellParams2& cp, float* acc,
                              bool keep_state) {  // emulator entry
    switch (mode) {
        case t64h::MODE_PQ4_NP: ri_update_mode<t64h::MODE_PQ4_NP, ROWS>(lane, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        case t64h::MODE_PQ4: ri_update_mode<t64h::MODE_PQ4, ROWS>(lane, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        case t64h::MODE_GAUSS_NP: ri_update_mode<t64h::MODE_GAUSS_NP, ROWS>(lane, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        case t64h::MODE_GAUSS: ri_update_mode<t64h::MODE_GAUSS, ROWS>(lane, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
        default: ri_update_mode<t64h::MODE_DYN, ROWS>(lane, ps, st, cells_out, field_out, pot_out, cp, acc, keep_state); break;
    }
}
// the packed line of the row pair from the two rows of new cells left in shared memory by ri_update(keep_state)
LNX_HD void rf_load_smem(int lane, const float* ps, float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 64; ++n2) v[n2] = make_float2(ps[lane + 32 * n2], ps[N + lane + 32 * n2]);
}


// ---------------------------------------------------------------------------------------------------------------------
// Real-row variant of the two rows kernels (round 2): ONE REAL ROW PER WARP instead of a packed row pair.  The row of 2048 reals is
// the complex line z[n] = x[2n] + i x[2n+1] of 1024 points = 32 x 32 (lane n1 transforms z[n1 + 32 n2] in 64 registers, twiddles
// W_1024^(n1 k2), one exchange through shared memory, lane k2 transforms the other axis), followed by the twiddled untangle
// X[k] = (Z[k] + conj Z[1024-k]) / 2 - i W_2048^k (Z[k] - conj Z[1024-k]) / 2.  Twice the warps with half the dependency chain each:
// a single 2048^2 world is one wave of warps whose chain length IS the kernel time.  Layouts (T[k][row], the
// kernel table, the partial sums of 16 rows per CTA) are unchanged, so lead_kernel and pass D serve both variants.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NC = N / 2;            // complex points of a packed real row
constexpr int RR_WARPS = 16;         // rows per CTA: the sixteen 8-byte pieces of one k are one 128-byte line of T
constexpr int NATS1 = 1058;          // complex values per warp buffer: >= 32 * 33 (exchange) and 1025 (raw half spectrum); even (128-bit row accesses), = 2 mod 16 (the cooperative 8-byte accesses of a warp fall on every bank exactly twice)
constexpr size_t RR_SMEM = (size_t)RR_WARPS * NATS1 * sizeof(float2);

// v[j] *= W_1024^(lane * k2) (INV: conjugate), k2 = br5(j): ten exact table values per lane, every twiddle one product of two
template <bool INV>
LNX_HD void twiddle32(int lane, float2* v, const float2* __restrict__ tab) {
    float2 ta[8], tc[4];
#pragma unroll
    for (int a = 1; a < 8; ++a) ta[a] = LNX_T64_LDG(tab + ((2 * lane * a) & (N - 1)));
#pragma unroll
    for (int c = 1; c < 4; ++c) tc[c] = LNX_T64_LDG(tab + ((16 * lane * c) & (N - 1)));
#pragma unroll
    for (int j = 1; j < 32; ++j) {
        const int k2 = br5(j), a = k2 & 7, c = k2 >> 3;
        float2 w;
        if (c == 0)
            w = ta[a];
        else if (a == 0)
            w = tc[c];
        else
            w = make_float2(ta[a].x * tc[c].x - ta[a].y * tc[c].y, ta[a].y * tc[c].x + ta[a].x * tc[c].y);
        v[j] = INV ? rot_inv(v[j], w.x, w.y) : rot_fwd(v[j], w.x, w.y);
    }
}
// forward: v[n2] = z[lane + 32 n2]  ->  u[j] = Z[lane + 32 br5(j)]
LNX_HD void f1_a(int lane, float2* v, const float2* __restrict__ tab) {
    fft_dif<32>(v);
    twiddle32<false>(lane, v, tab);
}
LNX_HD void f1_store(int lane, const float2* v, float2* ex) {
#pragma unroll
    for (int j = 0; j < 32; ++j) ex[br5(j) * EXS + lane] = v[j];
}
LNX_HD void f1_b(int lane, const float2* ex, float2* u) {
#pragma unroll
    for (int q = 0; q < 32; ++q) u[q] = ex[lane * EXS + q];
    fft_dif<32>(u);
}
LNX_HD void f1_nat_store(int lane, const float2* u, float2* nat) {
#pragma unroll
    for (int j = 0; j < 32; ++j) nat[lane + 32 * br5(j)] = u[j];
}
// inverse (un-normalised): Z in natural order -> v[n2] = z[lane + 32 n2]
LNX_HD void i1_nat_load(int lane, const float2* nat, float2* u) {
#pragma unroll
    for (int j = 0; j < 32; ++j) u[j] = nat[lane + 32 * br5(j)];
    ifft_dit<32>(u);
}
LNX_HD void i1_store(int lane, const float2* u, float2* ex) {
#pragma unroll
    for (int q = 0; q < 32; ++q) ex[lane * EXS + q] = u[q];
}
LNX_HD void i1_b(int lane, const float2* ex, float2* v, const float2* __restrict__ tab) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ex[br5(j) * EXS + lane];
    twiddle32<true>(lane, v, tab);
    ifft_dit<32>(v);
}
LNX_HD void rr_load(int lane, const float* __restrict__ row, float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = LNX_T64_LDG(reinterpret_cast<const float2*>(row) + lane + 32 * n2);
}
LNX_HD void rr_load_smem(int lane, const float* ps, float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = reinterpret_cast<const float2*>(ps)[lane + 32 * n2];
}
LNX_HD void rr_pot_store(int lane, const float2* v, float* ps) {
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) reinterpret_cast<float2*>(ps)[lane + 32 * n2] = v[n2];
}
// CTA-cooperative: thread -> (row buffer w = tid & 15, k = (tid >> 4) + 32 i).  Untangle and write T[k][row0 + w]; dst = T + row0
LNX_HD void rf16_untangle_store(int tid, const float2* nat_all, const float2* __restrict__ tab, float2* __restrict__ dst) {
    const int w = tid & 15, kq = tid >> 4;
    const float2* nat = nat_all + w * NATS1;
#pragma unroll 4
    for (int i = 0; i <= 32; ++i) {
        const int k = kq + 32 * i;
        if (i == 32 && kq != 0) break;  // k = 1024
        const float2 z = nat[k & (NC - 1)], zc = nat[(NC - k) & (NC - 1)], t = LNX_T64_LDG(tab + k);
        const float2 A = pk_add(z, make_float2(zc.x, -zc.y)), B = pk_add(z, make_float2(-zc.x, zc.y));
        const float2 Q = cmul(B, make_float2(-0.5f * t.y, -0.5f * t.x));  // -i/2 W_2048^k B
        dst[(size_t)k * N + w] = pk_fma(A, pk_bc(0.5f), Q);
    }
}
// raw half spectrum X[0..1024] of row w -> nat[w][k]; src = P + row0
LNX_HD void ri16_gather(int tid, const float2* __restrict__ src, float2* nat_all) {
    const int w = tid & 15, kq = tid >> 4;
    float2* nat = nat_all + w * NATS1;
    float2 x[33];
#pragma unroll
    for (int i = 0; i <= 32; ++i) {
        const int k = kq + 32 * i;
        if (i < 32 || kq == 0) x[i] = LNX_T64_LDG(src + (size_t)k * N + w);
    }
#pragma unroll
    for (int i = 0; i <= 32; ++i) {
        if (i < 32 || kq == 0) nat[kq + 32 * i] = x[i];
    }
}
// in place: nat[w][k] = Z[k] = (X[k] + conj X[1024-k]) + i W_2048^-k (X[k] - conj X[1024-k]), k = 0..1023; a thread owns both members of a pair
LNX_HD void ri16_tangle(int tid, float2* nat_all, const float2* __restrict__ tab) {
    const int w = tid & 15, kq = tid >> 4;
    float2* nat = nat_all + w * NATS1;
#pragma unroll 4
    for (int i = 0; i <= 16; ++i) {
        const int k = kq + 32 * i;
        if (i == 16 && kq != 0) break;  // k = 512
        const float2 x = nat[k], xc = nat[NC - k];
        if (k == 0) {  // X[0] and X[1024] are real (their imaginary parts are ignored, as a complex-to-real transform does)
            nat[0] = make_float2(x.x + xc.x, x.x - xc.x);
            continue;
        }
        const float2 t = LNX_T64_LDG(tab + k);
        const float2 A = pk_add(x, make_float2(xc.x, -xc.y)), B = pk_add(x, make_float2(-xc.x, xc.y));
        const float2 Q = cmul(B, make_float2(-t.y, t.x));  // i W_2048^-k B
        nat[k] = pk_add(A, Q);
        if (k != NC / 2) nat[NC - k] = pk_add(make_float2(A.x, -A.y), make_float2(-Q.x, Q.y));  // conj(A - Q)
    }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// kernels (one warp per CTA)
// ---------------------------------------------------------------------------------------------------------------------
struct Extra {            // what the generic argument structs do not carry
    const float2* tw;     // (cos, sin)(2 pi i / 2048), i < 2048
    const float2* ktab;   // [n_sols][1025][2048] kernel table in register order, pre-scaled by 1 / cells
};

// grid (128, 1, worlds), 8 warps: warp j of CTA c owns row pair 8 c + j; dynamic shared memory ROWS_SMEM bytes
constexpr size_t ROWS_SMEM = (size_t)ROWS_WARPS * NATS * sizeof(float2);
__global__ void __launch_bounds__(32 * ROWS_WARPS) rows_fwd_kernel(PassAArgs P, Extra X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* nat_all = reinterpret_cast<float2*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, p = blockIdx.x * ROWS_WARPS + wid, w = blockIdx.z;
    float2* sm = nat_all + wid * NATS;
    float2 v[64];
    rf_load(lane, P.state + (size_t)w * N * N + (size_t)(2 * p) * N, v);
    fs_fwd_a(lane, v, X.tw);
    fs_fwd_store(lane, v, sm);
    __syncwarp();
    fs_fwd_b(lane, sm, v);
    __syncwarp();
    rf_nat_store(lane, v, sm);
    __syncthreads();
    rf8_untangle_store(threadIdx.x, nat_all, P.spec + (size_t)w * SPEC + 2 * ROWS_WARPS * blockIdx.x);
}

// Programmatic dependent launch between the two kernels of a step (and between steps inside one captured graph).  A kernel launched
// with cudaLaunchAttributeProgrammaticStreamSerialization may become resident while its predecessor is still running; what it does
// before pdl_wait() must not touch anything the predecessor writes.  Both step kernels call pdl_launch() right AFTER pdl_wait(): their
// successor can then only start once the kernel BEFORE them has completed, i.e. at most two kernels of the chain overlap and a
// prologue may read what the kernel two places back wrote (rows_inv(t): the state rows of rows_inv(t - 1)).  In a launch without the
// attribute both instructions do nothing.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900  // (the CPU emulator build of this header targets nvcc's default architecture)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <class... KArgs, class... Args>
inline void launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);  // (errors surface through cudaGetLastError / the capture's end)
}

// grid (1025 columns + 1, 1, worlds).  CTA 1025 of world w finalises the statistics of the PREVIOUS step (pass D) if one is pending:
// nothing before rows_inv needs the carry it updates, so that latency-bound single-CTA job leaves the critical path.  The step index
// lives in the world's carry (WorldCarry::step / pending), so the three launches of a step have step-independent arguments and the
// time loop is ONE captured CUDA graph replayed max_run_iter times (the launches are 10-25 us each: issuing them one by one from the
// host was the bottleneck).
__global__ void __launch_bounds__(32) lead_kernel(PassBArgs P, Extra X, PassDArgs D) {
    __shared__ __align__(16) float2 sm[SMEM_C2];
    const int lane = threadIdx.x, k = blockIdx.x, w = blockIdx.z;
    if (k == HALF) {
        pdl_wait();
        pdl_launch();
        if (D.carry[w].pending) tiled::pass_d_body(D, w, D.carry[w].step);
        return;
    }
    const int sol = w / P.n_init;
    const float2* kt = X.ktab + (size_t)sol * SPEC + (size_t)k * N;
#pragma unroll
    for (int j = 0; j < 4; ++j)  // the 16 KB table line is needed after the forward transform: have it in L1 by then
        asm volatile("prefetch.global.L1 [%0];" ::"l"(kt + (j * 32 + lane) * 16));  // (a constant of the scan: before pdl_wait)
    pdl_wait();  // the spectrum of the rows kernel before this launch
    pdl_launch();
    float2 v[64];
    ld_load(lane, P.spec + (size_t)w * SPEC + (size_t)k * N, v);
    fs_fwd_a(lane, v, X.tw);
    fs_fwd_store(lane, v, sm);
    // the two halves of the table line are requested one phase before they are used (all 254 registers are busy otherwise and
    // the multiply waited for its loads)
    float2 ka[32], kb[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) ka[q] = __ldg(kt + q * 32 + lane);
    __syncwarp();
    fs_fwd_b(lane, sm, v);
#pragma unroll
    for (int q = 0; q < 32; ++q) kb[q] = __ldg(kt + (32 + q) * 32 + lane);
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = cmul(v[q], ka[q]);
    ifft_dit<32>(v);
#pragma unroll
    for (int q = 0; q < 32; ++q) v[32 + q] = cmul(v[32 + q], kb[q]);
    ifft_dit<32>(v + 32);
    __syncwarp();
    fs_inv_store(lane, v, sm);
    __syncwarp();
    fs_inv_b(lane, sm, v, X.tw);
    ld_store(lane, P.pot_spec + (size_t)w * SPEC + (size_t)k * N, v);
}

// grid (128, 1, worlds), 8 warps, dynamic shared memory ROWS_SMEM bytes.  next_spec != nullptr: fused step kernel — the updated row
// pair is still in the warp's shared memory, so it is transformed for the NEXT step right away (rows_fwd without its launch, its
// state read and its load latency); the time loop is then lead + this kernel, after one rows_fwd launch for the first step.
template <int MODE>
__global__ void __launch_bounds__(32 * ROWS_WARPS) rows_inv_kernel(PassCArgs P, Extra X, float2* next_spec) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* nat_all = reinterpret_cast<float2*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, p = blockIdx.x * ROWS_WARPS + wid, w = blockIdx.z;
    float2* sm = nat_all + wid * NATS;
    const int sol = w / P.n_init, init = w - sol * P.n_init;
    float* st = P.state + (size_t)w * N * N + (size_t)(2 * p) * N;
#pragma unroll
    for (int j = 0; j < 4; ++j)  // the state rows are needed after the transform: have them in L1 by then
        asm volatile("prefetch.global.L1 [%0];" ::"l"(st + (j * 32 + lane) * 32));  // (written two kernels back: legal before pdl_wait)
    pdl_wait();  // the potential spectrum and the carry of the lead launch before this one
    pdl_launch();
    ri8_gather(threadIdx.x, P.pot_spec + (size_t)w * SPEC + 2 * ROWS_WARPS * blockIdx.x, nat_all);
    __syncthreads();
    float2 v[64];
    ri_nat_load(lane, sm, v);
    fs_inv_a(v);
    __syncwarp();
    fs_inv_store(lane, v, sm);
    __syncwarp();
    fs_inv_b(lane, sm, v, X.tw);
    __syncwarp();
    ri_pot_store(lane, v, reinterpret_cast<float*>(sm));
    __syncwarp();
    const WorldCarry cr = P.carry[w];
    CellParams2 cp;
    cp.gf_id = P.gf_id[0];
    cp.state_fn = P.state_fn;
    cp.mean = P.mean;
    cp.gc = gf_prepare(cp.gf_id, P.gf_params[(size_t)sol * 2], P.gf_params[(size_t)sol * 2 + 1]);
    cp.wk = P.weights[sol];
    cp.wsum = cp.wk;
    cp.dt = P.dt[sol];
    cp.sh0 = cr.shift[0];
    cp.sh1 = cr.shift[1];
    cp.row0 = 2 * p;
    const size_t traj = ((size_t)sol * P.max_iter + cr.step) * P.n_init + init;
    const size_t toff = traj * ((size_t)N * N) + (size_t)(2 * p) * N;
    float acc[NP_T];
    const bool fuse = next_spec != nullptr && cr.step + 1 < P.max_iter;
    ri_update_mode<MODE>(lane, reinterpret_cast<float*>(sm), st, P.cells_out ? P.cells_out + toff : nullptr,
                         P.field_out ? P.field_out + toff : nullptr, P.potential_out ? P.potential_out + toff : nullptr, cp, acc, fuse);
#pragma unroll
    for (int i = 0; i < NP_T; ++i) {
        float x = acc[i];
        if (i < 5 + 3 * MAXD) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        }
        acc[i] = x;
    }
    // the eight row pairs of the CTA are summed here, so pass D (one small CTA inside the next lead launch) reads 128 slabs, not 1024;
    // the warp's sums are parked in shared memory now: they are not live across the transform below
    __shared__ float red[ROWS_WARPS * NP_T];
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NP_T; ++i) red[wid * NP_T + i] = acc[i];
    }
    if (fuse) {  // (uniform over the CTA)
        __syncwarp();
        rf_load_smem(lane, reinterpret_cast<const float*>(sm), v);
        fs_fwd_a(lane, v, X.tw);
        __syncwarp();
        fs_fwd_store(lane, v, sm);
        __syncwarp();
        fs_fwd_b(lane, sm, v);
        __syncwarp();
        rf_nat_store(lane, v, sm);
        __syncthreads();
        rf8_untangle_store(threadIdx.x, nat_all, next_spec + (size_t)w * SPEC + 2 * ROWS_WARPS * blockIdx.x);
    }
    __syncthreads();
    if (threadIdx.x < NP_T) {
        float x = 0.f;
#pragma unroll
        for (int j = 0; j < ROWS_WARPS; ++j) x += red[j * NP_T + threadIdx.x];
        P.partials[((size_t)w * (N / 2 / ROWS_WARPS) + blockIdx.x) * NP_T + threadIdx.x] = x;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) const_cast<WorldCarry*>(P.carry)[w].pending = 1;  // (only pass D, in a later launch, reads it)
}

// rows_inv in the compiled form of this plan (finite: LNX_RUN_ASSUME_FINITE)
// (pdl: launched as a programmatic dependent of the lead launch before it)
inline void launch_rows_inv(const PassCArgs& c, const Extra& x, float2* next_spec, unsigned nw, bool finite, bool pdl, cudaStream_t s) {
    const dim3 grid(N / 2 / ROWS_WARPS, 1, nw);
    switch (t64h::select_mode(c.gf_id[0], c.state_fn, finite)) {
        case t64h::MODE_PQ4: launch_maybe_pdl(rows_inv_kernel<t64h::MODE_PQ4>, grid, 32 * ROWS_WARPS, ROWS_SMEM, s, pdl, c, x, next_spec); break;
        case t64h::MODE_PQ4_NP: launch_maybe_pdl(rows_inv_kernel<t64h::MODE_PQ4_NP>, grid, 32 * ROWS_WARPS, ROWS_SMEM, s, pdl, c, x, next_spec); break;
        case t64h::MODE_GAUSS: launch_maybe_pdl(rows_inv_kernel<t64h::MODE_GAUSS>, grid, 32 * ROWS_WARPS, ROWS_SMEM, s, pdl, c, x, next_spec); break;
        case t64h::MODE_GAUSS_NP: launch_maybe_pdl(rows_inv_kernel<t64h::MODE_GAUSS_NP>, grid, 32 * ROWS_WARPS, ROWS_SMEM, s, pdl, c, x, next_spec); break;
        default: launch_maybe_pdl(rows_inv_kernel<t64h::MODE_DYN>, grid, 32 * ROWS_WARPS, ROWS_SMEM, s, pdl, c, x, next_spec); break;
    }
}
// (pdl: a programmatic dependent of the rows kernel of the previous step in the same graph)
inline void launch_lead(const PassBArgs& b, const Extra& x, const PassDArgs& d, unsigned nw, bool pdl, cudaStream_t s) {
    launch_maybe_pdl(lead_kernel, dim3(1026, 1, nw), 32, 0, s, pdl, b, x, d);
}

// ---- real-row variant: grid (128, 1, worlds), 16 warps; warp j of CTA c owns row 16 c + j; dynamic shared memory RR_SMEM bytes ----
__global__ void __launch_bounds__(32 * RR_WARPS) rows1_fwd_kernel(PassAArgs P, Extra X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* nat_all = reinterpret_cast<float2*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, row = blockIdx.x * RR_WARPS + wid, w = blockIdx.z;
    float2* sm = nat_all + wid * NATS1;
    float2 v[32];
    rr_load(lane, P.state + (size_t)w * N * N + (size_t)row * N, v);
    f1_a(lane, v, X.tw);
    f1_store(lane, v, sm);
    __syncwarp();
    f1_b(lane, sm, v);
    __syncwarp();
    f1_nat_store(lane, v, sm);
    __syncthreads();
    rf16_untangle_store(threadIdx.x, nat_all, X.tw, P.spec + (size_t)w * SPEC + RR_WARPS * blockIdx.x);
}
template <int MODE>
__global__ void __launch_bounds__(32 * RR_WARPS) rows1_inv_kernel(PassCArgs P, Extra X, float2* next_spec) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[RR_WARPS * NP_T];
    float2* nat_all = reinterpret_cast<float2*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, row = blockIdx.x * RR_WARPS + wid, w = blockIdx.z;
    float2* sm = nat_all + wid * NATS1;
    const int sol = w / P.n_init, init = w - sol * P.n_init;
    float* st = P.state + (size_t)w * N * N + (size_t)row * N;
#pragma unroll
    for (int j = 0; j < 2; ++j)  // the state row is needed after the transform: have it in L1 by then
        asm volatile("prefetch.global.L1 [%0];" ::"l"(st + (j * 32 + lane) * 32));
    ri16_gather(threadIdx.x, P.pot_spec + (size_t)w * SPEC + RR_WARPS * blockIdx.x, nat_all);
    __syncthreads();
    ri16_tangle(threadIdx.x, nat_all, X.tw);
    __syncthreads();
    float2 v[32];
    i1_nat_load(lane, sm, v);
    __syncwarp();
    i1_store(lane, v, sm);
    __syncwarp();
    i1_b(lane, sm, v, X.tw);
    __syncwarp();
    rr_pot_store(lane, v, reinterpret_cast<float*>(sm));
    __syncwarp();
    const WorldCarry cr = P.carry[w];
    CellParams2 cp;
    cp.gf_id = P.gf_id[0];
    cp.state_fn = P.state_fn;
    cp.mean = P.mean;
    cp.gc = gf_prepare(cp.gf_id, P.gf_params[(size_t)sol * 2], P.gf_params[(size_t)sol * 2 + 1]);
    cp.wk = P.weights[sol];
    cp.wsum = cp.wk;
    cp.dt = P.dt[sol];
    cp.sh0 = cr.shift[0];
    cp.sh1 = cr.shift[1];
    cp.row0 = row;
    const size_t traj = ((size_t)sol * P.max_iter + cr.step) * P.n_init + init;
    const size_t toff = traj * ((size_t)N * N) + (size_t)row * N;
    float acc[NP_T];
    const bool fuse = next_spec != nullptr && cr.step + 1 < P.max_iter;
    ri_update_mode<MODE, 1>(lane, reinterpret_cast<float*>(sm), st, P.cells_out ? P.cells_out + toff : nullptr,
                            P.field_out ? P.field_out + toff : nullptr, P.potential_out ? P.potential_out + toff : nullptr, cp, acc, fuse);
#pragma unroll
    for (int i = 0; i < NP_T; ++i) {
        float x = acc[i];
        if (i < 5 + 3 * MAXD) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        }
        acc[i] = x;
    }
    if (lane == 0) {  // the sixteen rows of the CTA are summed below: pass D reads 128 slabs
#pragma unroll
        for (int i = 0; i < NP_T; ++i) red[wid * NP_T + i] = acc[i];
    }
    if (fuse) {  // (uniform over the CTA)
        __syncwarp();
        rr_load_smem(lane, reinterpret_cast<const float*>(sm), v);
        f1_a(lane, v, X.tw);
        __syncwarp();
        f1_store(lane, v, sm);
        __syncwarp();
        f1_b(lane, sm, v);
        __syncwarp();
        f1_nat_store(lane, v, sm);
        __syncthreads();
        rf16_untangle_store(threadIdx.x, nat_all, X.tw, next_spec + (size_t)w * SPEC + RR_WARPS * blockIdx.x);
    }
    __syncthreads();
    if (threadIdx.x < NP_T) {
        float x = 0.f;
#pragma unroll
        for (int j = 0; j < RR_WARPS; ++j) x += red[j * NP_T + threadIdx.x];
        P.partials[((size_t)w * (N / RR_WARPS) + blockIdx.x) * NP_T + threadIdx.x] = x;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) const_cast<WorldCarry*>(P.carry)[w].pending = 1;  // (only pass D, in a later launch, reads it)
}
inline void launch_rows1_inv(const PassCArgs& c, const Extra& x, float2* next_spec, unsigned nw, bool finite, cudaStream_t s) {
    const dim3 grid(N / RR_WARPS, 1, nw);
    switch (t64h::select_mode(c.gf_id[0], c.state_fn, finite)) {
        case t64h::MODE_PQ4: (rows1_inv_kernel<t64h::MODE_PQ4>)<<<grid, 32 * RR_WARPS, RR_SMEM, s>>>(c, x, next_spec); break;
        case t64h::MODE_PQ4_NP: (rows1_inv_kernel<t64h::MODE_PQ4_NP>)<<<grid, 32 * RR_WARPS, RR_SMEM, s>>>(c, x, next_spec); break;
        case t64h::MODE_GAUSS: (rows1_inv_kernel<t64h::MODE_GAUSS>)<<<grid, 32 * RR_WARPS, RR_SMEM, s>>>(c, x, next_spec); break;
        case t64h::MODE_GAUSS_NP: (rows1_inv_kernel<t64h::MODE_GAUSS_NP>)<<<grid, 32 * RR_WARPS, RR_SMEM, s>>>(c, x, next_spec); break;
        default: (rows1_inv_kernel<t64h::MODE_DYN>)<<<grid, 32 * RR_WARPS, RR_SMEM, s>>>(c, x, next_spec); break;
    }
}

inline cudaError_t set_rows_inv_attributes() {
    cudaError_t e = cudaFuncSetAttribute(rows_inv_kernel<t64h::MODE_DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_inv_kernel<t64h::MODE_PQ4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_inv_kernel<t64h::MODE_PQ4_NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_inv_kernel<t64h::MODE_GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_inv_kernel<t64h::MODE_GAUSS_NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows1_inv_kernel<t64h::MODE_DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows1_inv_kernel<t64h::MODE_PQ4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows1_inv_kernel<t64h::MODE_PQ4_NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows1_inv_kernel<t64h::MODE_GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rows1_inv_kernel<t64h::MODE_GAUSS_NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM);
    return e;
}

// K_fft full complex [n_sols][nb_slots][2048][2048] (reference layout [m][k]) -> tab[sol][k][q * 32 + t] = K[m = freq_of(q, t)][k] * scale
__global__ void gather_ktab_kernel(const float2* __restrict__ K_fft, float2* __restrict__ tab, int nb_slots, int slot, float scale) {
    const int sol = blockIdx.z;
    const float2* src = K_fft + ((size_t)sol * nb_slots + slot) * ((size_t)N * N);
    float2* dst = tab + (size_t)sol * SPEC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < SPEC; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i >> 11), r = (int)(i & (N - 1));
        const float2 x = src[(size_t)freq_of(r >> 5, r & 31) * N + k];
        dst[i] = make_float2(x.x * scale, x.y * scale);
    }
}
#endif  // __CUDACC__

}  // namespace t2k
}  // namespace lnx
