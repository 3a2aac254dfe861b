// "R16" variant of the resident 128x128 pipeline: 512 compute threads, 16 complex values per thread.
//
// Same five phases / four exchanges as lnx_world128.cuh, but every real row is transformed on its own (in-register
// real FFT of 32 points = complex FFT16 + split), so no two-row packing / untangling is needed and each thread carries
// half the working set: ~96 registers, 16 compute warps per SM (4 per scheduler instead of 2) and a per-step code
// footprint that fits the SM's instruction cache.  See DESIGN.md §3.7.
//
//   P1'  thread (row r, residue l): real FFT32 over j of x[r][4j+l]                       -> Y_l[k1], k1 = 0..16
//   P2'  thread (r1, a, h): k1' = a (h=0) or 32-a (h=1): twiddle, radix-4 over l (2 of 4 outputs), radix-8 over the
//        8 rows r1+16i, column twiddle                                                   -> 2 spectral columns x 8 m2
//   P3'  thread (col, m2): radix-16 over r1, multiply by K, inverse radix-16
//   P4'  mirror of P2' (+ one 16-value shuffle exchange with the Hermitian partner lane), P5' mirror of P1'
#pragma once
#include "lnx_fft.cuh"

namespace lnx {
namespace r16 {

constexpr int WS = 128;
constexpr int NT = 512;       // compute threads per world
constexpr int REGION = 512;   // complex per exchange region (one per warp / row group r1)

struct Regs {
    float2 v[16];
};

// ---- in-register real FFT of 32 points -------------------------------------------------------------------------------
// forward: x[32] real (natural order) -> y[16]: y[k] = 2*Y[k] for k = 1..15, y[0] = (2*Y[0], 2*Y[16])   (both real)
template <int K>
LNX_HD void rfft32_split_fwd(const float2* z, float2* y) {  // z: FFT16 of (x[2m], x[2m+1]) at bit-reversed positions
    if constexpr (K <= 8) {
        const float2 zk = z[bitrev(K, 4)], zc = z[bitrev((16 - K) & 15, 4)];
        const float2 A = make_float2(zk.x + zc.x, zk.y - zc.y);   // Z[k] + conj Z[16-k]
        const float2 B = make_float2(zk.x - zc.x, zk.y + zc.y);   // Z[k] - conj Z[16-k]
        const float2 T = mul_tw<K + 8, 32, false>(B);             // -i W32^k B
        if constexpr (K == 0) {
            y[0] = make_float2(2.f * (zk.x + zk.y), 2.f * (zk.x - zk.y));
        } else if constexpr (K == 8) {
            y[8] = make_float2(A.x + T.x, A.y + T.y);
        } else {
            y[K] = make_float2(A.x + T.x, A.y + T.y);
            y[16 - K] = make_float2(A.x - T.x, T.y - A.y);        // conj(A - T)
        }
        rfft32_split_fwd<K + 1>(z, y);
    }
}
LNX_HD void rfft32_fwd(const float* x, float2* y) {
    float2 z[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) z[m] = make_float2(x[2 * m], x[2 * m + 1]);
    fft_dif<16>(z);
    rfft32_split_fwd<0>(z, y);
}
// inverse (un-normalised, exact inverse structure): y[16] as above but holding Y[k] (k=1..15) and (Y[0], Y[16]) -> x[32]
//   x[n] = sum_{k=0..31} Yfull[k] exp(+2 pi i n k / 32)
template <int K>
LNX_HD void rfft32_split_inv(const float2* y, float2* z) {
    if constexpr (K <= 8) {
        if constexpr (K == 0) {
            z[0] = make_float2(y[0].x + y[0].y, y[0].x - y[0].y);
        } else {
            const float2 yk = y[K], yc = y[16 - K];
            const float2 A = make_float2(yk.x + yc.x, yk.y - yc.y);  // Y[k] + conj Y[16-k]
            const float2 B = make_float2(yk.x - yc.x, yk.y + yc.y);  // Y[k] - conj Y[16-k]
            const float2 T = mul_tw<K + 8, 32, true>(B);             // i conj(W32^k) B
            z[bitrev(K, 4)] = make_float2(A.x + T.x, A.y + T.y);
            if constexpr (K != 8) z[bitrev(16 - K, 4)] = make_float2(A.x - T.x, T.y - A.y);
        }
        rfft32_split_inv<K + 1>(y, z);
    }
}
LNX_HD void rfft32_inv(const float2* y, float* x) {
    float2 z[16];
    rfft32_split_inv<0>(y, z);
    ifft_dit<16>(z);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        x[2 * m] = z[m].x;
        x[2 * m + 1] = z[m].y;
    }
}

// ---- thread index decompositions ------------------------------------------------------------------------------------
LNX_HD int t_group(int u) { return u >> 5; }            // r1 (one warp per row group)
LNX_HD int t_lane(int u) { return u & 31; }
LNX_HD int t_i(int u) { return (u >> 2) & 7; }          // P1/P5: row index inside the group
LNX_HD int t_l(int u) { return u & 3; }
LNX_HD int t_a(int u) { return (u >> 1) & 15; }         // P2/P4
LNX_HD int t_h(int u) { return u & 1; }
LNX_HD int t_col(int u) { return u >> 3; }              // P3
LNX_HD int t_m2(int u) { return u & 7; }
LNX_HD int cell_row(int u) { return t_group(u) + 16 * t_i(u); }
LNX_HD int k1_of(int a, int h) { return h == 0 ? a : (a == 0 ? 16 : 32 - a); }

// ---- shared-memory layouts -------------------------------------------------------------------------------------------
// E1' [i][k1][l]: k1 XOR (i & 3), and the two 16-byte halves of the l-quad swapped for rows i >= 4
LNX_HD int e1_addr(int i, int k1, int l) { return i * 64 + ((k1 ^ (i & 3)) << 2) + (l ^ (((i >> 2) & 1) << 1)); }
// E2' [col][m2] with the four 16-byte units of a column block permuted by g(col) (table found by search, tests check it)
constexpr unsigned long long E2_SWZ_LO = 0xfbebebea50505050ULL, E2_SWZ_HI = 0xfbebebea50505050ULL;  // tools/find_swizzle.py
LNX_HD int e2_swz(int col) { return (int)(((col < 32 ? E2_SWZ_LO : E2_SWZ_HI) >> (2 * (col & 31))) & 3ULL); }
LNX_HD int e2_addr(int col, int m2, int swz) { return col * 8 + ((((m2 >> 1) ^ swz)) << 1) + (m2 & 1); }


// ---- run-time twiddle tables (shared memory, filled once per kernel) -------------------------------------------------
//   TWA: float2 [3 l][32 (a,h)] (c, s) followed by float2 [3 l][32] (sigma*s, sigma*c): angle 2 pi l k1' / 128,
//        sigma = +1 (h=0) / -1 (h=1); lane-contiguous so that 64-bit loads are conflict free
//   TWR: float4 [16 r1][4]   : W128^(r1*m2) for m2 = 1,2 | 3,4 | 5,6 | 7,-   as (cos, sin) pairs
constexpr int TWA_F4 = 32 * 3;
constexpr int TWR_F4 = 16 * 4;
constexpr int TW_TABLE_F4 = TWA_F4 + TWR_F4;
LNX_HD void init_twiddle_table(int u, float4* table, const float2* tw128) {
    if (u < 32) {
        const int a = u >> 1, h = u & 1;
        const float sg = h ? -1.f : 1.f;
#pragma unroll
        float2* t2 = reinterpret_cast<float2*>(table);
        for (int l = 1; l < 4; ++l) {
            const float2 w = tw128[(l * k1_of(a, h)) & 127];
            t2[(l - 1) * 32 + u] = w;
            t2[96 + (l - 1) * 32 + u] = make_float2(sg * w.y, sg * w.x);
        }
    } else if (u < 48) {
        const int r1 = u - 32;
        float2 c[8];
#pragma unroll
        for (int m2 = 1; m2 < 8; ++m2) c[m2 - 1] = tw128[(r1 * m2) & 127];
        c[7] = make_float2(1.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) table[TWA_F4 + r1 * 4 + i] = make_float4(c[2 * i].x, c[2 * i].y, c[2 * i + 1].x, c[2 * i + 1].y);
    }
}
struct Twiddles {
    float2 row[3];   // (c, s) for l = 1..3
    float2 col[7];   // (cos, sin) for m2 = 1..7
};
LNX_HD void load_twiddles(int u, Twiddles& T, const float4* table) {
    const int lane = t_lane(u), r1 = t_group(u);
    const float2* t2 = reinterpret_cast<const float2*>(table);
#pragma unroll
    for (int i = 0; i < 3; ++i) T.row[i] = t2[i * 32 + lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = table[TWA_F4 + r1 * 4 + i];
        T.col[2 * i] = make_float2(t.x, t.y);
        if (i < 3) T.col[2 * i + 1] = make_float2(t.z, t.w);
    }
}
LNX_HD float2 tw_fwd(float2 d, float2 tw) { return rot_fwd(d, tw.x, tw.y); }
LNX_HD float2 tw_inv(float2 d, float2 tw) { return rot_inv(d, tw.x, tw.y); }

// =====================================================================================================================
// P1': x[32] (real, natural j) -> Y_l[k1] stored in E1'
// =====================================================================================================================
LNX_HD void phase1(int u, const float* x, float2* W) {
    float2 y[16];
    rfft32_fwd(x, y);
    float2* reg = W + t_group(u) * REGION;
    const int i = t_i(u), l = t_l(u);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) reg[e1_addr(i, k1, l)] = y[k1];
}

// position (inside the bit-reversed FFT8 output) of m2
LNX_HDC int pos8(int m2) { return bitrev(m2, 3); }

// =====================================================================================================================
// P2': rows i = 0..7 of the group: twiddle + half radix-4 over l -> columns k1', k1'+32; radix-8 over i; column twiddle
// =====================================================================================================================
struct P2State {
    float2 cA[8], cB[8];
};
LNX_HD void phase2_compute(int u, P2State& S2, const float2* W, const float4* twtab) {
    const int a = t_a(u), h = t_h(u);
    Twiddles T;
    load_twiddles(u, T, twtab);
    const bool a0 = a == 0;
    const float sg = a0 ? 0.f : (h ? -1.f : 1.f);  // imaginary sign (conjugate for h = 1); k1' in {0,16}: inputs are real
    const bool pick_y = a0 && h;                   // Y[16] travels in the imaginary slot of k1 = 0
    const float4* reg4 = reinterpret_cast<const float4*>(W + t_group(u) * REGION);
    float2* cA = S2.cA;
    float2* cB = S2.cB;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 q0 = reg4[e1_addr(i, a, 0) >> 1], q1 = reg4[e1_addr(i, a, 2) >> 1];
        // Y_l for l = 0..3
        const float2 Y0 = make_float2(pick_y ? q0.y : q0.x, q0.y * sg), Y1 = make_float2(pick_y ? q0.w : q0.z, q0.w * sg);
        const float2 Y2 = make_float2(pick_y ? q1.y : q1.x, q1.y * sg), Y3 = make_float2(pick_y ? q1.w : q1.z, q1.w * sg);
        // y_l = Y_l * W128^(l k1')   (Y already conjugated for h = 1 through sg)
        const float2 y1 = tw_fwd(Y1, T.row[0]);
        const float2 y2 = tw_fwd(Y2, T.row[1]);
        const float2 y3 = tw_fwd(Y3, T.row[2]);
        const float2 t0 = cadd(Y0, y2), t1 = csub(Y0, y2), t2 = cadd(y1, y3), d = csub(y1, y3);
        float2 XA = cadd(t0, t2);                                  // k2 = 0
        const float2 XB = make_float2(t1.x + d.y, t1.y - d.x);      // k2 = 1: t1 - i d
        if (a0 && !h) XA.y = t0.x - t2.x;                          // pack the (real) Nyquist column 64 with column 0
        cA[i] = XA;
        cB[i] = XB;
    }
    fft_dif<8>(cA);
    fft_dif<8>(cB);
#pragma unroll
    for (int m2 = 1; m2 < 8; ++m2) {
        cA[pos8(m2)] = tw_fwd(cA[pos8(m2)], T.col[m2 - 1]);
        cB[pos8(m2)] = tw_fwd(cB[pos8(m2)], T.col[m2 - 1]);
    }
}
// (a __syncwarp separates the two halves: every lane of the group must have consumed its E1' inputs before the region
//  is overwritten in its E2' view)
LNX_HD void phase2_store(int u, const P2State& S2, float2* W) {
    const int colA = k1_of(t_a(u), t_h(u)), colB = colA + 32;
    const int sA = e2_swz(colA), sB = e2_swz(colB);
    const float2* cA = S2.cA;
    const float2* cB = S2.cB;
    float4* st4 = reinterpret_cast<float4*>(W + t_group(u) * REGION);
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
        const float2 f = cA[pos8(2 * uu)], g = cA[pos8(2 * uu + 1)];
        st4[e2_addr(colA, 2 * uu, sA) >> 1] = make_float4(f.x, f.y, g.x, g.y);
        const float2 p = cB[pos8(2 * uu)], q = cB[pos8(2 * uu + 1)];
        st4[e2_addr(colB, 2 * uu, sB) >> 1] = make_float4(p.x, p.y, q.x, q.y);
    }
}

// =====================================================================================================================
// P3': thread (col, m2): radix-16 over r1 | multiply | inverse radix-16, in place
// =====================================================================================================================
LNX_HD void phase3_load_fft(int u, Regs& R, const float2* W) {
    const int off = e2_addr(t_col(u), t_m2(u), e2_swz(t_col(u)));
#pragma unroll
    for (int r1 = 0; r1 < 16; ++r1) R.v[r1] = W[r1 * REGION + off];
    fft_dif<16>(R.v);
}
LNX_HD void phase3_ifft_store(int u, Regs& R, float2* W) {
    ifft_dit<16>(R.v);
    const int off = e2_addr(t_col(u), t_m2(u), e2_swz(t_col(u)));
#pragma unroll
    for (int r1 = 0; r1 < 16; ++r1) W[r1 * REGION + off] = R.v[r1];
}
// Kt: float4 [8][512]: complex multipliers of slots (2i, 2i+1) of thread u, pre-scaled by 1/(2*128*128)
LNX_HD void phase3_multiply(int u, Regs& R, const float4* Kt) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 k = Kt[i * NT + u];
        R.v[2 * i] = cmul(R.v[2 * i], make_float2(k.x, k.y));
        R.v[2 * i + 1] = cmul(R.v[2 * i + 1], make_float2(k.z, k.w));
    }
}
// spectral index m (along the column transform) of slot `pos` of thread u
LNX_HD int p3_slot_m(int u, int pos) {
    int m1 = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) m1 |= ((pos >> b) & 1) << (3 - b);
    return t_m2(u) + 8 * m1;
}
// packed DC|Nyquist column (threads 0..7 = col 0, m2 = 0..7): G' = G*Kp + conj(G[-m])*Kq through a shared scratch,
// computed by the 32 lanes of warp 0 (4 products each).   scratch: float2 [2][8 threads][16 slots];  Kpq: float4 [16][8]
LNX_HD int bitrev4_rt(int x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }
LNX_HD void phase3_col0_stash(int u, const Regs& R, float2* scratch) {  // scratch [pos][t]: lane-contiguous
    if (u < 8) {
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) scratch[pos * 8 + u] = R.v[pos];
    }
}
LNX_HD void phase3_col0_compute(int u, float2* scratch, const float4* Kpq) {  // u < 32
    const int t = u & 7;  // owner thread = m2
    float2 g[4], gp[4];
    float4 k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pos = (u >> 3) + 4 * i;
        const int m1 = bitrev4_rt(pos);
        const int pm2 = (8 - t) & 7, pm1 = t == 0 ? ((16 - m1) & 15) : (15 - m1);
        g[i] = scratch[pos * 8 + t];
        gp[i] = scratch[bitrev4_rt(pm1) * 8 + pm2];
        k[i] = Kpq[pos * 8 + t];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x = cmul(g[i], make_float2(k[i].x, k[i].y));
        const float2 y = cmul(make_float2(gp[i].x, -gp[i].y), make_float2(k[i].z, k[i].w));
        scratch[128 + ((u >> 3) + 4 * i) * 8 + t] = cadd(x, y);
    }
}
LNX_HD void phase3_col0_fetch(int u, Regs& R, const float2* scratch) {
    if (u < 8) {
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) R.v[pos] = scratch[128 + pos * 8 + u];
    }
}

// =====================================================================================================================
// P4': inverse of P2'.  shfl = callback exchanging a float with lane ^ 1 (device: __shfl_xor_sync; emulator: two passes)
// =====================================================================================================================
struct P4State {       // registers alive across the lane exchange (the emulator runs the two halves separately)
    float2 cA[8], cB[8];   // inverse radix-8 outputs: [0..3] my rows (4h .. 4h+3), [4..7] the partner's rows
};
LNX_HD void phase4_load_ifft(int u, P4State& S, const float2* W, const float4* twtab) {
    const int a = t_a(u), h = t_h(u);
    Twiddles T;
    load_twiddles(u, T, twtab);
    const float4* reg4 = reinterpret_cast<const float4*>(W + t_group(u) * REGION);
    const int colA = k1_of(a, h), colB = colA + 32;
    const int sA = e2_swz(colA), sB = e2_swz(colB);
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
        const float4 f = reg4[e2_addr(colA, 2 * uu, sA) >> 1], g = reg4[e2_addr(colB, 2 * uu, sB) >> 1];
        S.cA[pos8(2 * uu)] = make_float2(f.x, f.y);
        S.cA[pos8(2 * uu + 1)] = make_float2(f.z, f.w);
        S.cB[pos8(2 * uu)] = make_float2(g.x, g.y);
        S.cB[pos8(2 * uu + 1)] = make_float2(g.z, g.w);
    }
    // conj column twiddle; odd m2 get an extra sign for h = 1 so that the inverse radix-8 comes out rotated by 4 rows:
    // register j then holds row (j + 4h) & 7, i.e. registers 0..3 are always "my" rows and 4..7 the partner's
    const float so = h ? -1.f : 1.f;
#pragma unroll
    for (int m2 = 1; m2 < 8; ++m2) {
        const float2 tw = (m2 & 1) ? make_float2(T.col[m2 - 1].x * so, T.col[m2 - 1].y * so) : T.col[m2 - 1];
        S.cA[pos8(m2)] = tw_inv(S.cA[pos8(m2)], tw);
        S.cB[pos8(m2)] = tw_inv(S.cB[pos8(m2)], tw);
    }
    ifft_dit<8>(S.cA);
    ifft_dit<8>(S.cB);
}
// pA/pB: the partner lane's cA[4..7] / cB[4..7] (its values for MY rows)
LNX_HD void phase4_finish_store(int u, const P4State& S, const float2* pA, const float2* pB, float2* W, const float4* twtab) {
    const int a = t_a(u), h = t_h(u);
    float4 row[3];  // (c, s, sigma s, sigma c)
    {
        const float2* t2 = reinterpret_cast<const float2*>(twtab);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float2 cs = t2[i * 32 + t_lane(u)], sg = t2[96 + i * 32 + t_lane(u)];
            row[i] = make_float4(cs.x, cs.y, sg.x, sg.y);
        }
    }
    float4* st4 = reinterpret_cast<float4*>(W + t_group(u) * REGION);
    if (a != 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = 4 * h + j;
            // z = X'[k1'], X'[k1'+32], X'[k1'+64] = conj(partner B), X'[k1'+96] = conj(partner A)
            const float2 z0 = S.cA[j], z1 = S.cB[j], z2 = make_float2(pB[j].x, -pB[j].y), z3 = make_float2(pA[j].x, -pA[j].y);
            const float2 t0 = cadd(z0, z2), t1 = csub(z0, z2), t2 = cadd(z1, z3), d = csub(z1, z3);
            const float2 U0 = cadd(t0, t2), U2 = csub(t0, t2);
            const float2 U1 = make_float2(t1.x - d.y, t1.y + d.x), U3 = make_float2(t1.x + d.y, t1.y - d.x);  // t1 +/- i d
            // Y'_l[a] = U_l * conj(w_l), conjugated for h = 1:  (Ur c - Ui s,  sigma (Ui c + Ur s))
            const float2 o0 = make_float2(U0.x, h ? -U0.y : U0.y);
            const float2 o1 = make_float2(U1.x * row[0].x - U1.y * row[0].y, U1.y * row[0].w + U1.x * row[0].z);
            const float2 o2 = make_float2(U2.x * row[1].x - U2.y * row[1].y, U2.y * row[1].w + U2.x * row[1].z);
            const float2 o3 = make_float2(U3.x * row[2].x - U3.y * row[2].y, U3.y * row[2].w + U3.x * row[2].z);
            st4[e1_addr(i, a, 0) >> 1] = make_float4(o0.x, o0.y, o1.x, o1.y);
            st4[e1_addr(i, a, 2) >> 1] = make_float4(o2.x, o2.y, o3.x, o3.y);
        }
    } else {
        // k1' = 0 (lane h = 0 owns columns 0|64 and 32) and k1' = 16 (lane h = 1 owns columns 16 and 48): both outputs are
        // real and share the k1 = 0 slot.  Each lane handles its 4 rows for BOTH k1' using the partner's columns.
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = 4 * h + j;
            const float2 c0 = h ? pA[j] : S.cA[j], c32 = h ? pB[j] : S.cB[j];   // (X'[0], X'[64]) packed, X'[32]
            const float2 c16 = h ? S.cA[j] : pA[j], c48 = h ? S.cB[j] : pB[j];
            // k1' = 0: z = X0, X32, X64, conj X32 (w = 1): real outputs
            const float e = c0.x + c0.y, f = c0.x - c0.y;
            const float y0_0 = e + 2.f * c32.x, y0_2 = e - 2.f * c32.x, y0_1 = f - 2.f * c32.y, y0_3 = f + 2.f * c32.y;
            // k1' = 16: z = X16, X48, conj X48, conj X16, w_l = W8^l; outputs real: take the real part of U_l conj(w_l)
            const float2 z0 = c16, z1 = c48, z2 = make_float2(c48.x, -c48.y), z3 = make_float2(c16.x, -c16.y);
            const float2 t0 = cadd(z0, z2), t1 = csub(z0, z2), t2 = cadd(z1, z3), d = csub(z1, z3);
            const float2 U0 = cadd(t0, t2), U2 = csub(t0, t2);
            const float2 U1 = make_float2(t1.x - d.y, t1.y + d.x), U3 = make_float2(t1.x + d.y, t1.y - d.x);
            constexpr float r = Tw128::c[16];  // 1/sqrt 2
            const float y16_0 = U0.x;
            const float y16_1 = (U1.x - U1.y) * r;      // Re(U1 (1+i)/sqrt2)
            const float y16_2 = -U2.y;                  // Re(U2 * i)
            const float y16_3 = -(U3.x + U3.y) * r;     // Re(U3 (-1+i)/sqrt2)
            st4[e1_addr(i, 0, 0) >> 1] = make_float4(y0_0, y16_0, y0_1, y16_1);
            st4[e1_addr(i, 0, 2) >> 1] = make_float4(y0_2, y16_2, y0_3, y16_3);
        }
    }
}

// =====================================================================================================================
// P5': Y'_l[k1] -> potential x[32] of row r at columns 4j + l
// =====================================================================================================================
LNX_HD void phase5(int u, float* x, const float2* W) {
    const float2* reg = W + t_group(u) * REGION;
    const int i = t_i(u), l = t_l(u);
    float2 y[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) y[k1] = reg[e1_addr(i, k1, l)];
    rfft32_inv(y, x);
}

}  // namespace r16
}  // namespace lnx
