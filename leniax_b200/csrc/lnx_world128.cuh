// Per-thread phases of the resident 128x128 Lenia step (host+device code).
//
// One CTA owns one world.  256 compute threads hold the 8192-point half spectrum (32 complex per thread) and walk
// five phases per convolution, exchanging data only through one 64 KB shared buffer `W`:
//
//   P1  rows : radix-32 DIF over j      (n = 4j + l)                      thread (p,l): packed rows p / p+64
//   P2  rows : twiddle + radix-4 over l, untangle the two packed real rows,
//       cols : radix-8 DIF over i       (r = r1 + 16 i) + twiddle         thread (r1,a): 8 rows x 4 spectral columns
//   P3  cols : radix-16 DIF over r1, multiply by the kernel spectrum,
//              radix-16 inverse DIT                                      thread (col,bidx): 2 x 16 column entries
//   P4  mirror of P2, P5 mirror of P1 -> potential of rows p (real part) and p+64 (imaginary part)
//
// Follows leniax/core.py:52-102 (get_potential_fft: fftn * K, real(ifftn)) restricted to 2-D 128x128 worlds; the
// real-input symmetry is exploited (two real rows per complex FFT, half spectrum, DC/Nyquist columns packed).
#pragma once
#include "lnx_fft.cuh"

namespace lnx {

constexpr int WS = 128;           // world side
constexpr int NT = 256;           // compute threads per world
constexpr int REGION = 512;       // complex elements per exchange region (one per r1 / row group)
constexpr int W_COMPLEX = 16 * REGION;

struct Regs {
    float2 v[32];  // working set (re-used by every phase)
};
// Run-time twiddles of P2/P4 (they depend on the thread's a / r1) live in a 1.8 KB shared table and are re-read at the
// start of both phases: keeping 26 values in registers for the whole kernel made ptxas spill them to local memory.
//   TW_A: float4 [16 a ][3]  = W128^(l*k1_s) for (s,l) = (0,1),(0,2) | (0,3),(1,1) | (1,2),(1,3)      (cos, sin) pairs
//   TW_R: float4 [16 r1][4]  = W128^(r1*m2)  for m2 = 1,2 | 3,4 | 5,6 | 7,-
constexpr int TW_A_F4 = 16 * 3;
constexpr int TW_R_F4 = 16 * 4;
constexpr int TW_TABLE_F4 = TW_A_F4 + TW_R_F4;
struct Twiddles {
    float2 twr[2][3];  // row twiddles   W128^(l * k1_s), l = 1..3
    float2 twc[7];     // column twiddles W128^(r1 * m2), m2 = 1..7
};

// ---- thread index decompositions -------------------------------------------------------------------------------
LNX_HD int t_group(int tid) { return tid >> 4; }        // G = r1 = p0, 0..15 (two groups per warp)
LNX_HD int t_sub(int tid) { return tid & 15; }          // a (P2/P4) or 4*q + l (P1/P5)
LNX_HD int t_col(int tid) { return tid >> 2; }          // P3: spectral column 0..63 (0 = packed DC|Nyquist)
LNX_HD int t_bidx(int tid) { return tid & 3; }          // P3: unit (m2 pair) 0..3

// k1 handled by P2/P4 thread `a` for s = 0/1
LNX_HD int k1_of(int a, int s) { return s == 0 ? a : (a == 0 ? 16 : 32 - a); }
// spectral column c (0..3) written by P2/P4 thread `a`
LNX_HD int col_of(int a, int c) {
    if (a == 0) return c == 0 ? 0 : (c == 1 ? 32 : (c == 2 ? 16 : 48));
    return c == 0 ? a : (c == 1 ? a + 32 : (c == 2 ? 32 - a : 64 - a));
}

// ---- shared-memory layouts (complex indices inside a 512-element region) -----------------------------------------
// E1 view [q][k1][l] with XOR swizzles: conflict-free for 64-bit accesses of (q,l) lanes at fixed k1 (P1/P5) and
// 128-bit accesses of lanes a -> k1 in {a, 32-a} at fixed q (P2/P4).
// The address is the XOR of a (q, l) part and a k1 part whose bit fields are disjoint or swizzled against each other:
//     bits 7-8: q   | bits 4-6: k1 bits 2-4 | bits 2-3: (k1 ^ q) & 3 | bit 1: (l >> 1) ^ (k1 >> 2) & 1 | bit 0: l & 1
// so a phase computes the part that depends on the THREAD once and every access is one XOR with a compile-time constant (the
// unswizzled bits even fold into the instruction's immediate offset) instead of four or five logic instructions per access.
LNX_HDC int e1_ql(int q, int l) { return q * 128 + ((q & 3) << 2) + (((l >> 1) & 1) << 1) + (l & 1); }
LNX_HDC int e1_k(int k1) { return ((k1 & 28) << 2) + ((k1 & 3) << 2) + (((k1 >> 2) & 1) << 1); }
LNX_HDC int e1_k_lo(int k1) { return e1_k(k1) & 0xE; }    // the bits swizzled against (q, l)
LNX_HDC int e1_k_hi(int k1) { return e1_k(k1) & ~0xE; }   // plain offset
LNX_HDC int e1_ql_lo(int q, int l) { return e1_ql(q, l) & 0xE; }
LNX_HDC int e1_ql_hi(int q, int l) { return e1_ql(q, l) & ~0xE; }
LNX_HDC int e1_addr(int q, int k1, int l) { return e1_ql(q, l) ^ e1_k(k1); }
// &base[idx ^ x] for a swizzle x < 16 elements, with the XOR applied to the BYTE offset: `idx * sizeof(T)` is computed once per phase
// and every access is one logic instruction away from it (index, scale and base were rebuilt for each access before: three
// instructions); the pointer stays "array + integer offset", so the compiler still sees shared memory
template <class T>
LNX_HD T* swz(T* base, int idx_bytes, int x) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + (idx_bytes ^ (x * (int)sizeof(T))));
}
template <class T>
LNX_HD const T* swz(const T* base, int idx_bytes, int x) {
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + (idx_bytes ^ (x * (int)sizeof(T))));
}
// E2 view [col][unit u][2]: unit = the pair of m2 handled together by a P3 thread.  = e2_col(col) ^ (u << 1)
LNX_HDC int e2_col(int col) { return col * 8 + (((col >> 1) & 3) << 1); }
LNX_HDC int e2_addr(int col, int u) { return e2_col(col) ^ (u << 1); }

// ---- twiddle table setup (once per kernel, threads 0..15 fill row `tid` of both tables) ----------------------------
LNX_HD void init_twiddle_table(int tid, float4* table, const float2* tw128 /* [128] = (cos, sin)(2 pi k / 128) */) {
    if (tid < 16) {
        float2 r[6];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int l = 1; l < 4; ++l) r[s * 3 + l - 1] = tw128[(l * k1_of(tid, s)) & 127];
#pragma unroll
        for (int i = 0; i < 3; ++i) table[tid * 3 + i] = make_float4(r[2 * i].x, r[2 * i].y, r[2 * i + 1].x, r[2 * i + 1].y);
        float2 c[8];
#pragma unroll
        for (int m2 = 1; m2 < 8; ++m2) c[m2 - 1] = tw128[(tid * m2) & 127];
        c[7] = make_float2(1.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) table[TW_A_F4 + tid * 4 + i] = make_float4(c[2 * i].x, c[2 * i].y, c[2 * i + 1].x, c[2 * i + 1].y);
    }
}
LNX_HD void load_twiddles(int tid, Twiddles& T, const float4* table) {
    const int a = t_sub(tid), r1 = t_group(tid);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float4 t = table[a * 3 + i];
        (&T.twr[0][0])[2 * i] = make_float2(t.x, t.y);
        (&T.twr[0][0])[2 * i + 1] = make_float2(t.z, t.w);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = table[TW_A_F4 + r1 * 4 + i];
        T.twc[2 * i] = make_float2(t.x, t.y);
        if (i < 3) T.twc[2 * i + 1] = make_float2(t.z, t.w);
    }
}
// multiply by W = (c, -s) (forward) or conj (inverse), tw = (c, s)
LNX_HD float2 tw_fwd(float2 d, float2 tw) { return rot_fwd(d, tw.x, tw.y); }
LNX_HD float2 tw_inv(float2 d, float2 tw) { return rot_inv(d, tw.x, tw.y); }

// =================================================================================================================
// P1: v[j] = (a[p][4j+l], a[p+64][4j+l]) already loaded by the caller.  radix-32 DIF, store E1.
// =================================================================================================================
template <int POS>
LNX_HD void p1_store(const Regs& R, float2* reg, int ql) {  // reg = W, ql = byte offset of element e1_ql(q, l) of this thread's region
    if constexpr (POS < 32) {
        constexpr int k1 = bitrev(POS, 5);
        swz(reg, ql, e1_k_lo(k1))[e1_k_hi(k1)] = R.v[POS];
        p1_store<POS + 1>(R, reg, ql);
    }
}
LNX_HD void phase1(int tid, Regs& R, float2* W) {
    const int sub = t_sub(tid);
    fft_dif<32>(R.v);
    p1_store<0>(R, W, (t_group(tid) * REGION + e1_ql(sub >> 2, sub & 3)) * (int)sizeof(float2));
}

// =================================================================================================================
// P2: load E1 -> (syncwarp by caller) -> row twiddle, radix-4, untangle, column radix-8, column twiddle, store E2
// =================================================================================================================
LNX_HD void phase2_load(int tid, Regs& R, const float2* W) {
    const int a = t_sub(tid);
    const float4* reg4 = reinterpret_cast<const float4*>(W);
    const int g4 = t_group(tid) * (REGION / 2);  // float4 index: bit 0 of the element index is l & 1 = 0
    const int ks[2] = {(g4 + (e1_k(k1_of(a, 0)) >> 1)) * (int)sizeof(float4), (g4 + (e1_k(k1_of(a, 1)) >> 1)) * (int)sizeof(float4)};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 t = swz(reg4, ks[s], e1_ql_lo(q, 2 * h) >> 1)[e1_ql_hi(q, 2 * h) >> 1];
                R.v[q * 8 + s * 4 + 2 * h] = make_float2(t.x, t.y);
                R.v[q * 8 + s * 4 + 2 * h + 1] = make_float2(t.z, t.w);
            }
        }
}

// untangle one (Z[k], Z[128-k]) pair of a packed row into the spectra of its two real rows (factor 2 kept)
LNX_HD void untangle_pair(float2 z, float2 zc, float2& A2, float2& B2) {
    A2 = pk_add(z, make_float2(zc.x, -zc.y));
    B2 = pk_add(make_float2(z.y, -z.x), pk_swap(zc));
}
// inverse of the above: from the spectra A', B' of two real rows rebuild Z'[k] and Z'[128-k]
LNX_HD void retangle_pair(float2 A, float2 B, float2& z, float2& zc) {
    z = pk_add(A, make_float2(-B.y, B.x));
    zc = pk_add(make_float2(A.x, -A.y), pk_swap(B));
}

// Lanes a = 0 (k1 in {0, 16}: columns {0|64 packed, 32, 16, 48}) pair the entries differently from lanes a > 0 (columns a, a+32, 32-a,
// 64-a).  Both cases run through the SAME instructions with per-lane selects: every warp holds two a = 0 lanes, so a branch made each
// warp execute both variants one after the other (and merge their registers with ~40 moves) in phase 2 and in phase 4.
LNX_HD float2 sel2(bool c, float2 x, float2 y) { return make_float2(c ? x.x : y.x, c ? x.y : y.y); }
LNX_HD void p2_untangle(bool a0, const float2* z /* [q*8 + s*4 + k2] */, float2* h /* [c*8 + i] */) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2* zq = z + q * 8;
        // a > 0: a <-> (32-a)+96, a+32 <-> (32-a)+64, 32-a <-> a+96, 64-a <-> a+64;   a = 0: 32 <-> 96, 16 <-> 112, 48 <-> 80
        float2 g0, g0b;
        untangle_pair(zq[0], zq[4 + 3], g0, g0b);
        h[0 * 8 + q] = sel2(a0, make_float2(2.f * zq[0].x, 2.f * zq[2].x), g0);
        h[0 * 8 + q + 4] = sel2(a0, make_float2(2.f * zq[0].y, 2.f * zq[2].y), g0b);
        untangle_pair(zq[1], sel2(a0, zq[3], zq[4 + 2]), h[1 * 8 + q], h[1 * 8 + q + 4]);
        untangle_pair(zq[4 + 0], sel2(a0, zq[4 + 3], zq[3]), h[2 * 8 + q], h[2 * 8 + q + 4]);
        untangle_pair(zq[4 + 1], sel2(a0, zq[4 + 2], zq[2]), h[3 * 8 + q], h[3 * 8 + q + 4]);
    }
}
LNX_HD void p4_retangle(bool a0, const float2* h /* [c*8 + i] */, float2* z /* [q*8 + s*4 + k2] */) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2* zq = z + q * 8;
        float2 u0, v0, v1, v2, v3;
        retangle_pair(h[0 * 8 + q], h[0 * 8 + q + 4], u0, v0);
        retangle_pair(h[1 * 8 + q], h[1 * 8 + q + 4], zq[1], v1);
        retangle_pair(h[2 * 8 + q], h[2 * 8 + q + 4], zq[4 + 0], v2);
        retangle_pair(h[3 * 8 + q], h[3 * 8 + q + 4], zq[4 + 1], v3);
        zq[0] = sel2(a0, make_float2(h[0 * 8 + q].x, h[0 * 8 + q + 4].x), u0);
        zq[2] = sel2(a0, make_float2(h[0 * 8 + q].y, h[0 * 8 + q + 4].y), v3);
        zq[3] = sel2(a0, v1, v2);
        zq[4 + 2] = sel2(a0, v3, v1);
        zq[4 + 3] = sel2(a0, v2, v0);
    }
}

// position (in the bit-reversed 8-point output) of the two m2 of unit u:  u0=(0,4) u1=(1,7) u2=(2,6) u3=(3,5)
LNX_HDC int unit_m2(int u, int e) { return u == 0 ? (e == 0 ? 0 : 4) : (e == 0 ? u : 8 - u); }
LNX_HDC int unit_pos(int u, int e) { return bitrev(unit_m2(u, e), 3); }

LNX_HD void phase2_compute_store(int tid, Regs& R, float2* W, const float4* twtab) {
    const int a = t_sub(tid);
    Twiddles T;
    load_twiddles(tid, T, twtab);
    // row twiddle + radix-4 over l (forward, W4 = -i)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float2* y = R.v + q * 8 + s * 4;
            const float2 y0 = y[0];
            const float2 y1 = tw_fwd(y[1], T.twr[s][0]);
            const float2 y2 = tw_fwd(y[2], T.twr[s][1]);
            const float2 y3 = tw_fwd(y[3], T.twr[s][2]);
            const float2 t0 = cadd(y0, y2), t1 = csub(y0, y2), t2 = cadd(y1, y3);
            const float2 d = csub(y1, y3);
            const float2 t3 = make_float2(d.y, -d.x);  // * (-i)
            y[0] = cadd(t0, t2);
            y[2] = csub(t0, t2);
            y[1] = cadd(t1, t3);
            y[3] = csub(t1, t3);
        }
    float2 h[32];
    p2_untangle(a == 0, R.v, h);
    float4* reg4 = reinterpret_cast<float4*>(W);
    const int g4 = t_group(tid) * (REGION / 2);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2* hc = h + c * 8;
        fft_dif<8>(hc);  // over i; output position pos <-> m2 = bitrev3(pos)
#pragma unroll
        for (int pos = 1; pos < 8; ++pos) hc[pos] = tw_fwd(hc[pos], T.twc[bitrev(pos, 3) - 1]);
        const int ec = (g4 + (e2_col(col_of(a, c)) >> 1)) * (int)sizeof(float4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 f = hc[unit_pos(u, 0)], g = hc[unit_pos(u, 1)];
            *swz(reg4, ec, u) = make_float4(f.x, f.y, g.x, g.y);
        }
    }
}

// =================================================================================================================
// P3: load E2 (all 16 regions), radix-16 DIF over r1  |  multiply  |  radix-16 inverse DIT, store back in place
//     v[h*16 + pos], pos <-> m1 = bitrev4(pos), m = m2(bidx,h) + 8*m1
// =================================================================================================================
LNX_HD void phase3_load_fft(int tid, Regs& R, const float2* W) {
    const float4* w4 = reinterpret_cast<const float4*>(W);
    const int off = e2_addr(t_col(tid), t_bidx(tid)) >> 1;
#pragma unroll
    for (int r1 = 0; r1 < 16; ++r1) {
        const float4 t = w4[r1 * (REGION / 2) + off];
        R.v[r1] = make_float2(t.x, t.y);
        R.v[16 + r1] = make_float2(t.z, t.w);
    }
    fft_dif<16>(R.v);
    fft_dif<16>(R.v + 16);
}
LNX_HD void phase3_ifft_store(int tid, Regs& R, float2* W) {
    ifft_dit<16>(R.v);
    ifft_dit<16>(R.v + 16);
    float4* w4 = reinterpret_cast<float4*>(W);
    const int off = e2_addr(t_col(tid), t_bidx(tid)) >> 1;
#pragma unroll
    for (int r1 = 0; r1 < 16; ++r1)
        w4[r1 * (REGION / 2) + off] = make_float4(R.v[r1].x, R.v[r1].y, R.v[16 + r1].x, R.v[16 + r1].y);
}

// slot of -m for the packed column (col 0): partner of slot (h,pos)
LNX_HDC int col0_partner(int bidx0, int slot) {
    const int h = slot >> 4, pos = slot & 15;
    if (bidx0 && h == 0) return bitrev((16 - bitrev(pos, 4)) & 15, 4);  // m2 = 0: m1 <-> -m1
    if (bidx0) return 16 + (15 - pos);                                  // m2 = 4: m1 <-> 15 - m1 (same h)
    return (1 - h) * 16 + (15 - pos);                                   // m2 = b <-> 8-b, m1 <-> 15 - m1
}

// Kt: float4 [16][256] = complex multipliers for slots (2i, 2i+1) of thread tid (already scaled by 1/(2*128*128)).
template <int I>
LNX_HD void p3_mul_generic(Regs& R, const float4* Kt, int tid) {
    if constexpr (I < 16) {
        const float4 k = Kt[I * NT + tid];
        R.v[2 * I] = cmul(R.v[2 * I], make_float2(k.x, k.y));
        R.v[2 * I + 1] = cmul(R.v[2 * I + 1], make_float2(k.z, k.w));
        p3_mul_generic<I + 1>(R, Kt, tid);
    }
}
// Packed DC|Nyquist column (threads 0..3): G' = G*Kp + conj(G[-m])*Kq.  Done through a 2 KB shared scratch by all 32
// lanes of warp 0 in a rolled loop (4 products per lane) so that the special case costs ~60 instructions of code instead
// of 400+ unrolled ones (the loop body of this kernel is instruction-fetch bound).
// Kpq: float4 [32 slots][4 threads] = (Kp, Kq).   scratch: float2 [2][4][32].
constexpr int KPQ_LANES = 4;
LNX_HD int bitrev4_rt(int x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }
LNX_HD int col0_partner_rt(bool b0, int s) {
    const int h = s >> 4, pos = s & 15;
    if (b0) return h == 0 ? bitrev4_rt((16 - bitrev4_rt(pos)) & 15) : 16 + (15 - pos);
    return (1 - h) * 16 + (15 - pos);
}
LNX_HD void phase3_col0_stash(int tid, const Regs& R, float2* scratch) {
    if (tid < 4) {  // 64-bit stores straight from the (re, im) register pairs: 128-bit ones cost four register moves each
        float2* d = scratch + tid * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) d[i] = R.v[i];
    }
}
LNX_HD void phase3_col0_compute(int tid, float2* scratch, const float4* Kpq) {  // tid < 32
    const int t = tid & 3;
    float2 g[4], gp[4];
    float4 k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // all loads first: the four products of a lane overlap their shared-memory latency
        const int sl = (tid >> 2) + 8 * i;
        g[i] = scratch[t * 32 + sl];
        gp[i] = scratch[t * 32 + col0_partner_rt(t == 0, sl)];
        k[i] = Kpq[sl * KPQ_LANES + t];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = cmul(g[i], make_float2(k[i].x, k[i].y));
        const float2 b = cmul(make_float2(gp[i].x, -gp[i].y), make_float2(k[i].z, k[i].w));
        scratch[128 + t * 32 + (tid >> 2) + 8 * i] = cadd(a, b);
    }
}
LNX_HD void phase3_col0_fetch(int tid, Regs& R, const float2* scratch) {
    if (tid < 4) {
        const float2* d = scratch + 128 + tid * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) R.v[i] = d[i];
    }
}
LNX_HD void phase3_multiply(int tid, Regs& R, const float4* Kt) { p3_mul_generic<0>(R, Kt, tid); }

// =================================================================================================================
// P4: load E2 (own region) -> inverse twiddle, inverse radix-8 over m2, retangle, inverse radix-4, twiddle, store E1
// =================================================================================================================
LNX_HD void phase4_load(int tid, Regs& R, const float2* W) {
    const int a = t_sub(tid);
    const float4* reg4 = reinterpret_cast<const float4*>(W);
    const int g4 = t_group(tid) * (REGION / 2);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int ec = (g4 + (e2_col(col_of(a, c)) >> 1)) * (int)sizeof(float4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 t = *swz(reg4, ec, u);
            R.v[c * 8 + unit_pos(u, 0)] = make_float2(t.x, t.y);
            R.v[c * 8 + unit_pos(u, 1)] = make_float2(t.z, t.w);
        }
    }
}
LNX_HD void phase4_compute_store(int tid, Regs& R, float2* W, const float4* twtab) {
    const int a = t_sub(tid);
    Twiddles T;
    load_twiddles(tid, T, twtab);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2* hc = R.v + c * 8;
#pragma unroll
        for (int pos = 1; pos < 8; ++pos) hc[pos] = tw_inv(hc[pos], T.twc[bitrev(pos, 3) - 1]);
        ifft_dit<8>(hc);  // -> natural i
    }
    float2 z[32];
    p4_retangle(a == 0, R.v, z);
    float4* reg4 = reinterpret_cast<float4*>(W);
    const int g4 = t_group(tid) * (REGION / 2);
    const int ks[2] = {(g4 + (e1_k(k1_of(a, 0)) >> 1)) * (int)sizeof(float4), (g4 + (e1_k(k1_of(a, 1)) >> 1)) * (int)sizeof(float4)};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float2* y = z + q * 8 + s * 4;
            const float2 t0 = cadd(y[0], y[2]), t1 = csub(y[0], y[2]), t2 = cadd(y[1], y[3]);
            const float2 d = csub(y[1], y[3]);
            const float2 t3 = make_float2(-d.y, d.x);  // * (+i)
            const float2 u0 = cadd(t0, t2);
            const float2 u2 = tw_inv(csub(t0, t2), T.twr[s][1]);
            const float2 u1 = tw_inv(cadd(t1, t3), T.twr[s][0]);
            const float2 u3 = tw_inv(csub(t1, t3), T.twr[s][2]);
            swz(reg4, ks[s], e1_ql_lo(q, 0) >> 1)[e1_ql_hi(q, 0) >> 1] = make_float4(u0.x, u0.y, u1.x, u1.y);
            swz(reg4, ks[s], e1_ql_lo(q, 2) >> 1)[e1_ql_hi(q, 2) >> 1] = make_float4(u2.x, u2.y, u3.x, u3.y);
        }
}

// =================================================================================================================
// P5: load E1, radix-32 inverse DIT -> v[j] = (potential[p][4j+l], potential[p+64][4j+l])
// =================================================================================================================
template <int POS>
LNX_HD void p5_load(Regs& R, const float2* reg, int ql) {
    if constexpr (POS < 32) {
        constexpr int k1 = bitrev(POS, 5);
        R.v[POS] = swz(reg, ql, e1_k_lo(k1))[e1_k_hi(k1)];
        p5_load<POS + 1>(R, reg, ql);
    }
}
LNX_HD void phase5_load(int tid, Regs& R, const float2* W) {
    const int sub = t_sub(tid);
    p5_load<0>(R, W, (t_group(tid) * REGION + e1_ql(sub >> 2, sub & 3)) * (int)sizeof(float2));
}
LNX_HD void phase5_ifft(Regs& R) { ifft_dit<32>(R.v); }

// =================================================================================================================
// State layout (thread-private, float4 index i*256 + tid): i = 0..7 row p, i = 8..15 row p+64; element e of float4 i
// is column 4*(4*(i&7) + e) + l.  Cell coordinates of thread tid:
// =================================================================================================================
LNX_HD int cell_row(int tid, int half) { return t_group(tid) + 16 * (t_sub(tid) >> 2) + 64 * half; }
LNX_HD int cell_col(int tid, int j) { return 4 * j + (t_sub(tid) & 3); }

// Kernel-spectrum table slot -> (m, m2, ...) used by the table builder (device gather kernel and the emulator)
LNX_HD int p3_slot_m(int tid, int slot) {
    const int bidx = t_bidx(tid), h = slot >> 4, pos = slot & 15;
    const int m2 = bidx == 0 ? (h == 0 ? 0 : 4) : (h == 0 ? bidx : 8 - bidx);
    int m1 = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) m1 |= ((pos >> b) & 1) << (3 - b);
    return m2 + 8 * m1;
}

}  // namespace lnx
