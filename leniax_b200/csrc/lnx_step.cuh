// Growth functions, state update, per-cell statistics partials and the per-step statistics finaliser
// (host+device; the CPU emulator in tests/ runs the same code).
//
// Reference: leniax/growth_functions.py:6-253, leniax/core.py:202-319 (weighted mean/sum, get_state*),
// leniax/statistics.py:36-126 (compute_stats), :134-205 + :287-333 (check_heuristics), leniax/utils.py:269-293.
#pragma once
#include <cstring>
#include "lnx_world128.cuh"

namespace lnx {

constexpr int MAX_C = 8;
constexpr int MAX_K = 32;
constexpr float EPS = 1e-7f;  // leniax/constant.py:7

enum GrowthFn { GF_POLY_QUAD4 = 0, GF_GAUSSIAN = 1, GF_GAUSSIAN_TARGET = 2, GF_STEP = 3, GF_STAIRCASE = 4, GF_TRIANGLE = 5, GF_IDENTITY = 6, GF_COUNT = 7 };
enum StateFn { SF_V1 = 0, SF_V2 = 1, SF_SIMPLE = 2, SF_COUNT = 3 };

// order of the per-step scalar statistics in the output tensor (leniax/statistics.py:102-115, channel_mass apart)
enum StatKey { ST_MASS = 0, ST_MASS_VOLUME, ST_MASS_DENSITY, ST_GROWTH, ST_GROWTH_VOLUME, ST_GROWTH_DENSITY, ST_MASS_SPEED,
               ST_MASS_ANGLE_SPEED, ST_MASS_GROWTH_DIST, ST_INERTIA, ST_POTENTIAL_VOLUME, ST_COUNT };

// per-thread partial sums reduced over the CTA each step
enum Partial { PT_CNT_A = 0, PT_G00, PT_CNT_G, PT_CNT_P, PT_MX_R, PT_MX_C, PT_MX2_R, PT_MX2_C, PT_GX_R, PT_GX_C, PT_M00_C0, PT_FIXED = PT_M00_C0 };

struct GfConst {
    float m, s, k0, k1, k2, k3;
};
LNX_HD GfConst gf_prepare(int gf, float m, float s) {
    GfConst g;
    g.m = m;
    g.s = s;
    g.k0 = g.k1 = g.k2 = g.k3 = 0.f;
    if (gf == GF_POLY_QUAD4) {
        g.k0 = 1.0f / (9.0f * (s * s));  // growth_functions.py:41
    } else if (gf == GF_GAUSSIAN || gf == GF_GAUSSIAN_TARGET) {
        g.k0 = 1.0f / s;
    } else if (gf == GF_STAIRCASE) {
        g.k0 = m - s;
        g.k1 = m - s / 2;
        g.k2 = m + s / 2;
        g.k3 = m + s;
    } else if (gf == GF_TRIANGLE) {
        g.k0 = m - s;
        g.k1 = m + s;
        g.k2 = 1.0f / (m - g.k0);
        g.k3 = 1.0f / (m - g.k1);
    }
    return g;
}

// NP = propagate NaN exactly like jnp.maximum / jnp.clip do (needed when s == 0 or a zero weight row can appear)
// XD = use true IEEE divisions where the reference divides (generic kernel); the fused kernel multiplies by reciprocals
template <int GF, bool NP, bool XD = false>
LNX_HD float growth(float X, const GfConst& g) {
    if constexpr (GF == GF_POLY_QUAD4) {
        const float t = X - g.m;
        float o = XD ? 1.0f - (t * t) / (9.0f * (g.s * g.s)) : 1.0f - (t * t) * g.k0;
        if constexpr (NP)
            o = (o < 0.f) ? 0.f : o;
        else
            o = fmaxf(o, 0.f);
        const float o2 = o * o;
        return 2.0f * (o2 * o2) - 1.0f;
    } else if constexpr (GF == GF_GAUSSIAN) {
        const float t = XD ? (X - g.m) / g.s : (X - g.m) * g.k0;
        return 2.0f * expf(-(t * t) / 2.0f) - 1.0f;
    } else if constexpr (GF == GF_GAUSSIAN_TARGET) {
        const float t = XD ? (X - g.m) / g.s : (X - g.m) * g.k0;
        return expf(-(t * t) / 2.0f);
    } else if constexpr (GF == GF_STEP) {
        return (fabsf(X - g.m) <= g.s) ? 1.0f : -1.0f;
    } else if constexpr (GF == GF_STAIRCASE) {
        float o = (X >= g.k0 && X < g.k1) ? 0.5f : 0.f;
        o += (X >= g.k1 && X <= g.k2) ? 1.0f : 0.f;
        o += (X > g.k2 && X <= g.k3) ? 0.5f : 0.f;
        return 2.0f * o - 1.0f;
    } else if constexpr (GF == GF_TRIANGLE) {
        float o = (X >= g.k0 && X < g.m) ? (XD ? (X - g.k0) / (g.m - g.k0) : (X - g.k0) * g.k2) : 0.f;
        o += (X >= g.m && X <= g.k1) ? (XD ? (X - g.k1) / (g.m - g.k1) : (X - g.k1) * g.k3) : 0.f;
        return 2.0f * o - 1.0f;
    } else {
        return X;
    }
}
// growth of a whole per-thread vector with the function selected OUTSIDE the (fully unrolled) loop: v stays in registers.
// poly_quad4 multiplies by the reciprocal of 9 s^2 (as the fused kernel does); the exponential / piecewise-linear
// functions keep the true divisions of the reference (they are the ones whose rounding showed in the aquarium fixture).
template <int GF, bool NP, int N>
LNX_HD void growth_vec(float2* v, const GfConst& g, float& cnt_p) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        cnt_p += (v[j].x > EPS ? 1.f : 0.f) + (v[j].y > EPS ? 1.f : 0.f);  // statistics.py:70
        v[j].x = growth<GF, NP, GF != GF_POLY_QUAD4>(v[j].x, g);
        v[j].y = growth<GF, NP, GF != GF_POLY_QUAD4>(v[j].y, g);
    }
}
template <bool NP, int N>
LNX_HD void growth_vec_dyn(int gf, float2* v, const GfConst& g, float& cnt_p) {
    switch (gf) {
        case GF_POLY_QUAD4: growth_vec<GF_POLY_QUAD4, NP, N>(v, g, cnt_p); break;
        case GF_GAUSSIAN: growth_vec<GF_GAUSSIAN, NP, N>(v, g, cnt_p); break;
        case GF_GAUSSIAN_TARGET: growth_vec<GF_GAUSSIAN_TARGET, NP, N>(v, g, cnt_p); break;
        case GF_STEP: growth_vec<GF_STEP, NP, N>(v, g, cnt_p); break;
        case GF_STAIRCASE: growth_vec<GF_STAIRCASE, NP, N>(v, g, cnt_p); break;
        case GF_TRIANGLE: growth_vec<GF_TRIANGLE, NP, N>(v, g, cnt_p); break;
        default: growth_vec<GF_IDENTITY, NP, N>(v, g, cnt_p); break;
    }
}
template <bool NP>
LNX_HD float growth_dyn(int gf, float X, const GfConst& g) {
    switch (gf) {
        case GF_POLY_QUAD4: return growth<GF_POLY_QUAD4, NP, true>(X, g);
        case GF_GAUSSIAN: return growth<GF_GAUSSIAN, NP, true>(X, g);
        case GF_GAUSSIAN_TARGET: return growth<GF_GAUSSIAN_TARGET, NP, true>(X, g);
        case GF_STEP: return growth<GF_STEP, NP, true>(X, g);
        case GF_STAIRCASE: return growth<GF_STAIRCASE, NP, true>(X, g);
        case GF_TRIANGLE: return growth<GF_TRIANGLE, NP, true>(X, g);
        default: return X;
    }
}

template <int SF, bool NP>
LNX_HD float state_update(float a, float f, float dt) {  // core.py:245-319
    if constexpr (SF == SF_V1) {
        const float n = a + dt * f;
        if constexpr (NP)
            return (n < 0.f) ? 0.f : ((n > 1.f) ? 1.f : n);
        else
            return fminf(fmaxf(n, 0.f), 1.f);
    } else if constexpr (SF == SF_V2) {
        return a * (1.0f - dt) + dt * f;
    } else {
        return a + dt * f;
    }
}
template <bool NP>
LNX_HD float state_update_dyn(int sf, float a, float f, float dt) {
    return sf == SF_V1 ? state_update<SF_V1, NP>(a, f, dt) : (sf == SF_V2 ? state_update<SF_V2, NP>(a, f, dt) : state_update<SF_SIMPLE, NP>(a, f, dt));
}

// centred coordinate of source index `idx` once the world is rolled by -shift (utils.py:269-293 + statistics.py:28-33)
LNX_HD float rolled_coord(int idx, int shift) { return (float)(((idx - shift) & (WS - 1)) - WS / 2); }

struct CellAcc {
    float sa0, sa1;   // sum of cells in row p / p+64 (all channels)
    float sg0, sg1;   // sum of positive field
    float cnt_a, cnt_g, cnt_p;  // counts kept as floats (exact up to 2^24): FSET.BF + FADD per cell
    float mxc, mx2c, gxc;       // column-coordinate moments
    LNX_HD void clear() { sa0 = sa1 = sg0 = sg1 = mxc = mx2c = gxc = cnt_a = cnt_g = cnt_p = 0.f; }
};

// statistics contribution of one column j of the thread's two rows: a0/a1 = cells, f0/f1 = field
LNX_HD void acc_cells(CellAcc& A, float xc, float a0, float a1, float f0, float f1) {
    A.sa0 += a0;
    A.sa1 += a1;
    A.cnt_a += (a0 > EPS ? 1.f : 0.f) + (a1 > EPS ? 1.f : 0.f);
    const float g0 = fmaxf(f0, 0.f), g1 = fmaxf(f1, 0.f);  // statistics.py:65
    A.sg0 += g0;
    A.sg1 += g1;
    A.cnt_g += (f0 > EPS ? 1.f : 0.f) + (f1 > EPS ? 1.f : 0.f);  // max(f, 0) > eps <=> f > eps
    const float as = a0 + a1, gs = g0 + g1;
    const float ax = as * xc;
    A.mxc += ax;
    A.mx2c += ax * xc;
    A.gxc += gs * xc;
}
LNX_HD float opaque(float x) {  // hide the integer origin of a value so ptxas keeps the arithmetic on the FP pipe
#ifdef __CUDA_ARCH__
    asm volatile("" : "+f"(x));
#endif
    return x;
}
LNX_HD float col_coord(float base /* (l - shift1) & 127 - 64 */, int j) {
    const float t = base + (float)(4 * j);
    return t >= 64.f ? t - 128.f : t;
}

// ---- fused single-channel single-kernel cell phase (the north-star fast path) ----
struct FusedConsts {
    GfConst gf;
    float c;   // w * (1 / sum_k W[0][k]) for weighted_mean, w for weighted_sum (core.py:202-242 with one kernel)
    float c2;  // 2 c
    float dt;  // 1 / T of this solution (runner.py:307)
};
LNX_HD FusedConsts fused_consts(int gf, float m, float s, float w, int mean, float dt) {
    FusedConsts K;
    K.gf = gf_prepare(gf, m, s);
    K.c = mean ? w * (1.0f / w) : w;
    K.c2 = 2.0f * K.c;
    K.dt = dt;
    return K;
}
// field = c * growth(X); for poly_quad4 the affine tail 2 o^4 - 1 and the weight fold into one FMA
template <int GF, bool NP>
LNX_HD float field_fused(float X, const FusedConsts& K) {
    if constexpr (GF == GF_POLY_QUAD4) {
        const float t = X - K.gf.m;
        float o = 1.0f - (t * t) * K.gf.k0;
        if constexpr (NP)
            o = (o < 0.f) ? 0.f : o;
        else
            o = fmaxf(o, 0.f);
        const float o2 = o * o;
        return K.c2 * (o2 * o2) - K.c;
    } else {
        return K.c * growth<GF, NP>(X, K.gf);
    }
}
LNX_HD float saturate01(float x) {
#ifdef __CUDA_ARCH__
    return __saturatef(x);
#else
    return fminf(fmaxf(x, 0.f), 1.f);
#endif
}
template <int GF, int SF, bool NP>
LNX_HD void cells_fused(int tid, const float2* pot /* [32] */, float4* A4, const FusedConsts& K, int shift0, int shift1,
                        float* part /* [NPART][256] */) {
    const int l = t_sub(tid) & 3;
    const float xr0 = rolled_coord(cell_row(tid, 0), shift0), xr1 = rolled_coord(cell_row(tid, 1), shift0);
    const float cbase = opaque((float)(((l - shift1) & (WS - 1)) - WS / 2));
    CellAcc A;
    A.clear();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 c0 = A4[i * NT + tid], c1 = A4[(8 + i) * NT + tid];
        const float a0[4] = {c0.x, c0.y, c0.z, c0.w}, a1[4] = {c1.x, c1.y, c1.z, c1.w};
        float n0[4], n1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * i + e;
            const float p0 = pot[j].x, p1 = pot[j].y;
            A.cnt_p += (p0 > EPS ? 1.f : 0.f) + (p1 > EPS ? 1.f : 0.f);  // statistics.py:70
            const float f0 = field_fused<GF, NP>(p0, K), f1 = field_fused<GF, NP>(p1, K);
            acc_cells(A, col_coord(cbase, j), a0[e], a1[e], f0, f1);
            if constexpr (SF == SF_V1 && !NP) {  // clip(a + dt f, 0, 1) as one saturating FMA (no NaN possible here)
                n0[e] = saturate01(a0[e] + K.dt * f0);
                n1[e] = saturate01(a1[e] + K.dt * f1);
            } else {
                n0[e] = state_update<SF, NP>(a0[e], f0, K.dt);
                n1[e] = state_update<SF, NP>(a1[e], f1, K.dt);
            }
        }
        A4[i * NT + tid] = make_float4(n0[0], n0[1], n0[2], n0[3]);
        A4[(8 + i) * NT + tid] = make_float4(n1[0], n1[1], n1[2], n1[3]);
    }
    part[PT_CNT_A * NT + tid] = A.cnt_a;
    part[PT_G00 * NT + tid] = A.sg0 + A.sg1;
    part[PT_CNT_G * NT + tid] = A.cnt_g;
    part[PT_CNT_P * NT + tid] = A.cnt_p;
    part[PT_MX_R * NT + tid] = xr0 * A.sa0 + xr1 * A.sa1;
    part[PT_MX_C * NT + tid] = A.mxc;
    part[PT_MX2_R * NT + tid] = (xr0 * xr0) * A.sa0 + (xr1 * xr1) * A.sa1;
    part[PT_MX2_C * NT + tid] = A.mx2c;
    part[PT_GX_R * NT + tid] = xr0 * A.sg0 + xr1 * A.sg1;
    part[PT_GX_C * NT + tid] = A.gxc;
    part[PT_M00_C0 * NT + tid] = A.sa0 + A.sa1;
}

// ---- cell phase of the TMEM kernel (lnx_world128_tm): packed over the thread's two rows ---------------------------------
// The state of the thread lives in a thread-private store of 8 chunks x 8 floats; chunk i holds columns 4(4i+e)+l, e = 0..3,
// as (row p, row p+64) pairs, i.e. exactly the (re, im) register pairs phase 1 consumes.  `Store` provides load(i, dst8) /
// wait_load(dst8) / store(i, src8) (TMEM on the device, a plain array in the emulator).  v[] holds the potential on entry
// and the NEW state on exit: the state is read once and written once per step and never re-loaded.
// All per-cell arithmetic runs on the (row p, row p+64) pairs with the packed FP32 instructions; the column coordinate of
// the rolled frame (utils.py:269-293 folded into statistics.py:28-33) and its square come from a 1 KB table rebuilt once
// per step by the warp that advances the shift carry, instead of three instructions per column in every thread.
LNX_HD int __float_as_int_hd(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_int(x);
#else
    int r;
    memcpy(&r, &x, 4);
    return r;
#endif
}
struct ArrayStore {  // host emulator / tests
    float* base;     // [64] of this thread
    LNX_HD void load(int i, float* d) const {
        for (int e = 0; e < 8; ++e) d[e] = base[8 * i + e];
    }
    LNX_HD void wait_load(float*) const {}
    LNX_HD void store(int i, const float* s) const {
        for (int e = 0; e < 8; ++e) base[8 * i + e] = s[e];
    }
};
// coordinate table: float4 xt[2][4][XT_STRIDE]: xt[0][l][i] = xc of columns 4(4i+e)+l, e = 0..3; xt[1] = their squares.
// XT_STRIDE = 9 float4 keeps the four l rows of a warp-wide 128-bit read in different banks.
constexpr int XT_STRIDE = 9;
constexpr int XT_F4 = 2 * 4 * XT_STRIDE;
LNX_HD float xt_coord(int l, int j, int shift1) { return (float)(((4 * j + l - shift1) & (WS - 1)) - WS / 2); }
// entries 4*idx .. 4*idx+3 of the table (idx = 0..31 <-> l = idx >> 3, i = idx & 7): one lane of the building warp
LNX_HD void xt_build(int idx, int shift1, float4* xt) {
    const int l = idx >> 3, i = idx & 7;
    const float x0 = xt_coord(l, 4 * i, shift1), x1 = xt_coord(l, 4 * i + 1, shift1), x2 = xt_coord(l, 4 * i + 2, shift1),
                x3 = xt_coord(l, 4 * i + 3, shift1);
    xt[l * XT_STRIDE + i] = make_float4(x0, x1, x2, x3);
    xt[(4 + l) * XT_STRIDE + i] = make_float4(x0 * x0, x1 * x1, x2 * x2, x3 * x3);
}
// Threshold counts on the integer/ALU pipe (the FP32 pipe is the busy one): x > thr gives the bit pattern of 1.0f
// (FSET.BF), and bit patterns are summed three at a time by IADD3.  1.0f = 127 * 2^23, so after n <= 511 hits the sum is
// ((127 n) mod 512) * 2^23 (mod 2^32) and n = (383 * (sum >> 23)) mod 512 because 127 * 383 = 1 (mod 512).
LNX_HD int gt_bits(float x, float thr) { return __float_as_int_hd(x > thr ? 1.0f : 0.0f); }
LNX_HD float count_from_bits(int sum) { return (float)((383 * (int)((unsigned)sum >> 23)) & 511); }
// c * growth(X) on a pair; poly_quad4 folds the affine tail and the weight into one FMA (as field_fused)
template <int GF, bool NP>
LNX_HD float2 field_fused_pk(float2 X, const FusedConsts& K) {
    if constexpr (GF == GF_POLY_QUAD4) {
        const float2 t = pk_add(X, pk_bc(-K.gf.m));
        float2 o = pk_fma(pk_mul(t, t), pk_bc(-K.gf.k0), pk_bc(1.0f));
        if constexpr (NP)
            o = make_float2((o.x < 0.f) ? 0.f : o.x, (o.y < 0.f) ? 0.f : o.y);
        else
            o = make_float2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f));
        const float2 o2 = pk_mul(o, o);
        return pk_fma(pk_mul(o2, o2), pk_bc(K.c2), pk_bc(-K.c));
    } else {
        return make_float2(K.c * growth<GF, NP, true>(X.x, K.gf), K.c * growth<GF, NP, true>(X.y, K.gf));  // true divisions, as the generic kernel
    }
}
template <int GF, int SF, bool NP, class Store>
LNX_HD void cells_fused_rs(int tid, float2* v /* [32] */, const Store& st, const FusedConsts& K, int shift0, const float4* xt,
                           float* part /* [NPART][256] */) {
    const int l = t_sub(tid) & 3;
    const float xr0 = rolled_coord(cell_row(tid, 0), shift0), xr1 = rolled_coord(cell_row(tid, 1), shift0);
    float2 sa = make_float2(0.f, 0.f), sg = sa, mx = sa, mx2 = sa, gx = sa;  // per-row sums: cells, positive field, moments
    int cnt_a = 0, cnt_g = 0, cnt_p = 0;  // at most 64 hits each per step
    float buf[2][8];
    st.load(0, buf[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float* a = buf[i & 1];
        st.wait_load(a);
        if (i + 1 < 8) st.load(i + 1, buf[(i + 1) & 1]);  // in flight while chunk i is processed
        const float4 x4 = xt[l * XT_STRIDE + i], q4 = xt[(4 + l) * XT_STRIDE + i];
        const float xc[4] = {x4.x, x4.y, x4.z, x4.w}, xc2[4] = {q4.x, q4.y, q4.z, q4.w};
        float n[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * i + e;
            const float2 A = make_float2(a[2 * e], a[2 * e + 1]), P = v[j];
            cnt_p += gt_bits(P.x, EPS) + gt_bits(P.y, EPS);  // statistics.py:70
            const float2 F = field_fused_pk<GF, NP>(P, K);
            sa = pk_add(sa, A);
            mx = pk_fma(A, pk_bc(xc[e]), mx);
            mx2 = pk_fma(A, pk_bc(xc2[e]), mx2);
            cnt_a += gt_bits(A.x, EPS) + gt_bits(A.y, EPS);
            const float2 G = make_float2(fmaxf(F.x, 0.f), fmaxf(F.y, 0.f));  // statistics.py:65
            sg = pk_add(sg, G);
            gx = pk_fma(G, pk_bc(xc[e]), gx);
            cnt_g += gt_bits(F.x, EPS) + gt_bits(F.y, EPS);  // max(f, 0) > eps <=> f > eps
            float2 N;
            if constexpr (SF == SF_V1 && !NP) {  // clip(a + dt f, 0, 1) as one saturating FMA (no NaN possible here)
                N = make_float2(saturate01(A.x + K.dt * F.x), saturate01(A.y + K.dt * F.y));
            } else {
                N = make_float2(state_update<SF, NP>(A.x, F.x, K.dt), state_update<SF, NP>(A.y, F.y, K.dt));
            }
            v[j] = N;
            n[2 * e] = N.x;
            n[2 * e + 1] = N.y;
        }
        st.store(i, n);
    }
    part[PT_CNT_A * NT + tid] = count_from_bits(cnt_a);
    part[PT_G00 * NT + tid] = sg.x + sg.y;
    part[PT_CNT_G * NT + tid] = count_from_bits(cnt_g);
    part[PT_CNT_P * NT + tid] = count_from_bits(cnt_p);
    part[PT_MX_R * NT + tid] = xr0 * sa.x + xr1 * sa.y;
    part[PT_MX_C * NT + tid] = mx.x + mx.y;
    part[PT_MX2_R * NT + tid] = (xr0 * xr0) * sa.x + (xr1 * xr1) * sa.y;
    part[PT_MX2_C * NT + tid] = mx2.x + mx2.y;
    part[PT_GX_R * NT + tid] = xr0 * sg.x + xr1 * sg.y;
    part[PT_GX_C * NT + tid] = gx.x + gx.y;
    part[PT_M00_C0 * NT + tid] = sa.x + sa.y;
}

// ---- per-world statistics carry + stop criteria (owned by lane 0 of the statistics warp) ----
struct StatsCarry {
    int shift[2];       // total_shift_idx          (runner.py:285-289)
    float centroid[2];  // mass_centroid carry
    float angle;        // mass_angle carry
    // check_heuristics carry (statistics.py:186-194)
    float should_continue, prev_mass, prev_sign, n_alive;
    float init_cm[MAX_C];
    int mono, vol;
    LNX_HD void reset() {
        shift[0] = shift[1] = 0;
        centroid[0] = centroid[1] = 0.f;
        angle = 0.f;
        should_continue = 1.f;
        prev_mass = prev_sign = n_alive = 0.f;
        mono = vol = 0;
#pragma unroll
        for (int c = 0; c < MAX_C; ++c) init_cm[c] = 0.f;
    }
};

// x mod 360 with the sign of the divisor (jnp `%`), for |x| of a few thousand at most (angles): one floor + FMA
LNX_HD float mod360(float x) {
    float r = x - 360.f * floorf(x * (1.0f / 360.f));
    if (r < 0.f) r += 360.f;
    if (r >= 360.f) r -= 360.f;
    return r;
}
// division used by the per-step statistics (once per world-step, on the statistics warp): approximate reciprocal based
// on the device (2 ulp) to keep the warp's code small — the statistics code is fetched every step by every SM
LNX_HD float sdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
LNX_HD int py_mod(int x, int m) {
    int r = x % m;
    return r < 0 ? r + m : r;
}
LNX_HD int trunc_to_int(float x) { return (x == x && fabsf(x) < 2.0e9f) ? (int)x : 0; }

// totals[PT_*] are the CTA-wide sums.  Writes the 11 scalar stats + C channel masses, updates the carry and the
// stop criteria.  Returns should_continue (0/1) for this step.  invR2 = 1/R^2, invR = 1/R, inv_dt = 1/dt (per plan).
LNX_HD float stats_finalize(const float* totals, int C, int t, float invR2, float invR, float inv_dt, StatsCarry& S,
                            float* out /* [ST_COUNT + C]: scalar stats then channel masses */) {
    float* cm_out = out + ST_COUNT;
    float m00 = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        m00 += totals[PT_M00_C0 + c];
        cm_out[c] = totals[PT_M00_C0 + c] * invR2;
    }
    const float g00 = totals[PT_G00];
    const float mass = m00 * invR2;
    const float mass_volume = totals[PT_CNT_A] * invR2;
    const float growth = g00 * invR2;
    const float growth_volume = totals[PT_CNT_G] * invR2;
    out[ST_MASS] = mass;
    out[ST_MASS_VOLUME] = mass_volume;
    out[ST_MASS_DENSITY] = sdiv(mass, mass_volume + EPS);
    out[ST_GROWTH] = growth;
    out[ST_GROWTH_VOLUME] = growth_volume;
    out[ST_GROWTH_DENSITY] = sdiv(growth, growth_volume + EPS);
    out[ST_POTENTIAL_VOLUME] = totals[PT_CNT_P] * invR2;

    const float im = sdiv(1.0f, m00 + EPS), ig = sdiv(1.0f, g00 + EPS);
    const float c0 = totals[PT_MX_R] * im, c1 = totals[PT_MX_C] * im;
    const float d0 = c0 - S.centroid[0], d1 = c1 - S.centroid[1];
    const float dist = sqrtf(d0 * d0 + d1 * d1);
    out[ST_MASS_SPEED] = dist * invR * inv_dt;
    const float angle = (atan2f(d1, d0) * 57.29577951308232f) * ((dist * invR > 0.001f) ? 1.f : 0.f);
    out[ST_MASS_ANGLE_SPEED] = (mod360(angle - S.angle + 540.f) - 180.f) * inv_dt;
    const float e0 = totals[PT_GX_R] * ig - c0, e1 = totals[PT_GX_C] * ig - c1;
    out[ST_MASS_GROWTH_DIST] = sqrtf(e0 * e0 + e1 * e1) * invR;
    const float iden = sdiv(1.0f, m00 * m00 + EPS);
    out[ST_INERTIA] = (totals[PT_MX2_R] - c0 * totals[PT_MX_R]) * iden + (totals[PT_MX2_C] - c1 * totals[PT_MX_C]) * iden;

    // carry (statistics.py:117-124)
    const int s0 = trunc_to_int(c0), s1 = trunc_to_int(c1);
    S.shift[0] = (S.shift[0] + s0) & (WS - 1);  // Python-sign modulo for a power-of-two world size
    S.shift[1] = (S.shift[1] + s1) & (WS - 1);
    S.centroid[0] = c0 - (float)s0;
    S.centroid[1] = c1 - (float)s1;
    S.angle = angle;

    // check_heuristics step t (statistics.py:144-183)
    if (t == 0) {
        S.prev_mass = mass;
        S.prev_sign = 0.f;
    }
    bool cond = true;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        if (t == 0) S.init_cm[c] = cm_out[c];
        cond = cond && (cm_out[c] >= EPS) && (cm_out[c] <= 3.f * S.init_cm[c]);
    }
    const float dm = mass - S.prev_mass;
    const float sign = (dm > 0.f) ? 1.f : ((dm < 0.f) ? -1.f : dm);  // jnp.sign: 0 -> 0, NaN -> NaN
    S.mono = S.mono * (sign == S.prev_sign ? 1 : 0) + 1;
    cond = cond && (S.mono <= 128);
    S.vol = S.vol * (mass_volume > 10.f ? 1 : 0) + 1;
    cond = cond && (S.vol <= 128);
    S.should_continue *= cond ? 1.f : 0.f;
    S.n_alive += S.should_continue;
    S.prev_mass = mass;
    S.prev_sign = sign;
    return S.should_continue;
}

}  // namespace lnx
