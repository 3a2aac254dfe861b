// Growth functions, state update, per-cell statistics partials and the per-step statistics finaliser
// (host+device; the CPU emulator in tests/ runs the same code).
//
// Reference: leniax/growth_functions.py:6-253, leniax/core.py:202-319 (weighted mean/sum, get_state*),
// leniax/statistics.py:36-126 (compute_stats), :134-205 + :287-333 (check_heuristics), leniax/utils.py:269-293.
#pragma once
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
    float cnt_a, cnt_g, cnt_p;
    float mxc, mx2c, gxc;  // column-coordinate moments
    LNX_HD void clear() { sa0 = sa1 = sg0 = sg1 = cnt_a = cnt_g = cnt_p = mxc = mx2c = gxc = 0.f; }
};

// statistics contribution of one column j of the thread's two rows: a0/a1 = cells, f0/f1 = field
LNX_HD void acc_cells(CellAcc& A, float xc, float a0, float a1, float f0, float f1) {
    A.sa0 += a0;
    A.sa1 += a1;
    A.cnt_a += (a0 > EPS ? 1.f : 0.f) + (a1 > EPS ? 1.f : 0.f);
    const float g0 = fmaxf(f0, 0.f), g1 = fmaxf(f1, 0.f);  // statistics.py:65
    A.sg0 += g0;
    A.sg1 += g1;
    A.cnt_g += (g0 > EPS ? 1.f : 0.f) + (g1 > EPS ? 1.f : 0.f);
    const float as = a0 + a1, gs = g0 + g1;
    const float ax = as * xc;
    A.mxc += ax;
    A.mx2c += ax * xc;
    A.gxc += gs * xc;
}
LNX_HD float col_coord(float base /* (l - shift1) & 127 - 64 */, int j) {
    const float t = base + (float)(4 * j);
    return t >= 64.f ? t - 128.f : t;
}

// ---- fused single-channel single-kernel cell phase (the north-star fast path) ----
struct FusedConsts {
    GfConst gf;
    float w;         // kernels_weight_per_channel[0][0]
    float inv_wsum;  // 1 / sum_k W[0][k] (weighted_mean) or 1 (weighted_sum)
    float dt;        // 1 / T of this solution (runner.py:307)
};
template <int GF, int SF, bool NP>
LNX_HD void cells_fused(int tid, const float2* pot /* [32] */, float4* A4, const FusedConsts& K, int shift0, int shift1,
                        float* part /* [NPART][256] */) {
    const int l = t_sub(tid) & 3;
    const float xr0 = rolled_coord(cell_row(tid, 0), shift0), xr1 = rolled_coord(cell_row(tid, 1), shift0);
    const float cbase = (float)(((l - shift1) & (WS - 1)) - WS / 2);
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
            const float f0 = (K.w * growth<GF, NP>(p0, K.gf)) * K.inv_wsum;
            const float f1 = (K.w * growth<GF, NP>(p1, K.gf)) * K.inv_wsum;
            acc_cells(A, col_coord(cbase, j), a0[e], a1[e], f0, f1);
            n0[e] = state_update<SF, NP>(a0[e], f0, K.dt);
            n1[e] = state_update<SF, NP>(a1[e], f1, K.dt);
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

LNX_HD float py_fmod(float x, float m) {  // sign of the divisor, like jnp `%`
    float r = fmodf(x, m);
    if (r != 0.f && ((r < 0.f) != (m < 0.f))) r += m;
    return r;
}
LNX_HD int py_mod(int x, int m) {
    int r = x % m;
    return r < 0 ? r + m : r;
}
LNX_HD int trunc_to_int(float x) { return (x == x && fabsf(x) < 2.0e9f) ? (int)x : 0; }

// totals[PT_*] are the CTA-wide sums.  Writes the 11 scalar stats + C channel masses, updates the carry and the
// stop criteria.  Returns should_continue (0/1) for this step.
LNX_HD float stats_finalize(const float* totals, int C, int t, float R, float dt, StatsCarry& S, float* out /* [ST_COUNT] */,
                            float* cm_out /* [C] */) {
    const float R2 = R * R;
    float m00 = 0.f;
    for (int c = 0; c < C; ++c) m00 += totals[PT_M00_C0 + c];
    const float g00 = totals[PT_G00];
    const float mass = m00 / R2;
    const float mass_volume = totals[PT_CNT_A] / R2;
    const float growth = g00 / R2;
    const float growth_volume = totals[PT_CNT_G] / R2;
    out[ST_MASS] = mass;
    out[ST_MASS_VOLUME] = mass_volume;
    out[ST_MASS_DENSITY] = mass / (mass_volume + EPS);
    out[ST_GROWTH] = growth;
    out[ST_GROWTH_VOLUME] = growth_volume;
    out[ST_GROWTH_DENSITY] = growth / (growth_volume + EPS);
    out[ST_POTENTIAL_VOLUME] = totals[PT_CNT_P] / R2;
    for (int c = 0; c < C; ++c) cm_out[c] = totals[PT_M00_C0 + c] / R2;

    const float c0 = totals[PT_MX_R] / (m00 + EPS), c1 = totals[PT_MX_C] / (m00 + EPS);
    const float d0 = c0 - S.centroid[0], d1 = c1 - S.centroid[1];
    const float dist = sqrtf(d0 * d0 + d1 * d1);
    out[ST_MASS_SPEED] = dist / R / dt;
    const float angle = (atan2f(d1, d0) * 57.29577951308232f) * ((dist / R > 0.001f) ? 1.f : 0.f);
    out[ST_MASS_ANGLE_SPEED] = (py_fmod(angle - S.angle + 540.f, 360.f) - 180.f) / dt;
    const float gc0 = totals[PT_GX_R] / (g00 + EPS), gc1 = totals[PT_GX_C] / (g00 + EPS);
    const float e0 = gc0 - c0, e1 = gc1 - c1;
    out[ST_MASS_GROWTH_DIST] = sqrtf(e0 * e0 + e1 * e1) / R;
    const float den = m00 * m00 + EPS;
    out[ST_INERTIA] = (totals[PT_MX2_R] - c0 * totals[PT_MX_R]) / den + (totals[PT_MX2_C] - c1 * totals[PT_MX_C]) / den;

    // carry (statistics.py:117-124)
    const int s0 = trunc_to_int(c0), s1 = trunc_to_int(c1);
    S.shift[0] = py_mod(S.shift[0] + s0, WS);
    S.shift[1] = py_mod(S.shift[1] + s1, WS);
    S.centroid[0] = c0 - (float)s0;
    S.centroid[1] = c1 - (float)s1;
    S.angle = angle;

    // check_heuristics step t (statistics.py:144-183)
    if (t == 0) {
        for (int c = 0; c < C; ++c) S.init_cm[c] = cm_out[c];
        S.prev_mass = mass;
        S.prev_sign = 0.f;
    }
    bool cond = true;
    for (int c = 0; c < C; ++c) cond = cond && (cm_out[c] >= EPS) && (cm_out[c] <= 3.f * S.init_cm[c]);
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
