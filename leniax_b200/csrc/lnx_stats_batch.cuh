// Lane-parallel statistics finaliser of the TMEM kernel: one warp turns up to 32 consecutive world-steps of CTA-wide
// totals into statistics rows at once (lane i <-> step tb + i).
//
// Same formulas as stats_finalize (lnx_step.cuh), i.e. leniax/statistics.py:65-124 (compute_stats) and :144-183
// (check_heuristics step) with the two counters of :287-333.  What is sequential in the reference's lax.scan becomes
//   * a neighbour exchange (previous centroid carry, previous angle, previous mass / sign): __shfl_up, lane 0 from the carry;
//   * the two "counter = counter * keep + 1" recurrences: distance to the most recent reset, found with ballot + clz;
//   * should_continue (a running product): first failing lane of a ballot.
// The expensive scalar chain (2 divisions, 2 square roots, atan2 per step) thus costs one warp ~300 instructions per 32
// steps instead of a dedicated warp per CTA, which is what frees the registers for two CTAs per SM.
#pragma once
#include "lnx_step.cuh"

namespace lnx {

constexpr int RING_ROWS = 32;
// ring row: the ten channel-independent totals (PT_CNT_A .. PT_GX_C), c0, c1 (mass centroid), then one mass sum per channel
constexpr int RING_C0 = PT_FIXED, RING_C1 = PT_FIXED + 1, RING_M00 = PT_FIXED + 2;
constexpr int RING_STRIDE_1 = 16;   // floats per row, single-channel kernel
constexpr int RING_STRIDE_C = 24;   // floats per row, up to MAX_C channels (12 + 8 = 20, padded to a multiple of 4)

struct BatchCarry {  // warp-uniform: every lane of the finalising warp holds the same copy
    float cc0, cc1;  // mass_centroid carry = c - trunc(c) of the last finalised step (statistics.py:124)
    float angle;     // mass_angle carry
    float prev_mass, prev_sign, should_continue, n_alive;
    float init_cm[MAX_C];
    int mono, vol;
    int rows;        // rows finalised so far (== index of the next step to finalise)
    __device__ __forceinline__ void reset() {
        cc0 = cc1 = angle = prev_mass = prev_sign = n_alive = 0.f;
#pragma unroll
        for (int c = 0; c < MAX_C; ++c) init_cm[c] = 0.f;
        should_continue = 1.f;
        mono = vol = rows = 0;
    }
};

// counter[i] of `c = c * keep + 1` over the lanes, given the ballot of lanes where keep == 0 and the carry-in counter
__device__ __forceinline__ int scan_counter(unsigned reset_mask, int lane, int carry) {
    const unsigned upto = reset_mask & (0xffffffffu >> (31 - lane));  // resets at lanes 0..lane
    return upto ? lane - (31 - __clz(upto)) + 1 : carry + lane + 1;
}

// ring: float [32][STRIDE] in shared memory, rows tb .. tb + n - 1 (tb is a multiple of 32, so ring row == lane).
// Writes the statistics rows to HBM and advances the carry.  C = number of channels (compile-time bound CMAX).
template <int CMAX, int STRIDE>
__device__ __forceinline__ void stats_finalize_batch(const float* ring, int n, int lane, int C, float* __restrict__ stats,
                                                     float* __restrict__ channel_mass, size_t plane, size_t idx0 /* index of step tb */,
                                                     size_t t_stride, float invR2, float invR, float inv_dt, BatchCarry& S) {
    const unsigned FULL = 0xffffffffu;
    const bool valid = lane < n;
    const unsigned vmask = n >= 32 ? FULL : ((1u << n) - 1u);
    const int tb = S.rows;
    float tot[STRIDE];
    {
        const float4* rp = reinterpret_cast<const float4*>(ring + lane * STRIDE);
#pragma unroll
        for (int q = 0; q < STRIDE / 4; ++q) {
            const float4 r = rp[q];
            tot[4 * q] = r.x;
            tot[4 * q + 1] = r.y;
            tot[4 * q + 2] = r.z;
            tot[4 * q + 3] = r.w;
        }
    }
    float m00 = 0.f, cm[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        cm[c] = 0.f;
        if (c < C) {
            m00 += tot[RING_M00 + c];
            cm[c] = tot[RING_M00 + c] * invR2;
        }
    }
    const float g00 = tot[PT_G00];
    const float mass = m00 * invR2;
    const float mass_volume = tot[PT_CNT_A] * invR2;
    const float growth = g00 * invR2;
    const float growth_volume = tot[PT_CNT_G] * invR2;
    const float ig = sdiv(1.0f, g00 + EPS);
    const float c0 = tot[RING_C0], c1 = tot[RING_C1];

    // neighbour exchange: carries of step t - 1
    float pc0 = __shfl_up_sync(FULL, c0, 1), pc1 = __shfl_up_sync(FULL, c1, 1);
    pc0 -= (float)trunc_to_int(pc0);
    pc1 -= (float)trunc_to_int(pc1);
    if (lane == 0) {
        pc0 = S.cc0;
        pc1 = S.cc1;
    }
    const float d0 = c0 - pc0, d1 = c1 - pc1;
    const float dist = sqrtf(d0 * d0 + d1 * d1);
    const float angle = (atan2f(d1, d0) * 57.29577951308232f) * ((dist * invR > 0.001f) ? 1.f : 0.f);
    float pangle = __shfl_up_sync(FULL, angle, 1);
    if (lane == 0) pangle = S.angle;
    const float e0 = tot[PT_GX_R] * ig - c0, e1 = tot[PT_GX_C] * ig - c1;
    const float iden = sdiv(1.0f, m00 * m00 + EPS);

    // check_heuristics (statistics.py:144-183)
    float pmass = __shfl_up_sync(FULL, mass, 1);
    if (lane == 0) pmass = (tb == 0) ? mass : S.prev_mass;
    const float dm = mass - pmass;
    const float sign = (dm > 0.f) ? 1.f : ((dm < 0.f) ? -1.f : dm);  // jnp.sign: 0 -> 0, NaN -> NaN
    float psign = __shfl_up_sync(FULL, sign, 1);
    if (lane == 0) psign = (tb == 0) ? 0.f : S.prev_sign;
    const int mono = scan_counter(__ballot_sync(FULL, !(sign == psign)) & vmask, lane, S.mono);
    const int vol = scan_counter(__ballot_sync(FULL, !(mass_volume > 10.f)) & vmask, lane, S.vol);
    bool cond = (mono <= 128) && (vol <= 128);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            const float first = __shfl_sync(FULL, cm[c], 0);
            if (tb == 0) S.init_cm[c] = first;
            cond = cond && (cm[c] >= EPS) && (cm[c] <= 3.f * S.init_cm[c]);
        }
    }
    const unsigned failed = __ballot_sync(FULL, !cond) & vmask;
    const int first_fail = failed ? __ffs(failed) - 1 : 32;
    const int alive_rows = S.should_continue != 0.f ? (first_fail < n ? first_fail : n) : 0;

    if (valid) {
        const size_t idx = idx0 + (size_t)lane * t_stride;
        stats[ST_MASS * plane + idx] = mass;
        stats[ST_MASS_VOLUME * plane + idx] = mass_volume;
        stats[ST_MASS_DENSITY * plane + idx] = sdiv(mass, mass_volume + EPS);
        stats[ST_GROWTH * plane + idx] = growth;
        stats[ST_GROWTH_VOLUME * plane + idx] = growth_volume;
        stats[ST_GROWTH_DENSITY * plane + idx] = sdiv(growth, growth_volume + EPS);
        stats[ST_MASS_SPEED * plane + idx] = dist * invR * inv_dt;
        stats[ST_MASS_ANGLE_SPEED * plane + idx] = (mod360(angle - pangle + 540.f) - 180.f) * inv_dt;
        stats[ST_MASS_GROWTH_DIST * plane + idx] = sqrtf(e0 * e0 + e1 * e1) * invR;
        stats[ST_INERTIA * plane + idx] = (tot[PT_MX2_R] - c0 * tot[PT_MX_R]) * iden + (tot[PT_MX2_C] - c1 * tot[PT_MX_C]) * iden;
        stats[ST_POTENTIAL_VOLUME * plane + idx] = tot[PT_CNT_P] * invR2;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) channel_mass[idx * C + c] = cm[c];
    }

    // carry out: values of the last valid lane
    const int last = n - 1;
    const float lc0 = __shfl_sync(FULL, c0, last), lc1 = __shfl_sync(FULL, c1, last);
    S.cc0 = lc0 - (float)trunc_to_int(lc0);
    S.cc1 = lc1 - (float)trunc_to_int(lc1);
    S.angle = __shfl_sync(FULL, angle, last);
    S.prev_mass = __shfl_sync(FULL, mass, last);
    S.prev_sign = __shfl_sync(FULL, sign, last);
    S.mono = __shfl_sync(FULL, mono, last);
    S.vol = __shfl_sync(FULL, vol, last);
    S.n_alive += (float)alive_rows;
    if (first_fail < n) S.should_continue = 0.f;
    S.rows = tb + n;
}

}  // namespace lnx
