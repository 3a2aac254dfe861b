// Cell phase of the R16 variant (one real row per thread): growth, update, statistics partials.
// Same arithmetic as cells_fused in lnx_step.cuh (reference: leniax/growth_functions.py, core.py:245-319,
// statistics.py:36-126); only the thread-to-cell mapping differs.
#pragma once
#include "lnx_step.cuh"
#include "lnx_w128r.cuh"

namespace lnx {
namespace r16 {

// x[32]: potential of row cell_row(u) at columns 4j + l.  A4: float4 [8][512] thread-private state (element e of
// float4 i4 is column 4*(4*i4 + e) + l).  part: float [NPART][512].
template <int GF, int SF, bool NP>
LNX_HD void cells_fused(int u, const float* x, float4* A4, const FusedConsts& K, int shift0, int shift1, float* part) {
    const int l = t_l(u);
    const float xr = rolled_coord(cell_row(u), shift0);
    const float cbase = opaque((float)(((l - shift1) & (WS - 1)) - WS / 2));
    float sa = 0.f, sg = 0.f, cnt_a = 0.f, cnt_g = 0.f, cnt_p = 0.f, mxc = 0.f, mx2c = 0.f, gxc = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        const float4 c = A4[i4 * NT + u];
        const float a[4] = {c.x, c.y, c.z, c.w};
        float n[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * i4 + e;
            const float p = x[j];
            cnt_p += p > EPS ? 1.f : 0.f;  // statistics.py:70
            const float f = field_fused<GF, NP>(p, K);
            const float g = fmaxf(f, 0.f);
            const float xc = col_coord(cbase, j);
            sa += a[e];
            cnt_a += a[e] > EPS ? 1.f : 0.f;
            sg += g;
            cnt_g += f > EPS ? 1.f : 0.f;
            const float ax = a[e] * xc;
            mxc += ax;
            mx2c += ax * xc;
            gxc += g * xc;
            if constexpr (SF == SF_V1 && !NP)
                n[e] = saturate01(a[e] + K.dt * f);
            else
                n[e] = state_update<SF, NP>(a[e], f, K.dt);
        }
        A4[i4 * NT + u] = make_float4(n[0], n[1], n[2], n[3]);
    }
    part[PT_CNT_A * NT + u] = cnt_a;
    part[PT_G00 * NT + u] = sg;
    part[PT_CNT_G * NT + u] = cnt_g;
    part[PT_CNT_P * NT + u] = cnt_p;
    part[PT_MX_R * NT + u] = xr * sa;
    part[PT_MX_C * NT + u] = mxc;
    part[PT_MX2_R * NT + u] = (xr * xr) * sa;
    part[PT_MX2_C * NT + u] = mx2c;
    part[PT_GX_R * NT + u] = xr * sg;
    part[PT_GX_C * NT + u] = gxc;
    part[PT_M00_C0 * NT + u] = sa;
}

}  // namespace r16
}  // namespace lnx
