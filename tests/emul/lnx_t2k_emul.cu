// CPU emulator of the 2048^2 four-step engine (TEST ONLY — not part of the product library).
// Runs the exact __host__ __device__ per-lane phase functions of leniax_b200/csrc/lnx_tiled2k.cuh lane by lane, with the
// kernels' __syncwarp points as loop boundaries.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../leniax_b200/csrc/lnx_tiled2k.cuh"

using namespace lnx;

namespace e2k {
using namespace lnx::t2k;

static std::vector<float2> make_tab() {
    std::vector<float2> tab(N);
    for (int i = 0; i < N; ++i) {
        const double a = 2.0 * 3.14159265358979323846 * i / N;
        tab[i] = make_float2((float)cos(a), (float)sin(a));
    }
    return tab;
}
// rows_fwd_kernel for CTA c (row pairs 8 c .. 8 c + 7)
static void rows_fwd(const float* state, float2* T, int c, const float2* tab) {
    std::vector<float2> nat_all(ROWS_WARPS * NATS), regs(32 * 64);
    for (int wid = 0; wid < ROWS_WARPS; ++wid) {
        const int p = c * ROWS_WARPS + wid;
        float2* sm = nat_all.data() + wid * NATS;
        for (int lane = 0; lane < 32; ++lane) {
            float2* v = regs.data() + lane * 64;
            rf_load(lane, state + (size_t)(2 * p) * N, v);
            fs_fwd_a(lane, v, tab);
            fs_fwd_store(lane, v, sm);
        }
        for (int lane = 0; lane < 32; ++lane) fs_fwd_b(lane, sm, regs.data() + lane * 64);
        for (int lane = 0; lane < 32; ++lane) rf_nat_store(lane, regs.data() + lane * 64, sm);
    }
    for (int tid = 0; tid < 32 * ROWS_WARPS; ++tid) rf8_untangle_store(tid, nat_all.data(), T + 2 * ROWS_WARPS * c);
}
// rows1_fwd_kernel (one real row per warp) for CTA c (rows 16 c .. 16 c + 15)
static void rows1_fwd(const float* state, float2* T, int c, const float2* tab) {
    std::vector<float2> nat_all(RR_WARPS * NATS1), regs(32 * 32);
    for (int wid = 0; wid < RR_WARPS; ++wid) {
        const int row = c * RR_WARPS + wid;
        float2* sm = nat_all.data() + wid * NATS1;
        for (int lane = 0; lane < 32; ++lane) {
            float2* v = regs.data() + lane * 32;
            rr_load(lane, state + (size_t)row * N, v);
            f1_a(lane, v, tab);
            f1_store(lane, v, sm);
        }
        for (int lane = 0; lane < 32; ++lane) f1_b(lane, sm, regs.data() + lane * 32);
        for (int lane = 0; lane < 32; ++lane) f1_nat_store(lane, regs.data() + lane * 32, sm);
    }
    for (int tid = 0; tid < 32 * RR_WARPS; ++tid) rf16_untangle_store(tid, nat_all.data(), tab, T + RR_WARPS * c);
}
// rows1_inv_kernel for CTA c; partials: one row of sums per ROW here (the kernel adds the sixteen rows of the CTA)
static void rows1_inv(float* state, const float2* Pm, int c, const float2* tab, const CellParams2& cp0, int mode, float* potential, float* field,
                      float* partials, float2* next_T) {
    std::vector<float2> nat_all(RR_WARPS * NATS1), regs(32 * 32);
    for (int tid = 0; tid < 32 * RR_WARPS; ++tid) ri16_gather(tid, Pm + RR_WARPS * c, nat_all.data());
    for (int tid = 0; tid < 32 * RR_WARPS; ++tid) ri16_tangle(tid, nat_all.data(), tab);
    for (int wid = 0; wid < RR_WARPS; ++wid) {
        const int row = c * RR_WARPS + wid;
        float2* smp = nat_all.data() + wid * NATS1;
        for (int lane = 0; lane < 32; ++lane) i1_nat_load(lane, smp, regs.data() + lane * 32);
        for (int lane = 0; lane < 32; ++lane) i1_store(lane, regs.data() + lane * 32, smp);
        for (int lane = 0; lane < 32; ++lane) i1_b(lane, smp, regs.data() + lane * 32, tab);
        for (int lane = 0; lane < 32; ++lane) rr_pot_store(lane, regs.data() + lane * 32, reinterpret_cast<float*>(smp));
        CellParams2 cp = cp0;
        cp.row0 = row;
        float tot[NP_T];
        for (int i = 0; i < NP_T; ++i) tot[i] = 0.f;
        const size_t off = (size_t)row * N;
        for (int lane = 0; lane < 32; ++lane) {
            float acc[NP_T];
            ri_update_by_mode<1>(mode, lane, reinterpret_cast<float*>(smp), state + off, nullptr, field + off, potential + off, cp, acc, next_T != nullptr);
            for (int i = 0; i < NP_T; ++i) tot[i] += acc[i];
        }
        for (int i = 0; i < NP_T; ++i) partials[row * NP_T + i] = tot[i];
        if (next_T) {
            for (int lane = 0; lane < 32; ++lane) {
                float2* v = regs.data() + lane * 32;
                rr_load_smem(lane, reinterpret_cast<const float*>(smp), v);
                f1_a(lane, v, tab);
            }
            for (int lane = 0; lane < 32; ++lane) f1_store(lane, regs.data() + lane * 32, smp);
            for (int lane = 0; lane < 32; ++lane) f1_b(lane, smp, regs.data() + lane * 32);
            for (int lane = 0; lane < 32; ++lane) f1_nat_store(lane, regs.data() + lane * 32, smp);
        }
    }
    if (next_T)
        for (int tid = 0; tid < 32 * RR_WARPS; ++tid) rf16_untangle_store(tid, nat_all.data(), tab, next_T + RR_WARPS * c);
}
// lead_kernel for column k; kt == nullptr: forward only, natural-order result to out[m * HALF + k]
static void lead(const float2* T, const float2* kt, float2* P, float2* fwd_out, int k, const float2* tab) {
    std::vector<float2> sm(SMEM_C2), regs(32 * 64);
    for (int lane = 0; lane < 32; ++lane) {
        float2* v = regs.data() + lane * 64;
        ld_load(lane, T + (size_t)k * N, v);
        fs_fwd_a(lane, v, tab);
        fs_fwd_store(lane, v, sm.data());
    }
    for (int lane = 0; lane < 32; ++lane) fs_fwd_b(lane, sm.data(), regs.data() + lane * 64);
    if (!kt) {
        for (int lane = 0; lane < 32; ++lane)
            for (int q = 0; q < 64; ++q) fwd_out[(size_t)freq_of(q, lane) * HALF + k] = regs[lane * 64 + q];
        return;
    }
    for (int lane = 0; lane < 32; ++lane) {
        float2* v = regs.data() + lane * 64;
        ld_mul(lane, v, kt + (size_t)k * N);
        fs_inv_a(v);
    }
    for (int lane = 0; lane < 32; ++lane) fs_inv_store(lane, regs.data() + lane * 64, sm.data());
    for (int lane = 0; lane < 32; ++lane) {
        float2* v = regs.data() + lane * 64;
        fs_inv_b(lane, sm.data(), v, tab);
        ld_store(lane, P + (size_t)k * N, v);
    }
}
}  // namespace e2k

extern "C" {

// world [2048][2048] -> natural-order half spectrum [2048][1025]
void lnx_t2k_emul_rfft2(const float* world, float2* spec, int real_rows) {
    using namespace e2k;
    const std::vector<float2> tab = make_tab();
    std::vector<float2> T(SPEC);
    if (real_rows)
        for (int c = 0; c < N / RR_WARPS; ++c) rows1_fwd(world, T.data(), c, tab.data());
    else
        for (int c = 0; c < N / 2 / ROWS_WARPS; ++c) rows_fwd(world, T.data(), c, tab.data());
    for (int k = 0; k < HALF; ++k) lead(T.data(), nullptr, nullptr, spec, k, tab.data());
}

// one Lenia step of one 2048^2 world, one channel / one kernel.  K_half: [2048][1025] complex (natural order, unscaled).
void lnx_t2k_emul_step(float* state, const float2* K_half, int gf_id, float m, float s, float wk, int mean, int state_fn, float dt,
                       const int* shift, float* potential, float* field, float* partials, float2* next_T, int finite, int real_rows) {
    using namespace e2k;
    // finite < 0: the per-cell selection form (MODE_DYN), else what the library picks for this plan
    const int mode = finite < 0 ? (int)lnx::t64h::MODE_DYN : lnx::t64h::select_mode(gf_id, state_fn, finite != 0);
    const std::vector<float2> tab = make_tab();
    std::vector<float2> T(SPEC), Pm(SPEC), kt(SPEC);
    const float scale = 1.0f / ((float)N * (float)N);
    for (size_t i = 0; i < SPEC; ++i) {  // gather_ktab_kernel
        const int k = (int)(i >> 11), r = (int)(i & (N - 1));
        const float2 x = K_half[(size_t)freq_of(r >> 5, r & 31) * HALF + k];
        kt[i] = make_float2(x.x * scale, x.y * scale);
    }
    if (real_rows)
        for (int c = 0; c < N / RR_WARPS; ++c) rows1_fwd(state, T.data(), c, tab.data());
    else
        for (int c = 0; c < N / 2 / ROWS_WARPS; ++c) rows_fwd(state, T.data(), c, tab.data());
    for (int k = 0; k < HALF; ++k) lead(T.data(), kt.data(), Pm.data(), nullptr, k, tab.data());
    if (real_rows) {  // partials: [2048 rows][NP_T] in this mode
        CellParams2 cp;
        cp.gf_id = gf_id;
        cp.state_fn = state_fn;
        cp.mean = mean;
        cp.gc = gf_prepare(gf_id, m, s);
        cp.wk = wk;
        cp.wsum = wk;
        cp.dt = dt;
        cp.sh0 = shift[0];
        cp.sh1 = shift[1];
        cp.row0 = 0;
        for (int c = 0; c < N / RR_WARPS; ++c) rows1_inv(state, Pm.data(), c, tab.data(), cp, mode, potential, field, partials, next_T);
        return;
    }
    for (int c = 0; c < N / 2 / ROWS_WARPS; ++c) {  // rows_inv_kernel, CTA c
      std::vector<float2> nat_all(ROWS_WARPS * NATS);
      for (int tid = 0; tid < 32 * ROWS_WARPS; ++tid) ri8_gather(tid, Pm.data() + 2 * ROWS_WARPS * c, nat_all.data());
      for (int wid = 0; wid < ROWS_WARPS; ++wid) {
        const int p = c * ROWS_WARPS + wid;
        float2* smp = nat_all.data() + wid * NATS;
        std::vector<float2> regs(32 * 64);
        for (int lane = 0; lane < 32; ++lane) {
            ri_nat_load(lane, smp, regs.data() + lane * 64);
            fs_inv_a(regs.data() + lane * 64);
        }
        for (int lane = 0; lane < 32; ++lane) fs_inv_store(lane, regs.data() + lane * 64, smp);
        for (int lane = 0; lane < 32; ++lane) fs_inv_b(lane, smp, regs.data() + lane * 64, tab.data());
        for (int lane = 0; lane < 32; ++lane) ri_pot_store(lane, regs.data() + lane * 64, reinterpret_cast<float*>(smp));
        CellParams2 cp;
        cp.gf_id = gf_id;
        cp.state_fn = state_fn;
        cp.mean = mean;
        cp.gc = gf_prepare(gf_id, m, s);
        cp.wk = wk;
        cp.wsum = wk;
        cp.dt = dt;
        cp.sh0 = shift[0];
        cp.sh1 = shift[1];
        cp.row0 = 2 * p;
        float tot[NP_T];
        for (int i = 0; i < NP_T; ++i) tot[i] = 0.f;
        const size_t off = (size_t)(2 * p) * N;
        for (int lane = 0; lane < 32; ++lane) {
            float acc[NP_T];
            ri_update_by_mode(mode, lane, reinterpret_cast<float*>(smp), state + off, nullptr, field + off, potential + off, cp, acc, next_T != nullptr);
            for (int i = 0; i < NP_T; ++i) tot[i] += acc[i];
        }
        for (int i = 0; i < NP_T; ++i) partials[p * NP_T + i] = tot[i];
        if (next_T) {  // fused tail of rows_inv_kernel: forward rows of the next step from the cells left in shared memory
            for (int lane = 0; lane < 32; ++lane) {
                float2* v = regs.data() + lane * 64;
                rf_load_smem(lane, reinterpret_cast<const float*>(smp), v);
                fs_fwd_a(lane, v, tab.data());
            }
            for (int lane = 0; lane < 32; ++lane) fs_fwd_store(lane, regs.data() + lane * 64, smp);
            for (int lane = 0; lane < 32; ++lane) fs_fwd_b(lane, smp, regs.data() + lane * 64);
            for (int lane = 0; lane < 32; ++lane) rf_nat_store(lane, regs.data() + lane * 64, smp);
        }
      }
      if (next_T)
          for (int tid = 0; tid < 32 * ROWS_WARPS; ++tid) rf8_untangle_store(tid, nat_all.data(), next_T + 2 * ROWS_WARPS * c);
    }
}

int lnx_t2k_emul_np() { return lnx::t2k::NP_T; }

}  // extern "C"
