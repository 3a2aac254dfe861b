// CPU emulator of the 64^3 thread-per-line engine (TEST ONLY — not part of the product library).
// Runs the exact __host__ __device__ per-lane phase functions of leniax_b200/csrc/lnx_tiled64.cuh lane by lane, with the
// kernels' __syncwarp points as loop boundaries.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../leniax_b200/csrc/lnx_tiled64.cuh"
#include "../../leniax_b200/csrc/lnx_tiled64h.cuh"

using namespace lnx;
using namespace lnx::t64;

// plane_fwd_kernel for one plane
static void emul_plane_fwd(const float* src, float2* dst) {
    std::vector<float> sm(SMEM_FLOATS, 0.f);
    std::vector<float2> regs(32 * 64);
    for (int lane = 0; lane < 32; ++lane) fwd_load(lane, src, sm.data());
    for (int lane = 0; lane < 32; ++lane) fwd_rows(lane, sm.data(), regs.data() + lane * 64);
    for (int lane = 0; lane < 32; ++lane) fwd_rows_store(lane, reinterpret_cast<float2*>(sm.data()), regs.data() + lane * 64);
    for (int lane = 0; lane < 32; ++lane) fwd_cols(lane, reinterpret_cast<const float2*>(sm.data()), dst);
}

extern "C" {

// world [64][64][64] -> natural-order half spectrum [64][64][33] (plane_fwd + lead in forward-only mode)
void lnx_t64_emul_rfftn(const float* world, float2* spec) {
    std::vector<float2> tmp((size_t)N * PLANE_SPEC);
    for (int l = 0; l < N; ++l) emul_plane_fwd(world + (size_t)l * PLANE_CELLS, tmp.data() + (size_t)l * PLANE_SPEC);
    for (int col = 0; col < COLS; ++col) {
        float2 v[64];
        lead_load_fwd(tmp.data() + col, v);
        lead_store_fwd(spec + col, v);
    }
}

// one Lenia step of one world, one channel / one kernel.  ktab: [64][64][33] complex, pre-scaled by 1 / 64^3.
// state is updated in place; potential / field: [64^3]; partials: [64 planes][NP_T] (lane-summed)
void lnx_t64_emul_step(float* state, const float2* ktab, int gf_id, float m, float s, float wk, int mean, int state_fn, float dt,
                       const int* shift, float* potential, float* field, float* partials, float2* next_spec) {
    std::vector<float2> spec((size_t)N * PLANE_SPEC), pot((size_t)N * PLANE_SPEC);
    for (int l = 0; l < N; ++l) emul_plane_fwd(state + (size_t)l * PLANE_CELLS, spec.data() + (size_t)l * PLANE_SPEC);
    for (int col = 0; col < COLS; ++col) {
        float2 v[64];
        lead_load_fwd(spec.data() + col, v);
        lead_mul_inv_store(ktab + col, pot.data() + col, v);
    }
    for (int l = 0; l < N; ++l) {
        std::vector<float> sm(SMEM_FLOATS, 0.f);
        std::vector<float2> regs(32 * 64);
        float2* pl = reinterpret_cast<float2*>(sm.data());
        for (int lane = 0; lane < 32; ++lane) inv_load(lane, pot.data() + (size_t)l * PLANE_SPEC, pl);
        for (int lane = 0; lane < 32; ++lane) inv_cols(lane, pl);
        for (int lane = 0; lane < 32; ++lane) inv_rows_load(lane, pl, regs.data() + lane * 64);
        CellParams cp;
        cp.gf_id = gf_id;
        cp.state_fn = state_fn;
        cp.mean = mean;
        cp.gc = gf_prepare(gf_id, m, s);
        cp.wk = wk;
        cp.wsum = wk;
        cp.dt = dt;
        cp.sh0 = shift[0];
        cp.sh1 = shift[1];
        cp.sh2 = shift[2];
        cp.l = l;
        for (int lane = 0; lane < 32; ++lane) {
            inv_rows(regs.data() + lane * 64);
            inv_pot_store(lane, regs.data() + lane * 64, sm.data(), potential + (size_t)l * PLANE_CELLS);
        }
        float tot[NP_T];
        for (int i = 0; i < NP_T; ++i) tot[i] = 0.f;
        for (int lane = 0; lane < 32; ++lane) {
            float acc[NP_T];
            inv_update_dispatch(lane, sm.data(), state + (size_t)l * PLANE_CELLS, nullptr, field + (size_t)l * PLANE_CELLS, cp, acc,
                                next_spec != nullptr);
            for (int i = 0; i < NP_T; ++i) tot[i] += acc[i];
        }
        if (next_spec) {  // fused tail of plane_inv_kernel: forward planes of the next step from the cells left in shared memory
            for (int lane = 0; lane < 32; ++lane) fwd_rows(lane, sm.data(), regs.data() + lane * 64);
            for (int lane = 0; lane < 32; ++lane) fwd_rows_store(lane, pl, regs.data() + lane * 64);
            for (int lane = 0; lane < 32; ++lane) fwd_cols(lane, pl, next_spec + (size_t)l * PLANE_SPEC);
        }
        for (int i = 0; i < NP_T; ++i) partials[l * NP_T + i] = tot[i];
    }
}

int lnx_t64_emul_np() { return NP_T; }

// the same step through the half-line kernels of lnx_tiled64h.cuh (lead_h_kernel + plane_step_kernel), thread by thread with the
// kernels' __syncthreads / shuffle points as loop boundaries.  spec: forward planes of `state` ([64][64][33], e.g. from
// lnx_t64_emul_rfftn's first stage = t64::plane_fwd); everything else as lnx_t64_emul_step.
void lnx_t64h_emul_step(float* state, const float2* ktab, int gf_id, float m, float s, float wk, int mean, int state_fn, float dt,
                        const int* shift, float* potential, float* field, float* partials, float2* next_spec, int finite) {
    namespace H = lnx::t64h;
    const int mode = H::select_mode(gf_id, state_fn, finite != 0);
    std::vector<float2> spec((size_t)N * PLANE_SPEC), pot((size_t)N * PLANE_SPEC);
    for (int l = 0; l < N; ++l) emul_plane_fwd(state + (size_t)l * PLANE_CELLS, spec.data() + (size_t)l * PLANE_SPEC);
    for (int col = 0; col < COLS; ++col) {  // lead_h_kernel: the two threads of a column, then the shuffle exchange
        float2 u[2][32];
        for (int h = 0; h < 2; ++h) {
            H::lead_fwd(h, spec.data() + col, u[h]);
            H::lead_mul_inv(h, ktab + col, u[h]);
        }
        for (int h = 0; h < 2; ++h)
            for (int n = 0; n < 32; ++n) pot[(size_t)(n + 32 * h) * COLS + col] = H::lead_combine(h, u[h][n], u[1 - h][n]);
    }
    for (int l = 0; l < N; ++l) {  // plane_step_kernel
        std::vector<float> sm(SMEM_FLOATS, 0.f);
        std::vector<float2> regs(64 * 32), zs(64);
        float2* pl = reinterpret_cast<float2*>(sm.data());
        for (int t = 0; t < 64; ++t) H::inv_load(t, pot.data() + (size_t)l * PLANE_SPEC, pl);
        for (int t = 0; t < 64; ++t) H::inv_pack(t, pl);
        for (int t = 0; t < 64; ++t) H::inv_col(t & 31, t >> 5, pl, regs.data() + t * 32);
        for (int t = 0; t < 64; ++t) H::inv_col_store(t & 31, t >> 5, pl, regs.data() + t * 32);
        for (int t = 0; t < 64; ++t) {
            H::inv_row_load(t, pl, regs.data() + t * 32);
            ifft_dit<32>(regs.data() + t * 32);
        }
        for (int t = 0; t < 64; ++t) H::inv_pot_store(t, regs.data() + t * 32, sm.data());
        CellParams cp;
        cp.gf_id = gf_id;
        cp.state_fn = state_fn;
        cp.mean = mean;
        cp.gc = gf_prepare(gf_id, m, s);
        cp.wk = wk;
        cp.wsum = wk;
        cp.dt = dt;
        cp.sh0 = shift[0];
        cp.sh1 = shift[1];
        cp.sh2 = shift[2];
        cp.l = l;
        float tot[NP_T];
        for (int i = 0; i < NP_T; ++i) tot[i] = 0.f;
        for (int t = 0; t < 64; ++t) {
            float acc[NP_T];
            H::update_dispatch(mode, t, sm.data(), state + (size_t)l * PLANE_CELLS, nullptr, field + (size_t)l * PLANE_CELLS,
                               potential + (size_t)l * PLANE_CELLS, cp, acc, next_spec != nullptr);
            for (int i = 0; i < NP_T; ++i) tot[i] += acc[i];
        }
        if (next_spec) {
            float2* dst = next_spec + (size_t)l * PLANE_SPEC;
            for (int t = 0; t < 64; ++t) H::fwd_row_load(t, sm.data(), regs.data() + t * 32);
            for (int t = 0; t < 64; ++t) H::fwd_row_store(t, pl, regs.data() + t * 32);
            for (int t = 0; t < 64; ++t) {
                H::fwd_col(t & 31, t >> 5, pl, regs.data() + t * 32);
                H::fwd_col_store(t & 31, t >> 5, regs.data() + t * 32, dst, zs.data());
            }
            for (int t = 0; t < 64; ++t) H::fwd_packed_store(t, zs.data(), dst);
        }
        for (int i = 0; i < NP_T; ++i) partials[l * NP_T + i] = tot[i];
    }
}

}  // extern "C"
