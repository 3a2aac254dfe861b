// CPU emulator of the resident-kernel phases (TEST ONLY — not part of the product library).
// Runs the exact __host__ __device__ per-thread code of leniax_b200/csrc/lnx_world128.cuh thread by thread, with the
// kernel's synchronisation points as loop boundaries, so index maps / hazards are validated without a GPU.
#include <cstring>
#include <vector>
#include <cmath>
#include "../../leniax_b200/csrc/lnx_step.cuh"

using namespace lnx;

static void make_tw(float2* tw) {
    for (int k = 0; k < 128; ++k) tw[k] = make_float2(Tw128::c[k], Tw128::s[k]);
}

// P3 with the warp-0 packed-column exchange; sub-phases split at the kernel's __syncwarp points
static void run_phase3(std::vector<Regs>& regs, std::vector<float2>& W, const float* Kt, const float* Kpq) {
    std::vector<float2> scratch(256);
    for (int t = 0; t < NT; ++t) phase3_load_fft(t, regs[t], W.data());
    for (int t = 0; t < 32; ++t) phase3_col0_stash(t, regs[t], scratch.data());
    for (int t = 0; t < 32; ++t) phase3_col0_compute(t, scratch.data(), reinterpret_cast<const float4*>(Kpq));
    for (int t = 0; t < NT; ++t) {
        phase3_multiply(t, regs[t], reinterpret_cast<const float4*>(Kt));
        if (t < 32) phase3_col0_fetch(t, regs[t], scratch.data());
        phase3_ifft_store(t, regs[t], W.data());
    }
}

extern "C" {

int lnx_emul_e1_addr(int q, int k1, int l) { return e1_addr(q, k1, l); }
int lnx_emul_e2_addr(int col, int u) { return e2_addr(col, u); }
int lnx_emul_k1_of(int a, int s) { return k1_of(a, s); }
int lnx_emul_col_of(int a, int c) { return col_of(a, c); }

// Kfull: interleaved complex [128][128] (m along rows, k along columns) = fftn(fftshift(kernel)).
void lnx_emul_build_kt(const float* Kfull, float* Kt /* float4[16][256] */, float* Kpq /* float4[32][4] */) {
    const float scale = 1.0f / (2.0f * 128.0f * 128.0f);
    for (int tid = 0; tid < NT; ++tid)
        for (int slot = 0; slot < 32; ++slot) {
            const int m = p3_slot_m(tid, slot), col = t_col(tid);
            float* dst = Kt + ((slot >> 1) * NT + tid) * 4 + (slot & 1) * 2;
            if (col == 0) {
                float* pq = Kpq + (slot * KPQ_LANES + tid) * 4;
                dst[0] = dst[1] = 0.f;
                const float* k0 = Kfull + (m * 128 + 0) * 2;
                const float* k64 = Kfull + (m * 128 + 64) * 2;
                pq[0] = (k0[0] + k64[0]) * 0.5f * scale;
                pq[1] = (k0[1] + k64[1]) * 0.5f * scale;
                pq[2] = (k0[0] - k64[0]) * 0.5f * scale;
                pq[3] = (k0[1] - k64[1]) * 0.5f * scale;
            } else {
                dst[0] = Kfull[(m * 128 + col) * 2] * scale;
                dst[1] = Kfull[(m * 128 + col) * 2 + 1] * scale;
            }
        }
}

// potential = real(ifft2(fft2(state) * K)) through the five phases.
void lnx_emul_potential(const float* state /* [128][128] */, const float* Kt, const float* Kpq, float* potential) {
    std::vector<Regs> regs(NT);
    std::vector<float2> W(W_COMPLEX);
    float2 tw[128];
    make_tw(tw);
    std::vector<float4> twtab(TW_TABLE_F4);
    for (int t = 0; t < NT; ++t) init_twiddle_table(t, twtab.data(), tw);
    for (int t = 0; t < NT; ++t) {  // P1
        for (int j = 0; j < 32; ++j)
            regs[t].v[j] = make_float2(state[cell_row(t, 0) * 128 + cell_col(t, j)], state[cell_row(t, 1) * 128 + cell_col(t, j)]);
        phase1(t, regs[t], W.data());
    }
    for (int t = 0; t < NT; ++t) phase2_load(t, regs[t], W.data());
    for (int t = 0; t < NT; ++t) phase2_compute_store(t, regs[t], W.data(), twtab.data());
    run_phase3(regs, W, Kt, Kpq);
    for (int t = 0; t < NT; ++t) phase4_load(t, regs[t], W.data());
    for (int t = 0; t < NT; ++t) phase4_compute_store(t, regs[t], W.data(), twtab.data());
    for (int t = 0; t < NT; ++t) {
        phase5_load(t, regs[t], W.data());
        phase5_ifft(regs[t]);
        for (int j = 0; j < 32; ++j) {
            potential[cell_row(t, 0) * 128 + cell_col(t, j)] = regs[t].v[j].x;
            potential[cell_row(t, 1) * 128 + cell_col(t, j)] = regs[t].v[j].y;
        }
    }
}


}  // extern "C"

// Full fused (1 channel, 1 kernel) run of `n_steps` steps with statistics, mirroring the CUDA kernel's phase order.
template <int GF, int SF, bool NP>
static void run_fused(const float* cells0, const float* Kt, const float* Kpq, float m, float s, float w, int mean, float T,
                      float R, float stats_dt, int n_steps, float* stats /*[ST_COUNT][n_steps]*/, float* cm /*[n_steps]*/,
                      float* N_out, float* final_cells, float* pot_out /* [n_steps][128][128] or null */) {
    std::vector<Regs> regs(NT);
    std::vector<float2> W(W_COMPLEX);
    std::vector<float4> A4(16 * NT);
    std::vector<float> part((PT_FIXED + 1) * NT);
    float2 tw[128];
    make_tw(tw);
    std::vector<float4> twtab(TW_TABLE_F4);
    for (int t = 0; t < NT; ++t) init_twiddle_table(t, twtab.data(), tw);
    for (int t = 0; t < NT; ++t)
        for (int i = 0; i < 16; ++i) {
            float e[4];
            for (int k = 0; k < 4; ++k) e[k] = cells0[cell_row(t, i >> 3) * 128 + cell_col(t, 4 * (i & 7) + k)];
            A4[i * NT + t] = make_float4(e[0], e[1], e[2], e[3]);
        }
    const FusedConsts K = fused_consts(GF, m, s, w, mean, 1.0f / T);
    StatsCarry S;
    S.reset();
    for (int step = 0; step < n_steps; ++step) {
        const int sh0 = S.shift[0], sh1 = S.shift[1];
        for (int t = 0; t < NT; ++t) {
            for (int i = 0; i < 8; ++i) {
                const float4 c0 = A4[i * NT + t], c1 = A4[(8 + i) * NT + t];
                regs[t].v[4 * i + 0] = make_float2(c0.x, c1.x);
                regs[t].v[4 * i + 1] = make_float2(c0.y, c1.y);
                regs[t].v[4 * i + 2] = make_float2(c0.z, c1.z);
                regs[t].v[4 * i + 3] = make_float2(c0.w, c1.w);
            }
            phase1(t, regs[t], W.data());
        }
        for (int t = 0; t < NT; ++t) phase2_load(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) phase2_compute_store(t, regs[t], W.data(), twtab.data());
        run_phase3(regs, W, Kt, Kpq);
        for (int t = 0; t < NT; ++t) phase4_load(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) phase4_compute_store(t, regs[t], W.data(), twtab.data());
        for (int t = 0; t < NT; ++t) phase5_load(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) {
            phase5_ifft(regs[t]);
            if (pot_out)
                for (int j = 0; j < 32; ++j) {
                    pot_out[(size_t)step * 16384 + cell_row(t, 0) * 128 + cell_col(t, j)] = regs[t].v[j].x;
                    pot_out[(size_t)step * 16384 + cell_row(t, 1) * 128 + cell_col(t, j)] = regs[t].v[j].y;
                }
            cells_fused<GF, SF, NP>(t, regs[t].v, A4.data(), K, sh0, sh1, part.data());
        }
        float totals[PT_FIXED + 1];
        for (int k = 0; k <= PT_FIXED; ++k) {  // same order as the statistics warp: lane partial sums, then xor-shuffle tree
            float lane[32];
            for (int ln = 0; ln < 32; ++ln) {
                float a = 0.f;
                for (int i = 0; i < 8; ++i) a += part[k * NT + ln + 32 * i];
                lane[ln] = a;
            }
            for (int off = 16; off >= 1; off >>= 1)
                for (int ln = 0; ln < 32; ++ln)
                    if ((ln & off) == 0) lane[ln] = lane[ln] + lane[ln ^ off];
            totals[k] = lane[0];
        }
        float row[ST_COUNT + MAX_C];
        stats_finalize(totals, 1, step, 1.0f / (R * R), 1.0f / R, 1.0f / stats_dt, S, row);
        for (int k = 0; k < ST_COUNT; ++k) stats[k * n_steps + step] = row[k];
        cm[step] = row[ST_COUNT];
    }
    *N_out = S.n_alive;
    for (int t = 0; t < NT; ++t)
        for (int i = 0; i < 16; ++i) {
            const float4 c = A4[i * NT + t];
            const float e[4] = {c.x, c.y, c.z, c.w};
            for (int k = 0; k < 4; ++k) final_cells[cell_row(t, i >> 3) * 128 + cell_col(t, 4 * (i & 7) + k)] = e[k];
        }
}

// Same run through the TMEM kernel's data flow (lnx_world128_tm): the state of a thread lives in a private 64-float store
// (tensor memory on the device), the new state stays in the registers between the cell phase and phase 1.
template <int GF, int SF, bool NP>
static void run_fused_rs(const float* cells0, const float* Kt, const float* Kpq, float m, float s, float w, int mean, float T, float R,
                         float stats_dt, int n_steps, float* stats, float* cm, float* N_out, float* final_cells) {
    std::vector<Regs> regs(NT);
    std::vector<float2> W(W_COMPLEX);
    std::vector<float> store(64 * NT);
    std::vector<float> part((PT_FIXED + 1) * NT);
    float2 tw[128];
    make_tw(tw);
    std::vector<float4> twtab(TW_TABLE_F4);
    for (int t = 0; t < NT; ++t) init_twiddle_table(t, twtab.data(), tw);
    for (int t = 0; t < NT; ++t)
        for (int i = 0; i < 8; ++i)
            for (int e = 0; e < 4; ++e) {
                const float a0 = cells0[cell_row(t, 0) * 128 + cell_col(t, 4 * i + e)], a1 = cells0[cell_row(t, 1) * 128 + cell_col(t, 4 * i + e)];
                store[64 * t + 8 * i + 2 * e] = a0;
                store[64 * t + 8 * i + 2 * e + 1] = a1;
                regs[t].v[4 * i + e] = make_float2(a0, a1);
            }
    const FusedConsts K = fused_consts(GF, m, s, w, mean, 1.0f / T);
    StatsCarry S;
    S.reset();
    for (int step = 0; step < n_steps; ++step) {
        const int sh0 = S.shift[0], sh1 = S.shift[1];
        std::vector<float4> xt(XT_F4);
        for (int idx = 0; idx < 32; ++idx) xt_build(idx, sh1, xt.data());
        for (int t = 0; t < NT; ++t) phase1(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) phase2_load(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) phase2_compute_store(t, regs[t], W.data(), twtab.data());
        run_phase3(regs, W, Kt, Kpq);
        for (int t = 0; t < NT; ++t) phase4_load(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) phase4_compute_store(t, regs[t], W.data(), twtab.data());
        for (int t = 0; t < NT; ++t) phase5_load(t, regs[t], W.data());
        for (int t = 0; t < NT; ++t) {
            phase5_ifft(regs[t]);
            const ArrayStore st{store.data() + 64 * t};
            cells_fused_rs<GF, SF, NP>(t, regs[t].v, st, K, sh0, xt.data(), part.data());
        }
        float totals[PT_FIXED + 1];
        for (int k = 0; k <= PT_FIXED; ++k) {
            float lane[32];
            for (int ln = 0; ln < 32; ++ln) {
                float a = 0.f;
                for (int i = 0; i < 8; ++i) a += part[k * NT + ln + 32 * i];
                lane[ln] = a;
            }
            for (int off = 16; off >= 1; off >>= 1)
                for (int ln = 0; ln < 32; ++ln)
                    if ((ln & off) == 0) lane[ln] = lane[ln] + lane[ln ^ off];
            totals[k] = lane[0];
        }
        float row[ST_COUNT + MAX_C];
        stats_finalize(totals, 1, step, 1.0f / (R * R), 1.0f / R, 1.0f / stats_dt, S, row);
        for (int k = 0; k < ST_COUNT; ++k) stats[k * n_steps + step] = row[k];
        cm[step] = row[ST_COUNT];
    }
    *N_out = S.n_alive;
    for (int t = 0; t < NT; ++t)
        for (int i = 0; i < 8; ++i)
            for (int e = 0; e < 4; ++e) {
                final_cells[cell_row(t, 0) * 128 + cell_col(t, 4 * i + e)] = store[64 * t + 8 * i + 2 * e];
                final_cells[cell_row(t, 1) * 128 + cell_col(t, 4 * i + e)] = store[64 * t + 8 * i + 2 * e + 1];
            }
}

extern "C" {
int lnx_emul_run_fused_rs(const float* cells0, const float* Kt, const float* Kpq, float m, float s, float w, int mean, float T, float R,
                          float stats_dt, int n_steps, float* stats, float* cm, float* N_out, float* final_cells) {
    run_fused_rs<GF_POLY_QUAD4, SF_V1, true>(cells0, Kt, Kpq, m, s, w, mean, T, R, stats_dt, n_steps, stats, cm, N_out, final_cells);
    return 0;
}
}

extern "C" {
int lnx_emul_run_fused(const float* cells0, const float* Kt, const float* Kpq, int gf, float m, float s, float w, int mean,
                       float T, int sf, float R, float stats_dt, int n_steps, float* stats, float* cm, float* N_out,
                       float* final_cells, float* pot_out) {
#define LNX_CASE(G, F)                                                                                                          \
    if (gf == G && sf == F) {                                                                                                   \
        run_fused<G, F, true>(cells0, Kt, Kpq, m, s, w, mean, T, R, stats_dt, n_steps, stats, cm, N_out, final_cells, pot_out); \
        return 0;                                                                                                               \
    }
    LNX_CASE(GF_POLY_QUAD4, SF_V1)
    LNX_CASE(GF_POLY_QUAD4, SF_V2)
    LNX_CASE(GF_GAUSSIAN, SF_V1)
    LNX_CASE(GF_GAUSSIAN_TARGET, SF_V2)
    LNX_CASE(GF_STEP, SF_V1)
    LNX_CASE(GF_STAIRCASE, SF_V1)
    LNX_CASE(GF_TRIANGLE, SF_V1)
    LNX_CASE(GF_IDENTITY, SF_SIMPLE)
#undef LNX_CASE
    return -1;
}

}  // extern "C"
