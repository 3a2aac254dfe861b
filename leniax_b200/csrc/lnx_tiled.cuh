// Tiled multi-pass engine: worlds that do not fit the resident 128x128 kernel (any power-of-two 2-D / 3-D size,
// e.g. BASELINE config D 2048x2048 and config E 64^3).  State and spectra live in HBM/L2; one Lenia step is
//
//   pass A  inner transforms   rows (two real rows per complex FFT) [+ axis-1 for 3-D planes]  -> half spectrum S_c
//   pass B  leading axis       FFT along axis 0, multiply by K_k, inverse FFT along axis 0     -> P_k
//   pass C  inner inverse      [axis-1 inverse,] rows inverse, growth, weighted mix, state update, statistics partials
//   pass D  statistics         reduce partials per world, finalise the 12 statistics, carry + stop criteria
//
// Algorithmic HBM traffic for 1 channel / 1 kernel: A 4+4, B 4+4+4, C 4+4+4 = 32 B per cell-update.
// Same reference functions as the resident kernels: leniax/core.py:52-102, :163-319, leniax/statistics.py:36-205.
#pragma once
#include "lnx_step.cuh"

namespace lnx {
namespace tiled {

constexpr int NMAX = 4096;     // longest supported axis
constexpr int TPB = 256;       // threads per block of every pass
constexpr int MAXD = 3;
constexpr int NP_T = 4 + 3 * MAXD + MAX_C;  // partials: cnt_a, g00, cnt_g, cnt_p, MX[3], MX2[3], GX[3], m00[C]

struct Geom {
    int nd;          // 2 or 3
    int dims[3];     // world dims, leading first (2-D: dims[0]=H, dims[1]=W)
    int L;           // leading axis length (pass B)
    int A1, A2;      // inner slab: A1 rows of A2 reals (2-D: A1 = 1)
    int logL, logA1, logA2;
    int half;        // A2/2 + 1 spectral columns
    int rows;        // L * A1 rows of length A2 in a world
    int slab_rows;   // rows handled by one CTA of pass A / C (3-D: A1; 2-D: 2..16)
    int n_slabs;     // rows / slab_rows
    long long cells;      // L * A1 * A2
    long long spec;       // rows * half complex values per (world, channel)
    int tc;          // pass B tile width (inner spectral columns per CTA)
    int any_size;    // 1: dims are not all powers of two - only the stand-alone statistics accept such a geometry (division / modulo
                     // instead of shifts / masks); the scan engines never see it
};

__device__ __forceinline__ int brev_n(int x, int logn) { return (int)(__brev((unsigned)x) >> (32 - logn)); }

// In-place FFT of `nl` lines of length n held in shared memory; element i of line j at s[i * is + j * js].
// Forward: DIF, natural order in -> bit-reversed order out.  Inverse: DIT, bit-reversed in -> natural out (unnormalised).
// LF = lines are the fast thread axis (js == 1).  tw[k] = (cos, sin)(2 pi k / NMAX).
//
// The radix-2 butterflies are grouped three stages at a time: a thread loads the 8 elements of a radix-8 butterfly into
// registers, runs the three stages on them with the packed FP32 instructions and stores them back, so a 2048-point line
// takes 4 shared-memory round trips and barriers instead of 11 (remaining stages: one radix-4 or radix-2 pass).
constexpr int LOG_NMAX = 12;
static_assert((1 << LOG_NMAX) == NMAX, "LOG_NMAX");

// one pass of LOGR stages; lsp = log2 of the smallest span of the pass
template <bool INV, bool LF, int LOGR>
__device__ __forceinline__ void radix_pass(float2* s, int n, int logn, int is, int nl, int js, const float2* tw, int log_tw, int lsp) {
    constexpr int R = 1 << LOGR;
    const int per_line = n >> LOGR, total = per_line * nl;
    const int sp = 1 << lsp;
    for (int b = threadIdx.x; b < total; b += blockDim.x) {
        int line, q;
        if (LF) {
            line = b % nl;
            q = b / nl;
        } else {
            q = b & (per_line - 1);
            line = b >> (logn - LOGR);
        }
        const int p = q & (sp - 1), g = q >> lsp;
        float2* base = s + (size_t)((g << (lsp + LOGR)) + p) * is + (size_t)line * js;
        const size_t st = (size_t)sp * is;
        float2 x[R];
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = base[m * st];
#pragma unroll
        for (int j = 0; j < LOGR; ++j) {
            // forward: largest span first; inverse: smallest span first
            const int lh = INV ? j : LOGR - 1 - j;  // log2 of the pair distance in units of m
            const int hs = 1 << lh;
            const int tshift = log_tw - 1 - (lsp + lh);  // twiddle index = pos * 2^log_tw / (2 * span), span = sp << lh
#pragma unroll
            for (int b2 = 0; b2 < R / 2; ++b2) {  // fixed trip count: every register index below is a compile-time constant
                const int u = b2 & (hs - 1), lo = ((b2 >> lh) << (lh + 1)) + u, hi = lo + hs;
                const float2 w = tw[(p + u * sp) << tshift];
                if (!INV) {
                    const float2 d = csub(x[lo], x[hi]);
                    x[lo] = cadd(x[lo], x[hi]);
                    x[hi] = rot_fwd(d, w.x, w.y);  // d * (cos - i sin)
                } else {
                    const float2 t = rot_inv(x[hi], w.x, w.y);  // y * (cos + i sin)
                    x[hi] = csub(x[lo], t);
                    x[lo] = cadd(x[lo], t);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) base[m * st] = x[m];
    }
}
// copy the part of the master table an FFT of length 2^logn needs into shared memory: stw[k] = (cos, sin)(2 pi k / 2^logn),
// k < 2^(logn-1).  (The butterflies read 7 twiddles per radix-8 group: from L1/L2 that was the top stall of every pass.)
__device__ __forceinline__ void load_twiddles(float2* stw, int logn, const float2* __restrict__ tw) {
    const int cnt = 1 << (logn - 1), sh = LOG_NMAX - logn;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) stw[k] = __ldg(tw + (k << sh));
}
template <bool INV, bool LF>
__device__ void block_fft(float2* s, int n, int logn, int is, int nl, int js, const float2* tw, int log_tw) {
    int done = 0;
    while (done < logn) {
        const int r = logn - done >= 3 ? 3 : logn - done;
        const int lsp = INV ? done : logn - done - r;
        if (r == 3)
            radix_pass<INV, LF, 3>(s, n, logn, is, nl, js, tw, log_tw, lsp);
        else if (r == 2)
            radix_pass<INV, LF, 2>(s, n, logn, is, nl, js, tw, log_tw, lsp);
        else
            radix_pass<INV, LF, 1>(s, n, logn, is, nl, js, tw, log_tw, lsp);
        __syncthreads();
        done += r;
    }
}

// Threads spread over a [rows][half] index space without an integer division by `half` (not a power of two): kw = smallest
// power of two >= half (at most the block size) lanes run along k, the remaining blockDim / kw along the rows.
struct RowK {
    int k0, kstep, r0, rstep;
};
__device__ __forceinline__ RowK rowk_map(int half) {
    int kw = 32;
    while (kw < half && kw < (int)blockDim.x) kw <<= 1;
    RowK m;
    m.k0 = threadIdx.x & (kw - 1);
    m.kstep = kw;
    m.r0 = threadIdx.x / kw;  // kw is a power of two: a shift
    m.rstep = blockDim.x / kw;
    return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// pass A: real rows -> half spectrum (natural order), optional axis-1 transform for 3-D planes
//   grid (n_slabs, C, worlds).  smem: zbuf [slab_rows/2][A2] complex, then (3-D) plane [A1][half] complex
// ---------------------------------------------------------------------------------------------------------------------
struct PassAArgs {
    const float* state;   // [worlds][C][rows][A2]
    float2* spec;         // [worlds][C][rows][half]
    const float2* tw;
    Geom g;
    int C;
    int world0;  // first world of this launch (the 64^3 line engine runs L2-sized batches of worlds; 0 elsewhere)
};
__global__ void __launch_bounds__(TPB) pass_a_kernel(PassAArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Geom& g = P.g;
    float2* z = reinterpret_cast<float2*>(smem_raw);                      // [pairs][A2]
    float2* pl = z + (size_t)(g.slab_rows / 2) * g.A2;                    // [slab_rows][half] (3-D only)
    const int log_tw = g.logA2 > g.logA1 ? g.logA2 : g.logA1;
    float2* stw = pl + (g.nd == 3 ? (size_t)g.A1 * g.half : 0);           // twiddles of the longest inner axis
    load_twiddles(stw, log_tw, P.tw);
    const int slab = blockIdx.x, c = blockIdx.y, w = blockIdx.z;
    const int pairs = g.slab_rows / 2, A2 = g.A2, half = g.half;
    const size_t row0 = (size_t)slab * g.slab_rows;
    const float* src = P.state + (((size_t)w * P.C + c) * g.rows + row0) * A2;
    for (int i = threadIdx.x * 4; i < pairs * A2; i += blockDim.x * 4) {  // 128-bit loads of both rows of the pair
        const int pr = i >> g.logA2, n = i & (A2 - 1);
        const float4 ra = *reinterpret_cast<const float4*>(src + (size_t)(2 * pr) * A2 + n);
        const float4 rb = *reinterpret_cast<const float4*>(src + (size_t)(2 * pr + 1) * A2 + n);
        float4* zp = reinterpret_cast<float4*>(z + i);
        zp[0] = make_float4(ra.x, rb.x, ra.y, rb.y);
        zp[1] = make_float4(ra.z, rb.z, ra.w, rb.w);
    }
    __syncthreads();
    block_fft<false, false>(z, A2, g.logA2, 1, pairs, A2, stw, log_tw);
    float2* dst = P.spec + (((size_t)w * P.C + c) * g.rows + row0) * half;
    const bool plane = g.nd == 3;
    const RowK rk = rowk_map(half);
    for (int pr = rk.r0; pr < pairs; pr += rk.rstep)
    for (int k = rk.k0; k < half; k += rk.kstep) {
        const float2 zk = z[(size_t)pr * A2 + brev_n(k, g.logA2)];
        const float2 zc = z[(size_t)pr * A2 + brev_n((A2 - k) & (A2 - 1), g.logA2)];
        const float2 a = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
        const float2 b = make_float2(0.5f * (zk.y + zc.y), 0.5f * (zc.x - zk.x));
        if (plane) {
            pl[(size_t)(2 * pr) * half + k] = a;
            pl[(size_t)(2 * pr + 1) * half + k] = b;
        } else {
            dst[(size_t)(2 * pr) * half + k] = a;
            dst[(size_t)(2 * pr + 1) * half + k] = b;
        }
    }
    if (plane) {
        __syncthreads();
        block_fft<false, true>(pl, g.A1, g.logA1, half, half, 1, stw, log_tw);  // along axis 1, `half` interleaved lines
        for (int m1 = rk.r0; m1 < g.A1; m1 += rk.rstep) {
            const int srow = brev_n(m1, g.logA1);
            for (int k = rk.k0; k < half; k += rk.kstep) dst[(size_t)m1 * half + k] = pl[(size_t)srow * half + k];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pass B: FFT along the leading axis, multiply by every kernel fed by this channel, inverse FFT
//   grid (ceil(M / tc), C, worlds).  smem: F [L][tc] + T [L][tc]
// ---------------------------------------------------------------------------------------------------------------------
struct PassBArgs {
    const float2* spec;    // [worlds][C][L][M]
    float2* pot_spec;      // [worlds][K][L][M]     (null in forward-only mode)
    const float2* ktab;    // [n_sols][K][L][M], pre-scaled by 1 / cells
    float2* fwd_out;       // forward-only mode (kernel-spectrum builder): [images][L][M]
    const float2* tw;
    Geom g;
    int C, K, n_init;
    int two_buf;           // 1: some channel feeds several kernels (forward spectrum kept in F, products in T); 0: in place
    int world0;
    int c_in[MAX_K];
};
__global__ void __launch_bounds__(TPB) pass_b_kernel(PassBArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Geom& g = P.g;
    const int L = g.L, tc = g.tc, ltc = __ffs(g.tc) - 1;  // tc is a power of two
    const long long M = g.spec / L;
    float2* F = reinterpret_cast<float2*>(smem_raw);
    float2* T = P.two_buf ? F + (size_t)L * tc : F;
    float2* stw = F + (size_t)(P.two_buf ? 2 : 1) * L * tc;
    load_twiddles(stw, g.logL, P.tw);
    const int c = blockIdx.y, w = blockIdx.z;
    const long long m0 = (long long)blockIdx.x * tc;
    const int wdt = (int)((M - m0) < tc ? (M - m0) : tc);
    const float2* src = P.spec + ((size_t)w * P.C + c) * g.spec + m0;
#pragma unroll 8
    for (int i = threadIdx.x; i < L * tc; i += blockDim.x) {  // independent loads: keep eight in flight per thread
        const int l = i >> ltc, j = i & (tc - 1);
        F[i] = j < wdt ? __ldg(src + (size_t)l * M + j) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    block_fft<false, true>(F, L, g.logL, tc, tc, 1, stw, g.logL);
    if (P.fwd_out) {
        float2* dst = P.fwd_out + ((size_t)w * P.C + c) * g.spec + m0;
        for (int i = threadIdx.x; i < L * tc; i += blockDim.x) {
            const int m = i >> ltc, j = i & (tc - 1);
            if (j < wdt) dst[(size_t)m * M + j] = F[(size_t)brev_n(m, g.logL) * tc + j];
        }
        return;
    }
    const int sol = w / P.n_init;
    for (int k = 0; k < P.K; ++k) {
        if (P.c_in[k] != c) continue;
        const float2* kt = P.ktab + ((size_t)sol * P.K + k) * g.spec + m0;
#pragma unroll 8
        for (int i = threadIdx.x; i < L * tc; i += blockDim.x) {
            const int pos = i >> ltc, j = i & (tc - 1);
            float2 v = make_float2(0.f, 0.f);
            if (j < wdt) {
                const float2 f = F[i], q = __ldg(kt + (size_t)brev_n(pos, g.logL) * M + j);
                v = make_float2(f.x * q.x - f.y * q.y, f.x * q.y + f.y * q.x);
            }
            T[i] = v;
        }
        __syncthreads();
        block_fft<true, true>(T, L, g.logL, tc, tc, 1, stw, g.logL);
        float2* dst = P.pot_spec + ((size_t)w * P.K + k) * g.spec + m0;
#pragma unroll 8
        for (int i = threadIdx.x; i < L * tc; i += blockDim.x) {
            const int l = i >> ltc, j = i & (tc - 1);
            if (j < wdt) dst[(size_t)l * M + j] = T[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pass C: inverse inner transforms of every kernel's potential spectrum, growth, mix, state update, statistics partials
//   grid (n_slabs, 1, worlds).  smem: pl [slab_rows][half] complex, z [pairs][A2] complex, field [C][slab_rows][A2]
// ---------------------------------------------------------------------------------------------------------------------
// Loads of data that another CTA of the SAME launch may have written (persistent whole-scan kernel of lnx_tiled64h.cuh): the L1 of an SM is
// not coherent with the other SMs' stores, so these go to L2 (ld.global.cg).  For the multi-launch engines it is the same streaming data.
#ifdef __CUDA_ARCH__
#define LNX_MUT_LD(p) __ldcg(p)
#else
#define LNX_MUT_LD(p) (*(p))
#endif
struct WorldCarry {   // per world, device memory
    int shift[3];
    float centroid[3];
    float angle;
    float should_continue, prev_mass, prev_sign, n_alive;
    int mono, vol;
    float init_cm[MAX_C];
    int step, pending;   // device-side step counter (graph-replayed loops, lnx_tiled2k.cuh): index of the step whose inverse pass runs
                         // next; pending = 1: that pass has run and its statistics wait to be finalised
};
struct PassCArgs {
    float* state;             // [worlds][C][rows][A2]  updated in place
    const float2* pot_spec;   // [worlds][K][rows][half]
    const float* gf_params;   // [n_sols][K][2]
    const float* weights;     // [n_sols][C][K]
    const float* dt;          // [n_sols]
    const WorldCarry* carry;  // [worlds]
    float* partials;          // [worlds][n_slabs][NP_T]
    float* cells_out;         // [worlds' trajectory slot] may be null: [C][cells]
    float* field_out;
    float* potential_out;     // [K][cells]
    const float2* tw;
    Geom g;
    int C, K, n_init, max_iter, t;
    int state_fn, mean;
    int world0;
    int gf_id[MAX_K];
};
__global__ void __launch_bounds__(TPB, 3) pass_c_kernel(PassCArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[NP_T][TPB / 32];
    const Geom& g = P.g;
    const int A2 = g.A2, half = g.half, R = g.slab_rows, pairs = R / 2;
    float2* pl = reinterpret_cast<float2*>(smem_raw);       // [R][half]
    float2* z = pl + (size_t)R * half;                       // [pairs][A2]
    float* field = reinterpret_cast<float*>(z + (size_t)pairs * A2);  // [C][R][A2]
    const int log_tw = g.logA2 > g.logA1 ? g.logA2 : g.logA1;
    float2* stw = reinterpret_cast<float2*>(field + (size_t)P.C * R * A2);
    load_twiddles(stw, log_tw, P.tw);
    const int slab = blockIdx.x, w = blockIdx.z;
    const int sol = w / P.n_init, init = w - sol * P.n_init;
    const size_t row0 = (size_t)slab * R;
    const bool plane = g.nd == 3;
    const int slab_cells = R * A2;
    const int t = P.t >= 0 ? P.t : P.carry[w].step;  // t < 0: graph-replayed loop, the step index lives in the world's carry
    const size_t traj = ((size_t)sol * P.max_iter + t) * P.n_init + init;  // world-step slot of the trajectory outputs
    for (int i = threadIdx.x; i < P.C * slab_cells; i += blockDim.x) field[i] = 0.f;
    float acc[NP_T];
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;

    for (int k = 0; k < P.K; ++k) {
        const float2* src = P.pot_spec + (((size_t)w * P.K + k) * g.rows + row0) * half;
        __syncthreads();
        if (plane) {
            const RowK rk = rowk_map(half);
            for (int kk = rk.k0; kk < half; kk += rk.kstep)
                for (int m1 = rk.r0; m1 < R; m1 += 4 * rk.rstep) {  // four independent rows per thread in flight
                    float2 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (m1 + j * rk.rstep < R) v[j] = __ldg(src + (size_t)(m1 + j * rk.rstep) * half + kk);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (m1 + j * rk.rstep < R) pl[(size_t)brev_n(m1 + j * rk.rstep, g.logA1) * half + kk] = v[j];
                }
            __syncthreads();
            block_fft<true, true>(pl, g.A1, g.logA1, half, half, 1, stw, log_tw);
        } else {
#pragma unroll 8
            for (int i = threadIdx.x; i < R * half; i += blockDim.x) pl[i] = __ldg(src + i);
            __syncthreads();
        }
        // retangle: Z'[k] = A + iB, Z'[N-k] = conj(A) + i conj(B), stored at bit-reversed positions for the DIT
        const RowK rk2 = rowk_map(half);
        for (int pr = rk2.r0; pr < pairs; pr += rk2.rstep)
        for (int kk = rk2.k0; kk < half; kk += rk2.kstep) {
            const float2 a = pl[(size_t)(2 * pr) * half + kk], b = pl[(size_t)(2 * pr + 1) * half + kk];
            z[(size_t)pr * A2 + brev_n(kk, g.logA2)] = make_float2(a.x - b.y, a.y + b.x);
            if (kk != 0 && kk != A2 / 2) z[(size_t)pr * A2 + brev_n(A2 - kk, g.logA2)] = make_float2(a.x + b.y, b.x - a.y);
        }
        __syncthreads();
        block_fft<true, false>(z, A2, g.logA2, 1, pairs, A2, stw, log_tw);
        const GfConst gc = gf_prepare(P.gf_id[k], P.gf_params[((size_t)sol * P.K + k) * 2], P.gf_params[((size_t)sol * P.K + k) * 2 + 1]);
        float* pout = P.potential_out ? P.potential_out + (traj * P.K + k) * g.cells + row0 * A2 : nullptr;
        float wk[MAX_C];
#pragma unroll
        for (int c = 0; c < MAX_C; ++c) wk[c] = c < P.C ? P.weights[((size_t)sol * P.C + c) * P.K + k] : 0.f;
        // four consecutive cells of a row per thread and iteration: 128-bit shared / global accesses, loads in flight together
        for (int i = threadIdx.x * 4; i < slab_cells; i += blockDim.x * 4) {
            const int r = i >> g.logA2, n = i & (A2 - 1);
            const float4* zp = reinterpret_cast<const float4*>(z + (size_t)(r >> 1) * A2 + n);
            const float4 z01 = zp[0], z23 = zp[1];
            const bool odd = r & 1;
            const float pot[4] = {odd ? z01.y : z01.x, odd ? z01.w : z01.z, odd ? z23.y : z23.x, odd ? z23.w : z23.z};
            float gv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[3] += pot[e] > EPS ? 1.f : 0.f;
                gv[e] = growth_dyn<true>(P.gf_id[k], pot[e], gc);
            }
            if (pout) *reinterpret_cast<float4*>(pout + i) = make_float4(pot[0], pot[1], pot[2], pot[3]);
#pragma unroll
            for (int c = 0; c < MAX_C; ++c)
                if (c < P.C && wk[c] != 0.f) {
                    float4* fp = reinterpret_cast<float4*>(field + (size_t)c * slab_cells + i);
                    float4 f = *fp;
                    f.x += wk[c] * gv[0];
                    f.y += wk[c] * gv[1];
                    f.z += wk[c] * gv[2];
                    f.w += wk[c] * gv[3];
                    *fp = f;
                }
        }
    }
    __syncthreads();
    // ---- update + statistics ----
    const WorldCarry cr = P.carry[w];
    const float dt = P.dt[sol];
    for (int c = 0; c < P.C; ++c) {
        float wsum = 0.f;
        for (int k = 0; k < P.K; ++k) wsum += P.weights[((size_t)sol * P.C + c) * P.K + k];
        float* st = P.state + (((size_t)w * P.C + c) * g.rows + row0) * A2;
        float* cout = P.cells_out ? P.cells_out + (traj * P.C + c) * g.cells + row0 * A2 : nullptr;
        float* fout = P.field_out ? P.field_out + (traj * P.C + c) * g.cells + row0 * A2 : nullptr;
        float m00 = 0.f;
#pragma unroll 2
        for (int i = threadIdx.x * 4; i < slab_cells; i += blockDim.x * 4) {
            const int r = i >> g.logA2, n0 = i & (A2 - 1);
            const size_t grow = row0 + r;            // global row index = l * A1 + a1
            const float4 fv = *reinterpret_cast<const float4*>(field + (size_t)c * slab_cells + i);
            const float4 av = *reinterpret_cast<const float4*>(st + i);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            float f4[4] = {fv.x, fv.y, fv.z, fv.w};
            float n4[4];
            // coordinates of these cells in the rolled (centred) world, statistics.py:28-33 + utils.py:269-293
            const int i0 = g.nd == 3 ? (int)(grow >> g.logA1) : (int)grow;
            const float x0 = (float)(((i0 - cr.shift[0]) & (g.dims[0] - 1)) - g.dims[0] / 2);
            const float x1row = (float)((((int)(grow & (g.A1 - 1)) - cr.shift[1]) & (g.dims[1] - 1)) - g.dims[1] / 2);  // 3-D only
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = n0 + e;
                float f = f4[e];
                if (P.mean) f = f / wsum;
                f4[e] = f;
                const float a = a4[e];
                n4[e] = state_update_dyn<true>(P.state_fn, a, f, dt);
                const float gp = fmaxf(f, 0.f);
                m00 += a;
                acc[0] += a > EPS ? 1.f : 0.f;
                acc[1] += gp;
                acc[2] += gp > EPS ? 1.f : 0.f;
                acc[4] += a * x0;
                acc[4 + MAXD] += a * x0 * x0;
                acc[4 + 2 * MAXD] += gp * x0;
                if (g.nd == 3) {
                    const float x2 = (float)(((n - cr.shift[2]) & (g.dims[2] - 1)) - g.dims[2] / 2);
                    acc[5] += a * x1row;
                    acc[5 + MAXD] += a * x1row * x1row;
                    acc[5 + 2 * MAXD] += gp * x1row;
                    acc[6] += a * x2;
                    acc[6 + MAXD] += a * x2 * x2;
                    acc[6 + 2 * MAXD] += gp * x2;
                } else {
                    const float x1 = (float)(((n - cr.shift[1]) & (g.dims[1] - 1)) - g.dims[1] / 2);
                    acc[5] += a * x1;
                    acc[5 + MAXD] += a * x1 * x1;
                    acc[5 + 2 * MAXD] += gp * x1;
                }
            }
            *reinterpret_cast<float4*>(st + i) = make_float4(n4[0], n4[1], n4[2], n4[3]);
            if (cout) *reinterpret_cast<float4*>(cout + i) = av;
            if (fout) *reinterpret_cast<float4*>(fout + i) = make_float4(f4[0], f4[1], f4[2], f4[3]);
        }
#pragma unroll
        for (int cc = 0; cc < MAX_C; ++cc)
            if (cc == c) acc[4 + 3 * MAXD + cc] = m00;
    }
    // block reduction of the partials
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NP_T; ++i) {
        float v = acc[i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NP_T) {
        float v = 0.f;
        for (int j = 0; j < TPB / 32; ++j) v += red[threadIdx.x][j];
        P.partials[((size_t)w * g.n_slabs + slab) * NP_T + threadIdx.x] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pass D: per-world statistics (leniax/statistics.py:36-126, 134-205 for nb_dims = 2 or 3)
// ---------------------------------------------------------------------------------------------------------------------
struct PassDArgs {
    const float* partials;
    WorldCarry* carry;
    float* stats;          // [ST_COUNT][n_sols][T][n_init]
    float* channel_mass;   // [n_sols][T][n_init][C]
    float* n_alive;
    Geom g;
    int C, n_sols, n_init, max_iter, t;
    int world0;
    float R, stats_dt;
};
constexpr int PASS_D_MAX_WARPS = 16;
// the work of one CTA (1..16 warps) for world w; also called from the four-step engine's lead kernel (lnx_tiled2k.cuh), where the
// statistics of step t are finalised by an extra CTA of step t+1's lead launch (nothing before rows_inv needs the carry)
__device__ __forceinline__ void pass_d_body(const PassDArgs& P, const int w, const int t) {
    __shared__ float tot[NP_T];
    __shared__ float red[NP_T][PASS_D_MAX_WARPS];
    const Geom& g = P.g;
    const int sol = w / P.n_init, init = w - sol * P.n_init;
    float acc[NP_T];
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;
    for (int s = threadIdx.x; s < g.n_slabs; s += blockDim.x) {
        const float* p = P.partials + ((size_t)w * g.n_slabs + s) * NP_T;
#pragma unroll
        for (int i = 0; i < NP_T; ++i) acc[i] += LNX_MUT_LD(p + i);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NP_T; ++i) {
        float v = acc[i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NP_T) {
        float v = 0.f;
        for (int j = 0; j < (int)(blockDim.x >> 5); ++j) v += red[threadIdx.x][j];
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    WorldCarry S;
    {
        static_assert(sizeof(WorldCarry) % sizeof(int) == 0, "WorldCarry is copied word by word");
        const int* src = reinterpret_cast<const int*>(P.carry + w);
        int* dst = reinterpret_cast<int*>(&S);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(WorldCarry) / sizeof(int)); ++i) dst[i] = LNX_MUT_LD(src + i);
    }
    const int nd = g.nd, C = P.C;
    // reference: R**2 is used for every "volume" normalisation whatever the dimension (statistics.py:70-78)
    const float R2 = P.R * P.R, R = P.R, dt = P.stats_dt;
    float m00 = 0.f, cm[MAX_C];
    for (int c = 0; c < C; ++c) {
        m00 += tot[4 + 3 * MAXD + c];
        cm[c] = tot[4 + 3 * MAXD + c] / R2;
    }
    const float g00 = tot[1];
    const float mass = m00 / R2, mass_volume = tot[0] / R2, growth = g00 / R2, growth_volume = tot[2] / R2;
    float out[ST_COUNT];
    out[ST_MASS] = mass;
    out[ST_MASS_VOLUME] = mass_volume;
    out[ST_MASS_DENSITY] = mass / (mass_volume + EPS);
    out[ST_GROWTH] = growth;
    out[ST_GROWTH_VOLUME] = growth_volume;
    out[ST_GROWTH_DENSITY] = growth / (growth_volume + EPS);
    out[ST_POTENTIAL_VOLUME] = tot[3] / R2;
    float cen[3] = {0.f, 0.f, 0.f}, dl[3] = {0.f, 0.f, 0.f};
    float dist2 = 0.f, gd2 = 0.f, inertia = 0.f;
    const float den = m00 * m00 + EPS;
    for (int d = 0; d < nd; ++d) {
        cen[d] = tot[4 + d] / (m00 + EPS);
        dl[d] = cen[d] - S.centroid[d];
        dist2 += dl[d] * dl[d];
        const float gcd = tot[4 + 2 * MAXD + d] / (g00 + EPS) - cen[d];
        gd2 += gcd * gcd;
        inertia += (tot[4 + MAXD + d] - cen[d] * tot[4 + d]) / den;
    }
    const float dist = sqrtf(dist2);
    out[ST_MASS_SPEED] = dist / R / dt;
    const float angle = (atan2f(dl[1], dl[0]) * 57.29577951308232f) * ((dist / R > 0.001f) ? 1.f : 0.f);  // dims 0 and 1 only
    out[ST_MASS_ANGLE_SPEED] = (mod360(angle - S.angle + 540.f) - 180.f) / dt;
    out[ST_MASS_GROWTH_DIST] = sqrtf(gd2) / R;
    out[ST_INERTIA] = inertia;
    for (int d = 0; d < nd; ++d) {
        const int s = trunc_to_int(cen[d]);
        S.shift[d] = py_mod(S.shift[d] + s, g.dims[d]);  // (= the mask for powers of two)
        S.centroid[d] = cen[d] - (float)s;
    }
    S.angle = angle;
    if (t == 0) {
        for (int c = 0; c < C; ++c) S.init_cm[c] = cm[c];
        S.prev_mass = mass;
        S.prev_sign = 0.f;
        S.should_continue = 1.f;
        S.n_alive = 0.f;
        S.mono = S.vol = 0;
    }
    bool cond = true;
    for (int c = 0; c < C; ++c) cond = cond && (cm[c] >= EPS) && (cm[c] <= 3.f * S.init_cm[c]);
    const float dm = mass - S.prev_mass;
    const float sign = (dm > 0.f) ? 1.f : ((dm < 0.f) ? -1.f : dm);
    S.mono = S.mono * (sign == S.prev_sign ? 1 : 0) + 1;
    cond = cond && (S.mono <= 128);
    S.vol = S.vol * (mass_volume > 10.f ? 1 : 0) + 1;
    cond = cond && (S.vol <= 128);
    S.should_continue *= cond ? 1.f : 0.f;
    S.n_alive += S.should_continue;
    S.prev_mass = mass;
    S.prev_sign = sign;
    S.step = t + 1;
    S.pending = 0;
    P.carry[w] = S;
    const size_t plane = (size_t)P.n_sols * P.max_iter * P.n_init;
    const size_t idx = ((size_t)sol * P.max_iter + t) * P.n_init + init;
    for (int k = 0; k < ST_COUNT; ++k) P.stats[k * plane + idx] = out[k];
    for (int c = 0; c < C; ++c) P.channel_mass[idx * C + c] = cm[c];
    P.n_alive[w] = S.n_alive;
}
__global__ void __launch_bounds__(32 * PASS_D_MAX_WARPS) pass_d_kernel(PassDArgs P) {  // 4..16 warps: see pass_d_threads()
    const int w = blockIdx.x + P.world0;
    pass_d_body(P, w, P.t >= 0 ? P.t : P.carry[w].step);
}

// ---------------------------------------------------------------------------------------------------------------------
// stand-alone compute_stats (leniax/statistics.py:36-126 called outside a scan, e.g. by user code): partial sums of given
// cells / field / potential arrays, finalised by pass_d_kernel
// ---------------------------------------------------------------------------------------------------------------------
struct StatsPartialArgs {
    const float* cells;      // [worlds][C][cells]
    const float* field;      // [worlds][C][cells]
    const float* potential;  // [worlds][K][cells]
    const WorldCarry* carry;
    float* partials;         // [worlds][n_slabs][NP_T]
    Geom g;
    int C, K;
};
__global__ void __launch_bounds__(TPB) stats_partials_kernel(StatsPartialArgs P) {
    __shared__ float red[NP_T][TPB / 32];
    const Geom& g = P.g;
    const int slab = blockIdx.x, w = blockIdx.z;
    const int slab_cells = g.slab_rows * g.A2;
    const size_t off = (size_t)slab * slab_cells;
    const WorldCarry cr = P.carry[w];
    float acc[NP_T];
#pragma unroll
    for (int i = 0; i < NP_T; ++i) acc[i] = 0.f;
    for (int k = 0; k < P.K; ++k) {
        const float* pt = P.potential + ((size_t)w * P.K + k) * g.cells + off;
        for (int i = threadIdx.x; i < slab_cells; i += blockDim.x) acc[3] += pt[i] > EPS ? 1.f : 0.f;
    }
    for (int c = 0; c < P.C; ++c) {
        const float* ce = P.cells + ((size_t)w * P.C + c) * g.cells + off;
        const float* fi = P.field + ((size_t)w * P.C + c) * g.cells + off;
        float m00 = 0.f;
        for (int i = threadIdx.x; i < slab_cells; i += blockDim.x) {
            const size_t grow = (size_t)slab * g.slab_rows + (g.any_size ? i / g.A2 : (i >> g.logA2));
            const int n = g.any_size ? i % g.A2 : (i & (g.A2 - 1));
            int idx[3];
            if (g.nd == 3) {
                idx[0] = (int)(g.any_size ? grow / g.A1 : (grow >> g.logA1));
                idx[1] = (int)(g.any_size ? grow % g.A1 : (grow & (g.A1 - 1)));
                idx[2] = n;
            } else {
                idx[0] = (int)grow;
                idx[1] = n;
                idx[2] = 0;
            }
            const float a = ce[i], gp = fmaxf(fi[i], 0.f);
            m00 += a;
            acc[0] += a > EPS ? 1.f : 0.f;
            acc[1] += gp;
            acc[2] += gp > EPS ? 1.f : 0.f;
            for (int d = 0; d < g.nd; ++d) {
                const float x = (float)(py_mod(idx[d] - cr.shift[d], g.dims[d]) - g.dims[d] / 2);  // (= the mask for powers of two)
                acc[4 + d] += a * x;
                acc[4 + MAXD + d] += a * x * x;
                acc[4 + 2 * MAXD + d] += gp * x;
            }
        }
        acc[4 + 3 * MAXD + c] = m00;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NP_T; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NP_T) {
        float v = 0.f;
        for (int j = 0; j < TPB / 32; ++j) v += red[threadIdx.x][j];
        P.partials[((size_t)w * g.n_slabs + slab) * NP_T + threadIdx.x] = v;
    }
}
// user-facing carry arrays <-> WorldCarry (total_shift_idx [N][nd] int32, mass_centroid [nd][N], mass_angle [N])
__global__ void carry_pack_kernel(WorldCarry* carry, const int* shift, const float* centroid, const float* angle, int n, int nd, bool unpack,
                                  int* shift_out, float* centroid_out, float* angle_out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    if (!unpack) {
        WorldCarry c;
        memset(&c, 0, sizeof(c));
        for (int d = 0; d < nd; ++d) {
            c.shift[d] = shift[w * nd + d];
            c.centroid[d] = centroid[d * n + w];
        }
        c.angle = angle[w];
        carry[w] = c;
    } else {
        const WorldCarry c = carry[w];
        for (int d = 0; d < nd; ++d) {
            shift_out[w * nd + d] = c.shift[d];
            centroid_out[d * n + w] = c.centroid[d];
        }
        angle_out[w] = c.angle;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// helpers: kernel table gather, Hermitian expansion (kernel-spectrum builder)
// ---------------------------------------------------------------------------------------------------------------------
// K_fft full complex [n_sols][nb_slots][cells] (reference layout) -> table [n_sols][K][rows][half] scaled by 1/cells
__global__ void gather_ktab_kernel(const float2* __restrict__ K_fft, float2* __restrict__ tab, Geom g, int K, int nb_slots,
                                   const int* __restrict__ slots_dev, float scale) {
    const long long n = (long long)g.spec;
    const int sol = blockIdx.z, k = blockIdx.y;
    const float2* src = K_fft + ((size_t)sol * nb_slots + slots_dev[k]) * g.cells;
    float2* dst = tab + ((size_t)sol * K + k) * g.spec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / g.half;
        const int kk = (int)(i - row * g.half);
        const float2 v = src[row * g.A2 + kk];
        dst[i] = make_float2(v.x * scale, v.y * scale);
    }
}
// half spectrum [images][rows][half] -> full spectrum [images][rows][A2] using X[-m] = conj(X[m]) over all axes
__global__ void expand_hermitian_kernel(const float2* __restrict__ half_spec, float2* __restrict__ full, Geom g) {
    const int img = blockIdx.z;
    const float2* src = half_spec + (size_t)img * g.spec;
    float2* dst = full + (size_t)img * g.cells;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.cells; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> g.logA2;
        const int kk = (int)(i & (g.A2 - 1));
        if (kk < g.half) {
            dst[i] = src[row * g.half + kk];
        } else {
            const int l = (int)(row >> g.logA1), a1 = (int)(row & (g.A1 - 1));
            const long long mrow = ((long long)((g.L - l) & (g.L - 1)) << g.logA1) + ((g.A1 - a1) & (g.A1 - 1));
            const float2 v = src[mrow * g.half + (g.A2 - kk)];
            dst[i] = make_float2(v.x, -v.y);
        }
    }
}

}  // namespace tiled
}  // namespace lnx
