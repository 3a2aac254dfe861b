// Direct-convolution potential (leniax/core.py:105-146, get_potential with fft=False) and the rest of core.update on
// top of it.  This is the reference's small-kernel / cross-check path, not the throughput path: the FFT
// kernels are.  Any 2-D world size (not only powers of two).
//
// lax.conv_general_dilated is a cross-correlation of the wrap-padded state (helpers.py:464-479: dim//2 cells before,
// dim//2 - 1 (even kernels) or dim//2 (odd) after), i.e.
//     potential[n][k][y][x] = sum_{i,j} state[n][c_in(k)][(y + i - kh/2) mod H][(x + j - kw/2) mod W] * K[slot(k)][i][j]
#pragma once
#include "lnx_step.cuh"

namespace lnx {
namespace conv {

struct ConvArgs {
    const float* state;     // [n][C][H][W]
    const float* kernels;   // [nb_slots][kh][kw]
    float* potential;       // [n][K][H][W]
    int C, K, H, W, kh, kw;
    int slot[MAX_K], c_in[MAX_K];
};
// grid (ceil(W/32), ceil(H/8), n*K), block (32, 8): one output cell per thread, kernel taps staged in shared memory in
// chunks of rows, state read through L1 (neighbouring lanes read neighbouring cells)
constexpr int TAP_CHUNK = 4096;
__global__ void __launch_bounds__(256) potential_kernel(ConvArgs P) {
    __shared__ float taps[TAP_CHUNK];
    const int k = blockIdx.z % P.K, n = blockIdx.z / P.K;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const bool live = x < P.W && y < P.H;
    const float* img = P.state + ((size_t)n * P.C + P.c_in[k]) * P.H * P.W;
    const float* ker = P.kernels + (size_t)P.slot[k] * P.kh * P.kw;
    const int rows_per_chunk = TAP_CHUNK / P.kw > 0 ? TAP_CHUNK / P.kw : 1;
    int x0 = (x - P.kw / 2) % P.W;
    if (x0 < 0) x0 += P.W;
    float acc = 0.f;
    for (int i0 = 0; i0 < P.kh; i0 += rows_per_chunk) {
        const int ni = min(rows_per_chunk, P.kh - i0);
        __syncthreads();
        for (int t = threadIdx.y * 32 + threadIdx.x; t < ni * P.kw; t += 256) taps[t] = ker[(size_t)i0 * P.kw + t];
        __syncthreads();
        if (live) {
            int yy = (y + i0 - P.kh / 2) % P.H;
            if (yy < 0) yy += P.H;
            for (int i = 0; i < ni; ++i) {
                const float* row = img + (size_t)yy * P.W;
                int xx = x0;
                for (int j = 0; j < P.kw; ++j) {
                    acc = fmaf(__ldg(row + xx), taps[i * P.kw + j], acc);
                    xx = xx + 1 == P.W ? 0 : xx + 1;
                }
                yy = yy + 1 == P.H ? 0 : yy + 1;
            }
        }
    }
    if (live) P.potential[(((size_t)n * P.K + k) * P.H + y) * P.W + x] = acc;
}

struct FieldArgs {
    const float* state;      // [n][C][cells]
    const float* potential;  // [n][K][cells]
    const float* gf_params;  // [K][2]
    const float* weights;    // [C][K]
    float* state_out;        // [n][C][cells]
    float* field_out;        // [n][C][cells]
    long long cells;
    int C, K, state_fn, mean;
    float dt;
    int gf_id[MAX_K];
};
// get_field (core.py:163-199) + weighted mean/sum (:202-242) + get_state* (:245-319), one thread per cell
__global__ void __launch_bounds__(256) field_update_kernel(FieldArgs P) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (i >= P.cells) return;
    float g[MAX_K];
#pragma unroll 1
    for (int k = 0; k < P.K; ++k) {
        const GfConst gc = gf_prepare(P.gf_id[k], P.gf_params[2 * k], P.gf_params[2 * k + 1]);
        g[k] = growth_dyn<true>(P.gf_id[k], P.potential[((size_t)n * P.K + k) * P.cells + i], gc);
    }
#pragma unroll 1
    for (int c = 0; c < P.C; ++c) {
        float f = 0.f, wsum = 0.f;
#pragma unroll 1
        for (int k = 0; k < P.K; ++k) {
            const float w = P.weights[c * P.K + k];
            wsum += w;
            if (w != 0.f) f += w * g[k];  // structural zeros are skipped like in the FFT kernels
        }
        if (P.mean) f = f / wsum;
        const size_t o = ((size_t)n * P.C + c) * P.cells + i;
        P.field_out[o] = f;
        P.state_out[o] = state_update_dyn<true>(P.state_fn, P.state[o], f, P.dt);
    }
}

}  // namespace conv
}  // namespace lnx
