// Throughput of the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2) against their scalar forms, in the mixes the
// Lenia FFT butterflies use.  Question answered: does a packed instruction cost one issue slot for two lane-operations, and
// what does the FMA pipe sustain?  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/f32x2_probe.bin tools/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int NV = 16;  // independent accumulators per thread (float) / 8 float2

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(const float* __restrict__ in, float* out, int iters) {
    float x[NV], y[NV], z[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        x[i] = in[threadIdx.x + 256 * i];
        y[i] = in[threadIdx.x + 256 * (i + NV)];
        z[i] = in[threadIdx.x + 256 * (i + 2 * NV)];
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if constexpr (MODE == 0) {  // scalar FADD, 2 register sources
#pragma unroll
                for (int i = 0; i < NV; ++i) x[i] = x[i] + y[i];
            } else if constexpr (MODE == 1) {  // FADD2
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    const float2 r = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(y[i], y[i + 1]));
                    x[i] = r.x;
                    x[i + 1] = r.y;
                }
            } else if constexpr (MODE == 2) {  // scalar FFMA, 3 register sources
#pragma unroll
                for (int i = 0; i < NV; ++i) x[i] = fmaf(x[i], y[i], z[i]);
            } else if constexpr (MODE == 3) {  // FFMA2
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    const float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(y[i], y[i + 1]), make_float2(z[i], z[i + 1]));
                    x[i] = r.x;
                    x[i + 1] = r.y;
                }
            } else if constexpr (MODE == 4) {  // scalar complex butterflies (a+b, a-b) on 4 complex pairs: 16 FADD
#pragma unroll
                for (int i = 0; i < NV; i += 4) {
                    const float ar = x[i], ai = x[i + 1], br = x[i + 2], bi = x[i + 3];
                    x[i] = ar + br;
                    x[i + 1] = ai + bi;
                    x[i + 2] = ar - br;
                    x[i + 3] = ai - bi;
                }
            } else if constexpr (MODE == 5) {  // packed butterflies: 8 FADD2 (subtraction through a negated operand)
#pragma unroll
                for (int i = 0; i < NV; i += 4) {
                    const float2 a = make_float2(x[i], x[i + 1]), b = make_float2(x[i + 2], x[i + 3]);
                    const float2 s = __fadd2_rn(a, b), d = __fadd2_rn(a, make_float2(-b.x, -b.y));
                    x[i] = s.x;
                    x[i + 1] = s.y;
                    x[i + 2] = d.x;
                    x[i + 3] = d.y;
                }
            } else if constexpr (MODE == 6) {  // scalar FFMA with one immediate-like constant (twiddle): x = x * c + y
#pragma unroll
                for (int i = 0; i < NV; ++i) x[i] = fmaf(x[i], 0.98078528f, y[i]);
            } else if constexpr (MODE == 7) {  // FFMA2 with a register-pair constant
                const float2 c = make_float2(z[0], z[1]);
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    const float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), c, make_float2(y[i], y[i + 1]));
                    x[i] = r.x;
                    x[i + 1] = r.y;
                }
            } else if constexpr (MODE == 8) {  // mix: 8 FADD2 + 8 scalar ALU ops (FSET-like compare+select) per rep
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    const float2 r = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(y[i], y[i + 1]));
                    x[i] = r.x;
                    x[i + 1] = r.y;
                    z[i] = fmaxf(z[i], r.x);
                }
            } else if constexpr (MODE == 9) {  // mix: 16 scalar FADD + 8 FMNMX
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    x[i] = x[i] + y[i];
                    x[i + 1] = x[i + 1] + y[i + 1];
                    z[i] = fmaxf(z[i], x[i]);
                }
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) acc += x[i] + z[i];
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE>
static int run(const char* name, int lane_ops_per_rep, const float* in, float* out, int sms) {
    const int iters = 20000, grid = 2 * sms;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    probe<MODE><<<grid, 256>>>(in, out, 200);
    CK(cudaEventRecord(e0));
    probe<MODE><<<grid, 256>>>(in, out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // per SMSP: 4 warps; lane-operations per warp per iteration = 4 reps * lane_ops_per_rep
    const double cycles = ms * 1e-3 * 1.965e9;
    const double warp_ops = 4.0 /*warps per SMSP*/ * iters * 4.0 * lane_ops_per_rep;
    printf("%-58s %8.3f ms  %.3f lane-op warp-instr-equivalents / cycle / SMSP\n", name, ms, warp_ops / cycles);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    float *in, *out;
    CK(cudaMalloc(&in, 256 * 3 * NV * 4));
    CK(cudaMalloc(&out, 2 * prop.multiProcessorCount * 256 * 4));
    CK(cudaMemset(in, 0, 256 * 3 * NV * 4));
    printf("%s, %d SMs; 2 CTAs x 8 warps per SM = 4 warps per scheduler; numbers assume 1.965 GHz\n", prop.name, prop.multiProcessorCount);
    const int s = prop.multiProcessorCount;
    if (run<0>("scalar FADD (2 reg sources)", 16, in, out, s)) return 1;
    if (run<1>("FADD2", 16, in, out, s)) return 1;
    if (run<2>("scalar FFMA (3 reg sources)", 16, in, out, s)) return 1;
    if (run<3>("FFMA2 (3 reg-pair sources)", 16, in, out, s)) return 1;
    if (run<4>("scalar butterflies (16 FADD)", 16, in, out, s)) return 1;
    if (run<5>("packed butterflies (8 FADD2)", 16, in, out, s)) return 1;
    if (run<6>("scalar FFMA with immediate multiplier", 16, in, out, s)) return 1;
    if (run<7>("FFMA2 with a shared register-pair multiplier", 16, in, out, s)) return 1;
    if (run<8>("8 FADD2 + 8 FMNMX (counted as 24 lane-ops)", 24, in, out, s)) return 1;
    if (run<9>("16 FADD + 8 FMNMX (24 lane-ops)", 24, in, out, s)) return 1;
    return 0;
}
