// Translation unit of the set-up kernels either side of the scan: everything a QD generation needs
// before lnx_run_scan, batched over all individuals in a handful of launches.
//   lnx_rasterize_kernels  circle_2d / ellipse_2d / oriented_ellipse_2d + the seven kernel functions (leniax/kernels.py:176-309,
//                          leniax/kernel_functions.py:7-261), one CTA per kernel, fp32 in the reference's operation order
//   lnx_kernel_spectrum    K = fftn(fftshift(centre-padded kernel)) (leniax/kernels.py:145-149, leniax/utils.py:231-263) as an exact
//                          separable DFT over the kernel's small support: fp64 accumulation, fp64 twiddles, one rounding to fp32.
//                          2 (2-D) or 3 (3-D) launches for the whole batch; any world size (no power-of-two restriction).
//   lnx_random_uniform     counter-based uniform [0, 1) numbers (SplitMix64 of seed + index)
//   lnx_init_perlin        leniax/perlin.py:16-71 + leniax/initializations.py:56-75 (+ loader.make_array_compressible), one CTA per world
//   lnx_init_uniform       leniax/initializations.py:24-30
#include <cmath>
#include <cstdio>

#include "lnx_internal.h"

namespace lnx {
namespace setup {

constexpr float KEPS = 1e-7f;  // leniax/constant.py:7

// ---------------------------------------------------------------------------------------------------------------------
// rasterisation
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }  // (no FMA contraction: XLA evaluates op by op)
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ float kernel_fn(const lnx_kernel_spec& s, float X) {
    const float p0 = s.kf_params[0], p1 = s.kf_params[1];
    switch (s.kf) {
        case LNX_KF_POLY_QUAD: {  // (4 X (1 - X)) ** q
            const float b = fmul(fmul(4.f, X), fsub(1.f, X));
            const float q = p0;
            if (q == floorf(q) && q >= 0.f && q <= 64.f) {  // integer_pow: exponentiation by squaring (lax.integer_pow)
                int e = (int)q;
                float acc = 1.f, base = b;
                bool first = true;
                while (e) {
                    if (e & 1) {
                        acc = first ? base : fmul(acc, base);
                        first = false;
                    }
                    e >>= 1;
                    if (e) base = fmul(base, base);
                }
                return acc;
            }
            return powf(b, q);
        }
        case LNX_KF_GAUSS_BUMP:  // exp(q (q - 1 / (X (1 - X) + eps)))
            return expf(fmul(p0, fsub(p0, fdiv(1.f, fadd(fmul(X, fsub(1.f, X)), KEPS)))));
        case LNX_KF_STEP:
            return (X >= p0 && X <= fsub(1.f, p0)) ? 1.f : 0.f;
        case LNX_KF_GAUSS: {  // exp(-((X - q) / (0.3 q))^2 / 2)
            const float t = fdiv(fsub(X, p0), fmul(0.3f, p0));
            return expf(fdiv(-fmul(t, t), 2.f));
        }
        case LNX_KF_THRESHOLD:
            return X >= p0 ? 1.f : 0.f;
        case LNX_KF_STAIRCASE: {
            const float m = p0, sg = p1, h = fdiv(sg, 2.f);
            float o = (X >= fsub(m, sg) && X < fsub(m, h)) ? 0.5f : 0.f;
            o = fadd(o, (X >= fsub(m, h) && X <= fadd(m, h)) ? 1.f : 0.f);
            return fadd(o, (X > fadd(m, h) && X <= fadd(m, sg)) ? 0.5f : 0.f);
        }
        default: {  // LNX_KF_TRIANGLE
            const float m = p0, sg = p1, left = fsub(m, sg), right = fadd(m, sg);
            float o = (X >= left && X < m) ? fdiv(fsub(X, left), fsub(m, left)) : 0.f;
            return fadd(o, (X >= m && X <= right) ? fdiv(fsub(X, right), fsub(m, right)) : 0.f);
        }
    }
}

__device__ float block_sum_double(double v, double* red) {  // deterministic: fixed tree over the 256 threads
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double t = red[0];
    __syncthreads();
    return (float)t;
}

// one CTA per kernel; out [n][side][side], the kernel of radius k embedded at offset side / 2 - k (zero elsewhere)
__global__ void __launch_bounds__(256) rasterize_kernel(const lnx_kernel_spec* specs, float R, int side, float* out) {
    __shared__ double red[256];
    const lnx_kernel_spec s = specs[blockIdx.x];
    float* img = out + (size_t)blockIdx.x * side * side;
    const int tid = threadIdx.x;
    for (int i = tid; i < side * side; i += 256) img[i] = 0.f;
    if (s.shape == LNX_KSHAPE_EMPTY) return;
    const float rR = (float)((double)s.r * (double)R);
    const int k = (int)ceil((double)s.r * (double)R), n = 2 * k, off = side / 2 - k;
    const bool ell = s.shape != LNX_KSHAPE_CIRCLE_2D;
    double acc = 0.0;
    __syncthreads();
    for (int p = tid; p < n * n; p += 256) {
        const int i = p / n, j = p - i * n;
        const float c0 = fdiv((float)(i - k), rR), c1 = fdiv((float)(j - k), rR);
        float dist, rc0 = 0.f;
        if (!ell) {
            dist = sqrtf(fadd(fmul(c0, c0), fmul(c1, c1)));
        } else {  // kernels.py:236-240: rotated coordinates, anisotropic distance
            rc0 = fadd(fmul(c0, s.cos_theta), fmul(c1, s.sin_theta));
            const float rc1 = fadd(fmul(-c0, s.sin_theta), fmul(c1, s.cos_theta));
            const float u = fdiv(rc0, s.a), v = fdiv(rc1, s.b);
            dist = sqrtf(fadd(fmul(u, u), fmul(v, v)));
        }
        const float B = fmul((float)s.nb_b, dist);
        int ring = (int)floorf(B);
        ring = ring < s.nb_b - 1 ? ring : s.nb_b - 1;
        const float X = fsub(B, floorf(B));  // B % 1, B >= 0
        float val = fmul(fmul(dist < 1.f ? 1.f : 0.f, kernel_fn(s, X)), s.bs[ring]);
        if (s.shape == LNX_KSHAPE_ORIENTED_ELLIPSE_2D) val = fmul(val, rc0);  // kernels.py:302
        img[(off + i) * side + off + j] = val;
        acc += s.shape == LNX_KSHAPE_ORIENTED_ELLIPSE_2D ? fabs((double)val) : (double)val;
    }
    const float total = block_sum_double(acc, red);  // kernel.sum() / |kernel|.sum(): fp64 accumulation, one rounding
    for (int p = tid; p < n * n; p += 256) {
        const int i = p / n, j = p - i * n;
        float v = fdiv(img[(off + i) * side + off + j], total);
        if (s.shape == LNX_KSHAPE_ELLIPSE_2D) {  // kernels.py:257-259: sign-like gradient along the rotated first axis
            const float c0 = fdiv((float)(i - k), rR), c1 = fdiv((float)(j - k), rR);
            const float rc0 = fadd(fmul(c0, s.cos_theta), fmul(c1, s.sin_theta));
            float g = rc0 < -0.01f ? -1.f : rc0;
            g = g > 0.01f ? 1.f : g;
            v = fmul(v, g);
        }
        img[(off + i) * side + off + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// exact spectrum of small-support kernels
// ---------------------------------------------------------------------------------------------------------------------
__global__ void twiddle_kernel(double2* tw, int N) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N) {
        double s, c;
        sincospi(-2.0 * (double)t / (double)N, &s, &c);  // exp(-2 pi i t / N)
        tw[t] = make_double2(c, s);
    }
}

// One separable DFT stage along an axis of length N whose n_in non-zero samples sit at positions (pos0 + i) mod N:
//   out[img][a][u][b] = sum_i in[img][a][i][b] * exp(-2 pi i u (pos0 + i) / N),  u < N.
// FIRST: input is the real fp32 kernel; LAST: output is rounded to complex64 (the reference's K dtype).
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(256) dft_stage_kernel(const void* in_, void* out_, int A, int n_in, int B, int N, int pos0, const double2* __restrict__ tw) {
    const size_t per_img_out = (size_t)A * N * B, per_img_in = (size_t)A * n_in * B;
    const size_t o = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (o >= per_img_out) return;
    const int b = (int)(o % B);
    const int u = (int)((o / B) % N);
    const int a = (int)(o / ((size_t)B * N));
    const size_t in_base = (size_t)blockIdx.y * per_img_in + (size_t)a * n_in * B + b;
    double re = 0.0, im = 0.0;
    int idx = (int)(((long long)u * (pos0 % N)) % N);  // u * position mod N, advanced by u per sample
    for (int i = 0; i < n_in; ++i) {
        const double2 w = tw[idx];
        if constexpr (FIRST) {
            const double x = (double)static_cast<const float*>(in_)[in_base + (size_t)i * B];
            re = fma(x, w.x, re);
            im = fma(x, w.y, im);
        } else {
            const double2 x = static_cast<const double2*>(in_)[in_base + (size_t)i * B];
            re = fma(x.x, w.x, fma(-x.y, w.y, re));
            im = fma(x.x, w.y, fma(x.y, w.x, im));
        }
        idx += u;
        if (idx >= N) idx -= N;
    }
    const size_t oi = (size_t)blockIdx.y * per_img_out + o;
    if constexpr (LAST)
        static_cast<float2*>(out_)[oi] = make_float2((float)re, (float)im);
    else
        static_cast<double2*>(out_)[oi] = make_double2(re, im);
}

// ---------------------------------------------------------------------------------------------------------------------
// random numbers and initial states
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t index) {  // 24 random mantissa bits: [0, 1)
    return (float)(splitmix64(seed + index * 0x9E3779B97F4A7C15ull) >> 40) * (1.0f / 16777216.0f);
}
__global__ void random_uniform_kernel(uint64_t seed, long long n, float* out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = uniform01(seed, (uint64_t)i);
}
__device__ __forceinline__ float quantize(float c) {  // loader.make_array_compressible (leniax/loader.py:16-30)
    const float q = 12543.f;  // NB_CHARS^2 - 1 = 112^2 - 1
    return fdiv((float)(int)rintf(fmul(c, q)), q);  // jnp.round: half to even
}
// cells[w][...] = quantize(u * maxvals[w])   (initializations.py:26-30: uniform(minval=0, maxval=maxvals))
__global__ void init_uniform_kernel(uint64_t seed, long long cells_per_world, const float* __restrict__ maxvals, float* out) {
    const long long w = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells_per_world; i += (long long)gridDim.x * blockDim.x) {
        const float u = uniform01(seed, (uint64_t)(w * cells_per_world + i));
        out[w * cells_per_world + i] = quantize(fmul(u, maxvals[w]));
    }
}

struct PerlinGeom {
    int H, W, res0, res1, d0, d1, diff0, diff1;
    float delta0, delta1;
};
__device__ __forceinline__ float interpolant(float t) {  // perlin.py:12-13
    return fmul(fmul(fmul(t, t), t), fadd(fmul(t, fsub(fmul(t, 6.f), 15.f)), 10.f));
}
__device__ __forceinline__ float perlin_value(const PerlinGeom& g, const float2* grad /* [res0][res1] (cos, sin) */, int y, int x) {
    // repeated wrap-padded gradient image G[yy][xx] = grad[(yy / d0) % res0][(xx / d1) % res1]; the four corners are the four
    // croppings of it (perlin.py:51-55)
    const int y0 = (y / g.d0) % g.res0, y1 = ((y + g.diff0) / g.d0) % g.res0;
    const int x0 = (x / g.d1) % g.res1, x1 = ((x + g.diff1) / g.d1) % g.res1;
    const float2 g00 = grad[y0 * g.res1 + x0], g10 = grad[y1 * g.res1 + x0], g01 = grad[y0 * g.res1 + x1], g11 = grad[y1 * g.res1 + x1];
    const float gy = fmul((float)y, g.delta0), gx = fmul((float)x, g.delta1);
    const float f0 = fsub(gy, floorf(gy)), f1 = fsub(gx, floorf(gx));  // % 1
    const float n00 = fadd(fmul(f0, g00.x), fmul(f1, g00.y));
    const float n10 = fadd(fmul(fsub(f0, 1.f), g10.x), fmul(f1, g10.y));
    const float n01 = fadd(fmul(f0, g01.x), fmul(fsub(f1, 1.f), g01.y));
    const float n11 = fadd(fmul(fsub(f0, 1.f), g11.x), fmul(fsub(f1, 1.f), g11.y));
    const float t0 = interpolant(f0), t1 = interpolant(f1);
    const float n0 = fadd(fmul(n00, fsub(1.f, t0)), fmul(t0, n10));
    const float n1 = fadd(fmul(n01, fsub(1.f, t0)), fmul(t0, n11));
    return fmul(1.41421356237309515f, fadd(fmul(fsub(1.f, t1), n0), fmul(t1, n1)));
}
// one CTA per world: noise, (x - min) / (max - min), * scaling, quantise.  raw != nullptr: write the plain noise there instead.
// angles == nullptr: the angles of world w = (individual s, initialisation i) are drawn here, 2 pi u(seeds[s], i * res0 * res1 + j)
// (the numbers lnx_random_uniform gives the [nb_init][res0][res1] array of that individual)
__global__ void __launch_bounds__(256) init_perlin_kernel(const float* __restrict__ angles, const uint64_t* __restrict__ seeds, int nb_init, PerlinGeom g,
                                                          const float* __restrict__ scaling, float* out, float* raw) {
    extern __shared__ float2 grad[];
    __shared__ float red[2][256];
    const int w = blockIdx.x, tid = threadIdx.x;
    const int nres = g.res0 * g.res1;
    for (int i = tid; i < nres; i += 256) {
        const float a = angles ? angles[(size_t)w * nres + i] : fmul(6.28318530717958648f, uniform01(seeds[w / nb_init], (uint64_t)(w % nb_init) * nres + i));
        grad[i] = make_float2(cosf(a), sinf(a));
    }
    __syncthreads();
    const int n = g.H * g.W;
    if (raw) {
        for (int p = tid; p < n; p += 256) raw[(size_t)w * n + p] = perlin_value(g, grad, p / g.W, p % g.W);
        return;
    }
    float mn = INFINITY, mx = -INFINITY;
    for (int p = tid; p < n; p += 256) {
        const float v = perlin_value(g, grad, p / g.W, p % g.W);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    red[0][tid] = mn;
    red[1][tid] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) {
            red[0][tid] = fminf(red[0][tid], red[0][tid + s]);
            red[1][tid] = fmaxf(red[1][tid], red[1][tid + s]);
        }
        __syncthreads();
    }
    mn = red[0][0];
    const float range = fsub(red[1][0], mn);  // max of (cells - min) (initializations.py:70-71)
    const float sc = scaling[w];
    for (int p = tid; p < n; p += 256) {
        const float v = perlin_value(g, grad, p / g.W, p % g.W);
        out[(size_t)w * n + p] = quantize(fmul(fdiv(fsub(v, mn), range), sc));
    }
}

}  // namespace setup
}  // namespace lnx

using namespace lnx::setup;

extern "C" {

int lnx_rasterize_kernels(int32_t n, const lnx_kernel_spec* specs, float R, int32_t side, float* out, void* stream) {
    if (n < 1 || !specs || !out || side < 2 || (side & 1) || !(R > 0.f)) return lnx_fail(LNX_ERR_INVALID, "lnx_rasterize_kernels: bad argument");
    for (int i = 0; i < n; ++i) {
        const lnx_kernel_spec& s = specs[i];
        if (s.shape == LNX_KSHAPE_EMPTY) continue;
        if (s.shape < LNX_KSHAPE_CIRCLE_2D || s.shape > LNX_KSHAPE_ORIENTED_ELLIPSE_2D)
            return lnx_fail(LNX_ERR_UNSUPPORTED, "lnx_rasterize_kernels: kernel %d: unknown shape %d", i, s.shape);
        if (s.kf < LNX_KF_POLY_QUAD || s.kf > LNX_KF_TRIANGLE) return lnx_fail(LNX_ERR_UNSUPPORTED, "lnx_rasterize_kernels: kernel %d: unknown kernel function %d", i, s.kf);
        if (s.nb_b < 1 || s.nb_b > LNX_MAX_RINGS) return lnx_fail(LNX_ERR_INVALID, "lnx_rasterize_kernels: kernel %d: nb_b must be in [1, %d]", i, LNX_MAX_RINGS);
        if (!(s.r > 0.f) || 2 * (int)std::ceil((double)s.r * (double)R) > side)
            return lnx_fail(LNX_ERR_INVALID, "lnx_rasterize_kernels: kernel %d: radius %d px does not fit side %d", i, (int)std::ceil((double)s.r * (double)R), side);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    lnx_kernel_spec* d = nullptr;
    LNX_CUDA(cudaMallocAsync(&d, sizeof(lnx_kernel_spec) * n, st));
    LNX_CUDA(cudaMemcpyAsync(d, specs, sizeof(lnx_kernel_spec) * n, cudaMemcpyHostToDevice, st));
    rasterize_kernel<<<n, 256, 0, st>>>(d, R, side, out);
    LNX_CUDA(cudaGetLastError());
    LNX_CUDA(cudaFreeAsync(d, st));
    return LNX_OK;
}

int lnx_kernel_spectrum(int32_t nb_dims, const int32_t* dims, int32_t n, const int32_t* support, const float* spatial, void* K_out, void* stream) {
    if (!dims || !support || !spatial || !K_out || n < 1 || nb_dims < 1 || nb_dims > 3) return lnx_fail(LNX_ERR_INVALID, "lnx_kernel_spectrum: bad argument");
    int D[3] = {1, 1, 1}, S[3] = {1, 1, 1};
    for (int d = 0; d < nb_dims; ++d) {  // right-aligned: a 2-D world is [1][H][W]
        D[3 - nb_dims + d] = dims[d];
        S[3 - nb_dims + d] = support[d];
        if (dims[d] < 1 || support[d] < 1 || support[d] > dims[d]) return lnx_fail(LNX_ERR_INVALID, "lnx_kernel_spectrum: support %d does not fit dimension %d", support[d], dims[d]);
    }
    if (n > 65535) return lnx_fail(LNX_ERR_INVALID, "lnx_kernel_spectrum: at most 65535 kernels per call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // stages run over the last axis first; the array between stages is [S0][S1 or D1][D2] ... in fp64
    const size_t tmp1 = (size_t)n * S[0] * S[1] * D[2], tmp2 = (size_t)n * S[0] * D[1] * D[2];
    double2 *t1 = nullptr, *t2 = nullptr, *tw = nullptr;
    int maxN = D[0] > D[1] ? D[0] : D[1];
    maxN = maxN > D[2] ? maxN : D[2];
    LNX_CUDA(cudaMallocAsync(&tw, sizeof(double2) * maxN, st));
    LNX_CUDA(cudaMallocAsync(&t1, sizeof(double2) * tmp1, st));
    LNX_CUDA(cudaMallocAsync(&t2, sizeof(double2) * tmp2, st));
    auto pos0 = [](int N, int s) { return ((N - s) / 2 + N / 2) % N; };  // centre padding (utils.py:231-263) then fftshift (kernels.py:147)
    auto blocks = [](size_t outputs) { return (unsigned)((outputs + 255) / 256); };
    int built = 0;
    auto twiddles = [&](int N) {
        if (built != N) twiddle_kernel<<<(N + 255) / 256, 256, 0, st>>>(tw, N);
        built = N;
    };
    // axis 2: [S0 * S1][S2] real -> [S0 * S1][D2]
    const bool only = nb_dims == 1;
    twiddles(D[2]);
    if (only)
        dft_stage_kernel<true, true><<<dim3(blocks((size_t)D[2]), n), 256, 0, st>>>(spatial, K_out, 1, S[2], 1, D[2], pos0(D[2], S[2]), tw);
    else
        dft_stage_kernel<true, false><<<dim3(blocks((size_t)S[0] * S[1] * D[2]), n), 256, 0, st>>>(spatial, t1, S[0] * S[1], S[2], 1, D[2], pos0(D[2], S[2]), tw);
    if (nb_dims == 2) {  // axis 1 (the leading world axis): [S1][D2] -> [D1][D2]
        twiddles(D[1]);
        dft_stage_kernel<false, true><<<dim3(blocks((size_t)D[1] * D[2]), n), 256, 0, st>>>(t1, K_out, 1, S[1], D[2], D[1], pos0(D[1], S[1]), tw);
    } else if (nb_dims == 3) {
        twiddles(D[1]);
        dft_stage_kernel<false, false><<<dim3(blocks((size_t)S[0] * D[1] * D[2]), n), 256, 0, st>>>(t1, t2, S[0], S[1], D[2], D[1], pos0(D[1], S[1]), tw);
        twiddles(D[0]);
        dft_stage_kernel<false, true><<<dim3(blocks((size_t)D[0] * D[1] * D[2]), n), 256, 0, st>>>(t2, K_out, 1, S[0], D[1] * D[2], D[0], pos0(D[0], S[0]), tw);
    }
    LNX_CUDA(cudaGetLastError());
    LNX_CUDA(cudaFreeAsync(t1, st));
    LNX_CUDA(cudaFreeAsync(t2, st));
    LNX_CUDA(cudaFreeAsync(tw, st));
    return LNX_OK;
}

int lnx_random_uniform(uint64_t seed, int64_t n, float* out, void* stream) {
    if (n < 1 || !out) return lnx_fail(LNX_ERR_INVALID, "lnx_random_uniform: bad argument");
    const long long blocks = (n + 255) / 256;
    random_uniform_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, n, out);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_init_uniform(uint64_t seed, int32_t n_worlds, int64_t cells_per_world, const float* maxvals, float* out, void* stream) {
    if (n_worlds < 1 || n_worlds > 65535 || cells_per_world < 1 || !maxvals || !out) return lnx_fail(LNX_ERR_INVALID, "lnx_init_uniform: bad argument");
    const long long blocks = (cells_per_world + 255) / 256;
    init_uniform_kernel<<<dim3((unsigned)(blocks < 256 ? blocks : 256), n_worlds), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, cells_per_world, maxvals, out);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

static int perlin_launch(int n_worlds, int H, int W, int res0, int res1, const float* angles, const uint64_t* seeds_dev, int nb_init, const float* scaling,
                         float* out, float* noise_out, cudaStream_t st) {
    PerlinGeom g;
    g.H = H, g.W = W, g.res0 = res0, g.res1 = res1;
    g.d0 = H / res0, g.d1 = W / res1;                                   // perlin.py:48
    g.diff0 = (res0 + 1) * g.d0 - H, g.diff1 = (res1 + 1) * g.d1 - W;   // :51
    g.delta0 = (float)((double)res0 / (double)H), g.delta1 = (float)((double)res1 / (double)W);  // :58
    if (g.diff0 < 1 || g.diff1 < 1) return lnx_fail(LNX_ERR_INVALID, "lnx_init_perlin: degenerate resolution");
    const size_t smem = sizeof(float2) * res0 * res1;
    if (smem > 48 * 1024) return lnx_fail(LNX_ERR_UNSUPPORTED, "lnx_init_perlin: resolution %d x %d too large", res0, res1);
    init_perlin_kernel<<<n_worlds, 256, smem, st>>>(angles, seeds_dev, nb_init, g, scaling, out, noise_out);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int lnx_init_perlin(int32_t n_worlds, int32_t H, int32_t W, int32_t res0, int32_t res1, const float* angles, const float* scaling, float* out,
                    float* noise_out, void* stream) {
    if (n_worlds < 1 || H < 1 || W < 1 || res0 < 1 || res1 < 1 || res0 > H || res1 > W || !angles || (!noise_out && (!scaling || !out)))
        return lnx_fail(LNX_ERR_INVALID, "lnx_init_perlin: bad argument");
    return perlin_launch(n_worlds, H, W, res0, res1, angles, nullptr, 1, scaling, out, noise_out, static_cast<cudaStream_t>(stream));
}

int lnx_init_perlin_seeded(int32_t n_seeds, const uint64_t* seeds, int32_t nb_init, int32_t H, int32_t W, int32_t res0, int32_t res1,
                           const float* scaling, float* out, void* stream) {
    if (n_seeds < 1 || nb_init < 1 || !seeds || H < 1 || W < 1 || res0 < 1 || res1 < 1 || res0 > H || res1 > W || !scaling || !out)
        return lnx_fail(LNX_ERR_INVALID, "lnx_init_perlin_seeded: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint64_t* d = nullptr;
    LNX_CUDA(cudaMallocAsync(&d, sizeof(uint64_t) * n_seeds, st));
    LNX_CUDA(cudaMemcpyAsync(d, seeds, sizeof(uint64_t) * n_seeds, cudaMemcpyHostToDevice, st));
    const int rc = perlin_launch(n_seeds * nb_init, H, W, res0, res1, nullptr, d, nb_init, scaling, out, nullptr, st);
    cudaFreeAsync(d, st);
    return rc;
}

}  // extern "C"
