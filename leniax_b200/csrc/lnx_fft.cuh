// In-register radix-2 FFT building blocks (compile-time indices and twiddles).
//
// Everything here is `__host__ __device__` so that the exact per-thread code of the CUDA kernels can be
// executed thread-by-thread on the CPU by the emulator in tests/ (index maps and hazards are checked without a GPU).
//
// Conventions: forward transform X[k] = sum_n x[n] exp(-2*pi*i*n*k/N) (same sign as numpy / jnp.fft.fftn used by the
// reference, leniax/core.py:81), inverse is un-normalised with exp(+...).  `fft_dif` maps natural order to
// bit-reversed order, `ifft_dit` maps bit-reversed order back to natural order, so a forward/pointwise/inverse chain
// never needs a reordering pass.
#pragma once
#include <cuda_runtime.h>

#define LNX_HD __host__ __device__ __forceinline__
#define LNX_HDC __host__ __device__ constexpr

namespace lnx {

#include "lnx_twiddle128.inc"

struct Tw128 {
    static constexpr float c[128] = LNX_COS128_INIT;
    static constexpr float s[128] = LNX_SIN128_INIT;
};

LNX_HDC int bitrev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
LNX_HDC int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// Complex arithmetic on (re, im) register pairs.  On the device every operation is one of sm_100's packed FP32 instructions
// (FADD2 / FMUL2 / FFMA2: two lane-operations per issue slot; operand swap, broadcast and per-half negation are free
// operand modifiers), which halves the instruction count of the butterflies — the kernels are issue- and instruction-fetch
// bound, not FP32-pipe bound (tools/f32x2_probe.cu).  The host versions (emulator) are the same formulas.
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define LNX_PACKED_F32 1
#else
#define LNX_PACKED_F32 0  // host pass of the emulator
#endif
LNX_HD float2 pk_add(float2 a, float2 b) {
#if LNX_PACKED_F32
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
LNX_HD float2 pk_mul(float2 a, float2 b) {
#if LNX_PACKED_F32
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
LNX_HD float2 pk_fma(float2 a, float2 b, float2 c) {
#if LNX_PACKED_F32
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
LNX_HD float2 pk_bc(float x) { return make_float2(x, x); }            // broadcast (a free operand modifier in SASS)
LNX_HD float2 pk_swap(float2 a) { return make_float2(a.y, a.x); }     // free modifier
LNX_HD float2 pk_neg(float2 a) { return make_float2(-a.x, -a.y); }    // free modifier

LNX_HD float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
LNX_HD float2 csub(float2 a, float2 b) { return pk_add(a, pk_neg(b)); }
// a * b = a * b.x + (a.y, a.x) * (-b.y, b.y)
LNX_HD float2 cmul(float2 a, float2 b) { return pk_fma(pk_swap(a), make_float2(-b.y, b.y), pk_mul(a, pk_bc(b.x))); }
// a * conj(b)
LNX_HD float2 cmulc(float2 a, float2 b) { return pk_fma(pk_swap(a), make_float2(b.y, -b.y), pk_mul(a, pk_bc(b.x))); }
// d * (c - i s): (d.x c + d.y s, d.y c - d.x s);  d * (c + i s): (d.x c - d.y s, d.y c + d.x s)
LNX_HD float2 rot_fwd(float2 d, float c, float s) { return pk_fma(pk_swap(d), make_float2(s, -s), pk_mul(d, pk_bc(c))); }
LNX_HD float2 rot_inv(float2 d, float c, float s) { return pk_fma(pk_swap(d), make_float2(-s, s), pk_mul(d, pk_bc(c))); }

// d * W_N^E with W_N = exp(-2*pi*i/N) (INV=false) or its conjugate (INV=true); E, N compile-time.
template <int E, int N, bool INV>
LNX_HD float2 mul_tw(float2 d) {
    constexpr int e = ((E % N) + N) % N;
    constexpr int idx = e * (128 / N);
    static_assert(128 % N == 0, "N must divide 128");
    constexpr float h = Tw128::c[16];  // sqrt(1/2)
    if constexpr (e == 0) {
        return d;
    } else if constexpr (4 * e == N) {  // -i (fwd) / +i (inv)
        return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    } else if constexpr (2 * e == N) {
        return pk_neg(d);
    } else if constexpr (4 * e == 3 * N) {  // +i (fwd) / -i (inv)
        return INV ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
    } else if constexpr (8 * e == N) {  // (1 - i)/sqrt2 fwd: (x + y, y - x) h;  (1 + i)/sqrt2 inv: (x - y, x + y) h
        return pk_mul(pk_add(d, INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x)), pk_bc(h));
    } else if constexpr (8 * e == 3 * N) {  // (-1 - i)/sqrt2 fwd: (y - x, -(x + y)) h;  (-1 + i)/sqrt2 inv: (-(x + y), x - y) h
        return pk_mul(pk_add(pk_neg(d), INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x)), pk_bc(h));
    } else if constexpr (8 * e == 5 * N) {  // (-1 + i)/sqrt2 fwd: (-(x + y), x - y) h;  inv: (y - x, -(x + y)) h
        return pk_mul(pk_add(pk_neg(d), INV ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x)), pk_bc(h));
    } else if constexpr (8 * e == 7 * N) {  // (1 + i)/sqrt2 fwd: (x - y, x + y) h;  inv: (x + y, y - x) h
        return pk_mul(pk_add(d, INV ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x)), pk_bc(h));
    } else {
        constexpr float c = Tw128::c[idx];
        constexpr float s = Tw128::s[idx];
        return INV ? rot_inv(d, c, s) : rot_fwd(d, c, s);
    }
}

// ---- forward DIF: natural order in, bit-reversed order out; elements at v[0], v[ST], ... ----
template <int N, int ST, int J>
LNX_HD void dif_stage(float2* v) {
    if constexpr (J < N / 2) {
        const float2 a = v[J * ST], b = v[(J + N / 2) * ST];
        v[J * ST] = cadd(a, b);
        v[(J + N / 2) * ST] = mul_tw<J, N, false>(csub(a, b));
        dif_stage<N, ST, J + 1>(v);
    }
}
template <int N, int ST = 1>
LNX_HD void fft_dif(float2* v) {
    if constexpr (N > 1) {
        dif_stage<N, ST, 0>(v);
        fft_dif<N / 2, ST>(v);
        fft_dif<N / 2, ST>(v + (N / 2) * ST);
    }
}

// ---- inverse DIT: bit-reversed order in, natural order out (un-normalised) ----
template <int N, int ST, int J>
LNX_HD void dit_stage(float2* v) {
    if constexpr (J < N / 2) {
        const float2 a = v[J * ST], b = mul_tw<J, N, true>(v[(J + N / 2) * ST]);
        v[J * ST] = cadd(a, b);
        v[(J + N / 2) * ST] = csub(a, b);
        dit_stage<N, ST, J + 1>(v);
    }
}
template <int N, int ST = 1>
LNX_HD void ifft_dit(float2* v) {
    if constexpr (N > 1) {
        ifft_dit<N / 2, ST>(v);
        ifft_dit<N / 2, ST>(v + (N / 2) * ST);
        dit_stage<N, ST, 0>(v);
    }
}

}  // namespace lnx
