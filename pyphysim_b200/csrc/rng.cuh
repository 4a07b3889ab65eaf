// rng.cuh — the shared counter-based random stream (device side).
// Contract (identical to oracle/philox.py):
//   key     = (seed lo32, seed hi32)
//   counter = (slot lo32, (stream << 16) | slot hi16, unit lo32, unit hi32),  slot = word / 4
//   word w of (stream, unit) = lane w % 4 of that Philox4x32-10 block
// Streams: 0 data symbols, 1 channel (Rayleigh H / Jakes phases), 2 noise.
#pragma once
#include "common.cuh"

namespace b200phy {

enum { STREAM_DATA = 0, STREAM_CHANNEL = 1, STREAM_NOISE = 2 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// The ten round keys of a seed, computed once on the host and passed as a kernel parameter: in the kernels
// they are constant-bank operands of the round's LOP3, so a block costs 2 IMAD.WIDE + 2 LOP3 per round (the
// per-thread key schedule was 18 more IADD3 per block, a third of the generator's instructions).
struct PhiloxKey {
    uint32_t k[20];
    PhiloxKey() = default;
    __host__ __device__ PhiloxKey(uint64_t seed) {
        uint32_t a = uint32_t(seed), b = uint32_t(seed >> 32);
        for (int i = 0; i < 10; ++i) {
            k[2 * i] = a;
            k[2 * i + 1] = b;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
    }
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKey &key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.k[2 * i], lo1, hi0 ^ c.w ^ key.k[2 * i + 1], lo0);
    }
    return c;
}

__device__ __forceinline__ uint4 rng_block(const PhiloxKey &key, uint32_t stream, uint64_t unit, uint64_t slot) {
    const uint4 c = make_uint4(uint32_t(slot), (stream << 16) | (uint32_t(slot >> 32) & 0xffffu),
                               uint32_t(unit), uint32_t(unit >> 32));
    return philox4x32_10(c, key);
}

__device__ __forceinline__ uint4 rng_block(uint64_t seed, uint32_t stream, uint64_t unit, uint64_t slot) {
    const uint4 c = make_uint4(uint32_t(slot), (stream << 16) | (uint32_t(slot >> 32) & 0xffffu),
                               uint32_t(unit), uint32_t(unit >> 32));
    return philox4x32_10(c, make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
}

__device__ __forceinline__ uint32_t lane_of(const uint4 &b, int l) {
    return l == 0 ? b.x : (l == 1 ? b.y : (l == 2 ? b.z : b.w));
}

// u(x) = (x + 0.5) * 2^-32.  f32: float(x) * 2^-32 + 2^-33 (the multiply is exact, so the fused and
// unfused forms round identically and the host float32 evaluation is bit-identical).
template <typename T> __device__ __forceinline__ T uniform01(uint32_t x);
template <> __device__ __forceinline__ float uniform01<float>(uint32_t x) {
    return __fmaf_rn(__uint2float_rn(x), 0x1p-32f, 0x1p-33f);
}
template <> __device__ __forceinline__ double uniform01<double>(uint32_t x) {
    return (double(x) + 0.5) * 0x1p-32;
}

// sin(pi x), cos(pi x) for x = k/2 + f with k an integer-valued float (|k| < 2^22) and |f| <= 0.2515:
// minimax polynomials in f^2 on the reduced interval (approximation error 1e-8 / 6e-10, 1.5 ulp after
// float evaluation - the accuracy of CUDA's sincospif at half its instruction count: no special values,
// no large-argument path, and the caller's own reduction supplies k and f).
__device__ __forceinline__ void sincospi_kf(float k, float f, float *s, float *c) {
    const int q = __float2int_rn(k);
    const float f2 = f * f;
    float p = fmaf(-5.915239453e-01f, f2, 2.549980879e+00f);
    p = fmaf(p, f2, -5.167712212e+00f);
    const float sn = fmaf(f2 * f, p, f * 3.14159274f);
    float g = fmaf(2.310452312e-01f, f2, -1.335041642e+00f);
    g = fmaf(g, f2, 4.058708668e+00f);
    g = fmaf(g, f2, -4.934802055e+00f);
    const float cs = fmaf(g, f2, 1.0f);
    const bool sw = q & 1;                       // odd quadrant: sin <-> cos
    const float ss = sw ? cs : sn, cc = sw ? sn : cs;
    *s = __uint_as_float(__float_as_uint(ss) ^ ((unsigned(q) << 30) & 0x80000000u));
    *c = __uint_as_float(__float_as_uint(cc) ^ ((unsigned(q + 1) << 30) & 0x80000000u));
}

// u in (0, 1): 2 u = k / 2 + f with k = rint(4 u) (both steps exact)
__device__ __forceinline__ void sincos2pi(float u, float *s, float *c) {
    const float x = u + u, k = rintf(x + x);
    sincospi_kf(k, fmaf(k, -0.5f, x), s, c);
}
__device__ __forceinline__ void sincos2pi(double u, double *s, double *c) { sincospi(2.0 * u, s, c); }

// sqrt(-ln u) for the Box-Muller radius.  u = (x + 0.5) 2^-32 rounded to float lies in [2^-33, 1]: positive, finite,
// never subnormal, so the library routines' special-case handling (zero, negative, infinity, subnormal scaling: 5 of
// logf's 22 instructions, and sqrtf's slow-path branch) is dead weight in a generator that runs twice per noise
// sample.  Float: ln u = e ln 2 + log1p(m - 1) with m in [2/3, 4/3) (minimax polynomial, < 1 ulp), then
// sqrt.approx (MUFU.SQRT, 1 ulp).  Double: the library routines.
__device__ __forceinline__ float boxmuller_radius(float u) {
    const int i = __float_as_int(u);
    const int e = (i - 0x3f2aaaab) & int(0xff800000);          // exponent that puts the mantissa in [2/3, 4/3)
    const float m = __int_as_float(i - e) - 1.0f, s = m * m;
    float r = -0.130310059f, t = 0.140869141f;
    r = fmaf(r, s, -0.121484190f);
    t = fmaf(t, s, 0.139814854f);
    r = fmaf(r, s, -0.166846052f);
    t = fmaf(t, s, 0.200120345f);
    r = fmaf(r, s, -0.249996200f);
    r = fmaf(t, m, r);
    r = fmaf(r, m, 0.333331972f);
    r = fmaf(r, m, -0.5f);
    r = fmaf(r, s, m);                                          // log1p(m)
    const float nl = -fmaf(float(e), 8.26295829e-8f, r);       // -(e / 2^23) ln 2 - log1p(m): e is a multiple of 2^23
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(fmaxf(nl, 0.0f)));
    return rad;
}
__device__ __forceinline__ double boxmuller_radius(double u) { return sqrt(-log(u)); }

// complex normal with E|c|^2 = 1 (randn_c, util/misc.py:327-355) by Box-Muller on two words
template <typename T> __device__ __forceinline__ cx<T> cnormal(uint32_t w0, uint32_t w1) {
    const T rad = boxmuller_radius(uniform01<T>(w0));
    T s, c;
    sincos2pi(uniform01<T>(w1), &s, &c);
    return {rad * c, rad * s};
}

// complex normal number j of (stream, unit): words (2j, 2j+1)
template <typename T, typename K>
__device__ __forceinline__ cx<T> cnormal_at(const K &seed, uint32_t stream, uint64_t unit, uint64_t j) {
    const uint4 b = rng_block(seed, stream, unit, j >> 1);
    return (j & 1) ? cnormal<T>(b.z, b.w) : cnormal<T>(b.x, b.y);
}

// Jakes phase: 2*pi*u (channels/fading_generators.py:413-414); single rounded multiply
template <typename T> __device__ __forceinline__ T phase_from_word(uint32_t x);
template <> __device__ __forceinline__ float phase_from_word<float>(uint32_t x) {
    return __fmul_rn(6.283185307179586f, uniform01<float>(x));
}
template <> __device__ __forceinline__ double phase_from_word<double>(uint32_t x) {
    return __dmul_rn(6.283185307179586, uniform01<double>(x));
}

}  // namespace b200phy
