// pb_math.h — deterministic fp32 building blocks shared by the sm_100a kernels and by
// the CPU checker in oracle/.
//
// Why this exists: every discrete output of the path (NMS bits, gate bits, auction
// argmax, track ids) is decided by comparing fp32 values.  To make those decisions
// bit-identical on the GPU and on the host, both sides evaluate every expression with
// the same IEEE-754 operations in the same order:
//   * device code is built with  nvcc --fmad=false  (no silent mul+add contraction),
//     host code with  g++ -ffp-contract=off ;
//   * division and sqrt are the correctly rounded IEEE ones on both sides
//     (nvcc defaults --prec-div=true --prec-sqrt=true, SSE divss/sqrtss);
//   * a fused multiply-add is used only where written explicitly (pb_fma), which is a
//     single correctly rounded operation on both sides;
//   * expf is replaced by pb_expf below (no libm / libdevice call), min/max by
//     pb_min/pb_max with one fixed NaN / signed-zero behaviour.
//
// The reference (naveedprojects/yolo-pose-cpp) calls CUDA's expf/fminf/fmaxf with FMA
// contraction enabled (src/cuda/gpu_postprocess.cu:156, gpu_tracker.cu:417,482).
// pb_expf is within 2 ulp of a correctly rounded exp (tests/test_pb_math.py), i.e. far
// inside the 1e-4 relative tolerance the parity contract states for floating point.
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD static inline
#endif

// One correctly rounded a*b+c.
PB_HD float pb_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

// max/min with a fixed rule: the first argument wins unless the second compares
// strictly greater/less.  NaN in b never wins; NaN in a is kept.  (Inputs on the path
// are finite; this only pins the behaviour so that both sides agree.)
PB_HD float pb_max(float a, float b) { return (b > a) ? b : a; }
PB_HD float pb_min(float a, float b) { return (b < a) ? b : a; }

PB_HD uint32_t pb_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PB_HD float pb_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// exp(x) in fp32.  Range reduction x = n*ln2 + r with the 1.5*2^23 rounding trick,
// degree-5 polynomial in r (Cephes expf coefficients), exponent assembled with integer
// arithmetic.  Inputs below -86 return 0 (exp(-86) ~ 4.5e-38: the reference would
// produce a value below 1e-37 there; every use on the path sums such terms with O(1)
// ones or compares them with thresholds >= 0.15).  Inputs above 88 are clamped.
PB_HD float pb_expf(float x) {
    if (!(x >= -86.0f)) return (x != x) ? x : 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float kMagic = 12582912.0f;                 // 1.5 * 2^23
    float t = pb_fma(x, 1.44269504088896341f, kMagic);
    float n = t - kMagic;                             // nearest integer to x*log2(e)
    float r = pb_fma(n, -0.693359375f, x);            // ln2 high part (exact product)
    r = pb_fma(n, 2.12194440e-4f, r);                 // ln2 low part
    float p = 1.9875691500e-4f;
    p = pb_fma(p, r, 1.3981999507e-3f);
    p = pb_fma(p, r, 8.3334519073e-3f);
    p = pb_fma(p, r, 4.1665795894e-2f);
    p = pb_fma(p, r, 1.6666665459e-1f);
    p = pb_fma(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = pb_fma(p, r2, r) + 1.0f;                // in (0.70, 1.42)
    int32_t ni = (int32_t)n;                          // -125 .. 127
    return pb_u2f(pb_f2u(e) + ((uint32_t)ni << 23));
}
