/*
 *  oracle/metrics_pinned.h — TEST INFRASTRUCTURE, not product code.
 *
 *  Portable C restatement of the distance arithmetic that `metric_punned_t` resolves to on an
 *  AVX-512 host (reference: /root/reference/include/usearch/index_plugins.hpp:1863-1916 routes
 *  the builtin metrics to SimSIMD). Each function below reproduces the *summation order* of the
 *  named SimSIMD kernel with scalar `fmaf`, so the result does not depend on the ISA of the
 *  machine the oracle runs on (the GPU box's host CPU is not the survey container's CPU).
 *
 *  f32, 16-lane order  — simsimd/include/simsimd/spatial.h:1520-1542 (l2sq_f32_skylake),
 *                        :1587-1615 (cos_f32_skylake), dot.h:1297-1318 (dot_f32_skylake),
 *                        horizontal reduce dot.h:1279-1284 (_simsimd_reduce_f32x16_skylake).
 *  cosine normalisation — spatial.h:1544-1585 uses `rsqrt14_pd` + one Newton step, which no
 *                        other device can reproduce bit-for-bit. The *pinned* definition used
 *                        for label parity is IEEE: 1 - ab * (1/sqrt(a2)) * (1/sqrt(b2)) in f64
 *                        with the same zero rules and the same clamp; it differs from the
 *                        native kernel by at most 1 ULP(f32) after the cast.
 *  ip                   — `1 - dot` (index_plugins.hpp:1916 invoke_simsimd_reverse).
 *  i8                   — exact integer sums (dot.h:1749-1775 dot_i8_ice; spatial.h:1880-1972).
 *  b1                   — exact popcounts (binary.h:92-105, :271-347).
 *
 *  Build with: -O2 -ffp-contract=off (no -ffast-math) so nothing is re-associated or fused
 *  other than the explicit fmaf calls.
 */
#ifndef USEARCH_B200_ORACLE_METRICS_PINNED_H
#define USEARCH_B200_ORACLE_METRICS_PINNED_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- f32, 16 virtual lanes ------------------------------------------------------------- */

static inline float pinned_reduce16_(float const v[16]) {
    /* r_i = (v[i] + v[i+8]) + (v[i+4] + v[i+12]); result = (r0 + r1) + (r2 + r3) */
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = (v[i] + v[i + 8]) + (v[i + 4] + v[i + 12]);
    return (r[0] + r[1]) + (r[2] + r[3]);
}

static inline float pinned_l2sq_f32(float const* a, float const* b, size_t n) {
    float acc[16] = {0};
    for (size_t i = 0; i < n; ++i) { /* tail lanes see zero padding == untouched accumulator */
        float x = a[i] - b[i];
        acc[i & 15] = fmaf(x, x, acc[i & 15]);
    }
    /* masked tail in the reference performs fma(0,0,acc) on the padded lanes: a no-op */
    return pinned_reduce16_(acc);
}

static inline float pinned_dot_f32_(float const* a, float const* b, size_t n) {
    float acc[16] = {0};
    for (size_t i = 0; i < n; ++i) acc[i & 15] = fmaf(a[i], b[i], acc[i & 15]);
    return pinned_reduce16_(acc);
}

static inline float pinned_ip_f32(float const* a, float const* b, size_t n) {
    /* index_plugins.hpp:1916 invoke_simsimd_reverse: `1 - invoke_simsimd()`, and invoke_simsimd
     * already returned a float, so the subtraction is done in f32 */
    return 1.0f - pinned_dot_f32_(a, b, n);
}

static inline float pinned_cos_normalize_f64(double ab, double a2, double b2) {
    if (a2 == 0 && b2 == 0) return 0.f;
    if (ab == 0) return 1.f;
    double ra = 1.0 / sqrt(a2);
    double rb = 1.0 / sqrt(b2);
    double r = 1.0 - (ab * ra) * rb;
    return r > 0 ? (float)r : 0.f;
}

static inline float pinned_cos_f32(float const* a, float const* b, size_t n) {
    float ab = pinned_dot_f32_(a, b, n);
    float a2 = pinned_dot_f32_(a, a, n);
    float b2 = pinned_dot_f32_(b, b, n);
    return pinned_cos_normalize_f64((double)ab, (double)a2, (double)b2);
}

/* ---- f16 / bf16, 8 virtual f32 lanes (Haswell order) ----------------------------------- */
/* spatial.h:1098-1146 (f16), :1152-1200 (bf16); reduce dot.h:857-869 widens to f64:
 * s_i = (double)v[i] + (double)v[i+4]; result = ((s0 + s2) + (s1 + s3)) — see
 * _simsimd_reduce_f64x4_haswell (dot.h:844-855): low pair + high pair, then hadd. */

static inline float pinned_f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { /* subnormal */
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline float pinned_bf16_to_f32(uint16_t h) {
    uint32_t bits = (uint32_t)h << 16;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline double pinned_reduce8_(float const v[8]) {
    double s[4];
    for (int i = 0; i < 4; ++i) s[i] = (double)v[i] + (double)v[i + 4];
    double lo = s[0] + s[2], hi = s[1] + s[3];
    return lo + hi;
}

static inline float pinned_cos_normalize_f32(float ab, float a2, float b2) {
    /* IEEE restatement of spatial.h:1050-1080 (which uses rsqrt_ps + Newton) */
    if (a2 == 0.0f && b2 == 0.0f) return 0.0f;
    if (ab == 0.0f) return 1.0f;
    float ra = 1.0f / sqrtf(a2);
    float rb = 1.0f / sqrtf(b2);
    float r = 1.0f - (ab * ra) * rb;
    return r > 0 ? r : 0.f;
}

#define PINNED_HALF_KERNELS_(name, conv)                                                       \
    static inline float pinned_l2sq_##name(uint16_t const* a, uint16_t const* b, size_t n) {   \
        float acc[8] = {0};                                                                    \
        for (size_t i = 0; i < n; ++i) {                                                       \
            float x = conv(a[i]) - conv(b[i]);                                                 \
            acc[i & 7] = fmaf(x, x, acc[i & 7]);                                               \
        }                                                                                      \
        return (float)pinned_reduce8_(acc);                                                    \
    }                                                                                          \
    static inline float pinned_ip_##name(uint16_t const* a, uint16_t const* b, size_t n) {     \
        float acc[8] = {0};                                                                    \
        for (size_t i = 0; i < n; ++i) acc[i & 7] = fmaf(conv(a[i]), conv(b[i]), acc[i & 7]);  \
        return 1.0f - (float)pinned_reduce8_(acc);                                             \
    }                                                                                          \
    static inline float pinned_cos_##name(uint16_t const* a, uint16_t const* b, size_t n) {    \
        float ab[8] = {0}, a2[8] = {0}, b2[8] = {0};                                           \
        for (size_t i = 0; i < n; ++i) {                                                       \
            float x = conv(a[i]), y = conv(b[i]);                                              \
            ab[i & 7] = fmaf(x, y, ab[i & 7]);                                                 \
            a2[i & 7] = fmaf(x, x, a2[i & 7]);                                                 \
            b2[i & 7] = fmaf(y, y, b2[i & 7]);                                                 \
        }                                                                                      \
        return pinned_cos_normalize_f32((float)pinned_reduce8_(ab), (float)pinned_reduce8_(a2), \
                                        (float)pinned_reduce8_(b2));                           \
    }

PINNED_HALF_KERNELS_(f16, pinned_f16_to_f32)
PINNED_HALF_KERNELS_(bf16, pinned_bf16_to_f32)

/* ---- i8: exact integer sums ------------------------------------------------------------- */

static inline float pinned_ip_i8(int8_t const* a, int8_t const* b, size_t n) {
    int32_t ab = 0;
    for (size_t i = 0; i < n; ++i) ab += (int32_t)a[i] * (int32_t)b[i];
    return 1.0f - (float)ab; /* dot_i8 -> f64 -> cast to f32 -> `1 - x` in f32 (index_plugins.hpp:1914-1916) */
}

static inline float pinned_l2sq_i8(int8_t const* a, int8_t const* b, size_t n) {
    int32_t d2 = 0;
    for (size_t i = 0; i < n; ++i) {
        int32_t x = (int32_t)a[i] - (int32_t)b[i];
        d2 += x * x;
    }
    return (float)d2;
}

static inline float pinned_cos_i8(int8_t const* a, int8_t const* b, size_t n) {
    int32_t ab = 0, a2 = 0, b2 = 0;
    for (size_t i = 0; i < n; ++i) {
        ab += (int32_t)a[i] * (int32_t)b[i];
        a2 += (int32_t)a[i] * (int32_t)a[i];
        b2 += (int32_t)b[i] * (int32_t)b[i];
    }
    /* spatial.h:1904-1972 hands the three i32 sums to the f32 normaliser */
    return pinned_cos_normalize_f32((float)ab, (float)a2, (float)b2);
}

/* ---- b1x8: exact popcounts --------------------------------------------------------------- */

static inline uint32_t pinned_popcount8_(uint8_t x) { return (uint32_t)__builtin_popcount(x); }

static inline float pinned_hamming_b1(uint8_t const* a, uint8_t const* b, size_t n_bytes) {
    uint32_t d = 0;
    for (size_t i = 0; i < n_bytes; ++i) d += pinned_popcount8_(a[i] ^ b[i]);
    return (float)d;
}

static inline float pinned_tanimoto_b1(uint8_t const* a, uint8_t const* b, size_t n_bytes) {
    /* jaccard_b8: 1 - and/or in f64, 1 when the union is empty (binary.h:99-105) */
    uint32_t and_ = 0, or_ = 0;
    for (size_t i = 0; i < n_bytes; ++i) {
        and_ += pinned_popcount8_(a[i] & b[i]);
        or_ += pinned_popcount8_(a[i] | b[i]);
    }
    return or_ ? (float)(1.0 - (double)and_ / (double)or_) : 1.f;
}

static inline float pinned_sorensen_b1(uint8_t const* a, uint8_t const* b, size_t n_bytes) {
    /* index_plugins.hpp:1452-1478 metric_sorensen_gt: 1 - 2*and / (|a| + |b|) in f32 */
    uint32_t and_ = 0, any_ = 0;
    for (size_t i = 0; i < n_bytes; ++i) {
        and_ += pinned_popcount8_(a[i] & b[i]);
        any_ += pinned_popcount8_(a[i]) + pinned_popcount8_(b[i]);
    }
    return 1.f - 2.f * (float)and_ / (float)any_;
}

#ifdef __cplusplus
}
#endif
#endif
