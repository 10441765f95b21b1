/* Host emulation of the WORD variants of the half-precision metrics (usearch_b200/csrc/metrics.cuh,
 * l2sq/ip/cos_halfw_t): 4 lanes, lane s owns accumulators 2s and 2s+1 and walks the 32-bit word s of every
 * 16-byte chunk; the f64 reduce is two xor-shuffles. Must give the bits of the pinned oracle, which keeps all 8
 * accumulators in one place. TEST INFRASTRUCTURE (includes oracle/metrics_pinned.h). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "metrics_pinned.h"

static float reduce_words(float lane_v[4][2]) { /* reduce_words_f64, evaluated as lane 0 sees it */
    double a[4], b[4], a2[4], b2[4];
    for (int s = 0; s < 4; ++s) a[s] = (double)lane_v[s][0], b[s] = (double)lane_v[s][1];
    for (int s = 0; s < 4; ++s) a2[s] = a[s] + a[s ^ 2], b2[s] = b[s] + b[s ^ 2];
    double fa[4], fb[4];
    for (int s = 0; s < 4; ++s) fa[s] = a2[s] + a2[s ^ 1], fb[s] = b2[s] + b2[s ^ 1];
    float out[4];
    for (int s = 0; s < 4; ++s) out[s] = (float)(fa[s] + fb[s]);
    for (int s = 1; s < 4; ++s)
        if (memcmp(&out[s], &out[0], 4) != 0) { printf("lanes disagree\n"); exit(1); }
    return out[0];
}

typedef float (*conv_t)(uint16_t);

static void emulate(uint16_t const* q, uint16_t const* b, size_t n, conv_t conv, float* l2sq, float* dot) {
    size_t chunks = (n + 7) / 8;
    float l[4][2] = {{0}}, d[4][2] = {{0}};
    for (int s = 0; s < 4; ++s)
        for (size_t j = 0; j < chunks; ++j)
            for (int e = 0; e < 2; ++e) {
                size_t i = j * 8 + (size_t)s * 2 + (size_t)e; /* word s of chunk j holds elements 2s, 2s+1 */
                float x = i < n ? conv(q[i]) : conv(0), y = i < n ? conv(b[i]) : conv(0); /* rows are zero-padded */
                float diff = x - y;
                l[s][e] = fmaf(diff, diff, l[s][e]);
                d[s][e] = fmaf(x, y, d[s][e]);
            }
    *l2sq = reduce_words(l);
    *dot = reduce_words(d);
}

static uint16_t random_half(int bf16) {
    float f = ((float)rand() / (float)RAND_MAX - 0.5f) * 4.0f;
    uint32_t x;
    memcpy(&x, &f, 4);
    if (bf16) return (uint16_t)(x >> 16);
    /* crude f32 -> f16 (truncate), any bit pattern of a normal half is a fine test input */
    uint32_t sign = (x >> 16) & 0x8000u, exp = (x >> 23) & 0xFF, mant = x & 0x7FFFFFu;
    int e = (int)exp - 127 + 15;
    if (e <= 0) return (uint16_t)sign;
    if (e >= 31) e = 30;
    return (uint16_t)(sign | ((uint32_t)e << 10) | (mant >> 13));
}

int main(void) {
    srand(7);
    size_t const dims[] = {1, 7, 8, 9, 33, 100, 256, 768, 1000};
    unsigned checked = 0;
    for (int bf16 = 0; bf16 < 2; ++bf16)
        for (size_t di = 0; di < sizeof(dims) / sizeof(dims[0]); ++di)
            for (int rep = 0; rep < 50; ++rep) {
                size_t n = dims[di];
                uint16_t q[1000], b[1000];
                for (size_t i = 0; i < n; ++i) q[i] = random_half(bf16), b[i] = random_half(bf16);
                float l2, dot, q2a, q2b, b2a, b2b;
                conv_t conv = bf16 ? pinned_bf16_to_f32 : pinned_f16_to_f32;
                emulate(q, b, n, conv, &l2, &dot);
                emulate(q, q, n, conv, &q2a, &q2b);
                emulate(b, b, n, conv, &b2a, &b2b);
                float want_l2 = bf16 ? pinned_l2sq_bf16(q, b, n) : pinned_l2sq_f16(q, b, n);
                float want_ip = bf16 ? pinned_ip_bf16(q, b, n) : pinned_ip_f16(q, b, n);
                float want_cos = bf16 ? pinned_cos_bf16(q, b, n) : pinned_cos_f16(q, b, n);
                float got_ip = 1.0f - dot;
                float got_cos = pinned_cos_normalize_f32(dot, q2b, b2b);
                if (memcmp(&l2, &want_l2, 4) || memcmp(&got_ip, &want_ip, 4) || memcmp(&got_cos, &want_cos, 4)) {
                    printf("MISMATCH bf16=%d n=%zu: l2 %a vs %a, ip %a vs %a, cos %a vs %a\n", bf16, n, l2, want_l2, got_ip, want_ip,
                           got_cos, want_cos);
                    return 1;
                }
                ++checked;
            }
    printf("HALF_WORDS_ORDER_OK %u cases\n", checked);
    return 0;
}
