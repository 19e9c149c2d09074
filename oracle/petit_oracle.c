/* See petit_oracle.h.  TEST INFRASTRUCTURE ONLY. */
#include "petit_oracle.h"

#include <math.h>
#include <string.h>

static const float kE2M1[16] = {0.0f, 0.5f, 1.0f, 1.5f, 2.0f, 3.0f, 4.0f, 6.0f,
                                -0.0f, -0.5f, -1.0f, -1.5f, -2.0f, -3.0f, -4.0f, -6.0f};

float petit_oracle_e2m1(unsigned code) { return kE2M1[code & 15]; }

float petit_oracle_e4m3(uint8_t bits) {
    int sign = bits >> 7, e = (bits >> 3) & 0xf, m = bits & 7;
    float v;
    if (e == 0xf && m == 7) return NAN;
    if (e == 0)
        v = ldexpf((float)m, -9);
    else
        v = ldexpf((float)(8 + m), e - 10);
    return sign ? -v : v;
}

float petit_oracle_e8m0(uint8_t bits) {
    uint32_t u = (uint32_t)bits << 23; /* bf16 (s << 7) widened to fp32 */
    float f;
    memcpy(&f, &u, 4);
    return f;
}

uint16_t petit_oracle_f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1);
    return (uint16_t)(u >> 16);
}

float petit_oracle_bf16_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

uint16_t petit_oracle_f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7e00);
    if (ax >= 0x47800000u) return (uint16_t)(sign | 0x7c00); /* >= 65536 -> inf */
    if (ax < 0x33000001u) return (uint16_t)sign;             /* < 2^-25 -> 0 */
    int e = (int)(ax >> 23) - 127;
    uint32_t man = (ax & 0x7fffffu) | 0x800000u;
    int shift = e < -14 ? (13 + (-14 - e)) : 13;
    uint32_t halfbit = 1u << (shift - 1);
    uint32_t rounded = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    if (rem > halfbit || (rem == halfbit && (rounded & 1))) rounded++;
    uint32_t h;
    if (e < -14)
        h = rounded; /* subnormal (may carry into the normal range, still right) */
    else
        h = ((uint32_t)(e + 15) << 10) + (rounded - 0x400u);
    if (h >= 0x7c00u) h = 0x7c00u;
    return (uint16_t)(sign | h);
}

float petit_oracle_f16_to_f32(uint16_t h) {
    int sign = h >> 15, e = (h >> 10) & 0x1f, m = h & 0x3ff;
    float v;
    if (e == 0x1f) v = m ? NAN : INFINITY;
    else if (e == 0) v = ldexpf((float)m, -24);
    else v = ldexpf((float)(1024 + m), e - 25);
    return sign ? -v : v;
}

static void dequant(float *out, const uint8_t *q, const uint8_t *scales, size_t n, size_t k,
                    size_t group, int mx) {
    for (size_t r = 0; r < n; ++r)
        for (size_t c = 0; c < k; ++c) {
            uint8_t byte = q[r * (k / 2) + c / 2];
            unsigned code = (c & 1) ? (byte >> 4) : (byte & 15);
            uint8_t sb = scales[r * (k / group) + c / group];
            float s = mx ? petit_oracle_e8m0(sb) : petit_oracle_e4m3(sb);
            out[r * k + c] = kE2M1[code] * s;
        }
}

void petit_oracle_dequant_nvfp4(float *out, const uint8_t *q, const uint8_t *scales, size_t n,
                                size_t k) {
    dequant(out, q, scales, n, k, 16, 0);
}

void petit_oracle_dequant_mxfp4(float *out, const uint8_t *q, const uint8_t *scales, size_t n,
                                size_t k) {
    dequant(out, q, scales, n, k, 32, 1);
}

void petit_oracle_gemm_f32(float *c, const float *a, const float *w, size_t m, size_t n,
                           size_t k) {
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) {
            const float *ar = a + i * k, *wr = w + j * k;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            size_t kk = 0;
            for (; kk + 8 <= k; kk += 8)
                for (int t = 0; t < 8; ++t) acc[t] += ar[kk + t] * wr[kk + t];
            float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            for (; kk < k; ++kk) s += ar[kk] * wr[kk];
            c[i * n + j] = s;
        }
}
