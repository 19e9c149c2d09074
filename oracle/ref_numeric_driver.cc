// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product.
//
// Driver around the REFERENCE's own host-side numeric helpers, compiled from where they lie
// (/root/reference/lib/tests/floating_points.h + lib/gemm/cpu/half_float.h, header-only, plain
// g++; nothing is copied into this repo).  It prints what the reference's
// ExhaustiveFp4DequantTest expects (quantization_utils_fp4_test.cc:240-264,344-365,388-394):
//     Element::from_fp32( fp8_e4m3_t::from_bits(s).to_fp32() * fp4_values[q] )
// for all 16 e2m1 codes x e4m3 bits 0x01..0x7E, Element = bf16_t and fp16_t, plus the raw
// e4m3 -> fp32 table and the reference's +-0-equal comparison, so that oracle/petit_oracle.py
// and the golden table can be checked against reference CODE, not only against a restatement.
// Built by oracle/Makefile into oracle/_ref/ref_numeric when /root/reference is present.
#include "tests/floating_points.h"

#include <cstdint>
#include <cstdio>
#include <cstring>

namespace num = causalflow::petit::tests::cpu_numeric;

static uint32_t f32_bits(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    return u;
}

int main() {
    // order and values of the table in the reference's GenerateNvOutput
    static const float fp4_values[] = {
        0.0f,  0.5f,  1.0f,  1.5f,  2.0f,  3.0f,  4.0f,  6.0f,
        -0.0f, -0.5f, -1.0f, -1.5f, -2.0f, -3.0f, -4.0f, -6.0f,
    };
    const unsigned lo = std::numeric_limits<num::fp8_e4m3_t>::denorm_min().to_bits();
    const unsigned hi = std::numeric_limits<num::fp8_e4m3_t>::max().to_bits();
    std::printf("range %u %u\n", lo, hi);
    for (unsigned s = 0; s < 256; ++s)
        std::printf("e4m3 %u %08x\n", s, f32_bits(num::fp8_e4m3_t::from_bits((uint8_t)s).to_fp32()));
    for (unsigned q = 0; q < 16; ++q)
        for (unsigned s = lo; s <= hi; ++s) {
            const float ref = num::fp8_e4m3_t::from_bits((uint8_t)s).to_fp32() * fp4_values[q];
            std::printf("deq %u %u %08x %04x %04x\n", q, s, f32_bits(ref),
                        (unsigned)num::bf16_t::from_fp32(ref).to_bits(),
                        (unsigned)num::fp16_t::from_fp32(ref).to_bits());
        }
    // the comparison the reference's tests use: +0 == -0, NaN != NaN
    std::printf("eq %d %d %d\n", (int)(num::bf16_t::from_bits(0x0000) == num::bf16_t::from_bits(0x8000)),
                (int)(num::bf16_t::from_bits(0x7fc0) == num::bf16_t::from_bits(0x7fc0)),
                (int)(num::fp16_t::from_bits(0x3c00) == num::fp16_t::from_bits(0x3c00)));
    return 0;
}
