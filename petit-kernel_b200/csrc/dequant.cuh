// In-register FP4 -> 16-bit dequantisation shared by the GEMM kernels and the
// dense dequant hooks (so the exhaustive bit-exactness tests exercise exactly the
// arithmetic the GEMM feeds to the tensor cores).
//
// Replaces the reference's bit-trick dequantisers
// (lib/gemm/rocm/quantization/dequant.cuh:38-401).  Contract kept: every
// dequantised weight e2m1 * scale is produced EXACTLY (<= 6 significant bits).
#pragma once

#include "fp4_gemm.h"
#include "sm100_ptx.cuh"

#include <cstdint>

namespace petit::dq {

using petit::gemm::kModeMxBf16;
using petit::gemm::kModeNvBf16;
using petit::gemm::kModeNvF16;
using petit::ptx::cvt_e2m1x8_to_f16x2x4;

// Dequantisation of one 16-byte chunk (32 weights of one row) into 16 packed
// 16-bit pairs, in k order (low half = even k).
// NVFP4 -> fp16:   w = f16(e2m1) * f16(scale)             (exact, <= 6 sig. bits)
// NVFP4 -> bf16:   t = bf16 bits of e2m1 * 2^-112  (f16 bits >> 3, sign moved)
//                  w = t * bf16(scale * 2^112)            (exact)
// MXFP4 -> bf16:   w * 4 = t * 2^(s-13)   [one multiply, s <= 140]
//                        = (t * 2^112) * 2^(s-125)      [otherwise]
//                  (the factor 4 keeps s = 0/1 out of the bf16 subnormals; its
//                  inverse is folded into the epilogue scale).
__device__ __forceinline__ uint32_t hmul2_f16(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// f16x2 bits of a normal-or-zero value with <= 7 mantissa bits -> bf16x2 bits
// of (value * 2^-112): shift the exponent/mantissa down by 3 and move the sign
// from bit 12 back to bit 15 (adding 0x7000 carries it up; the mask drops the
// carry trail).
// The shift runs on the FMA pipe as IMAD.HI (mad.hi with a 2^29 multiplier that
// arrives as a kernel argument, so ptxas cannot strength-reduce it to LEA.HI/SHF) because the ALU pipe (F2FP, LOP3, SHF, LEA: 64 lanes/clk/SM) is
// the measured bottleneck of the decode path.
__device__ __forceinline__ uint32_t f16x2_to_bf16x2_scaled(uint32_t h, uint32_t two29) {
    uint32_t t;
    asm("mad.hi.u32 %0, %1, %2, 0x70007000;" : "=r"(t) : "r"(h), "r"(two29));
    return t & 0x8fff8fffu;
}
template <int MODE>
__device__ __forceinline__ void dequant_chunk(const uint4 q, uint32_t mult0,
                                              uint32_t mult1, bool two_step,
                                              uint32_t two29, uint32_t (&out)[16]) {
    const uint32_t words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t h[4];
        cvt_e2m1x8_to_f16x2x4(words[w], h[0], h[1], h[2], h[3]);
        // NVFP4: group 16 -> words 0,1 use mult0, words 2,3 use mult1.
        // MXFP4: group 32 -> all words use mult0 (mult1 = second-step factor).
        const uint32_t mult = (MODE == kModeMxBf16) ? mult0 : (w < 2 ? mult0 : mult1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == kModeNvF16) {
                out[w * 4 + j] = hmul2_f16(h[j], mult);
            } else if (MODE == kModeNvBf16) {
                out[w * 4 + j] = hmul2_bf16(f16x2_to_bf16x2_scaled(h[j], two29), mult);
            } else {
                uint32_t t = f16x2_to_bf16x2_scaled(h[j], two29);
                if (two_step) t = hmul2_bf16(t, 0x77807780u); // * 2^112
                out[w * 4 + j] = hmul2_bf16(t, mult);
            }
        }
    }
}

// Scale byte -> packed multiplier (same value in both halves).
template <int MODE>
__device__ __forceinline__ uint32_t scale_multiplier(uint32_t byte, bool &two_step) {
    two_step = false;
    if (MODE == kModeNvF16) {
        // E5M3 byte is exactly the fp16 exponent+3 mantissa bits: bits [14:7].
        return byte * 0x00800080u;
    } else if (MODE == kModeNvBf16) {
        // bf16 bits of scale * 2^112: exponent field = e5 + 224.
        return byte ? byte * 0x00100010u + 0x70007000u : 0u;
    } else {
        // e8m0 byte s: one step 2^(s-13) (field s+114) when representable,
        // else second-step factor 2^(s-125) (field s+2).
        two_step = byte > 140;
        uint32_t field = two_step ? byte + 2 : byte + 114;
        field = field > 255 ? 255 : field;
        return field * 0x00800080u;
    }
}


// Power of two folded out of the A operand and applied in the epilogue.
template <int MODE> __host__ __device__ constexpr float epilogue_factor() {
    return MODE == kModeMxBf16 ? 0.25f : 1.0f; // 2^-2
}

} // namespace petit::dq
