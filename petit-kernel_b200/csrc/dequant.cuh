// In-register FP4 -> 16-bit dequantisation shared by the GEMM kernels and the
// dense dequant hooks (so the exhaustive bit-exactness tests exercise exactly the
// arithmetic the GEMM feeds to the tensor cores).
//
// Replaces the reference's bit-trick dequantisers
// (lib/gemm/rocm/quantization/dequant.cuh:38-401).  Contract kept: every
// dequantised weight e2m1 * scale is produced EXACTLY (<= 6 significant bits).
//
// Instruction budget (measured on B200, profiles/r01_probe_hw_facts.log): the ALU
// pipe (F2FP, LOP3, SHF, PRMT, LEA; 64 lanes/clk/SM) is the decode bottleneck, so
// everything that can run on the FMA pipe does:
//   NVFP4 -> fp16 : F2FP (ALU) + HMUL2 (FMA)                      per 2 weights
//   NVFP4 -> bf16 : F2FP (ALU) + IMAD.HI (FMA) + LOP3 (ALU) + HMUL2.BF16 (FMA)
//   MXFP4 -> bf16 : as NVFP4 -> bf16 (+ one HMUL2.BF16 for scales >= 2^14)
#pragma once

#include "fp4_gemm.h"
#include "sm100_ptx.cuh"

#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace petit::dq {

using petit::gemm::kModeMxBf16;
using petit::gemm::kModeNvBf16;
using petit::gemm::kModeNvF16;
using petit::ptx::cvt_e2m1x8_to_f16x2x4;

// Run-time constants that must not be visible to ptxas as immediates (it would
// strength-reduce the multiply-high back into ALU-pipe shifts).
struct Consts {
    uint32_t two29;  // 1 << 29
    uint64_t add64;  // 0x70007000 << 32
};

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
// a * (s.lo, s.lo) / a * (s.hi, s.hi): compiles to HMUL2 with the .H0_H0 /
// .H1_H1 operand swizzle, so one register carries the two group scales of a chunk.
template <bool kBf16, bool kHigh>
__device__ __forceinline__ uint32_t hmul2_bcast(uint32_t a, uint32_t s) {
    if (kBf16) {
        __nv_bfloat162 av = *reinterpret_cast<__nv_bfloat162 *>(&a);
        __nv_bfloat162 sv = *reinterpret_cast<__nv_bfloat162 *>(&s);
        __nv_bfloat162 r = __hmul2(
            av, __bfloat162bfloat162(kHigh ? __high2bfloat16(sv) : __low2bfloat16(sv)));
        return *reinterpret_cast<uint32_t *>(&r);
    } else {
        __half2 av = *reinterpret_cast<__half2 *>(&a);
        __half2 sv = *reinterpret_cast<__half2 *>(&s);
        __half2 r = __hmul2(av, __half2half2(kHigh ? __high2half(sv) : __low2half(sv)));
        return *reinterpret_cast<uint32_t *>(&r);
    }
}

// f16x2 bits of a normal-or-zero value with <= 7 mantissa bits -> bf16x2 bits of
// (value * 2^-112): shift exponent+mantissa down by 3 (IMAD.HI by 2^29) and move
// the sign from bit 12 back to bit 15 (the 0x7000 addend carries it up, the mask
// drops the carry trail).
__device__ __forceinline__ uint32_t f16x2_to_bf16x2_scaled(uint32_t h, const Consts &c) {
    const uint64_t p = (uint64_t)h * c.two29 + c.add64;
    return (uint32_t)(p >> 32) & 0x8fff8fffu;
}

// Scale bits of one 32-weight chunk -> multiplier register.
//   NVFP4: `bits` = two E5M3 bytes (groups of 16); result = (mult0, mult1) halves.
//   MXFP4: `bits` = one e8m0 byte; result = same multiplier in both halves.
//          `two_step` (must be true when mx_needs_two_step(bits); may be forced true,
//          e.g. warp-uniformly) selects the form with the extra * 2^112.
__device__ __forceinline__ bool mx_needs_two_step(uint32_t bits) { return (bits & 0xff) > 140; }

template <int MODE>
__device__ __forceinline__ uint32_t chunk_multiplier(uint32_t bits, bool two_step) {
    if (MODE == kModeMxBf16) {
        // A operand holds w * 4:  one step 2^(s-13) (field s+114) when representable,
        // else (t * 2^112) * 2^(s-125) (field s+2).
        const uint32_t s = bits & 0xff;
        uint32_t field = two_step ? s + 2 : s + 114;
        field = field > 255 ? 255 : field;
        return field * 0x00800080u;
    }
    // bytes (b0, b1) -> 16-bit lanes
    const uint32_t x = __byte_perm(bits, 0, 0x4140);
    if (MODE == kModeNvF16) {
        // E5M3 byte == fp16 exponent + top 3 mantissa bits: bits [14:7]
        return x << 7;
    }
    // bf16 bits of scale * 2^112: (byte << 4) + (224 << 7); a zero byte stays zero
    const uint32_t nz = (x + 0x00ff00ffu) & 0x01000100u; // bit 8 of each lane = (byte != 0)
    return x * 16u + nz * 0x70u;
}

// Dequantise one 16-byte chunk (32 weights of one row, k order, low nibble = even
// k) into 16 packed 16-bit pairs.
//   NVFP4 -> fp16:   w = f16(e2m1) * f16(scale)                    (exact)
//   NVFP4 -> bf16:   w = bf16(e2m1 * 2^-112) * bf16(scale * 2^112)  (exact)
//   MXFP4 -> bf16:   w * 4 (the 1/4 is folded into the epilogue scale)
template <int MODE>
__device__ __forceinline__ void dequant_chunk(const uint4 q, uint32_t mult, bool two_step,
                                              const Consts &c, uint32_t (&out)[16]) {
    const uint32_t words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t h[4];
        cvt_e2m1x8_to_f16x2x4(words[w], h[0], h[1], h[2], h[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == kModeNvF16) {
                out[w * 4 + j] = w < 2 ? hmul2_bcast<false, false>(h[j], mult)
                                       : hmul2_bcast<false, true>(h[j], mult);
            } else if (MODE == kModeNvBf16) {
                const uint32_t t = f16x2_to_bf16x2_scaled(h[j], c);
                out[w * 4 + j] = w < 2 ? hmul2_bcast<true, false>(t, mult)
                                       : hmul2_bcast<true, true>(t, mult);
            } else {
                uint32_t t = f16x2_to_bf16x2_scaled(h[j], c);
                if (two_step) t = hmul2_bf16(t, 0x77807780u); // * 2^112
                out[w * 4 + j] = hmul2_bf16(t, mult);
            }
        }
    }
}

// Power of two folded out of the A operand and applied in the epilogue.
template <int MODE> __host__ __device__ constexpr float epilogue_factor() {
    return MODE == kModeMxBf16 ? 0.25f : 1.0f; // 2^-2
}

} // namespace petit::dq
