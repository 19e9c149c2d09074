// In-register FP4 -> 16-bit dequantisation shared by the GEMM kernels and the
// dense dequant hooks (so the exhaustive bit-exactness tests exercise exactly the
// arithmetic the GEMM feeds to the tensor cores).
//
// Replaces the reference's bit-trick dequantisers
// (lib/gemm/rocm/quantization/dequant.cuh:38-401).  Contract kept: every
// dequantised weight e2m1 * scale is produced EXACTLY (<= 6 significant bits).
//
// The e2m1 code [s e1 e0 m] dropped into a 16-bit float with its sign at bit 15 and
// e1 e0 m at the two lowest exponent bits + top mantissa bit IS the value scaled by
// a power of two, subnormal and zero included:
//     bf16: bits [8:6]   -> value * 2^-126        fp16: bits [11:9] -> value * 2^-14
// layout.cuh::pack_word stores the bits of a word so that these patterns fall out
// with one shift/rotate + one LOP3 per PAIR of weights for bf16; HMUL2 by the block
// scale (times the inverse power of two) then normalises and scales in one exact
// step.  Instruction budget per 2 weights (B200: the 64-lane/clk ALU pipe is the
// decode bottleneck, see profiles/):
//     bf16: 1.75 ALU (LOP3, SHF) + 1.5 FMA-pipe (IMAD.SHL, HMUL2.BF16)
//     fp16: 2.75 ALU            + 2.5 FMA-pipe
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
using petit::gemm::kModeNvF16N;

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

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
    return __funnelshift_l(x, x, r);
}

// One packed word (8 weights) -> 4 bf16x2 registers holding value * 2^-126.
__device__ __forceinline__ void extract_bf16(uint32_t q, uint32_t (&x)[4]) {
    constexpr uint32_t kM = 0x81c081c0u;
    x[0] = q & kM;
    // Two plain shifts (IMAD.SHL, FMA pipe) and two rotates (SHF, ALU pipe) per word is
    // the measured optimum on B200: moving one more shift to either pipe slows the decode
    // loop (gate_up M=16: 57.6 us as is, 64.2 us with rotl14 -> (rotl10 << 4), 59.3 us
    // with every shift as a rotate).
    x[1] = (q << 4) & kM;
    x[2] = rotl32(q, 10) & kM;
    x[3] = (rotl32(q, 14) & 0x81808180u) | ((q << 6) & 0x00400040u);
}
// One packed word -> 4 f16x2 registers holding value * 2^-14.
__device__ __forceinline__ void extract_f16(uint32_t q, uint32_t (&x)[4]) {
    constexpr uint32_t kS = 0x80008000u, kG = 0x0e000e00u;
    x[0] = ((q << 3) & kG) | (q & kS);
    x[1] = ((q << 7) & kG) | ((q << 4) & kS);
    const uint32_t r = rotl32(q, 13);
    x[2] = (r & kG) | ((q << 10) & kS);
    x[3] = ((r << 4) & 0x0c000c00u) | ((q << 9) & 0x02000200u) | ((q << 14) & kS);
}

// Power of two folded out of the A operand and applied in the epilogue: the A
// operand holds  w * 2^-8 (NVFP4 bf16),  w * 2^-7 (NVFP4 fp16)  or  w * 4 (MXFP4).
template <int MODE> __host__ __device__ constexpr float epilogue_factor() {
    return MODE == kModeNvF16N ? 1.0f
                               : (MODE == kModeMxBf16 ? 0.25f : (MODE == kModeNvBf16 ? 256.0f : 128.0f));
}

__device__ __forceinline__ bool mx_needs_two_step(uint32_t bits) { return (bits & 0xff) > 125; }

// Scale bits of one 32-weight chunk -> multiplier register.
//   NVFP4: `bits` = two E5M3 bytes (groups of 16); result = (mult0, mult1) halves,
//          mult = scale * 2^118 (bf16) / scale * 2^7 (fp16); a zero byte gives 0.
//   MXFP4: `bits` = one e8m0 byte s; result = 2^(s+1) in both halves, or, when
//          `two_step` (required if mx_needs_two_step; may be forced warp-uniformly),
//          2^(s-125) to be applied after an exact * 2^126.
template <int MODE>
__device__ __forceinline__ uint32_t chunk_multiplier(uint32_t bits, bool two_step) {
    if (MODE == kModeMxBf16) {
        const uint32_t s = bits & 0xff;
        uint32_t field = two_step ? s + 2 : s + 128;
        field = field > 255 ? 255 : field;
        return field * 0x00800080u;
    }
    const uint32_t x = __byte_perm(bits, 0, 0x4140);     // bytes (b0, b1) -> 16-bit lanes
    if (MODE == kModeNvF16N) {
        // an E5M3 byte IS the top 8 bits (below the sign) of the fp16 with the same value:
        // exponent field e5, mantissa m -> the exact scale, 0 for a zero byte
        return x << 7;
    }
    const uint32_t nz = (x + 0x00ff00ffu) & 0x01000100u; // bit 8 of a lane = (byte != 0)
    if (MODE == kModeNvF16) {
        // fp16(scale * 2^7): exponent field e5 + 7, top 3 mantissa bits = m
        return x * 128u + nz * 0x1cu;
    }
    // bf16(scale * 2^118): exponent field e5 + 230 = (0x7300 >> 7) + e5
    return x * 16u + nz * 0x73u;
}

// Dequantise one 16-byte chunk (4 packed words = 32 weights of one row) into 16
// registers of 16-bit pairs in k order (low half = even k).
template <int MODE>
__device__ __forceinline__ void dequant_chunk(const uint4 q, uint32_t mult, bool two_step,
                                              uint32_t (&out)[16]) {
    const uint32_t words[4] = {q.x, q.y, q.z, q.w};
    if (MODE == kModeNvF16N) {
        // fp16-native layout: byte b of a word = elements (2b, 2b + 1), low nibble first.
        // One F2FP (cvt.rn.f16x2.e2m1x2) per pair gives the exact e2m1 values, one HMUL2 by
        // the exact fp16 scale the exact weight: 2 instructions per pair.
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t x[4];
            petit::ptx::cvt_e2m1x8_to_f16x2x4(words[w], x[0], x[1], x[2], x[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                out[w * 4 + j] = w < 2 ? hmul2_bcast<false, false>(x[j], mult)
                                       : hmul2_bcast<false, true>(x[j], mult);
        }
    } else {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t x[4];
        if (MODE == kModeNvF16)
            extract_f16(words[w], x);
        else
            extract_bf16(words[w], x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == kModeNvF16) {
                out[w * 4 + j] = w < 2 ? hmul2_bcast<false, false>(x[j], mult)
                                       : hmul2_bcast<false, true>(x[j], mult);
            } else if (MODE == kModeNvBf16) {
                out[w * 4 + j] = w < 2 ? hmul2_bcast<true, false>(x[j], mult)
                                       : hmul2_bcast<true, true>(x[j], mult);
            } else {
                uint32_t t = x[j];
                if (two_step) t = hmul2_bf16(t, 0x7e807e80u); // * 2^126 (exact)
                out[w * 4 + j] = hmul2_bf16(t, mult);
            }
        }
    }
    }
}

} // namespace petit::dq
