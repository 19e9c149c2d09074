// Packed ("Blackwell tile") layouts shared by the repack kernels and the GEMM.
//
// The reference shuffles weights into a 64(K) x 32(N) wave64 tile with a
// per-word nibble permutation (quantization_utils.cu:183-253, SURVEY Appendix
// A).  On B200 the consumer is different: one dequant *thread* owns one weight
// row (= one TMEM lane of the TS-form tcgen05.mma A operand), reads 16 bytes
// (32 consecutive k) per shared-memory load, and the whole (128 rows x 256 k)
// block must be one contiguous range so a single cp.async.bulk fetches it.
//
// Weights: [n_tile = n/128][k_tile = k/256][chunk = (k%256)/32][row = n%128][16 bytes]
//   Each 16-byte chunk is four 32-bit words of 8 consecutive k.  Inside a word the
//   bits of the 8 e2m1 codes are permuted (pack_word) so that the pair (k=2p, 2p+1)
//   comes out as a bf16x2 register -- sign at bit 15/31, exponent/mantissa bits at
//   [8:6]/[24:22], i.e. the value times 2^-126 -- with ONE shift or rotate and ONE
//   LOP3 (see dequant.cuh).  Measured on B200 the alternative (cvt.rn.f16x2.e2m1x2
//   on native nibbles + fp16->bf16 re-bias) needs an IMAD.HI per pair that costs
//   ~4.7 issue cycles on the FMA-heavy pipe and caps bf16 decode at ~47 % of HBM.
//   A tile that is cut by N (rows R < 128, R % 16 == 0) keeps the same order
//   with R rows, so the buffer is exactly N*K/2 bytes.
// NVFP4 scales (one byte per 16 k, re-encoded E5M3):
//   [n_tile][k_tile][ksub = (k%256)/64][row][4 bytes]
// MXFP4 scales (one e8m0 byte per 32 k):
//   [n_tile][k_tile][ksub = (k%256)/64][row][2 bytes]
#pragma once

#include <cstdint>

namespace petit::layout {

constexpr int kLayoutVersion = 2;
constexpr uint32_t kTileN = 128;   // weight rows per tile (tcgen05 M)
constexpr uint32_t kTileK = 256;   // k elements per scheduling unit
constexpr uint32_t kChunkK = 32;   // k elements per 16-byte chunk
constexpr uint32_t kSubK = 64;     // k elements per scale sub-block / TMA slab

__host__ __device__ inline uint32_t tile_rows(uint32_t n, uint32_t n_tile) {
    uint32_t rem = n - n_tile * kTileN;
    return rem < kTileN ? rem : kTileN;
}

// Byte offset of element-pair (n, k) (k even) in the packed weight buffer.
__host__ __device__ inline size_t weight_byte_offset(uint32_t n, uint32_t k,
                                                     uint32_t size_n,
                                                     uint32_t size_k) {
    uint32_t nt = n / kTileN, r = n % kTileN;
    uint32_t rows = tile_rows(size_n, nt);
    uint32_t kt = k / kTileK, kk = k % kTileK;
    uint32_t chunk = kk / kChunkK, within = (kk % kChunkK) / 2;
    return (size_t)nt * kTileN * (size_k / 2) + (size_t)kt * rows * (kTileK / 2) +
           (size_t)chunk * rows * 16 + (size_t)r * 16 + within;
}

// Byte offset of the scale of (n, group g) where group = k / group_size;
// bytes_per_sub = 4 (NVFP4, group 16) or 2 (MXFP4, group 32).
__host__ __device__ inline size_t scale_byte_offset(uint32_t n, uint32_t g,
                                                    uint32_t size_n,
                                                    uint32_t size_k,
                                                    uint32_t bytes_per_sub) {
    uint32_t nt = n / kTileN, r = n % kTileN;
    uint32_t rows = tile_rows(size_n, nt);
    uint32_t groups_per_tile_k = (kTileK / kSubK) * bytes_per_sub; // per row
    uint32_t kt = g / groups_per_tile_k, gg = g % groups_per_tile_k;
    uint32_t ksub = gg / bytes_per_sub, j = gg % bytes_per_sub;
    size_t groups_per_row = (size_t)size_k / kSubK * bytes_per_sub;
    return (size_t)nt * kTileN * groups_per_row +
           (size_t)kt * rows * groups_per_tile_k +
           (size_t)ksub * rows * bytes_per_sub + (size_t)r * bytes_per_sub + j;
}

// Bit position (0..31) of {sign, e1, e0, m} of element i (0..7) inside a packed
// word.  Pair p = i / 2, half h = i % 2 (0: low 16 bits).  The magnitude bits of
// pairs 2 and 3 live in the OTHER half-word so that a 32-bit rotate brings both
// elements of the pair home at once.
struct NibbleBits {
    uint8_t s, e1, e0, m;
};
__host__ __device__ inline NibbleBits packed_bits(uint32_t i) {
    const uint32_t p = i >> 1, own = (i & 1) * 16, other = 16 - own;
    switch (p) {
    case 0: return {(uint8_t)(15 + own), (uint8_t)(8 + own), (uint8_t)(7 + own), (uint8_t)(6 + own)};
    case 1: return {(uint8_t)(11 + own), (uint8_t)(4 + own), (uint8_t)(3 + own), (uint8_t)(2 + own)};
    case 2: return {(uint8_t)(5 + own), (uint8_t)(14 + other), (uint8_t)(13 + other), (uint8_t)(12 + other)};
    default: return {(uint8_t)(1 + own), (uint8_t)(10 + other), (uint8_t)(9 + other), (uint8_t)(0 + own)};
    }
}
// native word (nibble i = element i = [s e1 e0 m] at bits 4i+3..4i) -> packed word
__host__ __device__ inline uint32_t pack_word(uint32_t w) {
    uint32_t o = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        const NibbleBits b = packed_bits(i);
        const uint32_t nib = (w >> (4 * i)) & 0xf;
        o |= ((nib >> 3) & 1u) << b.s;
        o |= ((nib >> 2) & 1u) << b.e1;
        o |= ((nib >> 1) & 1u) << b.e0;
        o |= (nib & 1u) << b.m;
    }
    return o;
}
__host__ __device__ inline uint32_t unpack_word(uint32_t o) {
    uint32_t w = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        const NibbleBits b = packed_bits(i);
        const uint32_t nib = (((o >> b.s) & 1u) << 3) | (((o >> b.e1) & 1u) << 2) |
                             (((o >> b.e0) & 1u) << 1) | ((o >> b.m) & 1u);
        w |= nib << (4 * i);
    }
    return w;
}

// e4m3 (positive, finite) -> unsigned E5M3: scale = 2^(e5-15) * (1 + m/8).
// Exact for every e4m3 in [2^-9, 448]; 0 maps to 0.  The sign bit is dropped
// and NaN (0x7f) saturates to 448, mirroring the reference's "scales must be
// positive" contract (quantization_utils.cu:143-162).
__host__ __device__ inline uint8_t e4m3_to_e5m3(uint8_t s) {
    s &= 0x7f;
    if (s == 0x7f) s = 0x7e;
    uint32_t e = s >> 3, m = s & 7;
    if (e == 0) {
        if (m == 0) return 0;
        // subnormal: m/8 * 2^-6 ; normalise
        int shift = 0;
        while ((m & 8) == 0) {
            m <<= 1;
            ++shift;
        }
        m &= 7;
        // value = 2^(-6 - shift) * (1 + m/8)  -> E = -6 - shift, e5 = E + 15
        return (uint8_t)(((9 - shift) << 3) | m);
    }
    // value = 2^(e-7) * (1+m/8) -> e5 = e + 8
    return (uint8_t)(((e + 8) << 3) | m);
}

// Inverse (used by the unpack / round-trip hooks).
__host__ __device__ inline uint8_t e5m3_to_e4m3(uint8_t p) {
    if (p == 0) return 0;
    int e5 = p >> 3, m = p & 7;
    int e = e5 - 8;
    if (e >= 1) return (uint8_t)((e << 3) | m);
    // subnormal e4m3: value = 2^(e5-15)*(1+m/8) = M/8 * 2^-6
    int shift = 1 - e;              // 1..3
    int mm = (8 | m) >> shift;      // exact for encodable values
    return (uint8_t)mm;
}

} // namespace petit::layout
