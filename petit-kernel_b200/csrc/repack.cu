// Offline repack / unpack kernels and the dense dequantisation hooks.
//
// Replaces lib/gemm/rocm/quantization/fp4/quantization_utils.cu of the
// reference: RepackNvFp4ToPetitFp4Weights (:729-746), RepackNvFp4ToPetitFp4Scales
// (:748-760), RepackMxFp4ToPetitFp4Scales (:762-773) and the dense dequant
// launchers DequantNvFp4 / DequantMxFp4 / DequantPetitFp4 / DequantPetitMxFp4
// (:614-727).  The target layout is the Blackwell tile layout of layout.cuh, so
// the repack is a 16-byte block transpose (weights) and a 2/4-byte block
// transpose plus an exact e4m3 -> E5M3 re-encode (NVFP4 scales); inside each
// 32-bit word the 8 nibbles are bit-permuted into the bf16-native order of
// layout.cuh::pack_word (the role the reference's PetitFormat plays for MFMA).
//
// All kernels are pure data movement over N*K/2 bytes; they run once per layer
// at weight-load time.  Each thread moves one 16-byte chunk; reads of a warp
// cover 512 contiguous bytes of one weight row.
#include "dequant.cuh"
#include "layout.cuh"
#include "repack.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace petit::repack {

using namespace petit::layout;

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ weights
template <bool kUnpack, bool kNative16 = false>
__global__ void __launch_bounds__(kThreads)
repack_weights_kernel(uint4 *__restrict__ out, const uint4 *__restrict__ in, uint32_t size_n,
                      uint32_t size_k) {
    const uint32_t chunks_per_row = size_k / kChunkK;
    const uint64_t total = (uint64_t)size_n * chunks_per_row;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * kThreads) {
        const uint32_t n = (uint32_t)(i / chunks_per_row);
        const uint32_t ck = (uint32_t)(i % chunks_per_row);
        const size_t native = (size_t)n * chunks_per_row + ck;
        const size_t packed = weight_byte_offset(n, ck * kChunkK, size_n, size_k) / 16;
        if (kNative16) { // tile arrangement only: the words keep their native nibble order
            if (kUnpack) out[native] = in[packed];
            else out[packed] = in[native];
        } else if (kUnpack) {
            uint4 v = in[packed];
            out[native] = make_uint4(unpack_word(v.x), unpack_word(v.y), unpack_word(v.z),
                                     unpack_word(v.w));
        } else {
            uint4 v = in[native];
            out[packed] = make_uint4(pack_word(v.x), pack_word(v.y), pack_word(v.z),
                                     pack_word(v.w));
        }
    }
}

// ------------------------------------------------------------------- scales
// kBytes = 4: NVFP4 (group 16, e4m3 -> E5M3); kBytes = 2: MXFP4 (group 32, copy)
template <int kBytes, bool kUnpack>
__global__ void __launch_bounds__(kThreads)
repack_scales_kernel(uint8_t *__restrict__ out, const uint8_t *__restrict__ in, uint32_t size_n,
                     uint32_t size_k) {
    const uint32_t subs_per_row = size_k / kSubK;
    const uint64_t total = (uint64_t)size_n * subs_per_row;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * kThreads) {
        const uint32_t n = (uint32_t)(i / subs_per_row);
        const uint32_t sub = (uint32_t)(i % subs_per_row);
        const size_t native = ((size_t)n * subs_per_row + sub) * kBytes;
        const size_t packed = scale_byte_offset(n, sub * kBytes, size_n, size_k, kBytes);
#pragma unroll
        for (int j = 0; j < kBytes; ++j) {
            if (kUnpack) {
                uint8_t v = in[packed + j];
                out[native + j] = kBytes == 4 ? e5m3_to_e4m3(v) : v;
            } else {
                uint8_t v = in[native + j];
                out[packed + j] = kBytes == 4 ? e4m3_to_e5m3(v) : v;
            }
        }
    }
}

// ------------------------------------------------------ dense dequant hooks
// One thread per (row, 32-k chunk).  kPacked selects the input layout.  The
// arithmetic is dq::dequant_chunk -- the function the GEMM uses -- followed by
// the reference's 16-bit multiply with (16-bit)global_scale
// (quantization_utils.cu:563-585).
template <int MODE, bool kPacked>
__global__ void __launch_bounds__(kThreads)
dequant_dense_kernel(uint32_t *__restrict__ out, const uint8_t *__restrict__ w,
                     const uint8_t *__restrict__ sc, float global_scale, uint32_t size_n,
                     uint32_t size_k) {
    constexpr bool kIsMx = MODE == gemm::kModeMxBf16;
    constexpr bool kIsBf16 = MODE == gemm::kModeNvBf16 || MODE == gemm::kModeMxBf16;
    constexpr bool kNative16 = MODE == gemm::kModeNvF16N;
    constexpr uint32_t kGroup = kIsMx ? 32 : 16;
    constexpr uint32_t kScBytes = kIsMx ? 2 : 4;
    const uint32_t chunks_per_row = size_k / kChunkK;
    const uint64_t total = (uint64_t)size_n * chunks_per_row;

    uint32_t gs2;
    if (kIsBf16) {
        __nv_bfloat162 g = __float2bfloat162_rn(global_scale * dq::epilogue_factor<MODE>());
        gs2 = *reinterpret_cast<uint32_t *>(&g);
    } else {
        __half2 g = __float2half2_rn(global_scale * dq::epilogue_factor<MODE>());
        gs2 = *reinterpret_cast<uint32_t *>(&g);
    }

    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * kThreads) {
        const uint32_t n = (uint32_t)(i / chunks_per_row);
        const uint32_t ck = (uint32_t)(i % chunks_per_row);
        const uint32_t k0 = ck * kChunkK;
        uint4 q;
        uint32_t s0, s1;
        if (kPacked) {
            q = *reinterpret_cast<const uint4 *>(w + weight_byte_offset(n, k0, size_n, size_k));
            const size_t so = scale_byte_offset(n, k0 / kGroup, size_n, size_k, kScBytes);
            s0 = sc[so];
            s1 = kIsMx ? s0 : sc[so + 1];
        } else {
            q = *reinterpret_cast<const uint4 *>(w + (size_t)n * (size_k / 2) + k0 / 2);
            if (!kNative16)
                q = make_uint4(pack_word(q.x), pack_word(q.y), pack_word(q.z), pack_word(q.w));
            const size_t so = (size_t)n * (size_k / kGroup) + k0 / kGroup;
            s0 = kIsMx ? sc[so] : e4m3_to_e5m3(sc[so]);
            s1 = kIsMx ? s0 : e4m3_to_e5m3(sc[so + 1]);
        }
        const uint32_t bits = kIsMx ? s0 : (s0 | (s1 << 8));
        const bool two_step = kIsMx && dq::mx_needs_two_step(bits);
        const uint32_t mult = dq::chunk_multiplier<MODE>(bits, two_step);
        uint32_t v[16];
        dq::dequant_chunk<MODE>(q, mult, two_step, v);
        uint32_t *dst = out + ((size_t)n * size_k + k0) / 2;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            dst[j] = kIsBf16 ? dq::hmul2_bf16(v[j], gs2) : dq::hmul2_f16(v[j], gs2);
    }
}

inline unsigned grid_for(uint64_t total) {
    uint64_t blocks = (total + kThreads - 1) / kThreads;
    const uint64_t cap = 148ull * 16;
    return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? 0 : 3; }

} // namespace

bool shape_ok(unsigned size_k, unsigned size_n) {
    return size_k != 0 && size_n != 0 && size_k % kTileK == 0 && size_n % 16 == 0;
}

int weights(void *out, const void *in, unsigned size_k, unsigned size_n, bool unpack,
            cudaStream_t stream, bool native16) {
    if (!shape_ok(size_k, size_n)) return 1;
    const uint64_t total = (uint64_t)size_n * (size_k / kChunkK);
    if (native16 && unpack)
        repack_weights_kernel<true, true><<<grid_for(total), kThreads, 0, stream>>>(
            (uint4 *)out, (const uint4 *)in, size_n, size_k);
    else if (native16)
        repack_weights_kernel<false, true><<<grid_for(total), kThreads, 0, stream>>>(
            (uint4 *)out, (const uint4 *)in, size_n, size_k);
    else if (unpack)
        repack_weights_kernel<true><<<grid_for(total), kThreads, 0, stream>>>(
            (uint4 *)out, (const uint4 *)in, size_n, size_k);
    else
        repack_weights_kernel<false><<<grid_for(total), kThreads, 0, stream>>>(
            (uint4 *)out, (const uint4 *)in, size_n, size_k);
    return check_launch();
}

int scales(void *out, const void *in, unsigned size_k, unsigned size_n, bool mx, bool unpack,
           cudaStream_t stream) {
    if (!shape_ok(size_k, size_n)) return 1;
    const uint64_t total = (uint64_t)size_n * (size_k / kSubK);
    const unsigned g = grid_for(total);
    auto o = (uint8_t *)out;
    auto i = (const uint8_t *)in;
    if (mx) {
        if (unpack)
            repack_scales_kernel<2, true><<<g, kThreads, 0, stream>>>(o, i, size_n, size_k);
        else
            repack_scales_kernel<2, false><<<g, kThreads, 0, stream>>>(o, i, size_n, size_k);
    } else {
        if (unpack)
            repack_scales_kernel<4, true><<<g, kThreads, 0, stream>>>(o, i, size_n, size_k);
        else
            repack_scales_kernel<4, false><<<g, kThreads, 0, stream>>>(o, i, size_n, size_k);
    }
    return check_launch();
}

int dequant_dense(void *out, const void *w, const void *sc, float global_scale, int mode,
                  bool packed, unsigned size_k, unsigned size_n, cudaStream_t stream) {
    if (!shape_ok(size_k, size_n)) return -1;
    const uint64_t total = (uint64_t)size_n * (size_k / kChunkK);
    const unsigned g = grid_for(total);
    auto o = (uint32_t *)out;
    auto wp = (const uint8_t *)w;
    auto sp = (const uint8_t *)sc;
#define PETIT_DQ(MODE)                                                                        \
    if (packed)                                                                               \
        dequant_dense_kernel<MODE, true>                                                      \
            <<<g, kThreads, 0, stream>>>(o, wp, sp, global_scale, size_n, size_k); \
    else                                                                                      \
        dequant_dense_kernel<MODE, false>                                                     \
            <<<g, kThreads, 0, stream>>>(o, wp, sp, global_scale, size_n, size_k);
    switch (mode) {
    case gemm::kModeNvF16: PETIT_DQ(gemm::kModeNvF16) break;
    case gemm::kModeNvBf16: PETIT_DQ(gemm::kModeNvBf16) break;
    case gemm::kModeMxBf16: PETIT_DQ(gemm::kModeMxBf16) break;
    case gemm::kModeNvF16N: PETIT_DQ(gemm::kModeNvF16N) break;
    default: return -1;
    }
#undef PETIT_DQ
    return check_launch() == 0 ? 0 : -1;
}

} // namespace petit::repack
