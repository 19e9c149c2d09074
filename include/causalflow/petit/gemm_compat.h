/* C++ source-compatibility shim over the C ABI (petit.h).
 *
 * A translation unit written against the reference's C++ interface --
 *   lib/gemm/rocm/quantization/gemm.h:6-146   (namespace ...::rocm::quantization, SolutionId,
 *                                               PetitSolutionHints, fp4::GemmFp4Fp16Grid ...)
 *   lib/gemm/rocm/quantization/types.h:4-13   (DataType)
 *   lib/gemm/rocm/quantization/fp4/gemm_fp4.h:11-21 (dense dequant hooks)
 *   lib/hal/device.h:8-34                      (hal::Device / Platform / GetPlatform)
 * -- compiles against this header unchanged apart from the include path and `hipStream_t`
 * being `cudaStream_t`: same namespaces, same names, same argument order and meaning, same
 * return codes.  Everything is an inline forwarder to libpetit_b200.so; nothing here launches
 * anything itself.  Differences, all by construction of the B200 backend:
 *   - solutions are the five token-tile widths of one stream-K tcgen05 kernel, still carried in
 *     the reference's 64-bit SolutionId bit layout (ids from GemmGetSolutions are valid inputs,
 *     ids of the reference's MFMA tile shapes return kErrorKernelShape);
 *   - the Repack* functions return void like the reference but report a failed launch through
 *     std::runtime_error instead of ignoring it;
 *   - hal::Device returns int (0 or a cudaError_t) instead of absl::Status: abseil is not a
 *     dependency of this library.
 */
#ifndef CAUSALFLOW_PETIT_GEMM_COMPAT_H_
#define CAUSALFLOW_PETIT_GEMM_COMPAT_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime_api.h>
#include <memory>
#include <stddef.h>
#include <stdint.h>
#include <stdexcept>

/* The C ABI's own names (PetitSolutionHints, PETIT_DTYPE_* ...) live in causalflow::petit::cabi
 * here, so that `using namespace causalflow::petit::rocm::quantization;` -- what reference code
 * does -- never meets a second, global PetitSolutionHints.  (extern "C" functions keep their C
 * linkage inside a namespace.)  Include this header INSTEAD of petit.h in such a unit. */
#ifdef CAUSALFLOW_PETIT_PETIT_H_
#error "include causalflow/petit/gemm_compat.h instead of (not after) causalflow/petit/petit.h"
#endif
namespace causalflow::petit::cabi {
#include "causalflow/petit/petit.h"
} // namespace causalflow::petit::cabi

using hipStream_t = cudaStream_t; /* the one spelling the reference's signatures differ in */

namespace causalflow::petit::rocm::quantization {

/* types.h:4-13 */
enum class DataType {
    kDataTypeInt4 = cabi::PETIT_DTYPE_INT4,
    kDataTypeFp8e4m3 = cabi::PETIT_DTYPE_FP8_E4M3,
    kDataTypeFp8e8m0 = cabi::PETIT_DTYPE_FP8_E8M0,
    kDataTypeFp4e2m1 = cabi::PETIT_DTYPE_FP4_E2M1,
    kDataTypeFp16 = cabi::PETIT_DTYPE_FP16,
    kDataTypeBf16 = cabi::PETIT_DTYPE_BF16,
    kDataTypeFp8e5m2Fnuz = cabi::PETIT_DTYPE_FP8_E5M2_FNUZ,
    kDataTypeMxFp4e2m1 = cabi::PETIT_DTYPE_MXFP4_E2M1,
};

/* gemm.h:8-31 */
enum MatmulFeatures {
    kMatmulFeatures_Global = 0,
    kMatmulFeatures_Grid = 1,
    kMatmulFeatures_HighPrecision = 1 << 1,
};
enum MatmulElementB { kMatmulTypeBInt4, kMatmulTypeBNvFp4, kMatmulTypeBMxFp4 };
enum MatmulMfmaType { kMatmulMfmaTypeFp16, kMatmulMfmaTypeBf16, kMatmulMfmaTypeFp8 };
enum MatmulWarpPartition { kMatmulWarpPartition_NK, kMatmulWarpPartition_Cooperative };

/* gemm.h:33-105: same 64-bit layout, spelled with shifts instead of bit-fields so that the
 * representation does not depend on the compiler's bit-field ABI. */
struct SolutionId {
    uint64_t repr;

    constexpr unsigned long Repr() const { return (unsigned long)repr; }
    static constexpr SolutionId FromRepr(unsigned long r) { return SolutionId{(uint64_t)r}; }

    constexpr unsigned tile_m() const { return (unsigned)(repr >> 0) & 0xff; }
    constexpr unsigned tile_n() const { return (unsigned)(repr >> 8) & 0xff; }
    constexpr unsigned tile_k() const { return (unsigned)(repr >> 16) & 0xff; } /* units of 64 */
    constexpr MatmulFeatures features() const { return (MatmulFeatures)((repr >> 24) & 0xf); }
    constexpr MatmulElementB element_b() const { return (MatmulElementB)((repr >> 28) & 0xf); }
    constexpr MatmulMfmaType mfma_type() const { return (MatmulMfmaType)((repr >> 32) & 0xf); }
    constexpr unsigned warp_partition_m() const { return (unsigned)(repr >> 36) & 0xf; }
    constexpr unsigned warp_partition_n() const { return (unsigned)(repr >> 40) & 0xf; }
    constexpr unsigned warp_partition_k() const { return (unsigned)(repr >> 44) & 0xf; }
    constexpr MatmulWarpPartition warp_partition() const {
        return (MatmulWarpPartition)((repr >> 48) & 0xf);
    }

    static constexpr SolutionId MultiStage(MatmulFeatures features, MatmulElementB element_b,
                                           MatmulMfmaType mfma_type, unsigned tile_m,
                                           unsigned tile_n, unsigned tile_k,
                                           MatmulWarpPartition warp_partition,
                                           unsigned warp_partition_m, unsigned warp_partition_n,
                                           unsigned warp_partition_k) {
        return SolutionId{(uint64_t)(tile_m & 0xff) | ((uint64_t)(tile_n & 0xff) << 8) |
                          ((uint64_t)((tile_k / 4) & 0xff) << 16) | ((uint64_t)features << 24) |
                          ((uint64_t)element_b << 28) | ((uint64_t)mfma_type << 32) |
                          ((uint64_t)(warp_partition_m & 0xf) << 36) |
                          ((uint64_t)(warp_partition_n & 0xf) << 40) |
                          ((uint64_t)(warp_partition_k & 0xf) << 44) |
                          ((uint64_t)warp_partition << 48)};
    }
};
static_assert(sizeof(SolutionId) == 8, "");

static constexpr int kErrorProblemShape = PETIT_ERROR_PROBLEM_SHAPE; /* gemm.h:107-108 */
static constexpr int kErrorKernelShape = PETIT_ERROR_KERNEL_SHAPE;

/* gemm.h:112-117 */
struct PetitSolutionHints {
    DataType a_type;
    DataType b_type;
    DataType c_type;
    bool require_high_precision;
};

namespace detail {
inline cabi::PetitSolutionHints to_c(const PetitSolutionHints &h) {
    cabi::PetitSolutionHints c;
    c.a_type = (int32_t)h.a_type;
    c.b_type = (int32_t)h.b_type;
    c.c_type = (int32_t)h.c_type;
    c.require_high_precision = h.require_high_precision ? 1 : 0;
    return c;
}
inline void check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(what);
}
} // namespace detail

namespace fp4 {

/* gemm.h:120-124 */
inline int GemmFp4Fp16Grid(unsigned *c, const unsigned *a, const unsigned *b,
                           const unsigned *scales, const float *global_scale, const unsigned m,
                           const unsigned n, const unsigned k, const PetitSolutionHints &hints,
                           unsigned long solution_id, hipStream_t stream) {
    const cabi::PetitSolutionHints h = detail::to_c(hints);
    return cabi::petit_gemm_nvfp4_a16(c, a, b, scales, global_scale, m, n, k, &h, (uint64_t)solution_id,
                                stream);
}

/* gemm.h:126-130 */
inline int GemmMxFp4Fp16Grid(unsigned *c, const unsigned *a, const unsigned *b,
                             const unsigned *scales, const float *global_scale, const unsigned m,
                             const unsigned n, const unsigned k, const PetitSolutionHints &hints,
                             unsigned long solution_id, hipStream_t stream) {
    const cabi::PetitSolutionHints h = detail::to_c(hints);
    return cabi::petit_gemm_mxfp4_a16(c, a, b, scales, global_scale, m, n, k, &h, (uint64_t)solution_id,
                                stream);
}

/* gemm.h:132-133 */
inline int GemmGetSolutions(const PetitSolutionHints &hints, unsigned m, unsigned n, unsigned k,
                            SolutionId *sols, unsigned *n_sols) {
    const cabi::PetitSolutionHints h = detail::to_c(hints);
    return cabi::petit_get_solutions(&h, m, n, k, reinterpret_cast<uint64_t *>(sols), n_sols);
}

/* gemm.h:135-145 */
inline void RepackNvFp4ToPetitFp4Weights(unsigned *output, const unsigned *input, unsigned in_chan,
                                         unsigned out_chan, hipStream_t stream) {
    detail::check(cabi::petit_repack_fp4_weights(output, input, in_chan, out_chan, stream),
                  "RepackNvFp4ToPetitFp4Weights failed");
}
inline void RepackNvFp4ToPetitFp4Scales(unsigned *out_scales, const unsigned *scales,
                                        unsigned in_chan, unsigned out_chan, hipStream_t stream) {
    detail::check(cabi::petit_repack_nvfp4_scales(out_scales, scales, in_chan, out_chan, stream),
                  "RepackNvFp4ToPetitFp4Scales failed");
}
inline void RepackMxFp4ToPetitFp4Scales(unsigned *out_scales, const unsigned *scales,
                                        unsigned in_chan, unsigned out_chan, hipStream_t stream) {
    detail::check(cabi::petit_repack_mxfp4_scales(out_scales, scales, in_chan, out_chan, stream),
                  "RepackMxFp4ToPetitFp4Scales failed");
}

/* fp4/gemm_fp4.h:11-21 (default stream, like the reference) */
inline int DequantPetitFp4(unsigned *output, const unsigned *input, const unsigned *scales,
                           float global_scale, DataType out_type, unsigned k, unsigned n) {
    return cabi::petit_dequant_packed_nvfp4(output, input, scales, global_scale, (int)out_type, k, n,
                                      nullptr);
}
inline int DequantPetitMxFp4(unsigned *output, const unsigned *input, const unsigned *scales,
                             float global_scale, DataType out_type, unsigned k, unsigned n) {
    return cabi::petit_dequant_packed_mxfp4(output, input, scales, global_scale, (int)out_type, k, n,
                                      nullptr);
}
inline int DequantMxFp4(unsigned *output, const unsigned *input, const unsigned *scales,
                        float global_scale, DataType out_type, unsigned k, unsigned n) {
    return cabi::petit_dequant_mxfp4(output, input, scales, global_scale, (int)out_type, k, n, nullptr);
}
/* quantization_utils.cu:542-612 names the NVFP4 twin DequantNvFp4 */
inline int DequantNvFp4(unsigned *output, const unsigned *input, const unsigned *scales,
                        float global_scale, DataType out_type, unsigned k, unsigned n) {
    return cabi::petit_dequant_nvfp4(output, input, scales, global_scale, (int)out_type, k, n, nullptr);
}

} // namespace fp4
} // namespace causalflow::petit::rocm::quantization

/* lib/hal/device.h:8-34 with int status codes (0 = ok, else cudaError_t). */
namespace causalflow::petit::hal {

class Device {
  public:
    virtual int Malloc(void **ptr, size_t size) = 0;
    virtual int Free(void *ptr) = 0;
    virtual int Memset(void *ptr, int value, size_t size) = 0;
    virtual int CopyToDevice(void *dst, const void *src, size_t size) = 0;
    virtual int CopyToHost(void *dst, const void *src, size_t size) = 0;
    virtual int Synchronize() = 0;
    virtual ~Device() = default;

  protected:
    Device() = default;
};

class Platform {
  public:
    virtual int GetDevice(int id, std::unique_ptr<Device> *result) = 0;
    virtual ~Platform() = default;

  protected:
    Platform() = default;
};

namespace detail {
class CudaDevice final : public Device {
  public:
    explicit CudaDevice(int id) : id_(id) {}
    int Malloc(void **ptr, size_t size) override { return Bind() ? Bind() : cabi::petit_hal_malloc(ptr, size); }
    int Free(void *ptr) override { return Bind() ? Bind() : cabi::petit_hal_free(ptr); }
    int Memset(void *ptr, int value, size_t size) override {
        return Bind() ? Bind() : cabi::petit_hal_memset(ptr, value, size);
    }
    int CopyToDevice(void *dst, const void *src, size_t size) override {
        return Bind() ? Bind() : cabi::petit_hal_copy_to_device(dst, src, size);
    }
    int CopyToHost(void *dst, const void *src, size_t size) override {
        return Bind() ? Bind() : cabi::petit_hal_copy_to_host(dst, src, size);
    }
    int Synchronize() override { return Bind() ? Bind() : cabi::petit_hal_synchronize(); }

  private:
    int Bind() const { return cabi::petit_hal_set_device(id_); }
    int id_;
};
class CudaPlatform final : public Platform {
  public:
    int GetDevice(int id, std::unique_ptr<Device> *result) override {
        int count = 0;
        const int rc = cabi::petit_hal_device_count(&count);
        if (rc != 0) return rc;
        if (id < 0 || id >= count || !result) return 101; /* cudaErrorInvalidDevice */
        result->reset(new CudaDevice(id));
        return 0;
    }
};
} // namespace detail

/* The reference registers "rocm" (lib/hal/rocm/platform_rocm.cc:17-68); this backend answers to
 * "cuda" and, so that reference callers keep working unchanged, to "rocm" as well. */
inline Platform *GetPlatform(const char *name) {
    static detail::CudaPlatform platform;
    if (name && (!std::strcmp(name, "cuda") || !std::strcmp(name, "rocm"))) return &platform;
    return nullptr;
}

} // namespace causalflow::petit::hal

#endif /* CAUSALFLOW_PETIT_GEMM_COMPAT_H_ */
