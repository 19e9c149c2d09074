// A translation unit written the way a user of the reference's C++ interface writes it
// (lib/gemm/rocm/quantization/gemm.h, fp4/gemm_fp4.h, lib/hal/device.h): it must compile against
// include/causalflow/petit/gemm_compat.h and link against libpetit_b200.so.  Host-only calls run
// without a GPU; with "gpu" as argv[1] it also runs a small GEMM through the shim.
#include "causalflow/petit/gemm_compat.h"

#include <cstdio>
#include <cstring>
#include <vector>

using namespace causalflow::petit::rocm::quantization;
namespace hal = causalflow::petit::hal;

int main(int argc, char **argv) {
    PetitSolutionHints hints{DataType::kDataTypeBf16, DataType::kDataTypeFp4e2m1, DataType::kDataTypeBf16,
                             false};
    unsigned n_sols = 0;
    if (fp4::GemmGetSolutions(hints, 16, 4096, 4096, nullptr, &n_sols) != 0 || n_sols == 0) return 1;
    std::vector<SolutionId> sols(n_sols);
    if (fp4::GemmGetSolutions(hints, 16, 4096, 4096, sols.data(), &n_sols) != 0) return 2;
    for (const SolutionId &s : sols)
        if (s.features() != kMatmulFeatures_Grid || s.element_b() != kMatmulTypeBNvFp4 ||
            s.mfma_type() != kMatmulMfmaTypeBf16 || SolutionId::FromRepr(s.Repr()).repr != s.repr)
            return 3;
    // an MFMA tile shape of the reference is a well-formed id that this backend does not have
    const SolutionId ref_default = SolutionId::MultiStage(
        kMatmulFeatures_Grid, kMatmulTypeBNvFp4, kMatmulMfmaTypeFp16, 1, 4, 8, kMatmulWarpPartition_NK, 1, 2, 2);
    if (ref_default.tile_k() != 2) return 4;
    // m == 0 is a no-op that succeeds (gemm_fp4_fp16_grid.cc:42-44)
    if (fp4::GemmFp4Fp16Grid(nullptr, nullptr, nullptr, nullptr, nullptr, 0, 4096, 4096, hints, -1ul, nullptr) != 0)
        return 5;
    hints.b_type = DataType::kDataTypeInt4; // unsupported b_type -> -1 (algo_chooser.cc:20-23)
    if (fp4::GemmGetSolutions(hints, 16, 4096, 4096, nullptr, &n_sols) != -1) return 6;
    if (hal::GetPlatform("nope") != nullptr || hal::GetPlatform("cuda") == nullptr) return 7;
    std::printf("compat host checks ok (%zu solutions)\n", sols.size());
    if (argc > 1 && !std::strcmp(argv[1], "gpu")) {
        std::unique_ptr<hal::Device> dev;
        if (hal::GetPlatform("rocm")->GetDevice(0, &dev) != 0) return 10;
        const unsigned m = 16, n = 256, k = 512;
        void *a, *b, *bp, *s, *sp, *c, *gs;
        dev->Malloc(&a, m * k * 2); dev->Malloc(&b, n * k / 2); dev->Malloc(&bp, n * k / 2);
        dev->Malloc(&s, n * k / 16); dev->Malloc(&sp, n * k / 16); dev->Malloc(&c, m * n * 2);
        dev->Malloc(&gs, 4);
        dev->Memset(a, 0, m * k * 2); dev->Memset(b, 0x22, n * k / 2); dev->Memset(s, 0x38, n * k / 16);
        const float one = 1.0f;
        dev->CopyToDevice(gs, &one, 4);
        fp4::RepackNvFp4ToPetitFp4Weights((unsigned *)bp, (const unsigned *)b, k, n, nullptr);
        fp4::RepackNvFp4ToPetitFp4Scales((unsigned *)sp, (const unsigned *)s, k, n, nullptr);
        hints.b_type = DataType::kDataTypeFp4e2m1;
        const int rc = fp4::GemmFp4Fp16Grid((unsigned *)c, (const unsigned *)a, (const unsigned *)bp,
                                            (const unsigned *)sp, (const float *)gs, m, n, k, hints,
                                            sols[0].Repr(), nullptr);
        if (rc != 0 || dev->Synchronize() != 0) return 11;
        std::vector<unsigned short> host(m * n, 1);
        dev->CopyToHost(host.data(), c, m * n * 2);
        for (unsigned short v : host)
            if ((v & 0x7fff) != 0) return 12; // zero activations -> zero output
        std::printf("compat gpu GEMM ok\n");
    }
    return 0;
}
