// bench_matmul -- command-line benchmark / tuner with the flags and the result line of the
// reference's tools/benchmarks/matmul/main.cc (flags :17-35, result line :255-267, tune
// mode :269-325), over this repo's C ABI only (include/causalflow/petit/petit.h).
//
//   bench_matmul -m 16 -n 8192 -k 8192 -atype bf16 -ctype bf16 -btype nvfp4 -algo tune
//   bench_matmul -m 16 -n 8192 -k 8192 -atype bf16 -ctype bf16 -btype nvfp4 -algo <hex id>
//
// -algo ""      default solution (PETIT_SOLUTION_AUTO)
// -algo tune    time every solution petit_get_solutions lists, print the 5 fastest;
//               with -table FILE also append "<btype> <atype> m n k <hex>" for the fastest one,
//               the format PETIT_TUNE_TABLE / petit_tune_table_load feed to the default chooser
// -algo <hex>   the 8 bytes of a solution id as printed by tune mode
//
// Differences from the reference, on purpose: time is measured with CUDA events on the
// launching stream instead of the host clock, and by default the weights rotate over
// enough distinct copies that more than 2x the L2 is touched between reuses (-copies 1
// restores the reference's single hot buffer).  Only -backend petit exists.
#include "causalflow/petit/petit.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <map>
#include <random>
#include <string>
#include <vector>

namespace {

struct Flags {
    std::string backend = "petit", algo = "", atype = "fp16", ctype = "fp16", btype = "nvfp4";
    std::string table; // -algo tune: append the fastest solution here (petit_tune_table_load format)
    int m = 128, n = 4096, k = 4096, batch = 1, warmup = 10, repeat = 100, copies = 0;
};

bool parse_flags(int argc, char **argv, Flags *f) {
    std::map<std::string, std::string> kv;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.empty() || a[0] != '-') return false;
        a = a.substr(a.find_first_not_of('-'));
        const size_t eq = a.find('=');
        if (eq != std::string::npos) {
            kv[a.substr(0, eq)] = a.substr(eq + 1);
        } else if (i + 1 < argc) {
            kv[a] = argv[++i];
        } else {
            return false;
        }
    }
    for (auto &[key, v] : kv) {
        if (key == "backend") f->backend = v;
        else if (key == "algo") f->algo = v;
        else if (key == "atype") f->atype = v;
        else if (key == "ctype") f->ctype = v;
        else if (key == "btype") f->btype = v;
        else if (key == "m") f->m = std::atoi(v.c_str());
        else if (key == "n") f->n = std::atoi(v.c_str());
        else if (key == "k") f->k = std::atoi(v.c_str());
        else if (key == "batch") f->batch = std::atoi(v.c_str());
        else if (key == "warmup") f->warmup = std::atoi(v.c_str());
        else if (key == "repeat") f->repeat = std::atoi(v.c_str());
        else if (key == "copies") f->copies = std::atoi(v.c_str());
        else if (key == "table") f->table = v;
        else {
            std::fprintf(stderr, "Unknown flag: -%s\n", key.c_str());
            return false;
        }
    }
    return true;
}

#define CK(x)                                                                                 \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess) {                                                              \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, \
                         __LINE__);                                                           \
            std::exit(2);                                                                     \
        }                                                                                     \
    } while (0)

uint16_t f32_to_bf16(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
uint16_t f32_to_f16(float x) { // finite, |x| <= 2 only (what the generator draws)
    uint32_t u;
    std::memcpy(&u, &x, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int32_t e = (int32_t)((u >> 23) & 0xff) - 127 + 15;
    uint32_t m = u & 0x7fffffu;
    if (e <= 0) return (uint16_t)sign; // flush tiny values, irrelevant for timing
    uint32_t h = ((uint32_t)e << 10) | (m >> 13);
    if ((m & 0x1000u) && ((m & 0x2fffu) != 0)) ++h;
    return (uint16_t)(sign | h);
}

std::string hex_of(uint64_t id) { // the 8 bytes of SolutionId::Repr(), little endian
    char buf[17];
    for (int i = 0; i < 8; ++i) std::snprintf(buf + 2 * i, 3, "%02x", (unsigned)((id >> (8 * i)) & 0xff));
    return buf;
}
bool parse_hex(const std::string &s, uint64_t *id) {
    if (s.size() != 16) return false;
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) {
        unsigned b;
        if (std::sscanf(s.c_str() + 2 * i, "%2x", &b) != 1) return false;
        v |= (uint64_t)b << (8 * i);
    }
    *id = v;
    return true;
}

struct Problem {
    Flags f;
    bool mx = false;
    PetitSolutionHints hints{};
    void *a = nullptr, *c = nullptr;
    uint8_t *w = nullptr, *sc = nullptr;
    float *gs = nullptr;
    size_t wbytes = 0, sbytes = 0;
    int copies = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;

    int call(uint64_t id, int i) const {
        const uint8_t *wp = w + (size_t)(i % copies) * wbytes;
        const uint8_t *sp = sc + (size_t)(i % copies) * sbytes;
        return mx ? petit_gemm_mxfp4_a16(c, a, wp, sp, gs, f.m, f.n, f.k, &hints, id, stream)
                  : petit_gemm_nvfp4_a16(c, a, wp, sp, gs, f.m, f.n, f.k, &hints, id, stream);
    }
    // returns seconds for `repeat` launches, or < 0 on failure
    double run(uint64_t id) const {
        for (int i = 0; i < f.warmup; ++i)
            if (call(id, i) != PETIT_OK) return -1;
        CK(cudaStreamSynchronize(stream));
        CK(cudaEventRecord(e0, stream));
        for (int i = 0; i < f.repeat; ++i)
            if (call(id, i) != PETIT_OK) return -1;
        CK(cudaEventRecord(e1, stream));
        CK(cudaStreamSynchronize(stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        return ms * 1e-3;
    }
    void print(const std::string &algo, double seconds) const {
        const double ops = 2.0 * f.m * f.n * f.k * f.batch;
        std::printf("Matmul %dx%dx%d %s:%s. Backend: %s, batch: %d, algorithm: %s, %d times total "
                    "%.6f ms. %.4f TFLOPS\n",
                    f.m, f.n, f.k, f.atype.c_str(), f.ctype.c_str(), f.backend.c_str(), f.batch,
                    algo.c_str(), f.repeat, seconds * 1e3, ops * f.repeat / seconds / 1e12);
    }
};

} // namespace

int main(int argc, char **argv) {
    Flags f;
    if (!parse_flags(argc, argv, &f)) {
        std::fprintf(stderr,
                     "usage: bench_matmul [-backend petit] -m M -n N -k K [-warmup W] [-repeat R] "
                     "[-algo ''|tune|<hex>] [-atype fp16|bf16] [-ctype fp16|bf16] "
                     "[-btype nvfp4|mxfp4] [-batch 1] [-copies C] [-table FILE]\n");
        return 1;
    }
    if (f.backend != "petit") {
        std::fprintf(stderr, "Unknown backend: %s\n", f.backend.c_str());
        return 1;
    }
    if (f.btype != "nvfp4" && f.btype != "mxfp4") {
        std::fprintf(stderr, "Invalid b type for backend 'petit': %s. Supported: nvfp4, mxfp4\n",
                     f.btype.c_str());
        return 1;
    }
    if ((f.atype != "fp16" && f.atype != "bf16") || f.ctype != f.atype) {
        std::fprintf(stderr, "Invalid data type: a=%s c=%s. Supported: fp16, bf16 (a == c)\n",
                     f.atype.c_str(), f.ctype.c_str());
        return 1;
    }
    if (f.batch != 1) {
        std::fprintf(stderr, "petit backend: only -batch 1 is supported\n");
        return 1;
    }
    if (f.m <= 0 || f.n <= 0 || f.k <= 0 || f.k % 256 != 0 || f.n % 16 != 0) {
        std::fprintf(stderr, "petit backend needs k %% 256 == 0 and n %% 16 == 0\n");
        return 1;
    }

    Problem p;
    p.f = f;
    p.mx = f.btype == "mxfp4";
    const bool bf16 = f.atype == "bf16";
    p.hints.a_type = p.hints.c_type = bf16 ? PETIT_DTYPE_BF16 : PETIT_DTYPE_FP16;
    p.hints.b_type = p.mx ? PETIT_DTYPE_MXFP4_E2M1 : PETIT_DTYPE_FP4_E2M1;
    p.hints.require_high_precision = 0;

    const size_t nk = (size_t)f.n * f.k;
    p.wbytes = nk / 2;
    p.sbytes = p.mx ? nk / 32 : nk / 16;
    p.copies = f.copies > 0 ? f.copies
                            : (int)std::max<size_t>(2, (size_t)300e6 / (p.wbytes + p.sbytes) + 2);
    CK(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&p.e0));
    CK(cudaEventCreate(&p.e1));

    // inputs: A ~ U(-2, 2) (gemm_fp4_fp16_rocm_test.cc generator), random fp4 words,
    // positive scales, global_scale = 1
    std::mt19937 rng(42);
    std::uniform_real_distribution<float> ua(-2.f, 2.f);
    std::vector<uint16_t> ha((size_t)f.m * f.k);
    for (auto &v : ha) {
        const float x = ua(rng);
        v = bf16 ? f32_to_bf16(x) : f32_to_f16(x);
    }
    std::vector<uint32_t> hw(p.wbytes / 4);
    std::vector<uint8_t> hs(p.sbytes);
    CK(cudaMalloc(&p.a, ha.size() * 2));
    CK(cudaMalloc(&p.c, (size_t)f.m * f.n * 2));
    CK(cudaMalloc(&p.w, p.wbytes * p.copies));
    CK(cudaMalloc(&p.sc, p.sbytes * p.copies));
    CK(cudaMalloc(&p.gs, 4));
    const float one = 1.f;
    CK(cudaMemcpy(p.gs, &one, 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p.a, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    uint32_t *d_native_w;
    uint8_t *d_native_s;
    CK(cudaMalloc(&d_native_w, p.wbytes));
    CK(cudaMalloc(&d_native_s, p.sbytes));
    for (int cp = 0; cp < p.copies; ++cp) {
        for (auto &v : hw) v = rng();
        for (auto &v : hs)
            v = p.mx ? (uint8_t)(108 + rng() % 16)            // 2^-19 .. 2^-4
                     : (uint8_t)(0x20 + rng() % 0x30);        // e4m3 2^-3 .. ~13
        CK(cudaMemcpy(d_native_w, hw.data(), p.wbytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_native_s, hs.data(), p.sbytes, cudaMemcpyHostToDevice));
        int rc = petit_repack_fp4_weights((uint32_t *)(p.w + (size_t)cp * p.wbytes), d_native_w, f.k,
                                          f.n, p.stream);
        rc |= p.mx ? petit_repack_mxfp4_scales(p.sc + (size_t)cp * p.sbytes, d_native_s, f.k, f.n,
                                               p.stream)
                   : petit_repack_nvfp4_scales(p.sc + (size_t)cp * p.sbytes, d_native_s, f.k, f.n,
                                               p.stream);
        if (rc != PETIT_OK) {
            std::fprintf(stderr, "repack failed (%d)\n", rc);
            return 2;
        }
        CK(cudaStreamSynchronize(p.stream));
    }
    cudaFree(d_native_w);
    cudaFree(d_native_s);

    if (f.algo == "tune") {
        unsigned count = 0;
        if (petit_get_solutions(&p.hints, f.m, f.n, f.k, nullptr, &count) != 0) {
            std::fprintf(stderr, "Failed to tune the GEMM: unsupported types\n");
            return 2;
        }
        std::vector<uint64_t> sols(count);
        petit_get_solutions(&p.hints, f.m, f.n, f.k, sols.data(), &count);
        sols.resize(count);
        std::printf("Finished enumerating %u algorithms\n", count);
        std::vector<std::pair<double, uint64_t>> results;
        for (uint64_t id : sols) {
            const double t = p.run(id);
            if (t < 0) {
                std::fprintf(stderr, "Failed to run the matmul for repr: %s\n", hex_of(id).c_str());
                continue;
            }
            results.emplace_back(t, id);
        }
        std::sort(results.begin(), results.end());
        for (size_t i = 0; i < results.size() && i < 5; ++i)
            p.print(hex_of(results[i].second), results[i].first);
        if (!f.table.empty() && !results.empty()) {
            // one line per tuned problem; PETIT_TUNE_TABLE=<file> / petit_tune_table_load feed
            // it back to the default chooser
            if (std::FILE *tf = std::fopen(f.table.c_str(), "a")) {
                std::fprintf(tf, "%s %s %d %d %d %s\n", f.btype.c_str(), f.atype.c_str(), f.m, f.n,
                             f.k, hex_of(results[0].second).c_str());
                std::fclose(tf);
            } else {
                std::fprintf(stderr, "cannot append to %s\n", f.table.c_str());
                return 2;
            }
        }
        return results.empty() ? 2 : 0;
    }

    uint64_t id = PETIT_SOLUTION_AUTO;
    if (!f.algo.empty() && !parse_hex(f.algo, &id)) {
        std::fprintf(stderr, "Invalid -algo: expected 'tune' or 16 hex digits\n");
        return 1;
    }
    const double t = p.run(id);
    if (t < 0) {
        std::fprintf(stderr, "Failed to run the matmul\n");
        return 2;
    }
    p.print(f.algo, t);
    return 0;
}
