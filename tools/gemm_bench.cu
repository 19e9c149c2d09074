// Kernel-only timing of the C ABI GEMM on synthetic packed data (no parity
// check -- tests/ does that).  Rotates over enough distinct weight copies to
// defeat the 126 MB L2.  Usage: gemm_bench [nv|mx] [bf16|f16] [reps]
#include "causalflow/petit/petit.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

__global__ void fill_kernel(uint32_t *p, size_t n, uint32_t seed, uint32_t mask, uint32_t orv) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u + seed;
        x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; x *= 0x297a2d39u; x ^= x >> 15;
        p[i] = (x & mask) | orv;
    }
}

extern "C" void petit_debug_set_trace(unsigned long long *);

int main(int argc, char **argv) {
    bool mx = argc > 1 && !strcmp(argv[1], "mx");
    // "f16" = fp16 activations on default-layout weights, "f16n" = on the fp16-native layout
    const bool f16n = argc > 2 && !strcmp(argv[2], "f16n");
    bool bf16 = !(argc > 2 && (!strcmp(argv[2], "f16") || f16n));
    int reps = argc > 3 ? atoi(argv[3]) : 40;
    const char *only = argc > 4 ? argv[4] : "";
    struct Shape { const char *name; unsigned n, k; };
    // the four 70B shapes, their TP-8 shards, or any "NxK" given on the command line
    std::vector<Shape> shapes = {{"qkv", 10240, 8192},    {"o", 8192, 8192},
                                 {"gate_up", 57344, 8192}, {"down", 8192, 28672},
                                 {"qkv_tp8", 1280, 8192},  {"o_tp8", 8192, 1024},
                                 {"gate_up_tp8", 7168, 8192}, {"down_tp8", 8192, 3584}};
    unsigned cn = 0, ck = 0;
    if (sscanf(only, "%ux%u", &cn, &ck) == 2) shapes.push_back({only, cn, ck});
    std::vector<unsigned> ms = {1, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192};
    if (argc > 5) { ms.clear(); ms.push_back(atoi(argv[5])); }
    const double hbm_peak = 6535.7, tf_peak = 1600.2;
    float *d_gs;
    float one = 1.0f;
    CK(cudaMalloc(&d_gs, 4));
    CK(cudaMemcpy(d_gs, &one, 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (const Shape &s : shapes) {
        if (only[0] && strcmp(only, s.name) && strcmp(only, "all")) continue;
        const size_t wbytes = (size_t)s.n * s.k / 2, sbytes = (size_t)s.n * s.k / (mx ? 32 : 16);
        int copies = (int)std::max<size_t>(2, (size_t)400e6 / (wbytes + sbytes) + 1);
        if (getenv("PETIT_COPIES")) copies = atoi(getenv("PETIT_COPIES"));
        uint8_t *w, *sc;
        CK(cudaMalloc(&w, wbytes * copies));
        CK(cudaMalloc(&sc, sbytes * copies));
        fill_kernel<<<1184, 256>>>((uint32_t *)w, wbytes * copies / 4, 1, 0xffffffffu, 0);
        // scales: NV E5M3 bytes with e5 in [8,15] ; MX e8m0 in [108,123] (2^-19..2^-4)
        if (mx) fill_kernel<<<1184, 256>>>((uint32_t *)sc, sbytes * copies / 4, 2, 0x0f0f0f0fu, 0x6c6c6c6cu); // e8m0 108..123
        else fill_kernel<<<1184, 256>>>((uint32_t *)sc, sbytes * copies / 4, 2, 0x3f3f3f3fu, 0x40404040u);
        for (unsigned m : ms) {
            uint16_t *a, *c;
            CK(cudaMalloc(&a, (size_t)m * s.k * 2));
            CK(cudaMalloc(&c, (size_t)m * s.n * 2));
            // activations: small bf16/f16 values (exponent field mid-range)
            fill_kernel<<<1184, 256>>>((uint32_t *)a, (size_t)m * s.k / 2, 3,
                                       bf16 ? 0x807f807fu : 0x83ff83ffu, bf16 ? 0x3f003f00u : 0x38003800u);
            CK(cudaDeviceSynchronize());
            int t = bf16 ? PETIT_DTYPE_BF16 : PETIT_DTYPE_FP16;
            PetitSolutionHints hints = {t, mx ? PETIT_DTYPE_MXFP4_E2M1 : PETIT_DTYPE_FP4_E2M1, t, 0};
            auto call = [&](int i) {
                const uint8_t *wp = w + (size_t)(i % copies) * wbytes;
                const uint8_t *sp = sc + (size_t)(i % copies) * sbytes;
                PetitEpilogue epi = {nullptr, nullptr, PETIT_ACT_NONE,
                                     f16n ? PETIT_WEIGHT_LAYOUT_F16_NATIVE : PETIT_WEIGHT_LAYOUT_DEFAULT};
                int rc = mx ? petit_gemm_mxfp4_a16(c, a, wp, sp, d_gs, m, s.n, s.k, &hints,
                                                   PETIT_SOLUTION_AUTO, nullptr)
                            : petit_gemm_nvfp4_a16_ex(c, a, wp, sp, d_gs, m, s.n, s.k, &hints,
                                                      PETIT_SOLUTION_AUTO, &epi, nullptr, nullptr);
                if (rc) { printf("gemm rc=%d\n", rc); exit(1); }
            };
            int r = m >= 1024 ? std::max(4, reps / 8) : reps;
            for (int i = 0; i < 5; ++i) call(i);
            CK(cudaDeviceSynchronize());
            // back-to-back launches (includes launch gaps), then per-launch events
            CK(cudaEventRecord(e0));
            for (int i = 0; i < r; ++i) call(i);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms_total;
            cudaEventElapsedTime(&ms_total, e0, e1);
            double us = ms_total * 1e3 / r;
            double us_graph = 0;
            if (getenv("PETIT_GRAPH")) {
                cudaStream_t st;
                CK(cudaStreamCreate(&st));
                // warm the per-stream workspace outside capture
                int t2 = bf16 ? PETIT_DTYPE_BF16 : PETIT_DTYPE_FP16;
                PetitSolutionHints h2 = {t2, mx ? PETIT_DTYPE_MXFP4_E2M1 : PETIT_DTYPE_FP4_E2M1, t2, 0};
                auto call_s = [&](int i) {
                    const uint8_t *wp = w + (size_t)(i % copies) * wbytes;
                    const uint8_t *sp = sc + (size_t)(i % copies) * sbytes;
                    int rc = mx ? petit_gemm_mxfp4_a16(c, a, wp, sp, d_gs, m, s.n, s.k, &h2, PETIT_SOLUTION_AUTO, (petit_stream_t)st)
                                : petit_gemm_nvfp4_a16(c, a, wp, sp, d_gs, m, s.n, s.k, &h2, PETIT_SOLUTION_AUTO, (petit_stream_t)st);
                    if (rc) { printf("gemm rc=%d\n", rc); exit(1); }
                };
                call_s(0);
                CK(cudaStreamSynchronize(st));
                cudaGraph_t graph;
                cudaGraphExec_t exec;
                CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                for (int i = 0; i < r; ++i) call_s(i);
                CK(cudaStreamEndCapture(st, &graph));
                CK(cudaGraphInstantiate(&exec, graph, 0));
                CK(cudaGraphLaunch(exec, st));
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(e0, st));
                CK(cudaGraphLaunch(exec, st));
                CK(cudaEventRecord(e1, st));
                CK(cudaStreamSynchronize(st));
                float msg;
                cudaEventElapsedTime(&msg, e0, e1);
                us_graph = msg * 1e3 / r;
                printf("  [graph replay of %d launches: %.2f us/launch, %.0f GB/s]\n", r, us_graph,
                       ((double)wbytes + sbytes + 2.0 * m * s.k + 2.0 * m * s.n + 4) / us_graph * 1e-3);
                cudaGraphExecDestroy(exec);
                cudaGraphDestroy(graph);
                cudaStreamDestroy(st);
            }
            if (getenv("PETIT_TRACE")) {
                unsigned long long *d_tr;
                CK(cudaMalloc(&d_tr, (160 * 16 + 64 * 8 + 160) * 8));
                CK(cudaMemset(d_tr, 0, (160 * 16 + 64 * 8 + 160) * 8));
                petit_debug_set_trace(d_tr);
                for (int i = 0; i < 4; ++i) call(i); // back-to-back: the last launch's stamps survive
                CK(cudaDeviceSynchronize());
                petit_debug_set_trace(nullptr);
                std::vector<unsigned long long> tr(160 * 16 + 64 * 8);
                CK(cudaMemcpy(tr.data(), d_tr, tr.size() * 8, cudaMemcpyDeviceToHost));
                unsigned long long t0 = ~0ull;
                int nb = 0;
                for (int b = 0; b < 160; ++b) if (tr[b * 16]) { t0 = std::min(t0, tr[b * 16]); ++nb; }
                const char *names[9] = {"entry", "setup_done", "griddep_wait_done", "first_stage_landed",
                                        "dequant_done", "mma_issued_all", "last_acc_full", "epilogue_done", "exit"};
                printf("  trace over %d CTAs (us since first CTA entry): event min/avg/max\n", nb);
                for (int e = 0; e < 9; ++e) {
                    double mn = 1e30, mx2 = 0, sum = 0; int cnt = 0;
                    for (int b = 0; b < 160; ++b) {
                        unsigned long long v = tr[b * 16 + e];
                        if (!v || !tr[b * 16]) continue;
                        double d = (double)(v - t0) * 1e-3;
                        mn = std::min(mn, d); mx2 = std::max(mx2, d); sum += d; ++cnt;
                    }
                    if (cnt) printf("    %-20s %7.2f %7.2f %7.2f\n", names[e], mn, sum / cnt, mx2);
                }
                if (getenv("PETIT_TRACE_STAGES")) {
                    printf("  CTA0 per-stage (us since CTA0 entry): w_issue act_issue | dq:full a_empty st_done | mma:a_full committed\n");
                    unsigned long long c0 = tr[0];
                    for (int st = 0; st < 24; ++st) {
                        printf("   st%02d", st);
                        for (int e = 0; e < 7; ++e) {
                            unsigned long long v = tr[160 * 16 + st * 8 + e];
                            if (v) printf(" %7.2f", (double)(v - c0) * 1e-3); else printf("       -");
                        }
                        printf("\n");
                    }
                }
                cudaFree(d_tr);
            }
            if (getenv("PETIT_TRACE2")) {
                // two consecutive launches stamped into separate buffers on the common
                // globaltimer base: shows how launch i+1 overlaps the tail of launch i
                const size_t words = 160 * 16 + 64 * 8 + 160; // + one SM id per CTA
                unsigned long long *d_tr[2];
                for (int j = 0; j < 2; ++j) {
                    CK(cudaMalloc(&d_tr[j], words * 8));
                    CK(cudaMemset(d_tr[j], 0, words * 8));
                }
                call(0);
                call(1);
                petit_debug_set_trace(d_tr[0]);
                call(2);
                petit_debug_set_trace(d_tr[1]);
                call(3);
                petit_debug_set_trace(nullptr);
                call(4);
                CK(cudaDeviceSynchronize());
                std::vector<unsigned long long> tr[2];
                unsigned long long t0 = ~0ull;
                for (int j = 0; j < 2; ++j) {
                    tr[j].resize(words);
                    CK(cudaMemcpy(tr[j].data(), d_tr[j], words * 8, cudaMemcpyDeviceToHost));
                }
                for (int b = 0; b < 160; ++b) if (tr[0][b * 16]) t0 = std::min(t0, tr[0][b * 16]);
                const char *names[16] = {"entry", "setup_done", "griddep_wait_done", "first_stage_landed",
                                        "dequant_done", "mma_issued_all", "last_acc_full", "epilogue_done", "exit",
                                        "seg1_epilogue_begin", "seg1_epilogue_end", "lastseg_epi_begin", "lastseg_polled",
                                        "lastseg_prefetched", "lastseg_ldtm0_done", "lastseg_group0_stored"};
                for (int j = 0; j < 2; ++j) {
                    printf("  launch %d (us since launch-0 first entry): event min/avg/max\n", j);
                    for (int e = 0; e < 16; ++e) {
                        double mn = 1e30, mx2 = 0, sum = 0; int cnt = 0;
                        for (int b = 0; b < 160; ++b) {
                            unsigned long long v = tr[j][b * 16 + e];
                            if (!v || !tr[j][b * 16]) continue;
                            double d = (double)((long long)(v - t0)) * 1e-3;
                            mn = std::min(mn, d); mx2 = std::max(mx2, d); sum += d; ++cnt;
                        }
                        if (cnt) printf("    %-20s %7.2f %7.2f %7.2f\n", names[e], mn, sum / cnt, mx2);
                    }
                    cudaFree(d_tr[j]);
                }
                // PETIT_TRACE_DUMP=<file>: one CSV row per CTA and launch (all 16 stamps in us
                // since launch-0 first entry, 0 = not stamped) for offline straggler analysis
                if (const char *dump = getenv("PETIT_TRACE_DUMP")) {
                    FILE *f = fopen(dump, "a");
                    if (f) {
                        fprintf(f, "# %s %s %s M=%u N=%u K=%u: shape,m,launch,cta,smid", s.name,
                                mx ? "mx" : "nv", bf16 ? "bf16" : "f16", m, s.n, s.k);
                        for (int e = 0; e < 16; ++e) fprintf(f, ",%s", names[e]);
                        fprintf(f, "\n");
                        for (int j = 0; j < 2; ++j)
                            for (int b = 0; b < 160; ++b) {
                                if (!tr[j][b * 16]) continue;
                                fprintf(f, "%s,%u,%d,%d,%llu", s.name, m, j, b,
                                        tr[j][160 * 16 + 64 * 8 + b]);
                                for (int e = 0; e < 16; ++e) {
                                    unsigned long long v = tr[j][b * 16 + e];
                                    fprintf(f, ",%.2f", v ? (double)((long long)(v - t0)) * 1e-3 : 0.0);
                                }
                                fprintf(f, "\n");
                            }
                        fclose(f);
                    }
                }
            }
            double bytes = (double)wbytes + sbytes + 2.0 * m * s.k + 2.0 * m * s.n + 4;
            double flops = 2.0 * m * s.n * s.k;
            printf("%-7s %s %s M=%-5u  %9.2f us  %7.0f GB/s (%5.1f%% of %.0f)  %7.1f TFLOPS (%5.1f%% of %.0f)\n",
                   s.name, mx ? "mx" : "nv", bf16 ? "bf16" : "f16 ", m, us, bytes / us * 1e-3,
                   bytes / us * 1e-3 / hbm_peak * 100, hbm_peak, flops / us * 1e-6,
                   flops / us * 1e-6 / tf_peak * 100, tf_peak);
            cudaFree(a);
            cudaFree(c);
        }
        cudaFree(w);
        cudaFree(sc);
    }
    return 0;
}
