// Calibration row for the decode GEMMs: how fast can 148 persistent CTAs stream the same
// bytes through the same kind of shared-memory ring, launch after launch, chained by
// programmatic dependent launch -- with no dequantisation, no MMA, no reduction?  The
// difference between this floor and the GEMM is what the kernel's own work costs; the
// difference between this floor and bytes / HBM peak is what a launch costs on this part.
//
//   tools/stream_floor [reps]      (build: __graft_entry__.build() or
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/stream_floor tools/stream_floor.cu)
//
// Per launch: every CTA copies its contiguous share of `bytes` with cp.async.bulk into a ring
// of kStages x kStageBytes (the GEMM's 6 x 18 KB), a consumer warp frees the stages as they
// land; like the GEMM, the copies do not wait for the previous grid (weights are constants),
// and like the GEMM's token tile + epilogue the CTA then waits for the grid dependency
// (griddepcontrol.wait), reads 8 KB that the previous launch wrote and writes 8 KB of output.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int kStages = 6;
constexpr uint32_t kStageBytes = 18432; // one 128 x 256 unit: 16 KB of fp4 + 2 KB of scales
constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
stream_kernel(const uint8_t *__restrict__ src, uint64_t units, const float *__restrict__ dep_in,
              float *__restrict__ dep_out, int use_pdl) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + kStages;
    uint8_t *ring = smem + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (use_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    __syncthreads();
    const uint64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) { // producer
        for (uint64_t u = u0, it = 0; u < u1; ++u, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            if (it >= kStages) mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], kStageBytes);
            bulk_g2s(ring + (size_t)s * kStageBytes, src + u * kStageBytes, kStageBytes, &full[s]);
        }
    } else if (threadIdx.x == 32) { // consumer: a stage is free as soon as it has landed
        for (uint64_t u = u0, it = 0; u < u1; ++u, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            mbar_wait(&full[s], ph);
            mbar_arrive(&empty[s]);
        }
    }
    __syncthreads();
    // the dependent part of a GEMM launch: token tile in, output tile out
    if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    float acc = 0.f;
    for (int i = threadIdx.x; i < 2048; i += kThreads) acc += dep_in[(size_t)blockIdx.x * 2048 + i];
    for (int i = threadIdx.x; i < 2048; i += kThreads) dep_out[(size_t)blockIdx.x * 2048 + i] = acc + (float)i;
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 40;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double hbm_peak = 6535.7;
    struct Shape { const char *name; uint64_t n, k; };
    const Shape shapes[] = {{"qkv", 10240, 8192}, {"o", 8192, 8192}, {"gate_up", 57344, 8192}, {"down", 8192, 28672},
                            {"qkv_tp8", 1280, 8192}, {"gate_up_tp8", 7168, 8192}};
    const size_t smem = 1024 + (size_t)kStages * kStageBytes;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float *dep[2];
    CK(cudaMalloc(&dep[0], (size_t)sms * 2048 * 4));
    CK(cudaMalloc(&dep[1], (size_t)sms * 2048 * 4));
    CK(cudaMemset(dep[0], 0, (size_t)sms * 2048 * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (const Shape &s : shapes) {
        const uint64_t units = s.n / 128 * (s.k / 256), bytes = units * kStageBytes;
        const int copies = (int)std::max<uint64_t>(2, (uint64_t)400e6 / bytes + 1); // > 2x L2 between reuses
        uint8_t *w;
        CK(cudaMalloc(&w, bytes * copies));
        CK(cudaMemset(w, 1, bytes * copies));
        for (int pdl = 1; pdl >= 0; --pdl) {
            auto launch = [&](int i) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)std::min<uint64_t>(sms, units));
                cfg.blockDim = dim3(kThreads);
                cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = pdl;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                CK(cudaLaunchKernelEx(&cfg, stream_kernel, (const uint8_t *)(w + (size_t)(i % copies) * bytes), units,
                                      (const float *)dep[i & 1], dep[(i + 1) & 1], pdl));
            };
            for (int i = 0; i < 5; ++i) launch(i);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < reps; ++i) launch(i);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1e3 / reps;
            printf("%-12s %6.1f MB  %s  %7.2f us/launch  %6.0f GB/s (%5.1f%% of %.0f)   bytes/peak = %6.2f us\n", s.name,
                   bytes / 1e6, pdl ? "PDL-chained" : "serialised ", us, bytes / us * 1e-3,
                   bytes / us * 1e-3 / hbm_peak * 100, hbm_peak, bytes / hbm_peak * 1e-3);
        }
        CK(cudaFree(w));
    }
    return 0;
}
