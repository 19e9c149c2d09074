// Hardware fact-finding for the FP4 x 16-bit GEMM design.
// Answers, on a real B200:
//   cvt   : nibble/byte order of cvt.rn.f16x2.e2m1x2 / e4m3x2 / ue8m0x2,
//           HMUL2.BF16 behaviour on subnormal inputs
//   mma   : TS-form tcgen05.mma (A in TMEM written by tcgen05.st, B in smem
//           128B-swizzled) correctness, incl. MIXED A=f16 x B=bf16
//   tput  : issue throughput of candidate dequant instruction sequences
//   bw    : cp.async.bulk streaming bandwidth over all SMs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe probe.cu
#include "../../petit-kernel_b200/csrc/sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <vector>

using namespace petit::ptx;

#define CK(x)                                                                  \
    do {                                                                       \
        cudaError_t e_ = (x);                                                  \
        if (e_ != cudaSuccess) {                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_),         \
                   __FILE__, __LINE__);                                        \
            exit(1);                                                           \
        }                                                                      \
    } while (0)

// ---------------------------------------------------------------- cvt probe
__global__ void cvt_kernel(uint32_t *out) {
    unsigned t = threadIdx.x; // 0..255 : byte value
    uint32_t r0, r1, r2, r3;
    cvt_e2m1x8_to_f16x2x4(t | ((t ^ 0xff) << 8) | (0x21u << 16) | (0x43u << 24),
                          r0, r1, r2, r3);
    out[t * 8 + 0] = r0;
    out[t * 8 + 1] = r1;
    out[t * 8 + 2] = r2;
    out[t * 8 + 3] = r3;
    uint32_t lo, hi;
    cvt_e4m3x4_to_f16x2x2(t | (0x38u << 8) | (0x40u << 16) | (0x48u << 24), lo,
                          hi);
    out[t * 8 + 4] = lo;
    out[t * 8 + 5] = hi;
    cvt_ue8m0x4_to_bf16x2x2(t | (127u << 8) | (128u << 16) | (126u << 24), lo,
                            hi);
    out[t * 8 + 6] = lo;
    out[t * 8 + 7] = hi;
}

__global__ void bf16_subnormal_kernel(uint32_t *out) {
    // x = e2m1 magnitude code placed at bf16 bits [8:6] => value * 2^-126
    unsigned mag = threadIdx.x & 7;
    uint32_t x = (mag << 6) | (mag << 22);
    __nv_bfloat162 xv = *reinterpret_cast<__nv_bfloat162 *>(&x);
    // multiply by 2^119 * 1.75 (largest NV scale*2^119) and by 2^126
    __nv_bfloat162 s1 = __float2bfloat162_rn(ldexpf(1.0f, 126));
    __nv_bfloat162 s2 = __float2bfloat162_rn(ldexpf(1.75f, 119));
    __nv_bfloat162 p1 = __hmul2(xv, s1);
    __nv_bfloat162 p2 = __hmul2(xv, s2);
    out[threadIdx.x * 2 + 0] = *reinterpret_cast<uint32_t *>(&p1);
    out[threadIdx.x * 2 + 1] = *reinterpret_cast<uint32_t *>(&p2);
}

static float bf16_bits_to_float(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static void run_cvt() {
    uint32_t *d;
    CK(cudaMalloc(&d, 256 * 8 * 4));
    cvt_kernel<<<1, 256>>>(d);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> h(256 * 8);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    printf("== cvt.rn.f16x2.e2m1x2: byte -> (lo half, hi half)\n");
    for (int t : {0x00, 0x01, 0x10, 0x21, 0x7f, 0x8f, 0xf8, 0x99}) {
        __half2 v = *reinterpret_cast<__half2 *>(&h[t * 8]);
        printf("  byte 0x%02x -> lo=%g hi=%g (raw %08x); byte1(0x%02x)-> %08x "
               "byte2(0x21)->%08x byte3(0x43)->%08x\n",
               t, __low2float(v), __high2float(v), h[t * 8], t ^ 0xff,
               h[t * 8 + 1], h[t * 8 + 2], h[t * 8 + 3]);
    }
    printf("== e2m1 full table (low nibble):");
    for (int t = 0; t < 16; ++t) {
        __half2 v = *reinterpret_cast<__half2 *>(&h[t * 8]);
        printf(" %g", __low2float(v));
    }
    printf("\n== cvt.rn.f16x2.e4m3x2 (byte0=t, byte1=0x38=1.0; hi: 0x40=2, "
           "0x48=4)\n");
    for (int t : {0x00, 0x01, 0x08, 0x38, 0x7e, 0x7f, 0x80, 0xb8, 0xff}) {
        __half2 a = *reinterpret_cast<__half2 *>(&h[t * 8 + 4]);
        __half2 b = *reinterpret_cast<__half2 *>(&h[t * 8 + 5]);
        printf("  e4m3 0x%02x -> lo=%g hi=%g | second: lo=%g hi=%g\n", t,
               __low2float(a), __high2float(a), __low2float(b),
               __high2float(b));
    }
    printf("== cvt.rn.bf16x2.ue8m0x2 (byte0=t, byte1=127; hi: 128, 126)\n");
    for (int t : {0, 1, 2, 126, 127, 128, 237, 253, 254, 255}) {
        uint32_t a = h[t * 8 + 6], b = h[t * 8 + 7];
        printf("  ue8m0 %3d -> lo=%g (0x%04x) hi=%g | second: lo=%g hi=%g\n", t,
               bf16_bits_to_float(a & 0xffff), a & 0xffff,
               bf16_bits_to_float(a >> 16), bf16_bits_to_float(b & 0xffff),
               bf16_bits_to_float(b >> 16));
    }
    bf16_subnormal_kernel<<<1, 8>>>(d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d, 16 * 4, cudaMemcpyDeviceToHost));
    printf("== HMUL2.BF16 on (mag<<6) i.e. value*2^-126 (subnormal for mag=1): "
           "expect value, value*1.75*2^-7\n");
    for (int m = 0; m < 8; ++m)
        printf("  mag %d: x*2^126 = %g   x*1.75*2^119 = %g\n", m,
               bf16_bits_to_float(h[m * 2] & 0xffff),
               bf16_bits_to_float(h[m * 2 + 1] & 0xffff));
    CK(cudaFree(d));
}

// ---------------------------------------------------------------- mma probe
// One CTA, 128 threads. A: 128 x 64 (row = thread = TMEM lane), 16-bit,
// written with tcgen05.st (two elements per 32-bit column, low half first).
// B: NTOK x 64, 16-bit, K-major, 128B-swizzled rows in smem.
// D: 128 x NTOK fp32.
template <int NTOK>
__global__ void mma_ts_kernel(const uint16_t *A, const uint16_t *B, float *D,
                              uint32_t a_fmt, uint32_t b_fmt) {
    __shared__ __align__(1024) uint8_t sB[NTOK * 128];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const unsigned tid = threadIdx.x, warp = tid / 32;

    // B tile -> smem, swizzle 128B: row r (token), 16-byte chunk c -> c ^ (r%8)
    for (unsigned i = tid; i < NTOK * 8; i += blockDim.x) {
        unsigned r = i / 8, c = i % 8;
        uint4 v = reinterpret_cast<const uint4 *>(B)[r * 8 + c];
        *reinterpret_cast<uint4 *>(&sB[r * 128 + ((c ^ (r & 7)) * 16)]) = v;
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async(); // generic-proxy smem writes -> visible to tcgen05
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (warp * 32) << 16;
    // A: columns [32, 64) of the allocation; D: columns [0, NTOK)
    uint32_t v[16];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t lo = A[tid * 64 + half * 32 + 2 * j];
            uint32_t hi = A[tid * 64 + half * 32 + 2 * j + 1];
            v[j] = lo | (hi << 16);
        }
        tmem_st_x16(tmem + lane_base + 32 + half * 16, v);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = make_idesc_f16(a_fmt, b_fmt, 128, NTOK);
            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // A advances 8 columns per K=16; B advances 32 bytes in-atom
                mma_f16_ts(tmem, tmem + 32 + k * 8, bdesc + ((k * 32) >> 4),
                           idesc, k > 0);
            }
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < NTOK; c += 16) {
        tmem_ld_x16(tmem + lane_base + c, v);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j)
            D[tid * NTOK + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static uint16_t f2h(float f) {
    __half h = __float2half(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
static uint16_t f2b(float f) {
    __nv_bfloat16 h = __float2bfloat16(f);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
static float h2f(uint16_t b) {
    __half h;
    memcpy(&h, &b, 2);
    return __half2float(h);
}

template <int NTOK> static void run_mma_case(uint32_t a_fmt, uint32_t b_fmt) {
    std::vector<uint16_t> A(128 * 64), B(NTOK * 64);
    std::vector<float> Af(128 * 64), Bf(NTOK * 64);
    srand(7);
    for (int i = 0; i < 128 * 64; ++i) {
        float v = (float)((rand() % 13) - 6) * 0.25f;
        A[i] = a_fmt == kFmtF16 ? f2h(v) : f2b(v);
        Af[i] = v;
    }
    for (int i = 0; i < NTOK * 64; ++i) {
        // bf16-only magnitudes (1e6 overflows f16) when B is bf16
        float v = (float)((rand() % 17) - 8) * 0.125f;
        if (b_fmt == kFmtBF16 && (i % 5) == 0) v *= 1048576.0f;
        B[i] = b_fmt == kFmtF16 ? f2h(v) : f2b(v);
        Bf[i] = b_fmt == kFmtF16 ? h2f(B[i]) : bf16_bits_to_float(B[i]);
    }
    uint16_t *dA, *dB;
    float *dD;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&dD, 128 * NTOK * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * NTOK * 4));
    mma_ts_kernel<NTOK><<<1, 128>>>(dA, dB, dD, a_fmt, b_fmt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("  mma TS a_fmt=%u b_fmt=%u NTOK=%d: LAUNCH ERROR %s\n", a_fmt,
               b_fmt, NTOK, cudaGetErrorString(e));
        exit(2);
    }
    std::vector<float> D(128 * NTOK);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    int bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int t = 0; t < NTOK; ++t) {
            double ref = 0;
            for (int k = 0; k < 64; ++k)
                ref += (double)Af[r * 64 + k] * (double)Bf[t * 64 + k];
            double err = fabs(ref - D[r * NTOK + t]);
            if (!(err <= 1e-3 * fmax(1.0, fabs(ref)))) ++bad;
            if (err > maxerr) maxerr = err;
            if (fabs(ref) > maxref) maxref = fabs(ref);
        }
    printf("  mma TS A=%s B=%s NTOK=%d: bad=%d maxerr=%g maxref=%g  D[0,0]=%g "
           "D[5,3]=%g\n",
           a_fmt ? "bf16" : "f16", b_fmt ? "bf16" : "f16", NTOK, bad, maxerr,
           maxref, D[0], D[5 * NTOK + 3]);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
}

// ---------------------------------------------------------------- tput probe
// Each thread runs ITER iterations over 8 independent 32-bit words; the body
// is one candidate dequant sequence producing 4 output registers per word.
enum { SEQ_F2FP = 0, SEQ_HMUL_F16, SEQ_HMUL_BF16, SEQ_LOP3, SEQ_F2FP_HMUL,
       SEQ_F2FP_REBIAS_BF16, SEQ_MARLIN_BF16, SEQ_F2FP_FMA32_PACK, SEQ_COUNT };
static const char *kSeqName[] = {"F2FP.e2m1 only", "HMUL2.f16 only",
                                 "HMUL2.bf16 only", "LOP3 only",
                                 "F2FP+HMUL2.f16 (NV->f16)",
                                 "F2FP+shift/add/and+HMUL2.bf16",
                                 "marlin-style bf16 (lop,shf,lop,hmul)",
                                 "F2FP+2xFMA.f32.f16+pack.bf16x2"};

template <int SEQ>
__global__ void __launch_bounds__(256) tput_kernel(uint32_t *out, uint32_t seed,
                                                   int iters) {
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t acc = 0;
    const uint32_t sc16 = 0x3c003c00u;  // 1.0 f16x2
    const uint32_t scb = 0x3f803f80u;   // 1.0 bf16x2
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t w = q[i];
            uint32_t r0, r1, r2, r3;
            if (SEQ == SEQ_F2FP) {
                cvt_e2m1x8_to_f16x2x4(w, r0, r1, r2, r3);
            } else if (SEQ == SEQ_HMUL_F16) {
                asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r0) : "r"(w), "r"(sc16));
                asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r1) : "r"(w ^ 1), "r"(sc16));
                asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r2) : "r"(w ^ 2), "r"(sc16));
                asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(r3) : "r"(w ^ 3), "r"(sc16));
            } else if (SEQ == SEQ_HMUL_BF16) {
                asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r0) : "r"(w), "r"(scb));
                asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r1) : "r"(w ^ 1), "r"(scb));
                asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r2) : "r"(w ^ 2), "r"(scb));
                asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r3) : "r"(w ^ 3), "r"(scb));
            } else if (SEQ == SEQ_LOP3) {
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r0) : "r"(w), "r"(acc), "r"(seed));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r1) : "r"(w), "r"(acc), "r"(seed));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xca;" : "=r"(r2) : "r"(w), "r"(acc), "r"(seed));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x1e;" : "=r"(r3) : "r"(w), "r"(acc), "r"(seed));
            } else if (SEQ == SEQ_F2FP_HMUL) {
                cvt_e2m1x8_to_f16x2x4(w, r0, r1, r2, r3);
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r0) : "r"(sc16));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r1) : "r"(sc16));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r2) : "r"(sc16));
                asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r3) : "r"(sc16));
            } else if (SEQ == SEQ_F2FP_REBIAS_BF16) {
                cvt_e2m1x8_to_f16x2x4(w, r0, r1, r2, r3);
#define REBIAS(r)                                                              \
    r = (((r) >> 3) + 0x70007000u) & 0x8fff8fffu;                              \
    asm volatile("mul.rn.bf16x2 %0, %0, %1;" : "+r"(r) : "r"(scb));
                REBIAS(r0) REBIAS(r1) REBIAS(r2) REBIAS(r3)
            } else if (SEQ == SEQ_MARLIN_BF16) {
#define MARLIN(r, x)                                                           \
    r = ((x) & 0x80008000u) | (((x) & 0x70007000u) >> 6);                      \
    asm volatile("mul.rn.bf16x2 %0, %0, %1;" : "+r"(r) : "r"(scb));
                MARLIN(r0, w) MARLIN(r1, w << 4) MARLIN(r2, w << 8) MARLIN(r3, w << 12)
            } else {
                cvt_e2m1x8_to_f16x2x4(w, r0, r1, r2, r3);
#define FMAPACK(r)                                                             \
    {                                                                          \
        float f0, f1;                                                          \
        asm volatile("{.reg .b16 a,b,s,t; mov.b32 {a,b}, %2; mov.b32 {s,t}, %3;\n" \
                     "fma.rn.f32.f16 %0, a, s, 0f00000000;\n"                    \
                     "fma.rn.f32.f16 %1, b, s, 0f00000000;}"                     \
                     : "=f"(f0), "=f"(f1) : "r"(r), "r"(sc16));                 \
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f1), "f"(f0)); \
    }
                FMAPACK(r0) FMAPACK(r1) FMAPACK(r2) FMAPACK(r3)
            }
            acc ^= r0 ^ r1 ^ r2 ^ r3;
            q[i] = w + acc;
        }
    }
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

template <int SEQ> static void run_tput_one(int warps_per_sm) {
    uint32_t *d;
    CK(cudaMalloc(&d, 4096));
    const int iters = 2000;
    int threads = 256;
    int blocks_per_sm = warps_per_sm * 32 / threads;
    if (blocks_per_sm == 0) { blocks_per_sm = 1; threads = warps_per_sm * 32; }
    int grid = 148 * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    tput_kernel<SEQ><<<grid, threads>>>(d, 12345, 10);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    tput_kernel<SEQ><<<grid, threads>>>(d, 12345, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // weights per thread = iters * 8 words * 8 nibbles
    double weights = (double)grid * threads * iters * 64.0;
    double wps = weights / (ms * 1e-3);
    printf("  %-42s warps/SM=%2d  %.3f ms  %.1f Gweights/s/SM = %.1f "
           "weights/clk/SM @1.9GHz (roofline needs ~46)\n",
           kSeqName[SEQ], warps_per_sm, ms, wps / 148 / 1e9,
           wps / 148 / 1.9e9);
    cudaFree(d);
}

template <int SEQ> static void run_tput_seq() {
    run_tput_one<SEQ>(4);
    run_tput_one<SEQ>(8);
    run_tput_one<SEQ>(16);
}

// ---------------------------------------------------------------- bw probe
// grid CTAs, each streams a contiguous slice with cp.async.bulk into a ring of
// `stages` buffers of `stage_bytes`; a consumer warp touches 16 B per stage.
__global__ void __launch_bounds__(64) bw_kernel(const uint8_t *src,
                                                size_t bytes_per_cta,
                                                int stage_bytes, int stages,
                                                uint32_t *sink, int hint) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + 32;
    uint8_t *buf = smem + 1024;
    const unsigned warp = threadIdx.x / 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const uint8_t *p = src + (size_t)blockIdx.x * bytes_per_cta;
    const int n = (int)(bytes_per_cta / stage_bytes);
    if (warp == 0) {
        if (elect_one()) {
            uint64_t pol = policy_evict_first();
            for (int i = 0; i < n; ++i) {
                int s = i % stages;
                uint32_t ph = (i / stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                if (hint)
                    bulk_g2s_hint(buf + (size_t)s * stage_bytes,
                                  p + (size_t)i * stage_bytes, stage_bytes,
                                  &full[s], pol);
                else
                    bulk_g2s(buf + (size_t)s * stage_bytes,
                             p + (size_t)i * stage_bytes, stage_bytes,
                             &full[s]);
            }
        }
    } else {
        uint32_t acc = 0;
        for (int i = 0; i < n; ++i) {
            int s = i % stages;
            uint32_t ph = (i / stages) & 1;
            mbar_wait(&full[s], ph);
            acc += *reinterpret_cast<volatile uint32_t *>(
                buf + (size_t)s * stage_bytes + (threadIdx.x % 32) * 4);
            __syncwarp();
            if (threadIdx.x % 32 == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x1234567) sink[0] = acc;
    }
}

static void run_bw() {
    const size_t total = (size_t)2 << 30; // 2 GiB > L2
    uint8_t *src;
    uint32_t *sink;
    CK(cudaMalloc(&src, total));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(src, 1, total));
    CK(cudaFuncSetAttribute(bw_kernel,
                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                            220 * 1024));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    struct Cfg { int stage_kb, stages, ctas_per_sm, hint; };
    const Cfg cfgs[] = {{16, 4, 1, 0},  {16, 8, 1, 0},  {16, 12, 1, 0},
                        {18, 8, 1, 0},  {32, 4, 1, 0},  {32, 6, 1, 0},
                        {8, 16, 1, 0},  {8, 24, 1, 0},  {16, 6, 2, 0},
                        {16, 12, 1, 1}, {32, 6, 1, 1},  {64, 3, 1, 0},
                        {4, 32, 1, 0},  {16, 3, 4, 0}};
    for (const Cfg &c : cfgs) {
        int grid = 148 * c.ctas_per_sm;
        int stage_bytes = c.stage_kb * 1024;
        size_t per_cta = (total / grid / stage_bytes) * stage_bytes;
        size_t smem = 1024 + (size_t)stage_bytes * c.stages;
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            bw_kernel<<<grid, 64, smem>>>(src, per_cta, stage_bytes, c.stages,
                                          sink, c.hint);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double gb = (double)per_cta * grid / 1e9;
        printf("  bulk stream stage=%2dKB stages=%2d ctas/SM=%d hint=%d: %.3f "
               "ms  %.0f GB/s\n",
               c.stage_kb, c.stages, c.ctas_per_sm, c.hint, best,
               gb / (best * 1e-3));
    }
    // short transfers (decode-sized): 40 MB total per launch, distinct regions
    for (int mb : {40, 130, 260}) {
        int stage_bytes = 16 * 1024, stages = 12;
        size_t per_cta =
            ((size_t)mb * 1000000 / 148 / stage_bytes) * stage_bytes;
        size_t smem = 1024 + (size_t)stage_bytes * stages;
        float sum = 0;
        int reps = 8;
        for (int rep = 0; rep < reps; ++rep) {
            cudaEventRecord(e0);
            bw_kernel<<<148, 64, smem>>>(src + (size_t)rep * 268435456 % (total - 300000000),
                                         per_cta, stage_bytes, stages, sink, 0);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0) sum += ms;
        }
        float ms = sum / (reps - 1);
        printf("  short stream %d MB: %.2f us  %.0f GB/s (incl. launch "
               "ramp)\n",
               mb, ms * 1e3, (double)per_cta * 148 / 1e9 / (ms * 1e-3));
    }
    cudaFree(src);
    cudaFree(sink);
}

int main(int argc, char **argv) {
    const char *what = argc > 1 ? argv[1] : "all";
    bool all = !strcmp(what, "all");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, cc %d.%d, smem/block optin %zu\n", prop.name,
           prop.multiProcessorCount, prop.major, prop.minor,
           prop.sharedMemPerBlockOptin);
    if (all || !strcmp(what, "cvt")) run_cvt();
    if (all || !strcmp(what, "tput")) {
        printf("== instruction-sequence throughput\n");
        run_tput_seq<SEQ_F2FP>();
        run_tput_seq<SEQ_HMUL_F16>();
        run_tput_seq<SEQ_HMUL_BF16>();
        run_tput_seq<SEQ_LOP3>();
        run_tput_seq<SEQ_F2FP_HMUL>();
        run_tput_seq<SEQ_F2FP_REBIAS_BF16>();
        run_tput_seq<SEQ_MARLIN_BF16>();
        run_tput_seq<SEQ_F2FP_FMA32_PACK>();
    }
    if (all || !strcmp(what, "bw")) {
        printf("== cp.async.bulk streaming bandwidth\n");
        run_bw();
    }
    if (all || !strcmp(what, "mma")) {
        printf("== TS-form tcgen05.mma (A in TMEM)\n");
        run_mma_case<16>(kFmtF16, kFmtF16);
        run_mma_case<16>(kFmtBF16, kFmtBF16);
        run_mma_case<32>(kFmtBF16, kFmtBF16);
        run_mma_case<64>(kFmtF16, kFmtF16);
        printf("-- mixed formats (A=f16 weights, B=bf16 activations)\n");
        run_mma_case<16>(kFmtF16, kFmtBF16);
        run_mma_case<64>(kFmtF16, kFmtBF16);
        run_mma_case<16>(kFmtBF16, kFmtF16);
    }
    return 0;
}
