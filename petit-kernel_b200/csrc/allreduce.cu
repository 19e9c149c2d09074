// One-shot all-reduce of a small row-parallel GEMM output over NVLink peer memory
// (SURVEY.md section 8 row e / f2; nothing like it exists in the reference, whose benchmark
// list only carries the TP shard shapes, tools/benchmarks/matmul.py:13-25).
//
// Every rank's GEMM writes its partial [M, N] into a buffer that all peers have mapped
// (symmetric memory).  One kernel per rank then
//   1. tells every peer "my partial is complete" and waits for the same from all of them
//      (per-CTA flags in the peers' signal pads),
//   2. reads the matching 16-byte vectors of ALL ranks' buffers through NVLink and sums
//      them in fp32 in rank order -- the same order on every rank, so all ranks end up
//      with bit-identical results,
//   3. (end_barrier) tells every peer "I am done reading your buffer" and waits for the
//      same, so the next GEMM may overwrite the buffer as soon as this kernel has finished.
//      Callers that alternate between two buffers skip it: passing the start barrier of
//      call i+1 proves every peer finished call i, so buffer (i mod 2) is free for call i+2.
// The flags are monotonically increasing counters and the expected value lives in device
// memory (`epoch`), so the launch is CUDA-graph replayable and needs no host bookkeeping.
// It is launched with programmatic stream serialisation: the next GEMM's weight prefetch
// overlaps it, and it touches the partial only after griddepcontrol.wait.
#include "causalflow/petit/petit.h"

#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace petit::allreduce {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 64;
constexpr int kThreads = 256;
// signal pad (uint32 words) per rank: [phase 2][cta kMaxCtas][source rank kMaxWorld]
constexpr size_t kPadWords = 2 * kMaxCtas * kMaxWorld;

struct Args {
    const void *bufs[kMaxWorld]; // every rank's partial (bufs[rank] is the local one)
    uint32_t *pads[kMaxWorld];   // every rank's signal pad
    void *out;                   // local result
    uint32_t *epoch;             // local, [kMaxCtas] completed calls per CTA + [kStatusWord] status
    uint64_t vecs;               // 16-byte vectors to reduce
    unsigned long long timeout_ns;
    int rank, world;
    int end_barrier;             // 0: the caller double-buffers, see petit.h
    int fenced;                  // 1: release/acquire flags at system scope (PTX-model clean)
};

// Two flag protocols (PETIT_ALLREDUCE_FENCED / PETIT_AR_FENCED=1 selects the second):
//
//  relaxed (default): flags are RELAXED system-scope accesses, no fence on either side.
//    What makes this work is not the PTX memory model (which only promises visibility
//    through a release/acquire pattern) but three properties of the platform, each checked
//    by tests/tp_peer_allreduce_check.py on 2/4/8 real GPUs with data that changes every
//    call: (1) the partial was written by the PREVIOUS kernel on this stream;
//    griddepcontrol.wait / stream order returns only after that grid's writes are
//    performed at the GPU's L2, which is the point of coherence NVLink peers read from;
//    (2) the peer-data loads below are ld.relaxed.sys -- strong loads that are served by
//    the home L2 through NVLink and never by a (possibly stale) L1 line of the reader;
//    (3) a reader issues them only after its own flag load returned the expected value
//    (control dependency + bar.sync), so they reach the home L2 after the signal left it.
//  fenced: the signalling thread executes fence.acq_rel.sys before its red.relaxed.sys
//    (release pattern) and the polling thread fence.acq_rel.sys after the poll (acquire
//    pattern): formally synchronises-with; measured +4 us per barrier on B200/NVLink.
//  The end barrier only orders "my loads have returned" (their values were consumed by the
//  adds) before "the owner may overwrite"; nothing has to be published.
__device__ __forceinline__ void signal_add(uint32_t *p) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ uint32_t load_relaxed_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 load_relaxed_sys_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

constexpr int kStatusWord = kMaxCtas; // epoch[kStatusWord] != 0: a peer barrier timed out

// CTA-wide barrier with the same CTA of every other rank.  Thread t < world signals rank
// t and waits for rank t's signal.  A peer that does not show up within timeout_ns makes
// the CTA give up: it records the failure in the status word (petit_allreduce_status) and
// the kernel exits without reducing -- no trap, the context survives.  Returns false then.
__device__ __forceinline__ bool peer_barrier(const Args &a, int phase, uint32_t target) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads(); // every load / store of this CTA before the barrier has been issued
    if ((int)threadIdx.x < a.world) {
        const int t = threadIdx.x;
        const size_t slot = ((size_t)phase * kMaxCtas + blockIdx.x) * kMaxWorld;
        if (a.fenced) fence_sys();
        signal_add(a.pads[t] + slot + a.rank);
        const uint32_t *mine = a.pads[a.rank] + slot + t;
        unsigned long long t0 = 0;
        uint32_t spins = 0;
        while ((int32_t)(load_relaxed_sys(mine) - target) < 0) {
            if ((++spins & 0x3ff) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                if (now - t0 > a.timeout_ns) {
                    atomicExch(&timed_out, 1);
                    atomicExch(a.epoch + kStatusWord, 1u + (uint32_t)t);
                    break;
                }
            }
        }
        if (a.fenced) fence_sys();
    }
    __syncthreads();
    return timed_out == 0;
}

template <typename T2> __device__ __forceinline__ float2 to_f2(uint32_t v);
template <> __device__ __forceinline__ float2 to_f2<__nv_bfloat162>(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&v));
}
template <> __device__ __forceinline__ float2 to_f2<__half2>(uint32_t v) {
    return __half22float2(*reinterpret_cast<__half2 *>(&v));
}
template <typename T2> __device__ __forceinline__ uint32_t from_f2(float2 f);
template <> __device__ __forceinline__ uint32_t from_f2<__nv_bfloat162>(float2 f) {
    __nv_bfloat162 r = __float22bfloat162_rn(f);
    return *reinterpret_cast<uint32_t *>(&r);
}
template <> __device__ __forceinline__ uint32_t from_f2<__half2>(float2 f) {
    __half2 r = __float22half2_rn(f);
    return *reinterpret_cast<uint32_t *>(&r);
}

template <typename T2>
__global__ void __launch_bounds__(kThreads) oneshot_allreduce_kernel(Args a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory"); // the local partial is complete
    const uint32_t target = a.epoch[blockIdx.x] + 1;
    // a previous call already lost a peer: do not wait for the timeout again
    if (load_relaxed_sys(a.epoch + kStatusWord) != 0) return;
    if (!peer_barrier(a, 0, target)) return;

    for (uint64_t v = (uint64_t)blockIdx.x * kThreads + threadIdx.x; v < a.vecs;
         v += (uint64_t)gridDim.x * kThreads) {
        uint4 x[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < a.world) x[r] = load_relaxed_sys_v4(reinterpret_cast<const uint4 *>(a.bufs[r]) + v);
        float2 s[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r) {
            if (r < a.world) {
                const uint32_t w[4] = {x[r].x, x[r].y, x[r].z, x[r].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = to_f2<T2>(w[i]);
                    s[i].x += f.x;
                    s[i].y += f.y;
                }
            }
        }
        uint4 o;
        o.x = from_f2<T2>(s[0]);
        o.y = from_f2<T2>(s[1]);
        o.z = from_f2<T2>(s[2]);
        o.w = from_f2<T2>(s[3]);
        reinterpret_cast<uint4 *>(a.out)[v] = o;
    }

    if (a.end_barrier && !peer_barrier(a, 1, target)) return;
    if (threadIdx.x == 0) a.epoch[blockIdx.x] = target;
}

} // namespace petit::allreduce

extern "C" {

size_t petit_allreduce_pad_bytes(void) { return petit::allreduce::kPadWords * sizeof(uint32_t); }
size_t petit_allreduce_epoch_bytes(void) {
    return (petit::allreduce::kMaxCtas + 4) * sizeof(uint32_t);
}

int petit_allreduce_status(const void *epoch, petit_stream_t stream) {
    uint32_t st = 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(&st, static_cast<const uint32_t *>(epoch) + petit::allreduce::kStatusWord,
                        sizeof st, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return -1;
    return (int)st;
}

int petit_allreduce_oneshot(void *out, const void *const *peer_bufs, void *const *peer_pads,
                            void *epoch, int rank, int world, size_t numel, int dtype,
                            int end_barrier, petit_stream_t stream) {
    using namespace petit::allreduce;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return PETIT_ERROR_PROBLEM_SHAPE;
    if (dtype != PETIT_DTYPE_BF16 && dtype != PETIT_DTYPE_FP16) return PETIT_ERROR_PROBLEM_SHAPE;
    if (numel == 0) return PETIT_OK;
    if (numel % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return PETIT_ERROR_PROBLEM_SHAPE;
    Args a{};
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r] || !peer_pads[r] || (reinterpret_cast<uintptr_t>(peer_bufs[r]) & 15))
            return PETIT_ERROR_PROBLEM_SHAPE;
        a.bufs[r] = peer_bufs[r];
        a.pads[r] = static_cast<uint32_t *>(peer_pads[r]);
    }
    a.out = out;
    a.epoch = static_cast<uint32_t *>(epoch);
    a.vecs = numel / 8;
    a.rank = rank;
    a.world = world;
    a.end_barrier = end_barrier & PETIT_ALLREDUCE_END_BARRIER;
    static const int env_fenced = [] {
        const char *e = std::getenv("PETIT_AR_FENCED");
        return e ? std::atoi(e) : 0;
    }();
    a.fenced = ((end_barrier & PETIT_ALLREDUCE_FENCED) || env_fenced) ? 1 : 0;
    static const unsigned long long timeout_ns = [] {
        const char *e = std::getenv("PETIT_AR_TIMEOUT_MS");
        const long long ms = e ? std::atoll(e) : 4000;
        return (unsigned long long)(ms > 0 ? ms : 4000) * 1000000ull;
    }();
    a.timeout_ns = timeout_ns;
    // The grid must be the same on every rank and for every call on a pad (the flags are
    // per CTA): it depends on the element count only.
    const uint64_t want = (a.vecs + kThreads - 1) / kThreads;
    const unsigned grid = (unsigned)(want < (uint64_t)kMaxCtas ? want : (uint64_t)kMaxCtas);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = dtype == PETIT_DTYPE_BF16
                        ? cudaLaunchKernelEx(&cfg, oneshot_allreduce_kernel<__nv_bfloat162>, a)
                        : cudaLaunchKernelEx(&cfg, oneshot_allreduce_kernel<__half2>, a);
    return e == cudaSuccess ? PETIT_OK : PETIT_ERROR_CUDA;
}

} // extern "C"
