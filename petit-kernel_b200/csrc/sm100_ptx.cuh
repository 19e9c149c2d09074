// Thin inline-PTX wrappers for the sm_100a features the FP4 GEMM uses:
// mbarrier, 1-D bulk TMA (cp.async.bulk), tensor-map TMA, tcgen05 (TMEM
// alloc / st / ld / mma / commit) and the narrow-float converts.
//
// Role in the design: this replaces the reference's ISA wrapper layer
// (lib/gemm/rocm/amd_intrinsics.cuh:52-130 -- MFMA, buffer loads, v_perm) with
// its Blackwell counterpart. Nothing here is a translation of that file.
#pragma once

#include <cstdint>
#include <cuda.h>

namespace petit::ptx {

// ----------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- smem -> global bulk tensor store (epilogue) ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *smem_src, int c0,
                                             int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     map),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until at most N committed bulk groups still READ their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar,
                                                      uint32_t bytes) {
    asm volatile(
        "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
            smem_u32(bar)),
        "r"(bytes)
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ----------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------
// 1-D bulk copy global -> shared, completion signalled on an mbarrier.
// bytes must be a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src,
                                         uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::"
                 "bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Same with an L2 cache-policy hint (createpolicy result).
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst,
                                              const void *gmem_src,
                                              uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
                 : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;"
                 : "=l"(p));
    return p;
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

// 3-D tiled tensor-map load (c0 innermost).
__device__ __forceinline__ void tma_load_3d(void *smem_dst,
                                            const CUtensorMap *map,
                                            uint64_t *bar, int c0, int c1,
                                            int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile."
                 "mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst,
                                            const CUtensorMap *map,
                                            uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile."
                 "mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: tensor memory management
// ----------------------------------------------------------------------------
// Must be executed by one full warp. ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result,
                                           uint32_t ncols) {
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::
            "r"(smem_u32(smem_result)),
        "r"(ncols)
        : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::
                     : "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(
                     taddr),
                 "r"(ncols)
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// tcgen05.commit: the mbarrier receives one arrival once all tcgen05.mma
// issued so far by this thread have completed.
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::"
                 "cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 bit, N consecutive columns per thread. Thread i of the warp
// owns TMEM lane (lane_base + i); lane_base = 32 * (warp_id % 4) must be in
// bits [31:16] of taddr.
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr,
                                            const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, "
        "%7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
        "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr,
                                            uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, "
        "%7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]),
          "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, "
                 "%5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]),
                   "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05.mma, kind::f16
// ----------------------------------------------------------------------------
enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1 };

// Instruction descriptor (upper 32 bits of CuTe's 64-bit idescE): fp32
// accumulate, K-major A and B, no negate, dense.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t a_fmt,
                                                      uint32_t b_fmt,
                                                      uint32_t m, uint32_t n) {
    return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((n >> 3) << 17) |
           ((m >> 4) << 24);
}

// Shared-memory matrix descriptor for a K-major operand stored as 128-byte
// rows with the 128B swizzle (what TMA SWIZZLE_128B writes): 8-row groups are
// 1024 bytes apart (stride byte offset), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);  // start address
    d |= static_cast<uint64_t>(1) << 16;                 // LBO (unused)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;         // SBO
    d |= static_cast<uint64_t>(1) << 46;                 // version
    d |= static_cast<uint64_t>(2) << 61;                 // SWIZZLE_128B
    return d;
}

// D[tmem] (+)= A[tmem] * B[smem]   (TS form: A operand lives in TMEM)
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem,
                                           uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem),
                 "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]   (SS form)
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc,
                                           uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

// ----------------------------------------------------------------------------
// narrow-float converts (sm_100a)
// ----------------------------------------------------------------------------
// Four bytes of packed e2m1 pairs -> four f16x2 registers. Byte b of `q`
// holds elements (2b, 2b+1): low nibble -> low half.
__device__ __forceinline__ void cvt_e2m1x8_to_f16x2x4(uint32_t q, uint32_t &r0,
                                                      uint32_t &r1,
                                                      uint32_t &r2,
                                                      uint32_t &r3) {
    asm("{\n\t.reg .b8 b0, b1, b2, b3;\n\tmov.b32 {b0, b1, b2, b3}, %4;\n\t"
        "cvt.rn.f16x2.e2m1x2 %0, b0;\n\tcvt.rn.f16x2.e2m1x2 %1, b1;\n\t"
        "cvt.rn.f16x2.e2m1x2 %2, b2;\n\tcvt.rn.f16x2.e2m1x2 %3, b3;\n\t}"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "r"(q));
}

// Two e4m3 bytes (low 16 bits of each half of `s`) -> f16x2.
__device__ __forceinline__ void cvt_e4m3x4_to_f16x2x2(uint32_t s, uint32_t &lo,
                                                      uint32_t &hi) {
    asm("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %2;\n\t"
        "cvt.rn.f16x2.e4m3x2 %0, h0;\n\tcvt.rn.f16x2.e4m3x2 %1, h1;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"(s));
}

// Two ue8m0 bytes -> bf16x2 (2^(e-127)).
__device__ __forceinline__ void cvt_ue8m0x4_to_bf16x2x2(uint32_t s,
                                                        uint32_t &lo,
                                                        uint32_t &hi) {
    asm("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %2;\n\t"
        "cvt.rn.bf16x2.ue8m0x2 %0, h0;\n\tcvt.rn.bf16x2.ue8m0x2 %1, h1;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"(s));
}

} // namespace petit::ptx
