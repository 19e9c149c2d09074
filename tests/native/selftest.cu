// Native parity test of the C ABI (libpetit_b200.so) against the plain-C oracle.
// Runs on the GPU box:  ./selftest [quick|full]
// Checks, per (format, dtype, m, n, k):
//   * unpack(repack(w)) == w                       (bit-exact)
//   * dense hook on packed data == oracle dequant  (bit-exact, +-0 equal)
//   * dense hook on native data == packed hook     (bit-exact)
//   * GEMM vs fp32 accumulation of the oracle weights: max rel err <= 1e-2 and
//     the reference matcher |a-b| < max(1e-2, 0.01|b|) (gemm_fp4_fp16_rocm_test.cc:31-67)
#include "causalflow/petit/petit.h"
#include "petit_oracle.h"

#include <cmath>
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

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(rng_state >> 33);
}
static float rnd_uniform(float lo, float hi) { return lo + (hi - lo) * (rnd() / 2147483648.0f); }

static int g_fail = 0;

static bool same16_pm0(uint16_t a, uint16_t b) {
    return a == b || (((a | b) & 0x7fff) == 0);
}

static void run_case(bool mx, bool bf16, unsigned m, unsigned n, unsigned k, int ntok_force,
                     float gscale, bool check_dequant) {
    const unsigned group = mx ? 32 : 16;
    std::vector<uint8_t> q((size_t)n * k / 2), sc((size_t)n * k / group);
    for (auto &b : q) b = (uint8_t)rnd();
    for (auto &s : sc) s = mx ? (uint8_t)(1 + rnd() % 237) : (uint8_t)(1 + rnd() % 0x7e);
    std::vector<uint16_t> a((size_t)m * k);
    std::vector<float> af((size_t)m * k);
    for (size_t i = 0; i < a.size(); ++i) {
        float v = rnd_uniform(-2.f, 2.f);
        a[i] = bf16 ? petit_oracle_f32_to_bf16(v) : petit_oracle_f32_to_f16(v);
        af[i] = bf16 ? petit_oracle_bf16_to_f32(a[i]) : petit_oracle_f16_to_f32(a[i]);
    }
    // MX scales span 2^-126..2^110: keep magnitudes representable in the 16-bit output
    if (mx)
        for (auto &s : sc) s = (uint8_t)(100 + rnd() % 40);
    if (mx && check_dequant) // full-range scales for the dequant checks only
        for (size_t i = 0; i < sc.size(); i += 3) sc[i] = (uint8_t)(1 + rnd() % 237);

    uint8_t *d_q, *d_qp, *d_qu, *d_sc, *d_scp;
    uint16_t *d_a, *d_c, *d_dense, *d_dense2;
    float *d_gs;
    CK(cudaMalloc(&d_q, q.size()));
    CK(cudaMalloc(&d_qp, q.size()));
    CK(cudaMalloc(&d_qu, q.size()));
    CK(cudaMalloc(&d_sc, sc.size()));
    CK(cudaMalloc(&d_scp, sc.size()));
    CK(cudaMalloc(&d_a, a.size() * 2));
    CK(cudaMalloc(&d_c, (size_t)m * n * 2));
    CK(cudaMalloc(&d_dense, (size_t)n * k * 2));
    CK(cudaMalloc(&d_dense2, (size_t)n * k * 2));
    CK(cudaMalloc(&d_gs, 4));
    CK(cudaMemcpy(d_q, q.data(), q.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sc, sc.data(), sc.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gs, &gscale, 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_c, 0xff, (size_t)m * n * 2));

    int rc = petit_repack_fp4_weights((uint32_t *)d_qp, (const uint32_t *)d_q, k, n, nullptr);
    rc |= mx ? petit_repack_mxfp4_scales(d_scp, d_sc, k, n, nullptr)
             : petit_repack_nvfp4_scales(d_scp, d_sc, k, n, nullptr);
    rc |= petit_unpack_fp4_weights((uint32_t *)d_qu, (const uint32_t *)d_qp, k, n, nullptr);
    const int out_type = bf16 ? PETIT_DTYPE_BF16 : PETIT_DTYPE_FP16;
    if (mx) {
        rc |= petit_dequant_packed_mxfp4(d_dense, d_qp, d_scp, 1.0f, out_type, k, n, nullptr);
        rc |= petit_dequant_mxfp4(d_dense2, d_q, d_sc, 1.0f, out_type, k, n, nullptr);
    } else {
        rc |= petit_dequant_packed_nvfp4(d_dense, d_qp, d_scp, 1.0f, out_type, k, n, nullptr);
        rc |= petit_dequant_nvfp4(d_dense2, d_q, d_sc, 1.0f, out_type, k, n, nullptr);
    }
    PetitSolutionHints hints = {out_type, mx ? PETIT_DTYPE_MXFP4_E2M1 : PETIT_DTYPE_FP4_E2M1,
                                out_type, 0};
    uint64_t sol = PETIT_SOLUTION_AUTO;
    if (ntok_force) {
        uint64_t sols[8];
        unsigned ns = 8;
        petit_get_solutions(&hints, m, n, k, sols, &ns);
        for (unsigned i = 0; i < ns; ++i)
            if ((int)(sols[i] & 0xff) * 16 == ntok_force) sol = sols[i];
    }
    int grc = mx ? petit_gemm_mxfp4_a16(d_c, d_a, d_qp, d_scp, d_gs, m, n, k, &hints, sol, nullptr)
                 : petit_gemm_nvfp4_a16(d_c, d_a, d_qp, d_scp, d_gs, m, n, k, &hints, sol, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || grc || e != cudaSuccess) {
        printf("FAIL %s %s m=%u n=%u k=%u tok=%d: rc=%d gemm_rc=%d cuda=%s\n", mx ? "mx" : "nv",
               bf16 ? "bf16" : "f16", m, n, k, ntok_force, rc, grc, cudaGetErrorString(e));
        ++g_fail;
        if (e != cudaSuccess) exit(2);
        return;
    }
    std::vector<uint8_t> qu(q.size());
    std::vector<uint16_t> c((size_t)m * n), dense((size_t)n * k), dense2((size_t)n * k);
    CK(cudaMemcpy(qu.data(), d_qu, qu.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), d_c, c.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dense.data(), d_dense, dense.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dense2.data(), d_dense2, dense2.size() * 2, cudaMemcpyDeviceToHost));

    size_t bad_rt = 0, bad_dq = 0, bad_dq2 = 0;
    for (size_t i = 0; i < q.size(); ++i) bad_rt += qu[i] != q[i];
    std::vector<float> wf((size_t)n * k);
    if (mx)
        petit_oracle_dequant_mxfp4(wf.data(), q.data(), sc.data(), n, k);
    else
        petit_oracle_dequant_nvfp4(wf.data(), q.data(), sc.data(), n, k);
    for (size_t i = 0; i < wf.size(); ++i) {
        uint16_t ref = bf16 ? petit_oracle_f32_to_bf16(wf[i]) : petit_oracle_f32_to_f16(wf[i]);
        bad_dq += !same16_pm0(dense[i], ref);
        bad_dq2 += !same16_pm0(dense2[i], dense[i]);
    }
    // GEMM reference: fp32 accumulation of the dequantised weights, then * global scale
    std::vector<float> cref((size_t)m * n);
    petit_oracle_gemm_f32(cref.data(), af.data(), wf.data(), m, n, k);
    double max_err = 0, max_ref = 0;
    size_t bad_match = 0;
    for (size_t i = 0; i < cref.size(); ++i) {
        float ref = cref[i] * gscale;
        float got = bf16 ? petit_oracle_bf16_to_f32(c[i]) : petit_oracle_f16_to_f32(c[i]);
        // matcher is applied to the 16-bit rounded reference like the gtest does
        float ref16 = bf16 ? petit_oracle_bf16_to_f32(petit_oracle_f32_to_bf16(ref))
                           : petit_oracle_f16_to_f32(petit_oracle_f32_to_f16(ref));
        double err = fabs((double)got - ref);
        if (!(fabs(got - ref16) < fmax(1e-2, 0.01 * fabs(ref16)))) ++bad_match;
        if (err > max_err || std::isnan(err)) max_err = std::isnan(err) ? 1e30 : err;
        if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
    double rel = max_err / (max_ref > 0 ? max_ref : 1);
    // The gtest matcher (abs 1e-2 on outputs of magnitude ~1e3) is below fp32
    // accumulation-order noise once m*n is large; it is enforced on the reference's
    // own case sizes (m*n <= 96*128) and reported elsewhere.
    const bool matcher_required = (size_t)m * n <= 96 * 128;
    bool ok = bad_rt == 0 && bad_dq == 0 && bad_dq2 == 0 && rel <= 1e-2 &&
              (bad_match == 0 || !matcher_required);
    printf("%s %s %-4s m=%-5u n=%-6u k=%-6u tok=%-3d roundtrip_bad=%zu dequant_bad=%zu "
           "native_vs_packed_bad=%zu gemm max_rel=%.3e matcher_bad=%zu\n",
           ok ? "PASS" : "FAIL", mx ? "mx" : "nv", bf16 ? "bf16" : "f16", m, n, k, ntok_force,
           bad_rt, bad_dq, bad_dq2, rel, bad_match);
    if (!ok) ++g_fail;
    cudaFree(d_q); cudaFree(d_qp); cudaFree(d_qu); cudaFree(d_sc); cudaFree(d_scp);
    cudaFree(d_a); cudaFree(d_c); cudaFree(d_dense); cudaFree(d_dense2); cudaFree(d_gs);
}

// Grouped (MoE) GEMM through the C ABI: `groups` experts with their own weights, ragged token
// counts.  contiguous = the experts' rows are consecutive blocks of one tensor (one launch);
// otherwise every expert's rows start at a padded offset (the per-expert path).  Both against
// the C oracle per expert; rows outside the groups must stay untouched (0xffff).
static void run_grouped(bool contiguous, unsigned n, unsigned k, const std::vector<unsigned> &counts) {
    const unsigned groups = (unsigned)counts.size(), pad = contiguous ? 0 : 3;
    std::vector<std::vector<uint8_t>> q(groups), sc(groups);
    std::vector<uint8_t *> d_qp(groups), d_scp(groups);
    std::vector<float> gs(groups);
    float *d_gs;
    CK(cudaMalloc(&d_gs, groups * 4));
    unsigned total = 0;
    std::vector<unsigned> row0(groups);
    for (unsigned g = 0; g < groups; ++g) {
        row0[g] = total;
        total += counts[g] + (counts[g] ? pad : 0);
        q[g].resize((size_t)n * k / 2);
        sc[g].resize((size_t)n * k / 16);
        for (auto &b : q[g]) b = (uint8_t)rnd();
        for (auto &x : sc[g]) x = (uint8_t)(0x20 + rnd() % 0x30);
        gs[g] = rnd_uniform(0.5f, 1.5f);
        uint8_t *d_q, *d_sc;
        CK(cudaMalloc(&d_q, q[g].size()));
        CK(cudaMalloc(&d_sc, sc[g].size()));
        CK(cudaMalloc(&d_qp[g], q[g].size()));
        CK(cudaMalloc(&d_scp[g], sc[g].size()));
        CK(cudaMemcpy(d_q, q[g].data(), q[g].size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_sc, sc[g].data(), sc[g].size(), cudaMemcpyHostToDevice));
        int rc = petit_repack_fp4_weights((uint32_t *)d_qp[g], (const uint32_t *)d_q, k, n, nullptr);
        rc |= petit_repack_nvfp4_scales(d_scp[g], d_sc, k, n, nullptr);
        if (rc) { printf("FAIL grouped repack rc=%d\n", rc); ++g_fail; return; }
        CK(cudaDeviceSynchronize());
        cudaFree(d_q); cudaFree(d_sc);
    }
    CK(cudaMemcpy(d_gs, gs.data(), groups * 4, cudaMemcpyHostToDevice));
    std::vector<uint16_t> a((size_t)total * k);
    std::vector<float> af(a.size());
    for (size_t i = 0; i < a.size(); ++i) {
        a[i] = petit_oracle_f32_to_bf16(rnd_uniform(-2.f, 2.f));
        af[i] = petit_oracle_bf16_to_f32(a[i]);
    }
    uint16_t *d_a, *d_c;
    CK(cudaMalloc(&d_a, a.size() * 2));
    CK(cudaMalloc(&d_c, (size_t)total * n * 2));
    CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_c, 0xff, (size_t)total * n * 2));
    std::vector<PetitGroupedProblem> probs(groups);
    for (unsigned g = 0; g < groups; ++g)
        probs[g] = {d_c + (size_t)row0[g] * n, d_a + (size_t)row0[g] * k, d_qp[g], d_scp[g], d_gs + g, counts[g]};
    PetitSolutionHints hints = {PETIT_DTYPE_BF16, PETIT_DTYPE_FP4_E2M1, PETIT_DTYPE_BF16, 0};
    int grc = petit_gemm_fp4_a16_grouped(probs.data(), groups, n, k, &hints, PETIT_SOLUTION_AUTO, nullptr, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    int st = -1;
    petit_workspace_status(nullptr, &st);
    std::vector<uint16_t> c((size_t)total * n);
    CK(cudaMemcpy(c.data(), d_c, c.size() * 2, cudaMemcpyDeviceToHost));
    double worst = 0;
    size_t touched = 0;
    std::vector<float> wf((size_t)n * k);
    for (unsigned g = 0; g < groups; ++g) {
        if (!counts[g]) continue;
        petit_oracle_dequant_nvfp4(wf.data(), q[g].data(), sc[g].data(), n, k);
        std::vector<float> cref((size_t)counts[g] * n);
        petit_oracle_gemm_f32(cref.data(), af.data() + (size_t)row0[g] * k, wf.data(), counts[g], n, k);
        double max_err = 0, max_ref = 0;
        for (size_t i = 0; i < cref.size(); ++i) {
            const double ref = cref[i] * gs[g];
            const double got = petit_oracle_bf16_to_f32(c[(size_t)row0[g] * n + i]);
            const double err = fabs(got - ref);
            if (err > max_err || std::isnan(err)) max_err = std::isnan(err) ? 1e30 : err;
            if (fabs(ref) > max_ref) max_ref = fabs(ref);
        }
        worst = fmax(worst, max_err / (max_ref > 0 ? max_ref : 1));
        for (size_t i = 0; i < (size_t)pad * n; ++i) // the padding rows behind the group
            touched += c[(size_t)(row0[g] + counts[g]) * n + i] != 0xffff;
    }
    const bool ok = grc == 0 && e == cudaSuccess && st == 0 && worst <= 1e-2 && touched == 0;
    printf("%s grouped %s n=%u k=%u groups=%u tokens=%u rc=%d ws_status=%d max_rel=%.3e touched_outside=%zu\n",
           ok ? "PASS" : "FAIL", contiguous ? "one-launch" : "per-expert", n, k, groups, total, grc, st, worst,
           touched);
    if (!ok) ++g_fail;
    for (unsigned g = 0; g < groups; ++g) { cudaFree(d_qp[g]); cudaFree(d_scp[g]); }
    cudaFree(d_a); cudaFree(d_c); cudaFree(d_gs);
}

int main(int argc, char **argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0); // progress survives a timeout kill
    bool full = argc > 1 && !strcmp(argv[1], "full");
    // reference python cases (tests/ops/test_fp4_gemm_quark.py:27-35)
    run_case(false, true, 64, 128, 256, 0, 1.37f, true);
    run_case(false, true, 96, 64, 512, 0, 0.8f, true);
    run_case(true, true, 64, 128, 256, 0, 1.37f, true);
    run_case(true, true, 96, 96, 512, 0, 0.8f, true);
    run_case(false, false, 64, 128, 256, 0, 1.0f, true);
    // every token-tile variant, ragged M, partial n-tiles, split tiles
    const int toks[] = {16, 32, 64, 128, 256};
    for (int t : toks) {
        run_case(false, true, 1, 512, 1024, t, 1.0f, false);
        run_case(false, true, 17, 1280, 1024, t, 1.0f, false);
        run_case(true, true, 33, 320, 768, t, 1.0f, false);
        run_case(false, false, 7, 2048, 512, t, 1.0f, false);
    }
    run_case(false, true, 300, 1040, 768, 0, 1.0f, false);
    run_grouped(true, 1024, 2048, {3, 0, 17, 1, 16, 5, 40});
    run_grouped(false, 1024, 2048, {3, 0, 17, 1, 16, 5, 40});
    run_grouped(true, 528, 512, {1, 1, 2, 16});
    run_case(false, true, 566, 4096, 1024, 0, 1.0f, false);
    if (full) {
        run_case(false, true, 16, 10240, 8192, 0, 1.0f, true);
        run_case(false, true, 1, 8192, 8192, 0, 1.0f, false);
        run_case(true, true, 8, 8192, 8192, 0, 1.0f, true);
        run_case(false, false, 4, 10240, 8192, 0, 0.015625f, false); // keep fp16 outputs finite
        run_case(false, true, 1024, 8192, 8192, 0, 1.0f, false);
    }
    printf("%s (%d failures)\n", g_fail ? "SELFTEST FAILED" : "SELFTEST OK", g_fail);
    return g_fail ? 1 : 0;
}
