// C ABI (include/causalflow/petit/petit.h): argument checking, solution ids,
// the default-solution chooser, the stream-K workspace and the HAL shim.
//
// Replaces the reference's host dispatch layer: GemmFp4Fp16GridImpl / Dispatcher
// (fp4/gemm_fp4_fp16_grid.cc:11-95), GemmGetSolutions and
// ChooseDefaultFp4Fp16Solution (fp4/algo_chooser.cc:14-132) and lib/hal
// (device.h:8-34, rocm/platform_rocm.cc:17-68).  The reference instantiates 234
// MFMA tile shapes and looks them up in an unordered_map; here a solution is one
// of five token-tile widths of a single stream-K tcgen05 kernel, encoded in the
// reference's SolutionId bit layout (gemm.h:33-105) so ids stay opaque 64-bit
// integers for callers.
#include "causalflow/petit/petit.h"
#include "fp4_gemm.h"
#include "layout.cuh"
#include "repack.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <utility>
#include <vector>

namespace {

using namespace petit;

// ---- SolutionId bit layout (gemm.h:33-105) --------------------------------
constexpr uint64_t kFeatureGrid = 1;
constexpr uint64_t kElemNv = 1, kElemMx = 2;
constexpr uint64_t kMfmaF16 = 0, kMfmaBf16 = 1;
constexpr int kTokVariants[] = {16, 32, 64, 128, 256};
constexpr int kNumVariants = 5;

constexpr int stage_k_for(int ntok) { return ntok <= 64 ? 256 : (ntok == 128 ? 128 : 64); }

constexpr uint64_t make_solution(int ntok, uint64_t elem_b, uint64_t mfma) {
    return (uint64_t)(ntok / 16)                              // tile_m  [0,8)
           | ((uint64_t)(layout::kTileN / 16) << 8)            // tile_n  [8,16)
           | ((uint64_t)(stage_k_for(ntok) / 64) << 16)        // tile_k  [16,24) in units of 64
           | (kFeatureGrid << 24)                              // features
           | (elem_b << 28)                                    // element_b
           | (mfma << 32)                                      // mfma_type
           | (1ull << 36) | (4ull << 40) | (2ull << 44);       // warp partition m/n/k, NK
}

struct Decoded {
    int ntok;
    uint64_t elem_b, mfma;
};

bool decode_solution(uint64_t id, Decoded *d) {
    d->ntok = (int)(id & 0xff) * 16;
    d->elem_b = (id >> 28) & 0xf;
    d->mfma = (id >> 32) & 0xf;
    bool tok_ok = false;
    for (int v : kTokVariants) tok_ok |= v == d->ntok;
    if (!tok_ok) return false;
    if (d->elem_b != kElemNv && d->elem_b != kElemMx) return false;
    if (d->mfma != kMfmaF16 && d->mfma != kMfmaBf16) return false;
    return id == make_solution(d->ntok, d->elem_b, d->mfma);
}

// Default solution (the role of ChooseDefaultFp4Fp16Solution, algo_chooser.cc:64-132):
// the smallest token tile that holds M up to 128 tokens.  Beyond that, 256-token tiles
// run the main loop ~17 % faster than 128-token tiles (measured, Llama-70B shapes,
// M = 512..4096: 82-89 % vs 70-75 % of the bf16 peak) but pay a longer final epilogue
// and pad M to a multiple of 256, so they are chosen when the padded work still comes
// out ahead and every SM has enough 256-token units to amortise the tail.
int default_ntok(unsigned m, unsigned n, unsigned k) {
    if (const char *e = std::getenv("PETIT_FORCE_NTOK")) {
        int v = std::atoi(e);
        for (int t : kTokVariants)
            if (t == v) return v;
    }
    const unsigned long long n_tiles = (n + layout::kTileN - 1) / layout::kTileN;
    const unsigned long long k_tiles = k / layout::kTileK;
    // Shapes with few (n-tile x k-tile) units per SM (qkv, o_proj: < 28) are bound by the
    // per-launch tail, not by the main loop: there two or three 64-token tiles beat one 128- or
    // 256-token tile up to M = 192 (measured, profiles/r02_mid_m_tune.log: qkv M=128 28.9 vs
    // 32.6 us, M=192 38.7 vs 40.7; o M=192 33.0 vs 37.8).
    if (m > 64 && m <= 192 && n_tiles * k_tiles < 4096) return 64;
    for (int t : kTokVariants)
        if (t <= 128 && m <= (unsigned)t) return t;
    const unsigned long long pad256 = (m + 255ull) / 256 * 256, pad128 = (m + 127ull) / 128 * 128;
    const unsigned long long units256 = n_tiles * (pad256 / 256) * k_tiles;
    const bool faster = pad256 * 85 <= pad128 * 100;
    // (>= 27 units of 256 tokens per SM: o_proj M=512 56.8 us with 256-token tiles vs 62.5 with
    // 128, qkv 65.7 vs 75.9 -- profiles/r02_mid_m_tune.log; round 1's threshold of 40 predates
    // the faster exit path)
    return faster && units256 >= 27ull * 148 ? 256 : 128;
}

bool problem_shape_ok(unsigned n, unsigned k) {
    return n % 16 == 0 && k % layout::kTileK == 0;
}

// ---- tuned-solution table (petit_tune_table_*) --------------------------------
// key: (is_mx, a_type, m, n, k) -> solution id
using TuneKey = std::tuple<int, int, unsigned, unsigned, unsigned>;
std::mutex g_tune_mu;
std::map<TuneKey, uint64_t> g_tune;
std::atomic<bool> g_tune_nonempty{false};
std::once_flag g_tune_env_once;

bool hints_types_ok(const PetitSolutionHints *h, bool is_mx) {
    if (h->a_type != PETIT_DTYPE_FP16 && h->a_type != PETIT_DTYPE_BF16) return false;
    return !is_mx || h->a_type == PETIT_DTYPE_BF16;
}

// true if `id` is a solution petit_get_solutions lists for these types
bool solution_matches(uint64_t id, bool is_mx, int a_type) {
    Decoded d;
    if (!decode_solution(id, &d)) return false;
    if ((d.elem_b == kElemMx) != is_mx) return false;
    return (d.mfma == kMfmaBf16) == (a_type == PETIT_DTYPE_BF16);
}

int tune_load_file(const char *path) {
    std::FILE *f = std::fopen(path, "r");
    if (!f) return -1;
    std::vector<std::pair<TuneKey, uint64_t>> parsed;
    char line[256];
    bool ok = true;
    while (ok && std::fgets(line, sizeof line, f)) {
        if (char *hash = std::strchr(line, '#')) *hash = 0;
        char bt[16], at[16], hex[32];
        unsigned m, n, k;
        const int got = std::sscanf(line, "%15s %15s %u %u %u %31s", bt, at, &m, &n, &k, hex);
        if (got <= 0) continue; // blank / comment-only line
        const bool is_mx = !std::strcmp(bt, "mxfp4");
        const int a_type = !std::strcmp(at, "bf16") ? PETIT_DTYPE_BF16
                           : !std::strcmp(at, "fp16") ? PETIT_DTYPE_FP16 : -1;
        ok = got == 6 && (is_mx || !std::strcmp(bt, "nvfp4")) && a_type >= 0 &&
             std::strlen(hex) == 16;
        uint64_t id = 0;
        for (int b = 0; ok && b < 8; ++b) { // the id's 8 bytes, little endian
            unsigned byte = 0;
            ok = std::sscanf(hex + 2 * b, "%2x", &byte) == 1;
            id |= (uint64_t)byte << (8 * b);
        }
        ok = ok && solution_matches(id, is_mx, a_type);
        if (ok) parsed.push_back({TuneKey{is_mx, a_type, m, n, k}, id});
    }
    std::fclose(f);
    if (!ok) return -1;
    std::lock_guard<std::mutex> lock(g_tune_mu);
    for (auto &e : parsed) g_tune[e.first] = e.second;
    g_tune_nonempty = !g_tune.empty();
    return (int)parsed.size();
}

// 0 if no entry (or the table is empty: one relaxed load on the hot path)
uint64_t tune_lookup(bool is_mx, int a_type, unsigned m, unsigned n, unsigned k) {
    std::call_once(g_tune_env_once, [] {
        if (const char *e = std::getenv("PETIT_TUNE_TABLE")) tune_load_file(e);
    });
    if (!g_tune_nonempty.load(std::memory_order_relaxed)) return 0;
    std::lock_guard<std::mutex> lock(g_tune_mu);
    auto it = g_tune.find(TuneKey{is_mx, a_type, m, n, k});
    return it == g_tune.end() ? 0 : it->second;
}

// What PETIT_SOLUTION_AUTO resolves to (types and shape already validated).
Decoded default_solution(bool is_mx, int a_type, unsigned m, unsigned n, unsigned k) {
    Decoded d;
    if (uint64_t id = tune_lookup(is_mx, a_type, m, n, k)) {
        decode_solution(id, &d); // entries are validated when they are added
        return d;
    }
    d.ntok = default_ntok(m, n, k);
    d.elem_b = is_mx ? kElemMx : kElemNv;
    d.mfma = a_type == PETIT_DTYPE_BF16 ? kMfmaBf16 : kMfmaF16;
    return d;
}

// ---- per-(device, stream) stream-K workspace --------------------------------
struct Workspace {
    float *partials = nullptr;
    unsigned *counters = nullptr; // kMaxTiles tile counters + 1 status word
};
struct DeviceInfo {
    int num_sms = 0;
};
unsigned long long *g_trace = nullptr; // set by petit_debug_set_trace
std::mutex g_mu;
std::map<std::pair<int, cudaStream_t>, Workspace> g_ws;
std::deque<std::pair<int, cudaStream_t>> g_ws_order; // creation order, for the cap below
constexpr size_t kMaxWorkspaces = 16;                // x ~25 MB each
// Streams whose most recent petit call wrote packed weights / scales (repack_*): the weight
// producer warp of the GEMM does not execute griddepcontrol.wait (weights are constants), so
// a GEMM that directly follows such a call on the same stream is launched without
// programmatic dependent launch and therefore fully ordered behind it.
std::set<std::pair<int, cudaStream_t>> g_wrote_weights;

void note_weight_writer(cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(g_mu);
    g_wrote_weights.insert(std::make_pair(dev, stream));
}
// true (and forgotten) if the previous petit call on this stream wrote weights
bool take_weight_writer(cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    std::lock_guard<std::mutex> lock(g_mu);
    return g_wrote_weights.erase(std::make_pair(dev, stream)) != 0;
}
std::map<int, DeviceInfo> g_dev;

int get_context(cudaStream_t stream, Workspace *ws, int *num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PETIT_ERROR_CUDA;
    std::lock_guard<std::mutex> lock(g_mu);
    auto dit = g_dev.find(dev);
    if (dit == g_dev.end()) {
        DeviceInfo info;
        int major = 0;
        if (cudaDeviceGetAttribute(&info.num_sms, cudaDevAttrMultiProcessorCount, dev) !=
                cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
            return PETIT_ERROR_CUDA;
        if (major != 10) return PETIT_ERROR_CUDA; // sm_100a only; no fallback path
        if (info.num_sms > (int)gemm::kMaxGrid) info.num_sms = gemm::kMaxGrid;
        dit = g_dev.emplace(dev, info).first;
    }
    *num_sms = dit->second.num_sms;
    auto key = std::make_pair(dev, stream);
    auto it = g_ws.find(key);
    if (it == g_ws.end()) {
        // cudaMalloc is not capturable: step out of a possible stream capture.  The counters
        // are zeroed EAGERLY on an internal stream and waited for here, so the zeroing is
        // never recorded into a graph being captured on `stream` (it would then only run when
        // that one graph is replayed, and again on every replay) and is complete before any
        // later work -- eager or captured, on any stream -- can touch the workspace.
        cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
        cudaThreadExchangeStreamCaptureMode(&mode);
        Workspace w;
        cudaStream_t init = nullptr;
        cudaError_t e1 = cudaMalloc(&w.partials, gemm::workspace_partials_bytes());
        cudaError_t e2 = cudaMalloc(&w.counters, gemm::workspace_counters_bytes());
        cudaError_t e3 = cudaStreamCreateWithFlags(&init, cudaStreamNonBlocking);
        if (e2 == cudaSuccess && e3 == cudaSuccess)
            e3 = cudaMemsetAsync(w.counters, 0, gemm::workspace_counters_bytes(), init);
        if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(init);
        if (init) cudaStreamDestroy(init);
        cudaThreadExchangeStreamCaptureMode(&mode);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
            cudaFree(w.partials);
            cudaFree(w.counters);
            return PETIT_ERROR_CUDA;
        }
        // keep the footprint bounded: a process that cycles through many streams (torch's
        // stream pool) drops the workspaces of the least recently created streams
        if (g_ws.size() >= kMaxWorkspaces) {
            auto victim = g_ws_order.front();
            g_ws_order.pop_front();
            auto vit = g_ws.find(victim);
            if (vit != g_ws.end()) {
                // the victim's stream may still run a GEMM: let it drain before freeing
                cudaStreamCaptureMode m2 = cudaStreamCaptureModeRelaxed;
                cudaThreadExchangeStreamCaptureMode(&m2);
                cudaFree(vit->second.partials); // cudaFree synchronises the device
                cudaFree(vit->second.counters);
                cudaThreadExchangeStreamCaptureMode(&m2);
                g_ws.erase(vit);
            }
        }
        g_ws_order.push_back(key);
        it = g_ws.emplace(key, w).first;
    }
    *ws = it->second;
    return PETIT_OK;
}

int gemm_impl(void *c, const void *a, const void *b, const void *scales,
              const float *global_scale_dev, unsigned m, unsigned n, unsigned k,
              const PetitSolutionHints *hints, uint64_t solution_id, bool force_mx,
              cudaStream_t stream, const PetitFusedAllReduce *ar = nullptr,
              const PetitEpilogue *epi = nullptr, const gemm::GroupTable *table = nullptr) {
    if (ar) {
        // every rank must take part in every call: no early-out on empty shapes
        if (m == 0 || n == 0 || k == 0 || m > gemm::kArMaxTokens) return PETIT_ERROR_PROBLEM_SHAPE;
        if (ar->world < 2 || ar->world > (int)gemm::kArMaxWorld || ar->rank < 0 ||
            ar->rank >= ar->world || !ar->state)
            return PETIT_ERROR_PROBLEM_SHAPE;
        for (int r = 0; r < ar->world; ++r)
            if (!ar->recv[r] || (reinterpret_cast<uintptr_t>(ar->recv[r]) & 15))
                return PETIT_ERROR_PROBLEM_SHAPE;
    }
    if (m == 0 || n == 0 || k == 0) return PETIT_OK; // gemm_fp4_fp16_grid.cc:42-44
    if (!hints) return PETIT_ERROR_KERNEL_SHAPE;
    const bool is_mx = force_mx || hints->b_type == PETIT_DTYPE_MXFP4_E2M1;

    Decoded d;
    if (solution_id == PETIT_SOLUTION_AUTO) {
        if (!problem_shape_ok(n, k)) return PETIT_ERROR_PROBLEM_SHAPE;
        if (hints->a_type != PETIT_DTYPE_FP16 && hints->a_type != PETIT_DTYPE_BF16)
            return PETIT_ERROR_PROBLEM_SHAPE;
        d = default_solution(is_mx, hints->a_type, m, n, k);
    } else {
        if (!decode_solution(solution_id, &d)) return PETIT_ERROR_KERNEL_SHAPE;
        if (force_mx) d.elem_b = kElemMx; // gemm_fp4_fp16_grid.cc:91-94
    }
    if (d.elem_b == kElemMx) {
        // gemm_fp4_fp16_grid.cc:55-64
        if (k % 32 != 0) return PETIT_ERROR_PROBLEM_SHAPE;
        if (hints->a_type != PETIT_DTYPE_BF16 || hints->c_type != PETIT_DTYPE_BF16 ||
            d.mfma != kMfmaBf16)
            return PETIT_ERROR_KERNEL_SHAPE;
    }
    // an explicit id must agree with the activation type it will be fed
    if ((d.mfma == kMfmaBf16) != (hints->a_type == PETIT_DTYPE_BF16))
        return PETIT_ERROR_KERNEL_SHAPE;
    if (!problem_shape_ok(n, k)) return PETIT_ERROR_PROBLEM_SHAPE; // ConfigSelector::Invoke

    // TMA moves the activations in, the output out and the packed weights / scales
    // by bulk copy: all four base addresses must be 16-byte aligned (torch allocations
    // are; a view that starts at an odd element is not).
    if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(c) |
          reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(scales)) & 15) != 0)
        return PETIT_ERROR_PROBLEM_SHAPE;

    Workspace ws;
    int num_sms = 0;
    int err = get_context(stream, &ws, &num_sms);
    if (err != PETIT_OK) return err;

    gemm::GemmArgs args;
    args.a = a;
    args.w = static_cast<const uint8_t *>(b);
    args.sc = static_cast<const uint8_t *>(scales);
    args.global_scale = global_scale_dev;
    args.c = c;
    args.ws_partials = ws.partials;
    args.ws_counters = ws.counters;
    args.ws_status = ws.counters + gemm::kMaxTiles;
    {
        static const unsigned long long wd = [] {
            const char *e = std::getenv("PETIT_WATCHDOG_MS");
            const long long ms = e ? std::atoll(e) : 2000;
            return (unsigned long long)(ms > 0 ? ms : 2000) * 1000000ull;
        }();
        args.watchdog_ns = wd;
    }
    args.m = m;
    args.n = n;
    args.k = k;
    args.trace = g_trace;
    args.bias = nullptr;
    args.residual = nullptr;
    args.act_silu_mul = 0;
    if (epi) {
        if (epi->activation != PETIT_ACT_NONE && epi->activation != PETIT_ACT_SILU_MUL)
            return PETIT_ERROR_KERNEL_SHAPE;
        if (epi->activation == PETIT_ACT_SILU_MUL) {
            // gate / up rows interleaved per 128-row tile; no residual ([m, n] vs [m, n / 2]) and no
            // all-reduce (gate_up is column parallel); m in 16-token groups of up to 256 tokens
            if (n % layout::kTileN != 0 || epi->residual || ar) return PETIT_ERROR_PROBLEM_SHAPE;
            if (((n / 2) * 2) % 16 != 0) return PETIT_ERROR_PROBLEM_SHAPE;
        }
        // 2-byte elements, read with scalar loads: natural alignment is enough
        if (((reinterpret_cast<uintptr_t>(epi->bias) | reinterpret_cast<uintptr_t>(epi->residual)) & 1) != 0)
            return PETIT_ERROR_PROBLEM_SHAPE;
        args.bias = epi->bias;
        args.residual = epi->residual;
        args.act_silu_mul = epi->activation == PETIT_ACT_SILU_MUL ? 1u : 0u;
    }
    args.ar_world = 0;
    args.ar_rank = 0;
    args.ar_two_shot = 0;
    args.ar_state = nullptr;
    for (auto &p : args.ar_recv) p = nullptr;
    if (ar) {
        args.ar_world = (uint32_t)ar->world;
        // PETIT_AR_TWO_SHOT=0/1 overrides; must be the same on every rank
        static const int two_shot_env = [] {
            const char *e = std::getenv("PETIT_AR_TWO_SHOT");
            return e ? std::atoi(e) : -1;
        }();
        const bool can_two_shot = ar->world == 4 || ar->world == 8 || ar->world == 2;
        args.ar_two_shot = !can_two_shot ? 0u
                           : two_shot_env >= 0 ? (uint32_t)(two_shot_env != 0)
                                               : (ar->world > 4 ? 1u : 0u);
        // (measured, M = 16 layer set: world 4 one-shot 63.6 us per step vs two-shot 68.3 -- one
        // hop instead of two, 24 KB of packets per CTA; at world 8 the 56 KB per CTA cost more
        // than the second hop)
        args.ar_rank = (uint32_t)ar->rank;
        args.ar_state = static_cast<unsigned *>(ar->state);
        for (int r = 0; r < ar->world; ++r) args.ar_recv[r] = static_cast<uint8_t *>(ar->recv[r]);
    }
    {
        static const int pdl = [] {
            const char *e = std::getenv("PETIT_PDL");
            return e ? std::atoi(e) : 1;
        }();
        args.use_pdl = (uint32_t)pdl;
        if (take_weight_writer(stream)) args.use_pdl = 0;
        static const int cl = [] {
            const char *e = std::getenv("PETIT_CLUSTER");
            return e ? std::atoi(e) : 1;
        }();
        args.use_cluster = (uint32_t)cl;
        static const int dbg = [] {
            const char *e = std::getenv("PETIT_DEBUG_FLAGS");
            return e ? std::atoi(e) : 0;
        }();
        args.debug_flags = (uint32_t)dbg;
    }
    {
        static const int skew = [] {
            const char *e = std::getenv("PETIT_SKEW");
            return e ? std::atoi(e) : 0;
        }();
        args.skew_cycles = (uint32_t)skew;
    }
    int mode = d.elem_b == kElemMx
                   ? gemm::kModeMxBf16
                   : (d.mfma == kMfmaBf16 ? gemm::kModeNvBf16 : gemm::kModeNvF16);
    if (epi && epi->weight_layout != PETIT_WEIGHT_LAYOUT_DEFAULT) {
        // the fp16-native layout only fits NVFP4 weights multiplied with fp16 activations
        if (epi->weight_layout != PETIT_WEIGHT_LAYOUT_F16_NATIVE || mode != gemm::kModeNvF16)
            return PETIT_ERROR_KERNEL_SHAPE;
        mode = gemm::kModeNvF16N;
    }
    switch (table ? gemm::launch_grouped(mode, d.ntok, args, *table, num_sms, stream)
                  : gemm::launch(mode, d.ntok, args, num_sms, stream)) {
    case gemm::kLaunchOk: return PETIT_OK;
    case gemm::kLaunchBadShape: return PETIT_ERROR_PROBLEM_SHAPE;
    case gemm::kLaunchNoKernel: return PETIT_ERROR_KERNEL_SHAPE;
    default: return PETIT_ERROR_CUDA;
    }
}

int dequant_mode(int out_type, bool mx) {
    if (mx) return out_type == PETIT_DTYPE_BF16 ? gemm::kModeMxBf16 : -1;
    if (out_type == PETIT_DTYPE_BF16) return gemm::kModeNvBf16;
    if (out_type == PETIT_DTYPE_FP16) return gemm::kModeNvF16;
    return -1;
}

} // namespace

extern "C" {

int petit_gemm_nvfp4_a16(void *c, const void *a, const void *b, const void *scales,
                         const float *global_scale_dev, unsigned m, unsigned n, unsigned k,
                         const PetitSolutionHints *hints, uint64_t solution_id,
                         petit_stream_t stream) {
    return gemm_impl(c, a, b, scales, global_scale_dev, m, n, k, hints, solution_id, false,
                     reinterpret_cast<cudaStream_t>(stream));
}

int petit_gemm_mxfp4_a16(void *c, const void *a, const void *b, const void *scales,
                         const float *global_scale_dev, unsigned m, unsigned n, unsigned k,
                         const PetitSolutionHints *hints, uint64_t solution_id,
                         petit_stream_t stream) {
    return gemm_impl(c, a, b, scales, global_scale_dev, m, n, k, hints, solution_id, true,
                     reinterpret_cast<cudaStream_t>(stream));
}

int petit_gemm_nvfp4_a16_allreduce(void *c, const void *a, const void *b, const void *scales,
                                   const float *global_scale_dev, unsigned m, unsigned n,
                                   unsigned k, const PetitSolutionHints *hints,
                                   uint64_t solution_id, const PetitFusedAllReduce *ar,
                                   petit_stream_t stream) {
    if (!ar) return PETIT_ERROR_PROBLEM_SHAPE;
    return gemm_impl(c, a, b, scales, global_scale_dev, m, n, k, hints, solution_id, false,
                     reinterpret_cast<cudaStream_t>(stream), ar);
}

int petit_gemm_mxfp4_a16_allreduce(void *c, const void *a, const void *b, const void *scales,
                                   const float *global_scale_dev, unsigned m, unsigned n,
                                   unsigned k, const PetitSolutionHints *hints,
                                   uint64_t solution_id, const PetitFusedAllReduce *ar,
                                   petit_stream_t stream) {
    if (!ar) return PETIT_ERROR_PROBLEM_SHAPE;
    return gemm_impl(c, a, b, scales, global_scale_dev, m, n, k, hints, solution_id, true,
                     reinterpret_cast<cudaStream_t>(stream), ar);
}

int petit_gemm_nvfp4_a16_ex(void *c, const void *a, const void *b, const void *scales,
                            const float *global_scale_dev, unsigned m, unsigned n, unsigned k,
                            const PetitSolutionHints *hints, uint64_t solution_id,
                            const PetitEpilogue *epilogue, const PetitFusedAllReduce *ar,
                            petit_stream_t stream) {
    return gemm_impl(c, a, b, scales, global_scale_dev, m, n, k, hints, solution_id, false,
                     reinterpret_cast<cudaStream_t>(stream), ar, epilogue);
}

int petit_gemm_mxfp4_a16_ex(void *c, const void *a, const void *b, const void *scales,
                            const float *global_scale_dev, unsigned m, unsigned n, unsigned k,
                            const PetitSolutionHints *hints, uint64_t solution_id,
                            const PetitEpilogue *epilogue, const PetitFusedAllReduce *ar,
                            petit_stream_t stream) {
    return gemm_impl(c, a, b, scales, global_scale_dev, m, n, k, hints, solution_id, true,
                     reinterpret_cast<cudaStream_t>(stream), ar, epilogue);
}

/* Test hook (not in petit.h): the stream-K range cuts for `units` units in tiles of `k_tiles`
 * on `grid` CTAs; cuts[b] = first unit of CTA b, cuts[grid] = units.  Returns 0, -1 on bad input. */
int petit_debug_stream_k_cuts(unsigned units, unsigned k_tiles, unsigned grid, int lat, int late,
                              unsigned *cuts) {
    if (!cuts || grid == 0 || grid > gemm::kMaxGrid || k_tiles == 0 || units < grid) return -1;
    int8_t adj[gemm::kMaxGrid + 4];
    gemm::debug_stream_k_cuts(units, k_tiles, grid, lat, late, adj);
    for (unsigned b = 0; b <= grid; ++b)
        cuts[b] = (unsigned)((long long)((unsigned long long)units * b / grid) + adj[b]);
    return 0;
}

int petit_gemm_fp4_a16_grouped(const PetitGroupedProblem *problems, unsigned num_groups, unsigned n,
                               unsigned k, const PetitSolutionHints *hints, uint64_t solution_id,
                               const PetitEpilogue *epilogue, petit_stream_t stream) {
    if (!problems && num_groups) return PETIT_ERROR_PROBLEM_SHAPE;
    if (!hints) return PETIT_ERROR_KERNEL_SHAPE;
    if (epilogue && epilogue->residual) return PETIT_ERROR_PROBLEM_SHAPE; // per-group shapes differ
    const bool mx = hints->b_type == PETIT_DTYPE_MXFP4_E2M1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    // One launch for all groups when the groups' activations and outputs are consecutive row
    // blocks of one tensor (tokens sorted by expert), the token tile is a decode tile and the
    // tile table fits the kernel parameters; otherwise the groups are issued back to back.
    // PETIT_GROUPED_SINGLE=0 forces the second form (A/B runs, tests).
    const char *single_env = std::getenv("PETIT_GROUPED_SINGLE");
    const bool plain_epi = !epilogue || (!epilogue->bias && !epilogue->residual);
    const bool silu = epilogue && epilogue->activation == PETIT_ACT_SILU_MUL;
    const size_t out_n = silu ? n / 2 : n;
    if (!(single_env && std::atoi(single_env) == 0) && plain_epi && n != 0 && k != 0) {
        const size_t esz = 2; // fp16 / bf16
        const PetitGroupedProblem *first = nullptr, *prev = nullptr;
        unsigned max_m = 0, live = 0;
        uint64_t total_rows = 0;
        bool contiguous = true;
        for (unsigned g = 0; g < num_groups && contiguous; ++g) {
            const PetitGroupedProblem &p = problems[g];
            if (p.m == 0) continue;
            if (!first) first = &p;
            if (prev && (static_cast<const char *>(p.a) !=
                             static_cast<const char *>(prev->a) + (size_t)prev->m * k * esz ||
                         static_cast<char *>(p.c) != static_cast<char *>(prev->c) + (size_t)prev->m * out_n * esz))
                contiguous = false;
            if (((reinterpret_cast<uintptr_t>(p.b) | reinterpret_cast<uintptr_t>(p.scales)) & 15) ||
                !p.global_scale_dev)
                contiguous = false; // let the per-group path report it
            prev = &p;
            max_m = p.m > max_m ? p.m : max_m;
            total_rows += p.m;
            ++live;
        }
        int ntok = 0;
        uint64_t sid = solution_id;
        if (contiguous && live > 1 && total_rows < (1ull << 31)) {
            if (solution_id == PETIT_SOLUTION_AUTO) {
                if (hints->a_type == PETIT_DTYPE_FP16 || hints->a_type == PETIT_DTYPE_BF16) {
                    ntok = max_m <= 16 ? 16 : (max_m <= 32 ? 32 : 64);
                    sid = make_solution(ntok, mx ? kElemMx : kElemNv,
                                        hints->a_type == PETIT_DTYPE_BF16 ? kMfmaBf16 : kMfmaF16);
                }
            } else {
                Decoded d;
                if (decode_solution(solution_id, &d) && d.ntok <= 64) ntok = d.ntok;
            }
        }
        if (ntok) {
            gemm::GroupTable table;
            table.tiles = 0;
            table.pad = 0;
            uint32_t row = 0;
            bool fits = true;
            for (unsigned g = 0; g < num_groups && fits; ++g) {
                const PetitGroupedProblem &p = problems[g];
                for (unsigned r = 0; r < p.m; r += (unsigned)ntok) {
                    if (table.tiles == gemm::kMaxGroupTiles) {
                        fits = false;
                        break;
                    }
                    gemm::GroupEntry &e = table.e[table.tiles++];
                    e.w = static_cast<const uint8_t *>(p.b);
                    e.sc = static_cast<const uint8_t *>(p.scales);
                    e.gs = p.global_scale_dev;
                    e.row0 = row + r;
                    e.rows = p.m - r < (unsigned)ntok ? p.m - r : (unsigned)ntok;
                }
                row += p.m;
            }
            if (fits)
                return gemm_impl(first->c, first->a, first->b, first->scales, first->global_scale_dev,
                                 (unsigned)total_rows, n, k, hints, sid, mx, st, nullptr, epilogue,
                                 &table);
        }
    }
    for (unsigned g = 0; g < num_groups; ++g) {
        const PetitGroupedProblem &p = problems[g];
        if (p.m == 0) continue;
        const int rc = gemm_impl(p.c, p.a, p.b, p.scales, p.global_scale_dev, p.m, n, k, hints,
                                 solution_id, mx, st, nullptr, epilogue);
        if (rc != PETIT_OK) return rc;
    }
    return PETIT_OK;
}

size_t petit_fused_allreduce_recv_bytes(unsigned n) { return gemm::ar_recv_bytes(n); }
size_t petit_fused_allreduce_state_bytes(void) { return 4 * sizeof(unsigned); }

int petit_fused_allreduce_status(const void *state, petit_stream_t stream) {
    unsigned st = 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!state ||
        cudaMemcpyAsync(&st, static_cast<const unsigned *>(state) + 2, sizeof st,
                        cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return -1;
    return (int)st;
}

int petit_get_solutions(const PetitSolutionHints *hints, unsigned m, unsigned n, unsigned k,
                        uint64_t *sols, unsigned *n_sols) {
    (void)m;
    if (!hints || !n_sols) return -1;
    // algo_chooser.cc:20-23
    if (hints->b_type != PETIT_DTYPE_FP4_E2M1 && hints->b_type != PETIT_DTYPE_MXFP4_E2M1)
        return -1;
    const bool mx = hints->b_type == PETIT_DTYPE_MXFP4_E2M1;
    unsigned count = 0;
    const bool type_ok = (hints->a_type == PETIT_DTYPE_BF16) ||
                         (hints->a_type == PETIT_DTYPE_FP16 && !mx);
    if (type_ok && problem_shape_ok(n, k) && n != 0 && k != 0) {
        const uint64_t mfma = hints->a_type == PETIT_DTYPE_BF16 ? kMfmaBf16 : kMfmaF16;
        for (int i = 0; i < kNumVariants; ++i) {
            if (sols && count < *n_sols)
                sols[count] = make_solution(kTokVariants[i], mx ? kElemMx : kElemNv, mfma);
            ++count;
        }
    }
    *n_sols = count;
    return 0;
}

int petit_get_default_solution(const PetitSolutionHints *hints, unsigned m, unsigned n,
                               unsigned k, uint64_t *solution_id) {
    if (!hints || !solution_id) return -1;
    if (hints->b_type != PETIT_DTYPE_FP4_E2M1 && hints->b_type != PETIT_DTYPE_MXFP4_E2M1)
        return -1;
    const bool is_mx = hints->b_type == PETIT_DTYPE_MXFP4_E2M1;
    if (!hints_types_ok(hints, is_mx) || !problem_shape_ok(n, k) || m == 0 || n == 0 || k == 0)
        return PETIT_ERROR_PROBLEM_SHAPE;
    const Decoded d = default_solution(is_mx, hints->a_type, m, n, k);
    *solution_id = make_solution(d.ntok, d.elem_b, d.mfma);
    return PETIT_OK;
}

int petit_tune_table_set(const PetitSolutionHints *hints, unsigned m, unsigned n, unsigned k,
                         uint64_t solution_id) {
    if (!hints) return -1;
    if (hints->b_type != PETIT_DTYPE_FP4_E2M1 && hints->b_type != PETIT_DTYPE_MXFP4_E2M1)
        return -1;
    const bool is_mx = hints->b_type == PETIT_DTYPE_MXFP4_E2M1;
    if (!hints_types_ok(hints, is_mx)) return PETIT_ERROR_PROBLEM_SHAPE;
    const TuneKey key{is_mx, hints->a_type, m, n, k};
    std::lock_guard<std::mutex> lock(g_tune_mu);
    if (solution_id == PETIT_SOLUTION_AUTO) {
        g_tune.erase(key);
    } else {
        if (!solution_matches(solution_id, is_mx, hints->a_type)) return PETIT_ERROR_KERNEL_SHAPE;
        g_tune[key] = solution_id;
    }
    g_tune_nonempty = !g_tune.empty();
    return PETIT_OK;
}

int petit_tune_table_load(const char *path) { return path ? tune_load_file(path) : -1; }

void petit_tune_table_clear(void) {
    std::call_once(g_tune_env_once, [] {}); // a later lookup must not re-read $PETIT_TUNE_TABLE
    std::lock_guard<std::mutex> lock(g_tune_mu);
    g_tune.clear();
    g_tune_nonempty = false;
}

int petit_repack_fp4_weights(uint32_t *out, const uint32_t *in, unsigned in_chan,
                             unsigned out_chan, petit_stream_t stream) {
    note_weight_writer(reinterpret_cast<cudaStream_t>(stream));
    return repack::weights(out, in, in_chan, out_chan, false,
                           reinterpret_cast<cudaStream_t>(stream));
}

int petit_repack_fp4_weights_layout(uint32_t *out, const uint32_t *in, unsigned in_chan,
                                    unsigned out_chan, int weight_layout, petit_stream_t stream) {
    if (weight_layout != PETIT_WEIGHT_LAYOUT_DEFAULT && weight_layout != PETIT_WEIGHT_LAYOUT_F16_NATIVE)
        return PETIT_ERROR_PROBLEM_SHAPE;
    note_weight_writer(reinterpret_cast<cudaStream_t>(stream));
    return repack::weights(out, in, in_chan, out_chan, false, reinterpret_cast<cudaStream_t>(stream),
                           weight_layout == PETIT_WEIGHT_LAYOUT_F16_NATIVE);
}

int petit_unpack_fp4_weights_layout(uint32_t *out, const uint32_t *in_packed, unsigned in_chan,
                                    unsigned out_chan, int weight_layout, petit_stream_t stream) {
    if (weight_layout != PETIT_WEIGHT_LAYOUT_DEFAULT && weight_layout != PETIT_WEIGHT_LAYOUT_F16_NATIVE)
        return PETIT_ERROR_PROBLEM_SHAPE;
    return repack::weights(out, in_packed, in_chan, out_chan, true,
                           reinterpret_cast<cudaStream_t>(stream),
                           weight_layout == PETIT_WEIGHT_LAYOUT_F16_NATIVE);
}

int petit_dequant_packed_nvfp4_layout(void *out, const void *w_packed, const void *scales_packed,
                                      float global_scale, int out_type, unsigned k, unsigned n,
                                      int weight_layout, petit_stream_t stream) {
    int mode = dequant_mode(out_type, false);
    if (mode < 0) return -1;
    if (weight_layout == PETIT_WEIGHT_LAYOUT_F16_NATIVE) {
        if (mode != gemm::kModeNvF16) return -1;
        mode = gemm::kModeNvF16N;
    } else if (weight_layout != PETIT_WEIGHT_LAYOUT_DEFAULT) {
        return -1;
    }
    return repack::dequant_dense(out, w_packed, scales_packed, global_scale, mode, true, k, n,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int petit_unpack_fp4_weights(uint32_t *out, const uint32_t *in_packed, unsigned in_chan,
                             unsigned out_chan, petit_stream_t stream) {
    return repack::weights(out, in_packed, in_chan, out_chan, true,
                           reinterpret_cast<cudaStream_t>(stream));
}

int petit_repack_nvfp4_scales(void *out, const void *in, unsigned in_chan, unsigned out_chan,
                              petit_stream_t stream) {
    note_weight_writer(reinterpret_cast<cudaStream_t>(stream));
    return repack::scales(out, in, in_chan, out_chan, false, false,
                          reinterpret_cast<cudaStream_t>(stream));
}

int petit_repack_mxfp4_scales(void *out, const void *in, unsigned in_chan, unsigned out_chan,
                              petit_stream_t stream) {
    note_weight_writer(reinterpret_cast<cudaStream_t>(stream));
    return repack::scales(out, in, in_chan, out_chan, true, false,
                          reinterpret_cast<cudaStream_t>(stream));
}

int petit_dequant_nvfp4(void *out, const void *w, const void *scales, float global_scale,
                        int out_type, unsigned k, unsigned n, petit_stream_t stream) {
    const int mode = dequant_mode(out_type, false);
    if (mode < 0) return -1;
    return repack::dequant_dense(out, w, scales, global_scale, mode, false, k, n,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int petit_dequant_mxfp4(void *out, const void *w, const void *scales, float global_scale,
                        int out_type, unsigned k, unsigned n, petit_stream_t stream) {
    const int mode = dequant_mode(out_type, true);
    if (mode < 0) return -1;
    return repack::dequant_dense(out, w, scales, global_scale, mode, false, k, n,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int petit_dequant_packed_nvfp4(void *out, const void *w_packed, const void *scales_packed,
                               float global_scale, int out_type, unsigned k, unsigned n,
                               petit_stream_t stream) {
    const int mode = dequant_mode(out_type, false);
    if (mode < 0) return -1;
    return repack::dequant_dense(out, w_packed, scales_packed, global_scale, mode, true, k, n,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int petit_dequant_packed_mxfp4(void *out, const void *w_packed, const void *scales_packed,
                               float global_scale, int out_type, unsigned k, unsigned n,
                               petit_stream_t stream) {
    const int mode = dequant_mode(out_type, true);
    if (mode < 0) return -1;
    return repack::dequant_dense(out, w_packed, scales_packed, global_scale, mode, true, k, n,
                                 reinterpret_cast<cudaStream_t>(stream));
}

// ---- HAL shim (lib/hal/device.h:8-34) ---------------------------------------
int petit_hal_device_count(int *count) { return (int)cudaGetDeviceCount(count); }
int petit_hal_set_device(int device) { return (int)cudaSetDevice(device); }
int petit_hal_malloc(void **ptr, size_t bytes) { return (int)cudaMalloc(ptr, bytes); }
int petit_hal_free(void *ptr) { return (int)cudaFree(ptr); }
int petit_hal_memset(void *ptr, int value, size_t bytes) {
    return (int)cudaMemset(ptr, value, bytes);
}
int petit_hal_copy_to_device(void *dst, const void *src, size_t bytes) {
    return (int)cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
}
int petit_hal_copy_to_host(void *dst, const void *src, size_t bytes) {
    return (int)cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
}
int petit_hal_synchronize(void) { return (int)cudaDeviceSynchronize(); }

int petit_packed_layout_version(void) { return layout::kLayoutVersion; }

int petit_workspace_status(petit_stream_t stream, int *status) {
    if (!status) return -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PETIT_ERROR_CUDA;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned *word = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_ws.find(std::make_pair(dev, s));
        if (it == g_ws.end()) {
            *status = 0; // no GEMM has run on this stream yet
            return PETIT_OK;
        }
        word = it->second.counters + gemm::kMaxTiles;
    }
    unsigned v = 0;
    if (cudaMemcpyAsync(&v, word, sizeof v, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return PETIT_ERROR_CUDA;
    *status = (int)v;
    return PETIT_OK;
}

int petit_release_workspace(petit_stream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PETIT_ERROR_CUDA;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    Workspace w;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto key = std::make_pair(dev, s);
        auto it = g_ws.find(key);
        if (it == g_ws.end()) return PETIT_OK;
        w = it->second;
        g_ws.erase(it);
        for (auto o = g_ws_order.begin(); o != g_ws_order.end(); ++o)
            if (*o == key) {
                g_ws_order.erase(o);
                break;
            }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) return PETIT_ERROR_CUDA;
    cudaFree(w.partials);
    cudaFree(w.counters);
    return PETIT_OK;
}

// Debug hook (not in petit.h): device buffer of [grid][16] u64 globaltimer stamps.
void petit_debug_set_trace(unsigned long long *dev_buffer) { g_trace = dev_buffer; }

const char *petit_solution_name(uint64_t solution_id) {
    static const char *names[3][kNumVariants] = {
        {"sm100_streamk_nvfp4_f16_tok16", "sm100_streamk_nvfp4_f16_tok32",
         "sm100_streamk_nvfp4_f16_tok64", "sm100_streamk_nvfp4_f16_tok128",
         "sm100_streamk_nvfp4_f16_tok256"},
        {"sm100_streamk_nvfp4_bf16_tok16", "sm100_streamk_nvfp4_bf16_tok32",
         "sm100_streamk_nvfp4_bf16_tok64", "sm100_streamk_nvfp4_bf16_tok128",
         "sm100_streamk_nvfp4_bf16_tok256"},
        {"sm100_streamk_mxfp4_bf16_tok16", "sm100_streamk_mxfp4_bf16_tok32",
         "sm100_streamk_mxfp4_bf16_tok64", "sm100_streamk_mxfp4_bf16_tok128",
         "sm100_streamk_mxfp4_bf16_tok256"}};
    Decoded d;
    if (!decode_solution(solution_id, &d)) return "";
    int mode = d.elem_b == kElemMx ? 2 : (d.mfma == kMfmaBf16 ? 1 : 0);
    if (d.elem_b == kElemMx && d.mfma != kMfmaBf16) return "";
    for (int i = 0; i < kNumVariants; ++i)
        if (kTokVariants[i] == d.ntok) return names[mode][i];
    return "";
}

} // extern "C"
