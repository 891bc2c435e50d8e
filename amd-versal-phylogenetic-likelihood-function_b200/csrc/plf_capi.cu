// csrc/plf_capi.cu -- the C ABI of include/b200plf.h over the CUDA runtime.
//
// Replaces the XRT control code of the reference host (app/src/include.h:28-147 acap_info,
// app/src/host_mem.cpp:108-164,249-325): device open, device-only buffers, async write / run /
// read on per-instance queues.  Here an "instance" (NUM_ACCELERATORS, Makefile:29) is a set of
// device buffers plus one CUDA stream; the three PL kernels + AIE graph are one fused kernel.
//
// No exception leaves this file and there is no CPU fallback: every failure is a status code.
#include "../../include/b200plf.h"
#include "plf_kernels.cuh"
#include "plf_registry.h"

#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string g_last_error = "";
constexpr unsigned kSumSlots = 1024;
std::atomic<unsigned long long> g_launches{0};

struct Instance {
    bool allocated = false;
    size_t max_sites = 0;
    float *d_left = nullptr;
    float *d_right = nullptr;
    float *d_out = nullptr;
    unsigned char *d_scaler = nullptr;
    int *d_wgt = nullptr;
    bool use_wgt = false;
    unsigned long long *d_sum = nullptr;    // ring of kSumSlots counters, one per run (no per-run memset)
    unsigned sum_slot = 0;                  // slot of the last run
    unsigned long long runs = 0;
    double *d_check = nullptr;              // gen-discard checksum of the last run
    unsigned long long *h_sum = nullptr;    // pinned landing buffers for the read-backs
    double *h_check = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
    bool mark_set[4] = {false, false, false, false};
    std::string error;
};

}  // namespace

struct plf_ctx {
    int device = 0;
    int states = 4;                         // STATES knob: 4 (DNA) or 20 (protein)
    int layout = PLF_LAYOUT_COMB;
    int input_src = PLF_INPUT_MEM;
    int math = PLF_MATH_STRICT;
    int gen_sink = PLF_GEN_WRITE;
    int variant = 0, threads = 0, blocks_per_sm = 0;
    int num_sms = 0;
    float *d_gen = nullptr;                 // gen pattern: x1[16] x2[16] ev4[64] pl[64] pr[64]
    // streamed host path (plf_newview_stream): kStreamSlots chunk buffers + streams, allocated lazily
    static constexpr int kStreamSlots = 3;
    size_t stream_chunk = 0;                // sites per chunk the slots are sized for
    float *sd_x1[kStreamSlots] = {nullptr, nullptr, nullptr};
    float *sd_x2[kStreamSlots] = {nullptr, nullptr, nullptr};
    float *sd_x3[kStreamSlots] = {nullptr, nullptr, nullptr};
    unsigned char *sd_sc[kStreamSlots] = {nullptr, nullptr, nullptr};
    int *sd_wgt[kStreamSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t s_stream[kStreamSlots] = {nullptr, nullptr, nullptr};
    float *sd_mats = nullptr;               // EV[16] | P_left[64] | P_right[64]
    unsigned long long *sd_sum = nullptr, *sh_sum = nullptr;
    std::vector<Instance> inst;
    std::mutex err_mu;
    std::string error;
};

namespace {

int fail(plf_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        ctx->error = buf;
    }
    g_last_error = buf;
    return code;
}

#define PLF_CUDA(ctx, expr)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA,   \
                        "%s failed: %s", #expr, cudaGetErrorString(e__));                       \
    } while (0)

// ---- kernel variant registry: see plf_registry.h for the variant encoding ----------------------
// Default: tma ring with dynamic stage scheduling, 16 consumer warps, U=2 (256-site / 32 KB stages),
// 4 stages, 1 CTA per SM -- the fastest strict configuration at 8 Mi and 64 Mi sites per launch on
// B200 (profiles/r01_sweep.md, "static vs dynamic").
constexpr int kDefaultVariant = 1432;
constexpr int kDefaultThreads = 512;
// One shape for every launch size: at 1 Mi sites per launch (BASELINE configs[1]) the two-CTA-per-SM shapes (8 consumer
// warps, 96 registers, 96 KB rings: variants 2332, 2334, 2632 with 256 threads) measured within 1 % of this one, serial
// or spread over nine instance streams (profiles/r02_small_launch.md), so there is no size-dependent rule.
using plf::KernelSel;
using plf::NewviewFn;

KernelSel pick_kernel(int math, int variant, int threads)
{
    if (variant < 0 || variant > 9999) return KernelSel{};
    const int u = variant % 10, kind = (variant / 10) % 10, d = (variant / 100) % 10, b = variant / 1000;
    const bool fma = math == PLF_MATH_FMA;
    if (kind == 2) return fma ? plf::select_tma_fma(u, d, b, threads) : plf::select_tma_strict(u, d, b, threads);
    if (kind == 3) return fma ? plf::select_tma_dyn_fma(u, d, b, threads) : plf::select_tma_dyn_strict(u, d, b, threads);
    if ((kind == 0 || kind == 1) && d == 0)
        return fma ? plf::select_ldg_fma(u, kind, b, threads) : plf::select_ldg_strict(u, kind, b, threads);
    return KernelSel{};
}

int device_sms(int *sms)
{
    static std::atomic<int> cached[64];      // SM count per device ordinal, 0 = not read yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev >= 0 && dev < 64) {
        const int c = cached[dev].load(std::memory_order_relaxed);
        if (c > 0) {
            *sms = c;
            return 0;
        }
    }
    if (cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (dev >= 0 && dev < 64) cached[dev].store(*sms, std::memory_order_relaxed);
    return 0;
}

// Per (device, kernel): the > 48 KB shared-memory opt-in has been made and the occupancy is known.  Filled on the
// first launch, so the run path of a 30 us kernel does not pay three runtime queries per launch.
struct KernelState {
    bool prepared = false;
    int occupancy = 0;
};
std::mutex g_kstate_mu;
std::map<std::pair<int, const void *>, KernelState> g_kstate;

int kernel_state(plf_ctx *ctx, const KernelSel &k, KernelState *out)
{
    int dev = 0;
    PLF_CUDA(ctx, cudaGetDevice(&dev));
    const auto key = std::make_pair(dev, reinterpret_cast<const void *>(k.fn));
    {
        std::lock_guard<std::mutex> g(g_kstate_mu);
        auto it = g_kstate.find(key);
        if (it != g_kstate.end()) {
            *out = it->second;
            return PLF_OK;
        }
    }
    KernelState st;
    if (k.smem > 48u * 1024u)
        PLF_CUDA(ctx, cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem));
    st.prepared = true;
    PLF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st.occupancy, k.fn, k.threads, k.smem));
    if (st.occupancy <= 0) st.occupancy = 1;
    {
        std::lock_guard<std::mutex> g(g_kstate_mu);
        g_kstate[key] = st;
    }
    *out = st;
    return PLF_OK;
}

// Opt in to > 48 KB of dynamic shared memory once per kernel and device.
int prepare_kernel(plf_ctx *ctx, const KernelSel &k)
{
    KernelState st;
    return kernel_state(ctx, k, &st);
}

// Resolves defaults and the persistent grid size: SMs x resident blocks per SM, capped by the
// amount of work so that small inputs do not launch idle blocks.
int resolve_launch(plf_ctx *ctx, const plf_launch_opts *opts, size_t n, KernelSel *sel, int *grid)
{
    int math = opts ? opts->math_mode : PLF_MATH_STRICT;
    int variant = opts ? opts->variant : 0;
    int thr = opts ? opts->threads_per_block : 0;
    if (variant == 0) variant = kDefaultVariant;
    if (thr == 0) thr = kDefaultThreads;
    int bps = opts ? opts->blocks_per_sm : 0;
    if (math != PLF_MATH_STRICT && math != PLF_MATH_FMA)
        return fail(ctx, PLF_ERR_INVALID, "unknown math mode %d", math);
    KernelSel k = pick_kernel(math, variant, thr);
    if (!k.fn)
        return fail(ctx, PLF_ERR_INVALID, "unknown kernel variant %d / threads %d", variant, thr);
    int sms = 0;
    if (device_sms(&sms) != 0) return fail(ctx, PLF_ERR_CUDA, "no CUDA device");
    KernelState st;
    int rc = kernel_state(ctx, k, &st);
    if (rc != PLF_OK) return rc;
    if (bps <= 0) bps = st.occupancy;
    const size_t per_iter = (size_t)k.sites_per_block_iter;
    const size_t blocks_needed = (n + per_iter - 1) / per_iter;
    size_t g = (size_t)sms * (size_t)bps;
    if (g > blocks_needed) g = blocks_needed;
    if (g == 0 || (opts && (opts->flags & PLF_LAUNCH_SINGLE_CTA))) g = 1;
    *sel = k;
    *grid = (int)g;
    return PLF_OK;
}

// Per-stream device scratch: the work-counter pair of the dynamically scheduled kernels, the ticket and the
// per-block partial sums of the deterministic log-likelihood reduction, and a staging slot for host matrices.
// Launches on one stream are serialised, so a stream's launches can share one record; every (device, stream) pair
// gets its own on first use, so concurrent launches on different streams never share a work counter however many
// are in flight.  The slab holds kScratchSlots records per device; past that, streams share records round-robin.
constexpr unsigned kScratchSlots = 1024;
std::mutex g_scratch_mu;
plf::StreamScratch *g_scratch_dev[64] = {nullptr};
std::map<std::pair<int, cudaStream_t>, unsigned> g_scratch_of;
unsigned g_scratch_next[64] = {0};

}  // namespace

int plf::stream_scratch(cudaStream_t stream, plf::StreamScratch **out)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PLF_ERR_CUDA;
    std::lock_guard<std::mutex> g(g_scratch_mu);
    if (!g_scratch_dev[dev]) {
        plf::StreamScratch *d = nullptr;
        if (cudaMalloc(&d, kScratchSlots * sizeof(plf::StreamScratch)) != cudaSuccess) return PLF_ERR_NOMEM;
        if (cudaMemset(d, 0, kScratchSlots * sizeof(plf::StreamScratch)) != cudaSuccess) return PLF_ERR_CUDA;
        g_scratch_dev[dev] = d;
    }
    const auto key = std::make_pair(dev, stream);
    auto it = g_scratch_of.find(key);
    unsigned slot;
    if (it != g_scratch_of.end()) {
        slot = it->second;
    } else {
        slot = g_scratch_next[dev]++ % kScratchSlots;
        g_scratch_of[key] = slot;
    }
    *out = g_scratch_dev[dev] + slot;
    return PLF_OK;
}

namespace {

int work_pair(plf_ctx *ctx, cudaStream_t stream, unsigned long long **out)
{
    plf::StreamScratch *sc = nullptr;
    int rc = plf::stream_scratch(stream, &sc);
    if (rc != PLF_OK) return fail(ctx, rc, "per-stream scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    *out = sc->work;
    return PLF_OK;
}

// Ring-slot release mechanism of the bulk-copy kernels (plf_kernels.cuh, mbar_release_slot*): the fenced release is
// the default of the DRAM-bound kernels; PLF_SAFE_RELEASE=1 forces it everywhere (also the tree kernel), =0 forces the
// data-dependency release everywhere, so a field failure can be bisected without a rebuild.
// For the 20-state dispatcher, which picks the kernel (and with it the family default) itself: the caller's explicit
// choice, or plf::kAaReleaseUnset.
int release_request(const plf_launch_opts *opts)
{
    if (opts && (opts->flags & PLF_LAUNCH_FENCED_RELEASE)) return plf::kFlagFencedRelease;
    if (opts && (opts->flags & PLF_LAUNCH_DEP_RELEASE)) return 0;
    return plf::kAaReleaseUnset;
}

int release_flag(const plf_launch_opts *opts, bool default_fenced)
{
    if (opts && (opts->flags & PLF_LAUNCH_FENCED_RELEASE)) return plf::kFlagFencedRelease;
    if (opts && (opts->flags & PLF_LAUNCH_DEP_RELEASE)) return 0;
    return plf::fenced_release(default_fenced) ? plf::kFlagFencedRelease : 0;
}

int launch_newview(plf_ctx *ctx, const float *x1, const float *x2, float *x3, unsigned char *scaler,
                   const float *ev, const float *pl, const float *pr, const int *wgt, size_t n,
                   unsigned long long *sum, const plf_launch_opts *opts, cudaStream_t stream)
{
    if (n == 0) return PLF_OK;
    if (!x1 || !x2 || !x3 || !ev || !pl || !pr)
        return fail(ctx, PLF_ERR_INVALID, "newview: NULL device pointer");
    if (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)x3 | (uintptr_t)ev | (uintptr_t)pl |
         (uintptr_t)pr) & 15u)
        return fail(ctx, PLF_ERR_INVALID, "newview: CLV / matrix pointers must be 16-byte aligned");
    KernelSel k;
    int grid = 0;
    int rc = resolve_launch(ctx, opts, n, &k, &grid);
    if (rc != PLF_OK) return rc;
    unsigned long long *work = nullptr;
    if (k.dynamic) {
        rc = work_pair(ctx, stream, &work);
        if (rc != PLF_OK) return rc;
    }
    // Programmatic dependent launch: the kernel's prologue may overlap the tail of the previous kernel
    // in the stream (it blocks in griddepcontrol.wait before touching global memory).  PLF_PDL=0 disables.
    static const bool use_pdl = [] {
        const char *e = getenv("PLF_PDL");
        return !(e && e[0] == '0');
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)k.threads);
    cfg.dynamicSmemBytes = k.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (use_pdl && !(opts && (opts->flags & PLF_LAUNCH_NO_PDL))) ? 1 : 0;
    const int flags = ((opts && opts->ev_per_category) ? plf::kFlagEvPerCategory : 0) | release_flag(opts, true);
    PLF_CUDA(ctx, cudaLaunchKernelEx(&cfg, k.fn, reinterpret_cast<const float4 *>(x1),
                                     reinterpret_cast<const float4 *>(x2), reinterpret_cast<float4 *>(x3), scaler,
                                     ev, pl, pr, wgt, n, sum, flags, work));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLF_CUDA(ctx, cudaGetLastError());
    return PLF_OK;
}

// Constant site patterns of the gen movers (values are data: mm2sleft_genDNAwindowComb.cpp:44-49,
// mm2sright_genDNAwindowComb.cpp:45-50) and the header each AIE lane derives from them: in gen
// mode the 2 EV beats and the 4 branch beats in front of every window carry the same 128-bit
// word as the site beats (mm2sleft_genDNAwindowComb.cpp:70-84), so lane j sees
//   EV_j rows 0,1 = left[4j..4j+3], rows 2,3 = right[4j..4j+3]   (combine.cpp:17-19)
//   P^T rows = left[4j..4j+3]  =>  P_left[j][k][l] = left[4j+k]   (mmul_branch.cpp:27-29)
const float kGenLeft[16] = {0.2135f, 0.1427f, 0.4139f, 0.8301f, 0.2021f, 0.9124f, 0.6542f, 0.1235f,
                            0.4856f, 0.2242f, 0.1322f, 0.5223f, 0.8223f, 0.7741f, 0.9855f, 0.2024f};
const float kGenRight[16] = {0.123456f, 0.234567f, 0.345678f, 0.789543f, 0.456789f, 0.567890f,
                             0.678901f, 0.789012f, 0.890123f, 0.901234f, 0.012345f, 0.023456f,
                             0.034567f, 0.045678f, 0.056789f, 0.067890f};

void build_gen_pattern(float *x1, float *x2, float *ev4, float *pl, float *pr)
{
    for (int e = 0; e < 16; ++e) {
        x1[e] = kGenLeft[e];
        x2[e] = kGenRight[e];
    }
    for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 4; ++l) {
                ev4[16 * j + 4 * k + l] = (k < 2) ? kGenLeft[4 * j + l] : kGenRight[4 * j + l];
                pl[16 * j + 4 * k + l] = kGenLeft[4 * j + k];
                pr[16 * j + 4 * k + l] = kGenRight[4 * j + k];
            }
}

template <class M, bool DISCARD>
int launch_gen_t(plf_ctx *ctx, float *x3, unsigned char *scaler, const float *d_gen, size_t n,
                 unsigned long long *sum, double *checksum, int bps_opt, cudaStream_t stream)
{
    constexpr int U = 2, THREADS = 256;
    auto fn = plf::plf_newview_gen<M, U, DISCARD, THREADS, 1>;
    int sms = 0;
    if (device_sms(&sms) != 0) return fail(ctx, PLF_ERR_CUDA, "no CUDA device");
    int bps = bps_opt;
    if (bps <= 0) {
        PLF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, THREADS, 0));
        if (bps <= 0) bps = 1;
    }
    const size_t tiles = (n + 8 * U - 1) / (8 * U);
    const size_t need = (tiles + THREADS / 32 - 1) / (THREADS / 32);
    size_t g = (size_t)sms * bps;
    if (g > need) g = need;
    if (g == 0) g = 1;
    fn<<<(int)g, THREADS, 0, stream>>>(reinterpret_cast<float4 *>(x3), scaler, d_gen, d_gen + 16,
                                       d_gen + 32, d_gen + 96, d_gen + 160, n, sum, checksum);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLF_CUDA(ctx, cudaGetLastError());
    return PLF_OK;
}

int launch_gen(plf_ctx *ctx, float *x3, unsigned char *scaler, const float *d_gen, size_t n,
               unsigned long long *sum, double *checksum, int sink, const plf_launch_opts *opts,
               cudaStream_t stream)
{
    if (n == 0) return PLF_OK;
    const int math = opts ? opts->math_mode : PLF_MATH_STRICT;
    const int bps = opts ? opts->blocks_per_sm : 0;
    if (sink == PLF_GEN_WRITE) {
        if (!x3 || ((uintptr_t)x3 & 15u))
            return fail(ctx, PLF_ERR_INVALID, "gen: x3 must be a 16-byte aligned device pointer");
        return math == PLF_MATH_FMA
                   ? launch_gen_t<plf::MathFma, false>(ctx, x3, scaler, d_gen, n, sum, checksum, bps, stream)
                   : launch_gen_t<plf::MathStrict, false>(ctx, x3, scaler, d_gen, n, sum, checksum, bps, stream);
    }
    if (sink == PLF_GEN_DISCARD)
        return math == PLF_MATH_FMA
                   ? launch_gen_t<plf::MathFma, true>(ctx, x3, scaler, d_gen, n, sum, checksum, bps, stream)
                   : launch_gen_t<plf::MathStrict, true>(ctx, x3, scaler, d_gen, n, sum, checksum, bps, stream);
    return fail(ctx, PLF_ERR_INVALID, "gen: unknown sink %d", sink);
}

// One device-resident copy of the gen pattern per device, for the ctx-less entry point.
std::mutex g_gen_mu;
float *g_gen_dev[64] = {nullptr};

int gen_pattern_device(plf_ctx *ctx, float **out)
{
    int dev = 0;
    PLF_CUDA(ctx, cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(ctx, PLF_ERR_INVALID, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> g(g_gen_mu);
    if (!g_gen_dev[dev]) {
        float h[224];
        build_gen_pattern(h, h + 16, h + 32, h + 96, h + 160);
        float *d = nullptr;
        PLF_CUDA(ctx, cudaMalloc(&d, sizeof h));
        PLF_CUDA(ctx, cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
        g_gen_dev[dev] = d;
    }
    *out = g_gen_dev[dev];
    return PLF_OK;
}

// 20-state newview on device-resident operands (matrices included): the protein kernel of plf_protein.cu.
// opts->variant / threads_per_block of a context are DNA tuning knobs and do not apply; math and release flags do.
int launch_states(plf_ctx *ctx, const float *x1, const float *x2, float *x3, unsigned char *scaler, const float *ev,
                  const float *pl, const float *pr, const int *wgt, size_t n, unsigned long long *sum,
                  const plf_launch_opts *opts, cudaStream_t stream)
{
    const int math = opts ? opts->math_mode : PLF_MATH_STRICT;
    int rc = plf::launch_newview_aa(x1, x2, x3, scaler, ev, pl, pr, wgt, n, sum, math, 0, 0, release_request(opts), stream);
    if (rc != PLF_OK) return fail(ctx, rc, "20-state newview launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PLF_OK;
}

struct Mats144 {
    float v[144];
};
__global__ void plf_store_mats(const __grid_constant__ Mats144 m, float *__restrict__ dst)
{
    if (threadIdx.x < 144) dst[threadIdx.x] = m.v[threadIdx.x];
}

struct Mats3600 {
    float v[3600];
};
__global__ void plf_store_mats_aa(const __grid_constant__ Mats3600 m, float *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3600) dst[i] = m.v[i];
}

int check_inst(plf_ctx *ctx, unsigned inst, bool need_alloc, Instance **out)
{
    if (!ctx) return fail(nullptr, PLF_ERR_INVALID, "NULL context");
    if (inst >= ctx->inst.size())
        return fail(ctx, PLF_ERR_INVALID, "instance %u out of range (context has %zu)", inst,
                    ctx->inst.size());
    Instance *I = &ctx->inst[inst];
    if (need_alloc && !I->allocated)
        return fail(ctx, PLF_ERR_STATE, "instance %u has no buffers (call plf_instance_alloc)", inst);
    *out = I;
    return PLF_OK;
}

// Sizes in floats for the context's state count S: one site 4S, EV S^2, one child's P 4S^2; the packed buffers are
// [EV | P | CLV] (left, and right in the Comb layout) or [P | CLV] (right, Sep) -- host_mem.cpp:231-241 with 4 -> S.
size_t site_floats(const plf_ctx *ctx) { return 4u * (size_t)ctx->states; }
size_t ev_floats(const plf_ctx *ctx) { return (size_t)ctx->states * ctx->states; }
size_t p_floats(const plf_ctx *ctx) { return 4u * (size_t)ctx->states * ctx->states; }
size_t left_header(const plf_ctx *ctx) { return ev_floats(ctx) + p_floats(ctx); }
size_t right_header(const plf_ctx *ctx)
{
    return ctx->layout == PLF_LAYOUT_COMB ? left_header(ctx) : p_floats(ctx);
}

void free_instance(Instance &I)
{
    cudaFree(I.d_left);
    cudaFree(I.d_right);
    cudaFree(I.d_out);
    cudaFree(I.d_scaler);
    cudaFree(I.d_wgt);
    I.d_left = I.d_right = I.d_out = nullptr;
    I.d_scaler = nullptr;
    I.d_wgt = nullptr;
    I.use_wgt = false;
    I.allocated = false;
    I.max_sites = 0;
}

}  // namespace

extern "C" {

const char *plf_last_error(const plf_ctx *ctx)
{
    if (ctx) {
        plf_ctx *c = const_cast<plf_ctx *>(ctx);
        std::lock_guard<std::mutex> g(c->err_mu);
        g_last_error = c->error;
    }
    return g_last_error.c_str();
}

unsigned long long plf_launch_count(void) { return g_launches.load(); }

}  // extern "C"

void plf::count_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// -2: follow the environment (PLF_SAFE_RELEASE); -1 / 0 / 1 set by plf_set_release_mode
static std::atomic<int> g_release_mode{-2};

bool plf::fenced_release(bool family_default)
{
    static const int env = [] {
        const char *e = getenv("PLF_SAFE_RELEASE");
        return !e || !e[0] ? -1 : (e[0] == '0' ? 0 : 1);
    }();
    int v = g_release_mode.load(std::memory_order_relaxed);
    if (v == -2) v = env;
    return v < 0 ? family_default : v == 1;
}

extern "C" int plf_set_release_mode(int mode)
{
    if (mode < -1 || mode > 1) return fail(nullptr, PLF_ERR_INVALID, "release mode must be -1 (defaults), 0 (dependency) or 1 (fenced)");
    g_release_mode.store(mode, std::memory_order_relaxed);
    return PLF_OK;
}

extern "C" {

int plf_device_count(int *count)
{
    if (!count) return fail(nullptr, PLF_ERR_INVALID, "NULL count");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(nullptr, PLF_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return PLF_OK;
}

int plf_device_info(int device, char *name, size_t name_len, char *bdf, size_t bdf_len)
{
    cudaDeviceProp prop;
    PLF_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (name && name_len) snprintf(name, name_len, "%s", prop.name);
    if (bdf && bdf_len) {
        char tmp[32];
        PLF_CUDA(nullptr, cudaDeviceGetPCIBusId(tmp, sizeof tmp, device));
        snprintf(bdf, bdf_len, "%s", tmp);
    }
    return PLF_OK;
}

int plf_device_from_string(const char *s, int *device)
{
    if (!s || !device) return fail(nullptr, PLF_ERR_INVALID, "NULL argument");
    if (strchr(s, ':')) {
        PLF_CUDA(nullptr, cudaDeviceGetByPCIBusId(device, s));
        return PLF_OK;
    }
    char *end = nullptr;
    long v = strtol(s, &end, 10);
    if (end == s || *end != '\0' || v < 0)
        return fail(nullptr, PLF_ERR_INVALID, "'%s' is neither a PCI BDF nor a device ordinal", s);
    int count = 0;
    int rc = plf_device_count(&count);
    if (rc != PLF_OK) return rc;
    if (v >= count) return fail(nullptr, PLF_ERR_INVALID, "device %ld not present (%d devices)", v, count);
    *device = (int)v;
    return PLF_OK;
}

int plf_ctx_create(plf_ctx **out, int device, unsigned n_instances, int layout, int input_src)
{
    return plf_ctx_create_states(out, device, n_instances, layout, input_src, 4);
}

int plf_ctx_states(const plf_ctx *ctx) { return ctx ? ctx->states : 0; }

int plf_ctx_create_states(plf_ctx **out, int device, unsigned n_instances, int layout, int input_src, int states)
{
    if (!out) return fail(nullptr, PLF_ERR_INVALID, "NULL ctx out-pointer");
    *out = nullptr;
    if (states != 4 && states != 20)
        return fail(nullptr, PLF_ERR_INVALID, "STATES=%d is not supported (4 = DNA, 20 = protein)", states);
    if (states != 4 && input_src == PLF_INPUT_GEN)
        return fail(nullptr, PLF_ERR_INVALID, "INPUT_SRC=gen is defined for STATES=DNA only (the gen movers emit a 16-float site pattern)");
    if (n_instances == 0 || n_instances > 1024)
        return fail(nullptr, PLF_ERR_INVALID, "n_instances must be in 1..1024 (got %u)", n_instances);
    if (layout != PLF_LAYOUT_COMB && layout != PLF_LAYOUT_SEP)
        return fail(nullptr, PLF_ERR_INVALID, "unknown layout %d", layout);
    if (input_src != PLF_INPUT_MEM && input_src != PLF_INPUT_GEN)
        return fail(nullptr, PLF_ERR_INVALID, "unknown input source %d", input_src);
    int count = 0;
    int rc = plf_device_count(&count);
    if (rc != PLF_OK) return rc;
    if (count == 0) return fail(nullptr, PLF_ERR_CUDA, "no CUDA device present (no CPU fallback)");
    if (device < 0 || device >= count)
        return fail(nullptr, PLF_ERR_INVALID, "device %d not present (%d devices)", device, count);
    PLF_CUDA(nullptr, cudaSetDevice(device));
    plf_ctx *ctx = new (std::nothrow) plf_ctx;
    if (!ctx) return fail(nullptr, PLF_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->states = states;
    ctx->layout = layout;
    ctx->input_src = input_src;
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    ctx->inst.resize(n_instances);
    for (auto &I : ctx->inst) {
        cudaError_t e = cudaStreamCreateWithFlags(&I.stream, cudaStreamNonBlocking);
        for (int m = 0; m < 4 && e == cudaSuccess; ++m) e = cudaEventCreate(&I.marks[m]);
        if (e == cudaSuccess) e = cudaMalloc(&I.d_sum, kSumSlots * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(I.d_sum, 0, kSumSlots * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&I.d_check, sizeof(double));
        if (e == cudaSuccess) e = cudaMallocHost(&I.h_sum, sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMallocHost(&I.h_check, sizeof(double));
        if (e != cudaSuccess) {
            fail(nullptr, PLF_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
            plf_ctx_destroy(ctx);
            return PLF_ERR_CUDA;
        }
        *I.h_sum = 0;
        *I.h_check = 0.0;
    }
    {
        KernelSel k = pick_kernel(PLF_MATH_STRICT, kDefaultVariant, kDefaultThreads);
        cudaFuncAttributes attr;
        if (k.fn && cudaFuncGetAttributes(&attr, k.fn) == cudaSuccess) prepare_kernel(nullptr, k);
        cudaGetLastError();
    }
    if (input_src == PLF_INPUT_GEN) {
        rc = gen_pattern_device(nullptr, &ctx->d_gen);
        if (rc != PLF_OK) {
            plf_ctx_destroy(ctx);
            return rc;
        }
    }
    *out = ctx;
    return PLF_OK;
}

int plf_ctx_destroy(plf_ctx *ctx)
{
    if (!ctx) return PLF_OK;
    cudaSetDevice(ctx->device);
    for (auto &I : ctx->inst) {
        if (I.stream) cudaStreamSynchronize(I.stream);
        free_instance(I);
        cudaFree(I.d_sum);
        cudaFree(I.d_check);
        if (I.h_sum) cudaFreeHost(I.h_sum);
        if (I.h_check) cudaFreeHost(I.h_check);
        for (auto &m : I.marks)
            if (m) cudaEventDestroy(m);
        if (I.stream) cudaStreamDestroy(I.stream);
    }
    for (int k = 0; k < plf_ctx::kStreamSlots; ++k) {
        if (ctx->s_stream[k]) cudaStreamSynchronize(ctx->s_stream[k]);
        cudaFree(ctx->sd_x1[k]);
        cudaFree(ctx->sd_x2[k]);
        cudaFree(ctx->sd_x3[k]);
        cudaFree(ctx->sd_sc[k]);
        cudaFree(ctx->sd_wgt[k]);
        if (ctx->s_stream[k]) cudaStreamDestroy(ctx->s_stream[k]);
    }
    cudaFree(ctx->sd_mats);
    cudaFree(ctx->sd_sum);
    if (ctx->sh_sum) cudaFreeHost(ctx->sh_sum);
    delete ctx;
    return PLF_OK;
}

int plf_ctx_set_math(plf_ctx *ctx, int math_mode)
{
    if (!ctx) return fail(nullptr, PLF_ERR_INVALID, "NULL context");
    if (math_mode != PLF_MATH_STRICT && math_mode != PLF_MATH_FMA)
        return fail(ctx, PLF_ERR_INVALID, "unknown math mode %d", math_mode);
    ctx->math = math_mode;
    return PLF_OK;
}

int plf_ctx_set_gen_sink(plf_ctx *ctx, int sink)
{
    if (!ctx) return fail(nullptr, PLF_ERR_INVALID, "NULL context");
    if (sink != PLF_GEN_WRITE && sink != PLF_GEN_DISCARD)
        return fail(ctx, PLF_ERR_INVALID, "unknown gen sink %d", sink);
    ctx->gen_sink = sink;
    return PLF_OK;
}

int plf_ctx_set_tuning(plf_ctx *ctx, int variant, int threads_per_block, int blocks_per_sm)
{
    if (!ctx) return fail(nullptr, PLF_ERR_INVALID, "NULL context");
    const int v = variant ? variant : kDefaultVariant;
    const int t = threads_per_block ? threads_per_block : kDefaultThreads;
    if (!pick_kernel(PLF_MATH_STRICT, v, t).fn)
        return fail(ctx, PLF_ERR_INVALID, "unknown kernel variant %d / threads %d", variant,
                    threads_per_block);
    if (blocks_per_sm < 0 || blocks_per_sm > 32)
        return fail(ctx, PLF_ERR_INVALID, "blocks_per_sm %d out of range", blocks_per_sm);
    ctx->variant = variant;
    ctx->threads = threads_per_block;
    ctx->blocks_per_sm = blocks_per_sm;
    return PLF_OK;
}

unsigned plf_ctx_instances(const plf_ctx *ctx) { return ctx ? (unsigned)ctx->inst.size() : 0u; }

int plf_instance_alloc(plf_ctx *ctx, unsigned inst, size_t max_sites)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    if (max_sites == 0) return fail(ctx, PLF_ERR_INVALID, "max_sites must be > 0");
    if (max_sites > (SIZE_MAX / (site_floats(ctx) * 4)) - 64) return fail(ctx, PLF_ERR_INVALID, "max_sites too large");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (I->allocated) {
        PLF_CUDA(ctx, cudaStreamSynchronize(I->stream));
        free_instance(*I);
    }
    const size_t clv = max_sites * site_floats(ctx) * sizeof(float);
    cudaError_t e = cudaSuccess;
    if (ctx->input_src == PLF_INPUT_MEM) {
        e = cudaMalloc(&I->d_left, left_header(ctx) * sizeof(float) + clv);
        if (e == cudaSuccess) e = cudaMalloc(&I->d_right, right_header(ctx) * sizeof(float) + clv);
    }
    if (e == cudaSuccess) e = cudaMalloc(&I->d_out, clv);
    if (e == cudaSuccess) e = cudaMalloc(&I->d_scaler, max_sites);
    if (e != cudaSuccess) {
        free_instance(*I);
        cudaGetLastError();
        return fail(ctx, e == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA,
                    "instance %u: device allocation for %zu sites failed: %s", inst, max_sites,
                    cudaGetErrorString(e));
    }
    I->max_sites = max_sites;
    I->allocated = true;
    return PLF_OK;
}

int plf_instance_free(plf_ctx *ctx, unsigned inst)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaStreamSynchronize(I->stream));
    free_instance(*I);
    return PLF_OK;
}

static int write_packed(plf_ctx *ctx, unsigned inst, bool left, const float *src, size_t bytes,
                        size_t offset)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, true, &I);
    if (rc != PLF_OK) return rc;
    if (ctx->input_src != PLF_INPUT_MEM)
        return fail(ctx, PLF_ERR_STATE, "INPUT_SRC=gen instances have no input buffers");
    if (bytes == 0) return PLF_OK;
    if (!src) return fail(ctx, PLF_ERR_INVALID, "NULL host buffer");
    const size_t header = left ? left_header(ctx) : right_header(ctx);
    const size_t cap = (header + I->max_sites * site_floats(ctx)) * sizeof(float);
    if (offset > cap || bytes > cap - offset)
        return fail(ctx, PLF_ERR_INVALID, "write of %zu bytes at offset %zu exceeds the %s buffer (%zu bytes)",
                    bytes, offset, left ? "left" : "right", cap);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    char *dst = reinterpret_cast<char *>(left ? I->d_left : I->d_right) + offset;
    PLF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, I->stream));
    return PLF_OK;
}

int plf_write_left(plf_ctx *ctx, unsigned inst, const float *packed, size_t bytes, size_t offset)
{
    return write_packed(ctx, inst, true, packed, bytes, offset);
}

int plf_write_right(plf_ctx *ctx, unsigned inst, const float *packed, size_t bytes, size_t offset)
{
    return write_packed(ctx, inst, false, packed, bytes, offset);
}

int plf_write_wgt(plf_ctx *ctx, unsigned inst, const int *wgt, size_t count)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, true, &I);
    if (rc != PLF_OK) return rc;
    if (!wgt) {
        I->use_wgt = false;
        return PLF_OK;
    }
    if (count > I->max_sites)
        return fail(ctx, PLF_ERR_INVALID, "%zu weights exceed the instance capacity of %zu sites", count,
                    I->max_sites);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!I->d_wgt) PLF_CUDA(ctx, cudaMalloc(&I->d_wgt, I->max_sites * sizeof(int)));
    PLF_CUDA(ctx, cudaMemcpyAsync(I->d_wgt, wgt, count * sizeof(int), cudaMemcpyHostToDevice, I->stream));
    I->use_wgt = true;
    return PLF_OK;
}

int plf_run_async(plf_ctx *ctx, unsigned inst, size_t sites)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, true, &I);
    if (rc != PLF_OK) return rc;
    if (sites > I->max_sites)
        return fail(ctx, PLF_ERR_INVALID, "run of %zu sites exceeds the instance capacity of %zu", sites,
                    I->max_sites);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    // every run accumulates into its own zeroed counter; the ring is re-zeroed once per kSumSlots runs
    const unsigned slot = (unsigned)(I->runs % kSumSlots);
    if (slot == 0 && I->runs != 0)
        PLF_CUDA(ctx, cudaMemsetAsync(I->d_sum, 0, kSumSlots * sizeof(unsigned long long), I->stream));
    unsigned long long *d_sum = I->d_sum + slot;
    plf_launch_opts opts;
    opts.math_mode = ctx->math;
    opts.variant = ctx->variant;
    opts.threads_per_block = ctx->threads;
    opts.blocks_per_sm = ctx->blocks_per_sm;
    opts.ev_per_category = 0;
    opts.flags = 0;
    if (ctx->input_src == PLF_INPUT_MEM) {
        const float *ev = I->d_left;                                   // mem[0]
        const float *pl = I->d_left + ev_floats(ctx);                  // mem[1..4]
        const float *x1 = I->d_left + left_header(ctx);                // mem[5+i]
        const float *pr = ctx->layout == PLF_LAYOUT_COMB ? I->d_right + ev_floats(ctx) : I->d_right;
        const float *x2 = I->d_right + right_header(ctx);
        if (ctx->states == 4)
            rc = launch_newview(ctx, x1, x2, I->d_out, I->d_scaler, ev, pl, pr,
                                I->use_wgt ? I->d_wgt : nullptr, sites, d_sum, &opts, I->stream);
        else
            rc = launch_states(ctx, x1, x2, I->d_out, I->d_scaler, ev, pl, pr, I->use_wgt ? I->d_wgt : nullptr, sites, d_sum,
                               &opts, I->stream);
    } else {
        PLF_CUDA(ctx, cudaMemsetAsync(I->d_check, 0, sizeof(double), I->stream));
        rc = launch_gen(ctx, I->d_out, I->d_scaler, ctx->d_gen, sites, d_sum, I->d_check,
                        ctx->gen_sink, &opts, I->stream);
    }
    if (rc != PLF_OK) return rc;
    I->sum_slot = slot;
    ++I->runs;
    return PLF_OK;
}

int plf_wait(plf_ctx *ctx, unsigned inst)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaStreamSynchronize(I->stream));
    return PLF_OK;
}

int plf_read_out(plf_ctx *ctx, unsigned inst, float *dst, size_t bytes, size_t offset)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, true, &I);
    if (rc != PLF_OK) return rc;
    if (bytes == 0) return PLF_OK;
    if (!dst) return fail(ctx, PLF_ERR_INVALID, "NULL host buffer");
    const size_t cap = I->max_sites * site_floats(ctx) * sizeof(float);
    if (offset > cap || bytes > cap - offset)
        return fail(ctx, PLF_ERR_INVALID, "read of %zu bytes at offset %zu exceeds the out buffer (%zu bytes)",
                    bytes, offset, cap);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMemcpyAsync(dst, reinterpret_cast<char *>(I->d_out) + offset, bytes,
                                  cudaMemcpyDeviceToHost, I->stream));
    return PLF_OK;
}

int plf_read_scaler(plf_ctx *ctx, unsigned inst, char *dst, size_t bytes, size_t offset)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, true, &I);
    if (rc != PLF_OK) return rc;
    if (bytes == 0) return PLF_OK;
    if (!dst) return fail(ctx, PLF_ERR_INVALID, "NULL host buffer");
    if (offset > I->max_sites || bytes > I->max_sites - offset)
        return fail(ctx, PLF_ERR_INVALID, "read of %zu bytes at offset %zu exceeds the scaler buffer (%zu bytes)",
                    bytes, offset, I->max_sites);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMemcpyAsync(dst, I->d_scaler + offset, bytes, cudaMemcpyDeviceToHost, I->stream));
    return PLF_OK;
}

int plf_scaler_increment(plf_ctx *ctx, unsigned inst, long long *increment)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    if (!increment) return fail(ctx, PLF_ERR_INVALID, "NULL increment");
    if (I->runs == 0) {
        *increment = 0;
        return PLF_OK;
    }
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMemcpyAsync(I->h_sum, I->d_sum + I->sum_slot, sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, I->stream));
    rc = plf_wait(ctx, inst);
    if (rc != PLF_OK) return rc;
    *increment = (long long)*I->h_sum;
    return PLF_OK;
}

int plf_gen_checksum(plf_ctx *ctx, unsigned inst, double *checksum)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    if (!checksum) return fail(ctx, PLF_ERR_INVALID, "NULL checksum");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMemcpyAsync(I->h_check, I->d_check, sizeof(double), cudaMemcpyDeviceToHost, I->stream));
    rc = plf_wait(ctx, inst);
    if (rc != PLF_OK) return rc;
    *checksum = *I->h_check;
    return PLF_OK;
}

int plf_mark(plf_ctx *ctx, unsigned inst, int mark_id)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    if (mark_id < 0 || mark_id > 3) return fail(ctx, PLF_ERR_INVALID, "mark id %d out of range", mark_id);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaEventRecord(I->marks[mark_id], I->stream));
    I->mark_set[mark_id] = true;
    return PLF_OK;
}

int plf_elapsed_ms(plf_ctx *ctx, unsigned inst, int from_mark, int to_mark, float *ms)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    if (from_mark < 0 || from_mark > 3 || to_mark < 0 || to_mark > 3 || !ms)
        return fail(ctx, PLF_ERR_INVALID, "bad mark ids %d..%d", from_mark, to_mark);
    if (!I->mark_set[from_mark] || !I->mark_set[to_mark])
        return fail(ctx, PLF_ERR_STATE, "mark %d or %d was never recorded", from_mark, to_mark);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaEventSynchronize(I->marks[to_mark]));
    PLF_CUDA(ctx, cudaEventElapsedTime(ms, I->marks[from_mark], I->marks[to_mark]));
    return PLF_OK;
}

int plf_instance_device_ptrs(plf_ctx *ctx, unsigned inst, float **left, float **right, float **out,
                             unsigned char **scaler)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, true, &I);
    if (rc != PLF_OK) return rc;
    if (left) *left = I->d_left;
    if (right) *right = I->d_right;
    if (out) *out = I->d_out;
    if (scaler) *scaler = I->d_scaler;
    return PLF_OK;
}

int plf_instance_stream(plf_ctx *ctx, unsigned inst, void **stream)
{
    Instance *I = nullptr;
    int rc = check_inst(ctx, inst, false, &I);
    if (rc != PLF_OK) return rc;
    if (!stream) return fail(ctx, PLF_ERR_INVALID, "NULL stream out-pointer");
    *stream = I->stream;
    return PLF_OK;
}

int plf_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(nullptr, PLF_ERR_INVALID, "NULL out-pointer");
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, PLF_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return PLF_OK;
}

int plf_host_free(void *ptr)
{
    if (!ptr) return PLF_OK;
    PLF_CUDA(nullptr, cudaFreeHost(ptr));
    return PLF_OK;
}

int plf_host_register(void *ptr, size_t bytes)
{
    if (!ptr || !bytes) return fail(nullptr, PLF_ERR_INVALID, "NULL or empty range");
    PLF_CUDA(nullptr, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return PLF_OK;
}

int plf_host_unregister(void *ptr)
{
    if (!ptr) return PLF_OK;
    PLF_CUDA(nullptr, cudaHostUnregister(ptr));
    return PLF_OK;
}

int plf_newview_device(const float *x1, const float *x2, float *x3, unsigned char *scaler,
                       const float *ev, const float *p_left, const float *p_right, const int *wgt,
                       size_t n, unsigned long long *scaler_sum, const plf_launch_opts *opts,
                       void *stream)
{
    return launch_newview(nullptr, x1, x2, x3, scaler, ev, p_left, p_right, wgt, n, scaler_sum, opts,
                          static_cast<cudaStream_t>(stream));
}

int plf_newview_gen_device(float *x3, unsigned char *scaler, size_t n, unsigned long long *scaler_sum,
                           double *checksum, int sink, const plf_launch_opts *opts, void *stream)
{
    float *d_gen = nullptr;
    int rc = gen_pattern_device(nullptr, &d_gen);
    if (rc != PLF_OK) return rc;
    return launch_gen(nullptr, x3, scaler, d_gen, n, scaler_sum, checksum, sink, opts,
                      static_cast<cudaStream_t>(stream));
}

int plf_gen_pattern(float *x1, float *x2, float *ev4, float *p_left, float *p_right)
{
    if (!x1 || !x2 || !ev4 || !p_left || !p_right) return fail(nullptr, PLF_ERR_INVALID, "NULL argument");
    build_gen_pattern(x1, x2, ev4, p_left, p_right);
    return PLF_OK;
}

int plf_generate_device(float *x1, float *x2, size_t first_site, size_t n, uint64_t seed, void *stream)
{
    if (n == 0) return PLF_OK;
    if (!x1 || !x2 || (((uintptr_t)x1 | (uintptr_t)x2) & 15u))
        return fail(nullptr, PLF_ERR_INVALID, "generate: x1/x2 must be 16-byte aligned device pointers");
    int sms = 0;
    if (device_sms(&sms) != 0) return fail(nullptr, PLF_ERR_CUDA, "no CUDA device");
    const size_t n_vec4 = n * 4;
    size_t grid = (n_vec4 + 255) / 256;
    if (grid > (size_t)sms * 8) grid = (size_t)sms * 8;
    plf::plf_generate_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4 *>(x1), reinterpret_cast<float4 *>(x2), (uint64_t)first_site * 16u,
        n_vec4, seed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLF_CUDA(nullptr, cudaGetLastError());
    return PLF_OK;
}

int plf_generate_host(float *x1, float *x2, size_t first_site, size_t n, uint64_t seed)
{
    if (!x1 || !x2) return fail(nullptr, PLF_ERR_INVALID, "NULL argument");
    for (size_t e = 0; e < n * 16; ++e)
        plf::gen_pair(seed, (uint64_t)first_site * 16u + e, x1[e], x2[e]);
    return PLF_OK;
}

// Streamed host path: the whole round trip of one newview over host-resident,
// UNPACKED arrays, cut into chunks that flow through kStreamSlots device buffers on as many streams, so
// that the H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap.  It is the
// analogue of the reference's NO_INTERMEDIATE_RESULTS=1 round-trip mode (host_mem.cpp:327-382) without
// the host-side packing pass, and it lifts the device-memory limit (the reference's sweep list goes to
// 1e9 sites = 192 GB of CLVs, Makefile:16).
int plf_newview_stream(plf_ctx *ctx, const float *ev, const float *p_left, const float *p_right, const float *x1,
                       const float *x2, float *x3, char *scaler, const int *wgt, size_t n_sites, size_t chunk_sites,
                       long long *increment)
{
    if (!ctx) return fail(nullptr, PLF_ERR_INVALID, "NULL context");
    if (!ev || !p_left || !p_right) return fail(ctx, PLF_ERR_INVALID, "stream: NULL matrix");
    if (n_sites && (!x1 || !x2 || !x3)) return fail(ctx, PLF_ERR_INVALID, "stream: NULL host CLV");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (chunk_sites == 0) {
        // auto: about 16 chunks per call, between 256 Ki sites (16 MiB per CLV copy: still far above the per-copy
        // overhead) and 2 Mi sites (128 MiB).  The pipeline drains for one chunk's D2H at the end of a call, so a
        // 7 Mi-site call cut into four 2 Mi chunks ran at 2/3 of the PCIe rate; sixteen chunks lose 6 %.
        const size_t scale = site_floats(ctx) / 16;                   // the limits are byte sizes: 20-state sites are 5x larger
        chunk_sites = (n_sites + 15) / 16;
        if (chunk_sites < ((size_t)256 << 10) / scale) chunk_sites = ((size_t)256 << 10) / scale;
        if (chunk_sites > ((size_t)2 << 20) / scale) chunk_sites = ((size_t)2 << 20) / scale;
    }
    if (chunk_sites > n_sites && n_sites > 0) chunk_sites = n_sites;
    chunk_sites = (chunk_sites + 255) & ~(size_t)255;
    constexpr int K = plf_ctx::kStreamSlots;
    const size_t sf = site_floats(ctx), evf = ev_floats(ctx), pf = p_floats(ctx);
    if (!ctx->sd_mats) {
        PLF_CUDA(ctx, cudaMalloc(&ctx->sd_mats, (evf + 2 * pf) * sizeof(float)));
        PLF_CUDA(ctx, cudaMalloc(&ctx->sd_sum, sizeof(unsigned long long)));
        PLF_CUDA(ctx, cudaMallocHost(&ctx->sh_sum, sizeof(unsigned long long)));
        for (int k = 0; k < K; ++k) PLF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_stream[k], cudaStreamNonBlocking));
    }
    if (ctx->stream_chunk < chunk_sites) {                          // (re)size the chunk buffers
        for (int k = 0; k < K; ++k) {
            PLF_CUDA(ctx, cudaStreamSynchronize(ctx->s_stream[k]));
            cudaFree(ctx->sd_x1[k]);
            cudaFree(ctx->sd_x2[k]);
            cudaFree(ctx->sd_x3[k]);
            cudaFree(ctx->sd_sc[k]);
            cudaFree(ctx->sd_wgt[k]);
            ctx->sd_x1[k] = ctx->sd_x2[k] = ctx->sd_x3[k] = nullptr;
            ctx->sd_sc[k] = nullptr;
            ctx->sd_wgt[k] = nullptr;
        }
        ctx->stream_chunk = 0;
        const size_t clv = chunk_sites * sf * sizeof(float);
        for (int k = 0; k < K; ++k) {
            PLF_CUDA(ctx, cudaMalloc(&ctx->sd_x1[k], clv));
            PLF_CUDA(ctx, cudaMalloc(&ctx->sd_x2[k], clv));
            PLF_CUDA(ctx, cudaMalloc(&ctx->sd_x3[k], clv));
            PLF_CUDA(ctx, cudaMalloc(&ctx->sd_sc[k], chunk_sites));
            PLF_CUDA(ctx, cudaMalloc(&ctx->sd_wgt[k], chunk_sites * sizeof(int)));
        }
        ctx->stream_chunk = chunk_sites;
    }
    cudaStream_t s0 = ctx->s_stream[0];
    std::vector<float> mats(evf + 2 * pf);
    memcpy(mats.data(), ev, evf * sizeof(float));
    memcpy(mats.data() + evf, p_left, pf * sizeof(float));
    memcpy(mats.data() + evf + pf, p_right, pf * sizeof(float));
    PLF_CUDA(ctx, cudaMemcpyAsync(ctx->sd_mats, mats.data(), mats.size() * sizeof(float), cudaMemcpyHostToDevice, s0));
    PLF_CUDA(ctx, cudaMemsetAsync(ctx->sd_sum, 0, sizeof(unsigned long long), s0));
    PLF_CUDA(ctx, cudaStreamSynchronize(s0));                        // `mats` is a stack temporary
    plf_launch_opts opts;
    opts.math_mode = ctx->math;
    opts.variant = ctx->variant;
    opts.threads_per_block = ctx->threads;
    opts.blocks_per_sm = ctx->blocks_per_sm;
    opts.ev_per_category = 0;
    opts.flags = 0;
    // A failure in the middle of the pipeline must not return while copies to or from the caller's host arrays are
    // still in flight on the other streams: drain all of them first (secondary errors of the drain are ignored).
    std::vector<cudaEvent_t> tev;                                    // PLF_STREAM_TRACE events (below)
    auto drain = [&](int rc) {
        for (int b = 0; b < K; ++b) cudaStreamSynchronize(ctx->s_stream[b]);
        for (cudaEvent_t e : tev)
            if (e) cudaEventDestroy(e);
        tev.clear();
        cudaGetLastError();
        return rc;
    };
#define PLF_CUDA_DRAIN(expr)                                                                                         \
    do {                                                                                                             \
        cudaError_t e__ = (expr);                                                                                    \
        if (e__ != cudaSuccess) return drain(fail(ctx, PLF_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__))); \
    } while (0)
    // PLF_STREAM_TRACE=<file> (debug; there is no nsys in this image): four events per chunk on its stream -- before the
    // H2D copies, after them, after the kernel, after the D2H copies -- written after the call as one line per chunk in
    // ms since the first event.  tools/stream_timeline.py turns that into the three-way overlap of the pipeline.
    const char *trace_path = getenv("PLF_STREAM_TRACE");
    auto mark = [&](cudaStream_t st) {
        if (!trace_path) return;
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) == cudaSuccess) cudaEventRecord(e, st);
        tev.push_back(e);
    };
    size_t k = 0;
    for (size_t lo = 0; lo < n_sites; lo += chunk_sites, ++k) {
        const int b = (int)(k % K);
        cudaStream_t st = ctx->s_stream[b];
        const size_t cnt = n_sites - lo < chunk_sites ? n_sites - lo : chunk_sites;
        const size_t bytes = cnt * sf * sizeof(float);
        mark(st);
        PLF_CUDA_DRAIN(cudaMemcpyAsync(ctx->sd_x1[b], x1 + lo * sf, bytes, cudaMemcpyHostToDevice, st));
        PLF_CUDA_DRAIN(cudaMemcpyAsync(ctx->sd_x2[b], x2 + lo * sf, bytes, cudaMemcpyHostToDevice, st));
        if (wgt) PLF_CUDA_DRAIN(cudaMemcpyAsync(ctx->sd_wgt[b], wgt + lo, cnt * sizeof(int), cudaMemcpyHostToDevice, st));
        mark(st);
        int rc = ctx->states == 4
                     ? launch_newview(ctx, ctx->sd_x1[b], ctx->sd_x2[b], ctx->sd_x3[b], ctx->sd_sc[b], ctx->sd_mats,
                                      ctx->sd_mats + evf, ctx->sd_mats + evf + pf, wgt ? ctx->sd_wgt[b] : nullptr, cnt,
                                      ctx->sd_sum, &opts, st)
                     : launch_states(ctx, ctx->sd_x1[b], ctx->sd_x2[b], ctx->sd_x3[b], ctx->sd_sc[b], ctx->sd_mats,
                                     ctx->sd_mats + evf, ctx->sd_mats + evf + pf, wgt ? ctx->sd_wgt[b] : nullptr, cnt,
                                     ctx->sd_sum, &opts, st);
        if (rc != PLF_OK) return drain(rc);
        mark(st);
        PLF_CUDA_DRAIN(cudaMemcpyAsync(x3 + lo * sf, ctx->sd_x3[b], bytes, cudaMemcpyDeviceToHost, st));
        if (scaler) PLF_CUDA_DRAIN(cudaMemcpyAsync(scaler + lo, ctx->sd_sc[b], cnt, cudaMemcpyDeviceToHost, st));
        mark(st);
    }
#undef PLF_CUDA_DRAIN
    for (int b = 0; b < K; ++b) PLF_CUDA(ctx, cudaStreamSynchronize(ctx->s_stream[b]));
    if (trace_path && !tev.empty()) {
        if (FILE *f = fopen(trace_path, "w")) {
            fprintf(f, "# chunk slot sites  h2d_begin h2d_end kernel_end d2h_end   (ms since the first event; chunk = %zu sites, %zu B/site in, %zu B/site out)\n",
                    chunk_sites, 2 * sf * sizeof(float), sf * sizeof(float) + 1);
            for (size_t c = 0; c * 4 + 3 < tev.size(); ++c) {
                float t[4] = {0, 0, 0, 0};
                for (int j = 0; j < 4; ++j)
                    if (tev[0] && tev[c * 4 + j]) cudaEventElapsedTime(&t[j], tev[0], tev[c * 4 + j]);
                const size_t lo = c * chunk_sites;
                fprintf(f, "%zu %zu %zu  %.4f %.4f %.4f %.4f\n", c, c % K, n_sites - lo < chunk_sites ? n_sites - lo : chunk_sites, t[0], t[1], t[2], t[3]);
            }
            fclose(f);
        }
        for (cudaEvent_t e : tev)
            if (e) cudaEventDestroy(e);
    }
    PLF_CUDA(ctx, cudaMemcpy(ctx->sh_sum, ctx->sd_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (increment) *increment = (long long)*ctx->sh_sum;
    return PLF_OK;
}

int plf_evaluate_device(const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                        const float *diag, size_t n, double *lnl, void *stream)
{
    return plf_evaluate_states_device(4, x1, x2, cnt1, cnt2, wgt, diag, n, lnl, stream);
}

int plf_evaluate_states_device(int states, const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                               const float *diag, size_t n, double *lnl, void *stream)
{
    if (states != 4 && states != 20)
        return fail(nullptr, PLF_ERR_INVALID, "STATES=%d is not supported (4 = DNA, 20 = protein)", states);
    if (n == 0) return PLF_OK;
    if (!x1 || !x2 || !diag || !lnl) return fail(nullptr, PLF_ERR_INVALID, "evaluate: NULL device pointer");
    if (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)diag) & 15u)
        return fail(nullptr, PLF_ERR_INVALID, "evaluate: CLV / diag pointers must be 16-byte aligned");
    int rc = plf::launch_evaluate(states, x1, x2, cnt1, cnt2, wgt, diag, n, lnl, static_cast<cudaStream_t>(stream));
    if (rc != PLF_OK) return fail(nullptr, rc, "evaluate kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PLF_OK;
}

int plf_newview_states_device(int states, const float *x1, const float *x2, float *x3, unsigned char *scaler,
                              const float *ev, const float *p_left, const float *p_right, const int *wgt, size_t n,
                              unsigned long long *scaler_sum, const plf_launch_opts *opts, void *stream)
{
    if (states != 4 && states != 20)
        return fail(nullptr, PLF_ERR_INVALID, "STATES=%d is not supported (4 = DNA, 20 = protein)", states);
    if (n == 0) return PLF_OK;
    if (!x1 || !x2 || !x3 || !ev || !p_left || !p_right) return fail(nullptr, PLF_ERR_INVALID, "newview: NULL pointer");
    if (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)x3) & 15u)
        return fail(nullptr, PLF_ERR_INVALID, "newview: CLV pointers must be 16-byte aligned");
    const int math = opts ? opts->math_mode : PLF_MATH_STRICT;
    if (math != PLF_MATH_STRICT && math != PLF_MATH_FMA) return fail(nullptr, PLF_ERR_INVALID, "unknown math mode %d", math);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (states == 4) {
        // The matrices are HOST arrays in this entry point.  They travel BY VALUE as the argument of a one-block
        // upload kernel that writes them into the stream's scratch record (so they are consumed before the call
        // returns, pinned or not, and the call can be captured into a CUDA graph: no allocation, no host copy),
        // and the DNA kernel reads them from there.  Stream order keeps consecutive calls on one stream apart.
        plf::StreamScratch *sc = nullptr;
        int rc = plf::stream_scratch(st, &sc);
        if (rc != PLF_OK) return fail(nullptr, rc, "per-stream scratch allocation failed");
        Mats144 m;
        memcpy(m.v, ev, 16 * sizeof(float));
        memcpy(m.v + 16, p_left, 64 * sizeof(float));
        memcpy(m.v + 80, p_right, 64 * sizeof(float));
        plf_store_mats<<<1, 160, 0, st>>>(m, sc->mats);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        PLF_CUDA(nullptr, cudaGetLastError());
        return launch_newview(nullptr, x1, x2, x3, scaler, sc->mats, sc->mats + 16, sc->mats + 80, wgt, n, scaler_sum, opts, st);
    }
    if (opts && opts->ev_per_category) return fail(nullptr, PLF_ERR_INVALID, "ev_per_category is a DNA gen-mode option");
    const int aa_flags = release_request(opts) | ((opts && (opts->flags & PLF_LAUNCH_SINGLE_CTA)) ? plf::kAaSingleCta : 0);
    // HOST matrices: by value into the stream's staging record (as for S = 4), then the kernel reads device memory
    plf::StreamScratch *sc = nullptr;
    int rc = plf::stream_scratch(st, &sc);
    if (rc != PLF_OK) return fail(nullptr, rc, "per-stream scratch allocation failed");
    {
        Mats3600 m;
        memcpy(m.v, ev, 400 * sizeof(float));
        memcpy(m.v + 400, p_left, 1600 * sizeof(float));
        memcpy(m.v + 2000, p_right, 1600 * sizeof(float));
        plf_store_mats_aa<<<4, 1024, 0, st>>>(m, sc->mats);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        PLF_CUDA(nullptr, cudaGetLastError());
    }
    rc = plf::launch_newview_aa(x1, x2, x3, scaler, sc->mats, sc->mats + 400, sc->mats + 2000, wgt, n, scaler_sum, math,
                                opts ? opts->variant : 0, opts ? opts->threads_per_block : 0, aa_flags, st);
    if (rc == PLF_ERR_INVALID)
        return fail(nullptr, rc, "no 20-state kernel for variant %d / threads %d", opts ? opts->variant : 0,
                    opts ? opts->threads_per_block : 0);
    if (rc != PLF_OK) return fail(nullptr, rc, "20-state newview launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PLF_OK;
}

int plf_generate_states_device(int states, float *x1, float *x2, size_t first_site, size_t n, uint64_t seed, void *stream)
{
    if (states != 4 && states != 20) return fail(nullptr, PLF_ERR_INVALID, "STATES=%d is not supported (4 = DNA, 20 = protein)", states);
    if (n == 0) return PLF_OK;
    if (!x1 || !x2 || (((uintptr_t)x1 | (uintptr_t)x2) & 15u))
        return fail(nullptr, PLF_ERR_INVALID, "generate: x1/x2 must be 16-byte aligned device pointers");
    int rc = plf::launch_generate_states(states, x1, x2, first_site, n, seed, static_cast<cudaStream_t>(stream));
    if (rc != PLF_OK) return fail(nullptr, rc, "generator launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PLF_OK;
}

int plf_generate_states_host(int states, float *x1, float *x2, size_t first_site, size_t n, uint64_t seed)
{
    if (states != 4 && states != 20) return fail(nullptr, PLF_ERR_INVALID, "STATES=%d is not supported (4 = DNA, 20 = protein)", states);
    if (!x1 || !x2) return fail(nullptr, PLF_ERR_INVALID, "NULL argument");
    plf::generate_states_host(states, x1, x2, first_site, n, seed);
    return PLF_OK;
}

int plf_states_kernel_info(int states, int math_mode, int variant, int threads_per_block, int *regs_per_thread,
                           int *block_threads, size_t *smem_bytes, int *tile_sites)
{
    if (states != 20) return fail(nullptr, PLF_ERR_INVALID, "plf_states_kernel_info: STATES=%d (use plf_kernel_info for DNA)", states);
    int rc = plf::aa_kernel_info(math_mode, variant, threads_per_block, regs_per_thread, block_threads, smem_bytes, tile_sites);
    if (rc == PLF_ERR_INVALID) return fail(nullptr, rc, "no 20-state kernel for variant %d / threads %d", variant, threads_per_block);
    if (rc != PLF_OK) return fail(nullptr, rc, "cudaFuncGetAttributes failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PLF_OK;
}

int plf_device_malloc(int device, void **ptr, size_t bytes)
{
    if (!ptr) return fail(nullptr, PLF_ERR_INVALID, "NULL out-pointer");
    *ptr = nullptr;
    PLF_CUDA(nullptr, cudaSetDevice(device));
    PLF_CUDA(nullptr, cudaMalloc(ptr, bytes ? bytes : 1));
    return PLF_OK;
}

int plf_device_free(void *ptr)
{
    if (ptr) PLF_CUDA(nullptr, cudaFree(ptr));
    return PLF_OK;
}

int plf_memcpy_h2d(void *dst_device, const void *src_host, size_t bytes, void *stream)
{
    if (bytes == 0) return PLF_OK;
    if (!dst_device || !src_host) return fail(nullptr, PLF_ERR_INVALID, "memcpy: NULL pointer");
    PLF_CUDA(nullptr, cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return PLF_OK;
}

int plf_memcpy_d2h(void *dst_host, const void *src_device, size_t bytes, void *stream)
{
    if (bytes == 0) return PLF_OK;
    if (!dst_host || !src_device) return fail(nullptr, PLF_ERR_INVALID, "memcpy: NULL pointer");
    PLF_CUDA(nullptr, cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return PLF_OK;
}

int plf_memset_device(void *dst_device, int value, size_t bytes, void *stream)
{
    if (bytes == 0) return PLF_OK;
    if (!dst_device) return fail(nullptr, PLF_ERR_INVALID, "memset: NULL pointer");
    PLF_CUDA(nullptr, cudaMemsetAsync(dst_device, value, bytes, static_cast<cudaStream_t>(stream)));
    return PLF_OK;
}

int plf_stream_sync(void *stream)
{
    PLF_CUDA(nullptr, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return PLF_OK;
}

// NVTX ranges for the hosts (the reference wraps its run loop in xrt::profile::user_range("roundtrip_exec_time"),
// host_mem.cpp:273,282,395).  NVTX v3 is header-only: without a profiler attached the calls cost a few nanoseconds.
int plf_range_push(const char *name)
{
    if (!name) return fail(nullptr, PLF_ERR_INVALID, "NULL range name");
    nvtxRangePushA(name);
    return PLF_OK;
}

int plf_range_pop(void)
{
    nvtxRangePop();
    return PLF_OK;
}

// Bare host-link probe (no kernels): what the platform gives plain pinned copies in both directions at once.
int plf_probe_host_link(int device, void *host_in, size_t h2d_bytes, void *host_out, size_t d2h_bytes, int reps, int pieces,
                        double *seconds)
{
    if (!seconds || reps < 1 || pieces < 1 || pieces > 64) return fail(nullptr, PLF_ERR_INVALID, "probe: bad arguments");
    PLF_CUDA(nullptr, cudaSetDevice(device));
    char *h_in = static_cast<char *>(host_in), *h_out = static_cast<char *>(host_out), *d_in = nullptr, *d_out = nullptr;
    const bool own_in = h2d_bytes && !h_in, own_out = d2h_bytes && !h_out;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaError_t e = cudaSuccess;
    if (own_in) e = cudaMallocHost(&h_in, h2d_bytes);
    if (e == cudaSuccess && own_out) e = cudaMallocHost(&h_out, d2h_bytes);
    if (e == cudaSuccess && h2d_bytes) e = cudaMalloc(&d_in, h2d_bytes);
    if (e == cudaSuccess && d2h_bytes) e = cudaMalloc(&d_out, d2h_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess && own_in) memset(h_in, 1, h2d_bytes);
    if (e == cudaSuccess && own_out) memset(h_out, 1, d2h_bytes);
    auto round = [&]() {
        for (int k = 0; k < pieces && e == cudaSuccess; ++k) {
            const size_t ci = (h2d_bytes / pieces) & ~(size_t)255, co = (d2h_bytes / pieces) & ~(size_t)255;
            if (h2d_bytes) e = cudaMemcpyAsync(d_in + k * ci, h_in + k * ci, k == pieces - 1 ? h2d_bytes - k * ci : ci, cudaMemcpyHostToDevice, s_in);
            if (e == cudaSuccess && d2h_bytes)
                e = cudaMemcpyAsync(h_out + k * co, d_out + k * co, k == pieces - 1 ? d2h_bytes - k * co : co, cudaMemcpyDeviceToHost, s_out);
        }
    };
    double dt = 0.0;
    if (e == cudaSuccess) {
        round();                                               // warm-up round
        if (e == cudaSuccess) e = cudaStreamSynchronize(s_in);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s_out);
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int r = 0; r < reps && e == cudaSuccess; ++r) round();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s_in);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s_out);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    cudaFree(d_in);
    cudaFree(d_out);
    if (own_in && h_in) cudaFreeHost(h_in);
    if (own_out && h_out) cudaFreeHost(h_out);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, e == cudaErrorMemoryAllocation ? PLF_ERR_NOMEM : PLF_ERR_CUDA, "host-link probe failed: %s", cudaGetErrorString(e));
    }
    *seconds = dt;
    return PLF_OK;
}

int plf_kernel_info(int variant, int math_mode, int *regs_per_thread, int *threads_per_block,
                    int *blocks_per_sm, int *num_sms)
{
    const int v = variant ? variant : kDefaultVariant;
    const int t = threads_per_block && *threads_per_block ? *threads_per_block : kDefaultThreads;
    KernelSel k = pick_kernel(math_mode, v, t);
    if (!k.fn) return fail(nullptr, PLF_ERR_INVALID, "unknown kernel variant %d / threads %d", variant, t);
    cudaFuncAttributes attr;
    PLF_CUDA(nullptr, cudaFuncGetAttributes(&attr, k.fn));
    int rc = prepare_kernel(nullptr, k);
    if (rc != PLF_OK) return rc;
    if (regs_per_thread) *regs_per_thread = attr.numRegs;
    if (threads_per_block) *threads_per_block = k.threads;
    if (blocks_per_sm)
        PLF_CUDA(nullptr, cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k.fn, k.threads, k.smem));
    if (num_sms && device_sms(num_sms) != 0) return fail(nullptr, PLF_ERR_CUDA, "no CUDA device");
    return PLF_OK;
}

}  // extern "C"
