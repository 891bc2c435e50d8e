// csrc/plf_evaluate.cu -- root log-likelihood across a branch: the step right after the newview
// path in RAxML-like codes.  NOT part of the reference repository: the
// algorithm is evaluateGTRGAMMA of standard-RAxML (evaluateGenericSpecial.c), which the reference's
// plf() was derived from (README.md:188-189,207-208), restated for the reference's CLV layout:
//
//   lnL = sum_i wgt[i] * ( log(0.25 * | sum_{j,k} x1[i,j,k] * x2[i,j,k] * diag[j,k] |)
//                          + (cnt1[i] + cnt2[i]) * log(2^-32) )
//
// for S = 4 (DNA) or S = 20 (protein) states per category k, four categories j.
// x1, x2 are the CLVs at the two ends of the branch (eigen-space, as newview leaves them), diag[j,k]
// = exp(lambda_k * rate_j * t) for the branch, cnt1/cnt2 the accumulated per-site scaler counts.
// One thread per (site, category) as in the newview kernels: a 128-bit load per child, four fp64
// products, two shuffles for the per-site sum; fp64 log and accumulation.  HBM-bound: 128 B/site (+ 4 B per
// count vector and for wgt).
// The sum is REPRODUCIBLE run to run: every block leaves its partial sum in the stream's scratch record, the
// last block to finish (a ticket counter) adds the partials in block order with a fixed reduction tree and
// makes the single addition to *lnl.  The grid is a function of (n, SM count) only, so the same input on
// the same device always takes the same summation order -- a likelihood that drives a tree search must not
// flicker in its last digits.
#include "../../include/b200plf.h"
#include "plf_kernels.cuh"
#include "plf_registry.h"

#include <atomic>
#include <cstdint>
#include <cstdlib>

namespace plf {

constexpr int kEvalThreads = 256;
constexpr size_t kEvalRingMinSites = (size_t)1 << 18;      // below ~2 stages per SM the load-use-load kernel is as fast

// S states per category (4: DNA, one 128-bit load per child and element; 20: protein, five), U independent
// (site, category) elements per thread and iteration.  diag is [category][state], 4*S floats.
template <int S, int U>
__global__ void __launch_bounds__(kEvalThreads, S == 4 ? (U == 1 ? 6 : 3) : 2)
plf_evaluate_kernel(const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                    const int *__restrict__ cnt1, const int *__restrict__ cnt2,
                    const int *__restrict__ wgt, const float *__restrict__ diag, size_t n,
                    double *__restrict__ lnl, StreamScratch *__restrict__ scratch)
{
    constexpr int Q = S / 4;                    // 128-bit words per (site, category)
    const int lane = threadIdx.x & 31;
    const int cat = lane & 3;
    float4 dg[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) dg[q] = __ldg(reinterpret_cast<const float4 *>(diag) + cat * Q + q);
    const double log_min = -32.0 * 0.69314718055994530942;        // log(2^-32)
    const size_t n_vec = n * 4;                 // (site, category) elements
    const size_t n_pad = (n_vec + 31) & ~(size_t)31;   // whole warps: the 4 lanes of a site stay together
    const size_t stride = (size_t)gridDim.x * kEvalThreads;
    double acc = 0.0;
    // All 2*U*Q 128-bit loads of an iteration are issued before the first use (ncu on the one-element version:
    // long_scoreboard 12 warps per issue, DRAM 56 %: too few bytes in flight).
    for (size_t v0 = (size_t)blockIdx.x * kEvalThreads + threadIdx.x; v0 < n_pad; v0 += U * stride) {
        float4 a[U][Q], b[U][Q];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
            live[u] = v < n_vec;
            if (live[u]) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    a[u][q] = ld_stream(x1 + v * Q + q);
                    b[u][q] = ld_stream(x2 + v * Q + q);
                }
            }
        }
        // Per element: the four category lanes of a site add up their partial sums (two shuffles).  With U == 4 the
        // logarithm of element u is then taken by the lane whose category is u: one log per LANE for four elements,
        // instead of one per warp instruction with 8 of 32 lanes active (the log is two thirds of the instructions).
        double mine = 0.0;
        size_t my_site = 0;
        bool my_live = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
            double t = 0.0;
            if (live[u]) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    t += (double)a[u][q].x * (double)b[u][q].x * (double)dg[q].x + (double)a[u][q].y * (double)b[u][q].y * (double)dg[q].y +
                         (double)a[u][q].z * (double)b[u][q].z * (double)dg[q].z + (double)a[u][q].w * (double)b[u][q].w * (double)dg[q].w;
            }
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            if (cat == (U == 4 ? u : 0)) {
                mine = t;
                my_site = v >> 2;
                my_live = live[u];
            }
            if (U != 4 && my_live && cat == 0) {          // one element per iteration: lane 0 of the site finishes it here
                double term = log(0.25 * fabs(mine));
                int c = 0;
                if (cnt1) c += __ldg(cnt1 + my_site);
                if (cnt2) c += __ldg(cnt2 + my_site);
                term += (double)c * log_min;
                acc += (wgt ? (double)__ldg(wgt + my_site) : 1.0) * term;
            }
        }
        if (U == 4 && my_live) {
            double term = log(0.25 * fabs(mine));
            int c = 0;
            if (cnt1) c += __ldg(cnt1 + my_site);
            if (cnt2) c += __ldg(cnt2 + my_site);
            term += (double)c * log_min;
            acc += (wgt ? (double)__ldg(wgt + my_site) : 1.0) * term;
        }
    }
    __shared__ double warp_acc[kEvalThreads / 32];
    __shared__ bool is_last;
    auto block_sum = [&](double v) {             // fixed tree: shuffles within the warp, then across the warps
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) warp_acc[threadIdx.x >> 5] = v;
        __syncthreads();
        double t = threadIdx.x < kEvalThreads / 32 ? warp_acc[threadIdx.x] : 0.0;
        if (threadIdx.x < 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        }
        return t;                                // valid in thread 0
    };
    const double mine = block_sum(acc);
    if (threadIdx.x == 0) {
        scratch->partials[blockIdx.x] = mine;
        __threadfence();
        is_last = atomicAdd(&scratch->ticket, 1ull) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double part = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kEvalThreads) part += __ldcg(&scratch->partials[b]);
    const double total = block_sum(part);
    if (threadIdx.x == 0) {
        atomicAdd(lnl, total);                   // the launch's only addition
        scratch->ticket = 0ull;                  // clean for the next launch on this stream
    }
}

// The same sum for S = 4 with the operands fed by the bulk-copy engine through an mbarrier ring, as in the newview kernels
// (plf_newview_tma): one producer lane keeps DEPTH stages of 768 sites (2 x 48 KB) in flight whatever the consumers are
// doing, where the load-use-load loop above leaves the memory system idle while a warp takes its logarithms (ncu on that
// kernel: DRAM 72 %, 9 warps per issue slot waiting on loads).  Static stage schedule (stage st belongs to block st mod
// grid), so the summation order -- and the result -- is still a function of (n, SM count) only.
constexpr int kRingWarps = 24, kRingU = 4, kRingDepth = 2;
constexpr int kRingThreads = (kRingWarps + 1) * 32;
constexpr int kRingStage = kRingWarps * 8 * kRingU;             // 768 sites per stage
constexpr size_t kRingSmem = (size_t)kRingDepth * kRingStage * 64 * 2 + 2 * kRingDepth * sizeof(uint64_t);

__global__ void __launch_bounds__(kRingThreads, 1)
plf_evaluate_ring(const float4 *__restrict__ x1, const float4 *__restrict__ x2, const int *__restrict__ cnt1,
                  const int *__restrict__ cnt2, const int *__restrict__ wgt, const float *__restrict__ diag, size_t n,
                  double *__restrict__ lnl, StreamScratch *__restrict__ scratch, int flags)
{
    constexpr int U = kRingU, STAGE = kRingStage, STAGE_F4 = STAGE * 4, TILE = 8 * U, DEPTH = kRingDepth;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s1 = reinterpret_cast<float4 *>(smem_raw);                 // [DEPTH][STAGE_F4]
    float4 *s2 = s1 + (size_t)DEPTH * STAGE_F4;
    uint64_t *full = reinterpret_cast<uint64_t *>(s2 + (size_t)DEPTH * STAGE_F4);
    uint64_t *empty = full + DEPTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t n_stages = (n + STAGE - 1) / STAGE;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            mbar_init(&full[d], 1);
            mbar_init(&empty[d], kRingWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    double acc = 0.0;
    if (warp == kRingWarps) {
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (size_t st = blockIdx.x; st < n_stages; st += gridDim.x) {
                mbar_wait(&empty[slot], phase ^ 1u);
                const size_t s0 = st * STAGE, left = n - s0;
                const uint32_t bytes = (uint32_t)(left < (size_t)STAGE ? left : (size_t)STAGE) * 64u;
                mbar_arrive_expect_tx(&full[slot], 2u * bytes);
                bulk_g2s(s1 + slot * STAGE_F4, x1 + s0 * 4, bytes, &full[slot]);
                bulk_g2s(s2 + slot * STAGE_F4, x2 + s0 * 4, bytes, &full[slot]);
                if (++slot == DEPTH) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        const int cat = lane & 3;
        const float4 dg = __ldg(reinterpret_cast<const float4 *>(diag) + cat);
        const double log_min = -32.0 * 0.69314718055994530942;    // log(2^-32)
        const uint32_t tile_off = warp * (TILE * 4) + lane;
        uint32_t slot = 0, phase = 0;
        for (size_t st = blockIdx.x; st < n_stages; st += gridDim.x) {
            // the lane with category u finishes element u of its four: its site within the range
            const size_t my_site = st * STAGE + (size_t)warp * TILE + (lane >> 2) + 8 * cat;
            const bool my_live = my_site < n;
            int c = 0, w = 1;                                     // asked for before the wait: 4 B per site each
            if (my_live) {
                if (cnt1) c += __ldg(cnt1 + my_site);
                if (cnt2) c += __ldg(cnt2 + my_site);
                if (wgt) w = __ldg(wgt + my_site);
            }
            const float4 *t1 = s1 + slot * STAGE_F4 + tile_off;
            const float4 *t2 = s2 + slot * STAGE_F4 + tile_off;
            mbar_wait(&full[slot], phase);
            float4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = t1[32 * u];
                b[u] = t2[32 * u];
            }
            unsigned dep = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) dep ^= __float_as_uint(a[u].x) ^ __float_as_uint(b[u].w);
            release_slot(&empty[slot], lane, dep, flags);
            if (++slot == DEPTH) {
                slot = 0;
                phase ^= 1u;
            }
            double mine = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // rows past the end of the site range were not copied: whatever the slot holds there is never used
                double t = (double)a[u].x * (double)b[u].x * (double)dg.x + (double)a[u].y * (double)b[u].y * (double)dg.y +
                           (double)a[u].z * (double)b[u].z * (double)dg.z + (double)a[u].w * (double)b[u].w * (double)dg.w;
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                if (cat == u) mine = t;
            }
            if (my_live) acc += (double)w * (log(0.25 * fabs(mine)) + (double)c * log_min);
        }
    }
    __shared__ double warp_acc[kRingThreads / 32];
    __shared__ bool is_last;
    auto block_sum = [&](double v) {             // fixed tree: shuffles within the warp, then across the warps
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) warp_acc[warp] = v;
        __syncthreads();
        double t = threadIdx.x < kRingThreads / 32 ? warp_acc[threadIdx.x] : 0.0;
        if (threadIdx.x < 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        }
        return t;                                // valid in thread 0
    };
    const double mine = block_sum(acc);
    if (threadIdx.x == 0) {
        scratch->partials[blockIdx.x] = mine;
        __threadfence();
        is_last = atomicAdd(&scratch->ticket, 1ull) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double part = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kRingThreads) part += __ldcg(&scratch->partials[b]);
    const double total = block_sum(part);
    if (threadIdx.x == 0) {
        atomicAdd(lnl, total);                   // the launch's only addition
        scratch->ticket = 0ull;                  // clean for the next launch on this stream
    }
}

static int launch_evaluate_ring(const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                                const float *diag, size_t n, double *lnl, cudaStream_t stream, int sms)
{
    static std::atomic<int> prepared[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PLF_ERR_CUDA;
    if (!prepared[dev].load(std::memory_order_acquire)) {
        if (cudaFuncSetAttribute(plf_evaluate_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingSmem) != cudaSuccess)
            return PLF_ERR_CUDA;
        prepared[dev].store(1, std::memory_order_release);
    }
    size_t grid = (n + kRingStage - 1) / kRingStage;
    if (grid > (size_t)sms) grid = sms;
    StreamScratch *scratch = nullptr;
    if (int rc = stream_scratch(stream, &scratch)) return rc;
    plf_evaluate_ring<<<(int)grid, kRingThreads, kRingSmem, stream>>>(reinterpret_cast<const float4 *>(x1),
                                                                     reinterpret_cast<const float4 *>(x2), cnt1, cnt2, wgt, diag,
                                                                     n, lnl, scratch, fenced_release(true) ? kFlagFencedRelease : 0);
    count_launches(1);
    return cudaGetLastError() == cudaSuccess ? PLF_OK : PLF_ERR_CUDA;
}

template <int S, int U>
static int launch_evaluate_t(const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                             const float *diag, size_t n, double *lnl, cudaStream_t stream, int sms)
{
    constexpr int bps = S == 4 ? (U == 1 ? 6 : 3) : 2;
    size_t grid = (n * 4 + kEvalThreads * U - 1) / (kEvalThreads * U);
    if (grid > (size_t)sms * bps) grid = (size_t)sms * bps;
    if (grid > (size_t)kEvalMaxBlocks) grid = kEvalMaxBlocks;
    if (grid == 0) return PLF_OK;
    StreamScratch *scratch = nullptr;
    if (int rc = stream_scratch(stream, &scratch)) return rc;
    plf_evaluate_kernel<S, U><<<(int)grid, kEvalThreads, 0, stream>>>(reinterpret_cast<const float4 *>(x1),
                                                                      reinterpret_cast<const float4 *>(x2), cnt1, cnt2, wgt,
                                                                      diag, n, lnl, scratch);
    count_launches(1);
    return cudaGetLastError() == cudaSuccess ? PLF_OK : PLF_ERR_CUDA;
}

int launch_evaluate(int states, const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                    const float *diag, size_t n, double *lnl, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return PLF_ERR_CUDA;
    // the ring kernel needs 16-byte aligned CLVs (bulk copies) and enough stages to fill its pipeline
    if (states == 4 && n >= kEvalRingMinSites && (((uintptr_t)x1 | (uintptr_t)x2) & 15u) == 0 && !getenv("PLF_EVAL_NO_RING"))
        return launch_evaluate_ring(x1, x2, cnt1, cnt2, wgt, diag, n, lnl, stream, sms);
    if (states == 4) return launch_evaluate_t<4, 4>(x1, x2, cnt1, cnt2, wgt, diag, n, lnl, stream, sms);
    if (states == 20) return launch_evaluate_t<20, 1>(x1, x2, cnt1, cnt2, wgt, diag, n, lnl, stream, sms);
    return PLF_ERR_INVALID;
}

}  // namespace plf
