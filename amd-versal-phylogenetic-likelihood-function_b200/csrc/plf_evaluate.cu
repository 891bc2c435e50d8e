// csrc/plf_evaluate.cu -- root log-likelihood across a branch: the step right after the newview
// path in RAxML-like codes.  NOT part of the reference repository: the
// algorithm is evaluateGTRGAMMA of standard-RAxML (evaluateGenericSpecial.c), which the reference's
// plf() was derived from (README.md:188-189,207-208), restated for the reference's CLV layout:
//
//   lnL = sum_i wgt[i] * ( log(0.25 * | sum_{j,k} x1[i,j,k] * x2[i,j,k] * diag[j,k] |)
//                          + (cnt1[i] + cnt2[i]) * log(2^-32) )
//
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

namespace plf {

constexpr int kEvalThreads = 256;
constexpr int kEvalUnroll = 4;

__global__ void __launch_bounds__(kEvalThreads, 3)
plf_evaluate_kernel(const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                    const int *__restrict__ cnt1, const int *__restrict__ cnt2,
                    const int *__restrict__ wgt, const float *__restrict__ diag, size_t n,
                    double *__restrict__ lnl, StreamScratch *__restrict__ scratch)
{
    const int lane = threadIdx.x & 31;
    const int cat = lane & 3;
    const float4 dg = __ldg(reinterpret_cast<const float4 *>(diag) + cat);
    const double d0 = dg.x, d1 = dg.y, d2 = dg.z, d3 = dg.w;
    const double log_min = -32.0 * 0.69314718055994530942;        // log(2^-32)
    const size_t n_vec = n * 4;
    const size_t n_pad = (n_vec + 31) & ~(size_t)31;   // whole warps: the 4 lanes of a site stay together
    const size_t stride = (size_t)gridDim.x * kEvalThreads;
    double acc = 0.0;
    // kEvalUnroll independent (site, category) elements per thread and iteration: all 2*kEvalUnroll 128-bit loads are
    // issued before the first use, so a thread keeps 128 B in flight instead of 32 (the kernel is latency-bound
    // otherwise: 5.3 TB/s with one element per iteration).
    for (size_t v0 = (size_t)blockIdx.x * kEvalThreads + threadIdx.x; v0 < n_pad; v0 += kEvalUnroll * stride) {
        float4 a[kEvalUnroll], b[kEvalUnroll];
        bool live[kEvalUnroll];
#pragma unroll
        for (int u = 0; u < kEvalUnroll; ++u) {
            const size_t v = v0 + u * stride;
            live[u] = v < n_vec;
            if (live[u]) {
                a[u] = ld_stream(x1 + v);
                b[u] = ld_stream(x2 + v);
            }
        }
#pragma unroll
        for (int u = 0; u < kEvalUnroll; ++u) {
            const size_t v = v0 + u * stride;
            if (v >= n_pad) break;                      // warp-uniform: n_pad and the strides are multiples of 32
            double t = 0.0;
            if (live[u])
                t = (double)a[u].x * (double)b[u].x * d0 + (double)a[u].y * (double)b[u].y * d1 +
                    (double)a[u].z * (double)b[u].z * d2 + (double)a[u].w * (double)b[u].w * d3;
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            if (live[u] && cat == 0) {
                const size_t s = v >> 2;
                double term = log(0.25 * fabs(t));
                int c = 0;
                if (cnt1) c += __ldg(cnt1 + s);
                if (cnt2) c += __ldg(cnt2 + s);
                term += (double)c * log_min;
                acc += (wgt ? (double)__ldg(wgt + s) : 1.0) * term;
            }
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

int launch_evaluate(const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                    const float *diag, size_t n, double *lnl, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return PLF_ERR_CUDA;
    size_t grid = (n * 4 + kEvalThreads * kEvalUnroll - 1) / (kEvalThreads * kEvalUnroll);
    if (grid > (size_t)sms * 3) grid = (size_t)sms * 3;
    if (grid > (size_t)kEvalMaxBlocks) grid = kEvalMaxBlocks;
    if (grid == 0) return PLF_OK;
    StreamScratch *scratch = nullptr;
    if (int rc = stream_scratch(stream, &scratch)) return rc;
    plf_evaluate_kernel<<<(int)grid, kEvalThreads, 0, stream>>>(reinterpret_cast<const float4 *>(x1),
                                                                reinterpret_cast<const float4 *>(x2), cnt1, cnt2,
                                                                wgt, diag, n, lnl, scratch);
    count_launches(1);
    return cudaGetLastError() == cudaSuccess ? PLF_OK : PLF_ERR_CUDA;
}

}  // namespace plf
