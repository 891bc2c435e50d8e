// csrc/plf_sel_tma.cu -- instantiations of the bulk-copy / mbarrier ring ("tma") newview kernel.
// Compiled twice: -DPLF_SEL_MATH=MathStrict -DPLF_SEL_NAME=select_tma_strict and the FMA pair.
#include "plf_kernels.cuh"
#include "plf_registry.h"

namespace plf {
namespace {

using M = PLF_SEL_MATH;

template <int U, int WARPS, int DEPTH, int MINB>
KernelSel sel_one()
{
    KernelSel k;
    k.smem = tma_smem_bytes<U, WARPS, DEPTH>();
    if (k.smem * MINB > 227u * 1024u) return KernelSel{};
#ifdef PLF_SEL_DYNAMIC
    k.fn = plf_newview_tma_dyn<M, U, WARPS, DEPTH, MINB>;
    k.dynamic = true;
#else
    k.fn = plf_newview_tma<M, U, WARPS, DEPTH, MINB>;
#endif
    k.threads = (WARPS + 1) * 32;
    k.sites_per_block_iter = WARPS * 8 * U;
    return k;
}

template <int U, int WARPS, int MINB>
KernelSel sel_depth(int d)
{
    switch (d) {
    case 2: return sel_one<U, WARPS, 2, MINB>();
    case 3: return sel_one<U, WARPS, 3, MINB>();
    case 0: case 4: return sel_one<U, WARPS, 4, MINB>();
    case 6: return sel_one<U, WARPS, 6, MINB>();
    default: return KernelSel{};
    }
}

// consumer warps 4 / 8 / 16 with the launch bounds that still fit 2048 threads per SM
template <int U>
KernelSel sel_u(int d, int b, int threads)
{
    if (b == 0) b = 1;
    switch (threads) {
    case 128:
        switch (b) {
        case 1: return sel_depth<U, 4, 1>(d);
        case 2: return sel_depth<U, 4, 2>(d);
        case 3: return sel_depth<U, 4, 3>(d);
        default: return KernelSel{};
        }
    case 256:
        switch (b) {
        case 1: return sel_depth<U, 8, 1>(d);
        case 2: return sel_depth<U, 8, 2>(d);
        default: return KernelSel{};
        }
    case 512: return b == 1 ? sel_depth<U, 16, 1>(d) : KernelSel{};
    // 15 / 11 consumer warps + the producer warp = 16 / 12 warps: register allocation is per 4 warps, so
    // 17 warps are charged as 20 (96 registers per thread) while 16 get 128 and 12 get 168
    case 480: return b == 1 ? sel_depth<U, 15, 1>(d) : KernelSel{};
    case 352: return b == 1 ? sel_depth<U, 11, 1>(d) : KernelSel{};
    default: return KernelSel{};
    }
}

}  // namespace

KernelSel PLF_SEL_NAME(int u, int d, int b, int threads)
{
    switch (u) {
    case 1: return sel_u<1>(d, b, threads);
    case 2: return sel_u<2>(d, b, threads);
    case 4: return sel_u<4>(d, b, threads);
    default: return KernelSel{};
    }
}

}  // namespace plf
