// csrc/plf_sel_ldg.cu -- instantiations of the register-staged ("ldg") newview kernel.
// Compiled twice: -DPLF_SEL_MATH=MathStrict -DPLF_SEL_NAME=select_ldg_strict and the FMA pair.
#include "plf_kernels.cuh"
#include "plf_registry.h"

namespace plf {
namespace {

using M = PLF_SEL_MATH;

template <int U, bool STREAM, int MINB>
KernelSel sel_threads(int threads)
{
    KernelSel k;
    switch (threads) {
    case 128: k.fn = plf_newview_ldg<M, U, STREAM, 128, MINB>; break;
    case 256: k.fn = plf_newview_ldg<M, U, STREAM, 256, MINB>; break;
    case 512: if (MINB == 1) k.fn = plf_newview_ldg<M, U, STREAM, 512, 1>; break;
    default: break;
    }
    k.threads = threads;
    k.sites_per_block_iter = (threads / 32) * 8 * U;
    return k;
}

template <int U>
KernelSel sel_u(int kind, int b, int threads)
{
    if (kind == 1) return b <= 1 ? sel_threads<U, false, 1>(threads) : KernelSel{};
    switch (b) {
    case 0: case 1: return sel_threads<U, true, 1>(threads);
    case 2: return sel_threads<U, true, 2>(threads);
    case 3: return sel_threads<U, true, 3>(threads);
    default: return KernelSel{};
    }
}

}  // namespace

KernelSel PLF_SEL_NAME(int u, int kind, int b, int threads)
{
    switch (u) {
    case 1: return sel_u<1>(kind, b, threads);
    case 2: return sel_u<2>(kind, b, threads);
    case 4: return sel_u<4>(kind, b, threads);
    default: return KernelSel{};
    }
}

}  // namespace plf
