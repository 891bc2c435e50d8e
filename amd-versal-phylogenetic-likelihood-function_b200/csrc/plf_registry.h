// csrc/plf_registry.h -- kernel variant registry shared by the selector translation units.
//
// variant = 1000*B + 100*D + 10*K + U
//   U : 128-bit loads per child per thread per tile (1, 2 or 4)
//   K : 0 = "ldg" register-staged kernel with streaming load/store policy
//       1 = "ldg" with plain (cached) loads/stores
//       2 = "tma" bulk-copy / mbarrier ring kernel, static stage schedule
//       3 = "tma-dyn": the same ring with dynamic stage scheduling (global work counter)
//   D : tma only: ring depth in stages (2, 3, 4, 6; 0 -> 4)
//   B : minimum resident blocks per SM given to __launch_bounds__ (0 -> 1); caps registers/thread
// threads_per_block is the number of COMPUTE threads: 128, 256 or 512 (tma adds one producer warp).
// The instantiations are spread over four translation units (ldg/tma x strict/fma) so that the
// library builds in parallel.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace plf {

using NewviewFn = void (*)(const float4 *, const float4 *, float4 *, unsigned char *, const float *,
                           const float *, const float *, const int *, size_t, unsigned long long *, int,
                           unsigned long long *);

struct KernelSel {
    NewviewFn fn = nullptr;
    int threads = 0;                 // launch block size
    size_t smem = 0;                 // dynamic shared memory
    int sites_per_block_iter = 0;    // sites one block consumes per loop iteration
    bool dynamic = false;            // needs a zeroed work-counter pair
};

// math: 0 strict, 1 fma.  Return fn == nullptr for combinations that are not compiled in.
KernelSel select_ldg_strict(int u, int kind, int b, int threads);
KernelSel select_ldg_fma(int u, int kind, int b, int threads);
KernelSel select_tma_strict(int u, int d, int b, int threads);
KernelSel select_tma_fma(int u, int d, int b, int threads);
KernelSel select_tma_dyn_strict(int u, int d, int b, int threads);
KernelSel select_tma_dyn_fma(int u, int d, int b, int threads);

// root log-likelihood kernel (plf_evaluate.cu); returns a plf_status
int launch_evaluate(int states, const float *x1, const float *x2, const int *cnt1, const int *cnt2, const int *wgt,
                    const float *diag, size_t n, double *lnl, cudaStream_t stream);

// 20-state (protein) newview and the S-state stimulus generator (plf_protein.cu); return a plf_status
int launch_newview_aa(const float *x1, const float *x2, float *x3, unsigned char *scaler, const float *ev,
                      const float *pl, const float *pr, const int *wgt, size_t n, unsigned long long *scaler_sum,
                      int math, int variant, int threads, int flags, cudaStream_t stream, const int *cnt1 = nullptr,
                      const int *cnt2 = nullptr, int *cnt3 = nullptr);
constexpr int kAaReleaseUnset = 1 << 17;  // launch flag of launch_newview_aa: no explicit slot-release choice, use the kernel's own default
constexpr int kAaSingleCta = 1 << 16;     // launch flag of launch_newview_aa: one block (stress-test hook); low bits = kernel flags
// the tensor-core (tcgen05 / TMEM, 3xTF32) 20-state kernel of plf_protein_tc.cu: tolerance mode only
int launch_newview_aa_tc(const float *x1, const float *x2, float *x3, unsigned char *scaler, const float *ev, const float *pl,
                         const float *pr, const int *wgt, size_t n, unsigned long long *scaler_sum, int flags, cudaStream_t stream,
                         const int *cnt1 = nullptr, const int *cnt2 = nullptr, int *cnt3 = nullptr);
int aa_tc_kernel_info(int *regs, int *block_threads, size_t *smem, int *tile_sites);
constexpr size_t kAaTensorCoreMinSites = 16384;   // FMA-mode calls at least this long default to the tensor-core kernel
constexpr int kAaVariantTensorCore = 9;   // opts->variant of the 20-state entry points: the tcgen05 kernel (PLF_MATH_FMA only)
int aa_kernel_info(int math, int variant, int threads, int *regs, int *block_threads, size_t *smem, int *tile_sites);
int launch_generate_states(int states, float *x1, float *x2, size_t first_site, size_t n, uint64_t seed, cudaStream_t stream);
void generate_states_host(int states, float *x1, float *x2, size_t first_site, size_t n, uint64_t seed);

// process-wide count of kernel launches issued by this library (plf_launch_count)
void count_launches(unsigned long long n);

// Per-(device, stream) scratch record in device memory (plf_capi.cu): launches on one stream are serialised, so
// they can share it; launches on different streams never do.
constexpr int kEvalMaxBlocks = 2048;
struct StreamScratch {
    unsigned long long work[2];           // work-counter pair of the dynamically scheduled kernels (self-cleaning)
    unsigned long long ticket;            // block ticket of the log-likelihood reduction (self-cleaning)
    unsigned long long pad;
    float mats[3600];                     // EV | P_left | P_right staged from HOST arrays (S = 4: 16|64|64, S = 20: 400|1600|1600)
    double partials[kEvalMaxBlocks];      // per-block partial sums, reduced in block order by the last block
};
int stream_scratch(cudaStream_t stream, StreamScratch **out);     // returns a plf_status

// Slot-release mechanism of the ring kernels: PLF_SAFE_RELEASE=1 forces the fenced release everywhere, =0 the
// data-dependency release everywhere; unset, every kernel family keeps its own default.
bool fenced_release(bool family_default);

}  // namespace plf
