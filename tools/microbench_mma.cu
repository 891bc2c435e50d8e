// tools/microbench_mma.cu -- what a SMALL tcgen05.mma costs (developer tool).
//
// The 20-state kernel's products are M = 128 sites x N = 40 x K = 8 per instruction: far below the shapes the tensor
// core is built for.  This program issues chains of 64 such MMAs from lane 0 of one or several warps of one CTA per SM
// and times each chain with clock64 from its first issue to the arrival of its tcgen05.commit on an mbarrier:
//
//   form     TS = A from tensor memory (what the kernel uses), SS = A from shared memory through a descriptor
//   acc      1 = every MMA accumulates into the same TMEM columns (a dependent chain, what one product is);
//            4 = round-robin over independent accumulators
//   issuers  warps issuing their own chain at the same time (or lanes of ONE warp)
//
// Measured on B200 (profiles/r02_microbench_mma.txt):
//   * issued by an ELECTED lane of a converged warp (what CUTLASS does, what the kernel does since v13): 20 cycles per
//     MMA up to N = 40, then N / 2 cycles (34 at N = 64, 66 at 128, 130 at 256: 2048 multiply-adds per cycle, the tf32
//     peak), the same for a dependent chain and for four independent accumulators; several issuing warps share that
//     rate (20 cycles per MMA in aggregate);
//   * issued from `if (lane == 0)`: 74 cycles per MMA whatever N <= 128, M, operand form or dependency -- nvcc wraps
//     every UTCHMMA of such a branch in an ELECT / R2UR.BROADCAST loop because the operands live in uniform registers.
//     Chains of different warps then overlap (37 cycles per MMA in aggregate with two, 14 with six).
// Operand contents are zeros; only the timing matters.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/microbench_mma tools/microbench_mma.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

template <bool SS, bool CONVERGED>
__global__ void __launch_bounds__(256, 1) bench(int n, int chain, int accs, int issuers, int lanes_mode, int m, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    // B: [n rows x 8 k] K-major no-swizzle: 2 chunks of n x 16 B;  A (SS form): [128 rows x 8 k] likewise
    unsigned char *b = smem, *a = smem + 8192;
    for (int i = threadIdx.x; i < (8192 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // zero the A columns (TS form): 8 columns per lane
    {
        const uint32_t row = tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + 448u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" :: "r"(row), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // issuer w: warp w lane 0 (lanes_mode 0) or lane w of warp 0 (lanes_mode 1); its accumulators start at column 64 * w
    // CONVERGED: the whole warp w runs the loop and elect.sync picks the issuing lane (uniform control flow: the MMA's
    // operands stay in uniform registers, no ELECT / R2UR.BROADCAST loop around every UTCHMMA)
    const int w = CONVERGED ? __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) : lanes_mode ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
    bool leader = true;
    if (CONVERGED) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
        leader = pred != 0;
    }
    const bool is_issuer = CONVERGED ? (w < issuers) : lanes_mode ? (threadIdx.x < (unsigned)issuers) : ((threadIdx.x & 31) == 0 && w < issuers);
    if (is_issuer) {
        uint64_t &bar = bars[w];
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
        const uint64_t bd = desc(smem_u32(b), (uint32_t)n * 16u, 128u);
        const uint64_t ad = desc(smem_u32(a), 128u * 16u, 128u);
        uint32_t phase = 0;
        long long best = 1ll << 60;
        for (int rep = 0; rep < 20; ++rep) {
            const long long t0 = clock64();
#pragma unroll 8
            for (int i = 0; i < chain; ++i) {
                const uint32_t d = tmem + (uint32_t)(w * 64 + (i & (accs - 1)) * 64);       // accumulators 64 columns apart (accs: power of two)
                const uint32_t acc = i >= accs ? 1u : 0u;                  // the first MMA into each accumulator overwrites
                if (!leader) continue;
                if (SS) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                                 :: "r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0u) : "memory");
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                                 :: "r"(d), "r"(tmem + 448u), "l"(bd), "r"(idesc), "r"(acc), "r"(0u) : "memory");
                }
            }
            if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            if (CONVERGED) __syncwarp();
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                         :: "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1u;
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (leader) out[blockIdx.x * 8 + w] = best;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
    }
}

static double run(bool ss, int n, int chain, int accs, int issuers, int lanes_mode, int m, long long *d, int sms, double *worst)
{
    const int smem = 8192 + 4096;
    std::vector<long long> h(sms * 8);
    if (lanes_mode == 2)
        bench<false, true><<<sms, 256, smem>>>(n, chain, accs, issuers, 0, m, d);
    else if (ss)
        bench<true, false><<<sms, 256, smem>>>(n, chain, accs, issuers, lanes_mode, m, d);
    else
        bench<false, false><<<sms, 256, smem>>>(n, chain, accs, issuers, lanes_mode, m, d);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        exit(1);
    }
    cudaMemcpy(h.data(), d, sms * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    std::vector<double> per;
    for (int b = 0; b < sms; ++b) {
        long long mx = 0;
        for (int w = 0; w < issuers; ++w) mx = std::max(mx, h[b * 8 + w]);
        per.push_back((double)mx / chain);
    }
    std::sort(per.begin(), per.end());
    *worst = per.back();
    return per[sms / 2];
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *d = nullptr;
    cudaMalloc(&d, sms * 8 * sizeof(long long));
    const int chain = 64;
    double worst;
    printf("# %d SMs, chains of %d MMAs, K = 8, kind::tf32; cycles per MMA = (first issue .. commit arrival) / %d, best of 20, median over SMs\n",
           sms, chain, chain);
    printf("## one issuing thread: operand form, N, independent accumulators (M = 128)\n");
    for (int ss = 0; ss < 2; ++ss)
        for (int n : {8, 40, 64, 128, 256})
            for (int accs : {1, 4}) {
                if (n > 64 && accs > 1) continue;
                const double c = run(ss, n, chain, accs, 1, 0, 128, d, sms, &worst);
                printf("form %s  N %3d  accumulators %d  cycles/MMA %.1f\n", ss ? "SS" : "TS", n, accs, c);
            }
    printf("## M = 64\n");
    for (int n : {8, 40, 64}) printf("form TS  M 64  N %3d  cycles/MMA %.1f\n", n, run(false, n, chain, 1, 1, 0, 64, d, sms, &worst));
    printf("## several issuing threads at once, each with its own accumulator (TS, M = 128, N = 40): cycles per MMA of the SLOWEST chain\n");
    for (int issuers : {1, 2, 3, 4, 6, 8}) {
        const double c = run(false, 40, chain, 1, issuers, 0, 128, d, sms, &worst);
        printf("issuers %d (lane 0 of %d warps)   cycles/MMA per chain %.1f   -> aggregate %.1f cycles per MMA\n", issuers, issuers, c, c / issuers);
    }
    for (int issuers : {2, 4}) {
        const double c = run(false, 40, chain, 1, issuers, 1, 128, d, sms, &worst);
        printf("issuers %d (lanes of ONE warp)     cycles/MMA per chain %.1f   -> aggregate %.1f cycles per MMA\n", issuers, c, c / issuers);
    }
    printf("## the same chains issued by an ELECTED lane of a converged warp (TS, M = 128): the tensor core itself\n");
    for (int n : {8, 40, 64, 128, 256})
        for (int accs : {1, 4}) {
            if (n > 64 && accs > 1) continue;
            printf("converged  N %3d  accumulators %d  cycles/MMA %.1f\n", n, accs, run(false, n, chain, accs, 1, 2, 128, d, sms, &worst));
        }
    for (int issuers : {2, 4}) {
        const double c = run(false, 40, chain, 1, issuers, 2, 128, d, sms, &worst);
        printf("converged  issuers %d (N = 40)   cycles/MMA per chain %.1f   -> aggregate %.1f cycles per MMA\n", issuers, c, c / issuers);
    }
    return 0;
}
