// csrc/plf_protein_tc.cu -- 20-state newview on the 5th-generation tensor cores (tcgen05 + TMEM), tolerance mode.
//
// The 20-state path (the reference's STATES knob) is the one place in scope with a real
// contraction: per rate category and child, [sites x 20] . [20 x 20], and once more for the EV back-transform
// (app/src/plf.cpp:29-50 with 4 -> 20 states).  The CUDA-core kernel of plf_protein.cu is bound by shared-memory
// operand bandwidth and the fp32 pipe (4.7 G sites/s in FMA mode).  Here the three products run as
//
//        3xTF32:   A.B  ~=  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo,     x_hi = x truncated to tf32, x_lo = x - x_hi
//
// on tcgen05.mma (kind::tf32, M = 128 sites, fp32 accumulation in TMEM), each product as ONE accumulation chain of
// five MMAs (see kN below).  Each operand keeps 22 significant bits, the dropped lo.lo term is 2^-22 relative:
// fp32-class accuracy (worst observed 2e-6), but NOT the reference's rounding sequence -- bit-exactness is impossible
// on tensor cores, so this kernel exists for PLF_MATH_FMA only (<= 1e-5 relative, the tests state it);
// PLF_MATH_STRICT keeps the FMUL2/FADD kernel.  The tensor core reads fp32 denormal operands as zero (measured,
// tools/tc_denormal_probe.py); such an entry is >= 2^94 below its own site's rescaling threshold.
//
// Shape (one CTA per SM, 384 threads):
//   warps 0-3 / 4-7   two WORKER GROUPS of 128 threads; thread t of a group owns site row t of the group's current
//                     128-site tile = TMEM lane t.  A group is a software-pipelined sequence over (tile, category)
//                     steps (convert -> MMA -> product -> MMA -> read back); the two groups work on alternating tiles,
//                     so one group's CUDA-core phases overlap the other's tensor-core phases.
//   warps 8 / 9       one PRODUCER lane per group: TMA tensor copies (cp.async.bulk.tensor.2d, SASS UTMALDG) of
//                     {20 floats x 128 sites} boxes -- one (category, child) operand, 10 KB, row pitch 80 B, which makes
//                     the row-per-lane LDS.128 reads conflict-free -- into a 4-box ring per group.  Rows past the end
//                     of the site range are zero-filled by the TMA unit.
//   warps 10 / 11     one MMA-ISSUER warp per group: waits (converged) for the workers' "operand is in TMEM" mbarriers, an
//                     elected lane issues the five MMAs of a product back to back and commits to the mbarrier the
//                     workers wait on.
//   A operands        never touch shared memory again: a worker splits its row into hi/lo in registers and writes
//                     [hi(20) | lo(20)] to TMEM (tcgen05.st); the MMAs take A from TMEM and B (the nine constant
//                     matrices, pre-split, 40 x 40 blocks in the K-major no-swizzle canonical layout, 56 KB) from
//                     shared memory.
//   per category      a = x1.P_l^T, b = x2.P_r^T (10 MMAs) -> tcgen05.ld -> p = a*b in registers, split, tcgen05.st ->
//                     x3 = p.EV (5 MMAs) -> tcgen05.ld into registers; after the fourth category the thread holds its
//                     site's 80 results: threshold test, x 2^32, a staging row per site, one {80 x 32} TMA tensor store
//                     per warp (SASS UTMASTG), scaler byte, scaler count.
// TMEM: 512 columns = 2 groups x (3 accumulators x 40 + 3 A-operand regions x 40) (+ 16 unused per group).
// How it got here (3.3 -> 6.4 G sites/s) and what bounds it now: profiles/r02_protein_tc.md.
#include "../../include/b200plf.h"
#include "plf_kernels.cuh"
#include "plf_registry.h"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace plf {

namespace tc {

constexpr int kS = 20;                       // states
constexpr int kSite = 80;                    // floats per site
constexpr int kTile = 128;                   // sites per tile = MMA M = TMEM lanes
constexpr int kBoxBytes = kS * kTile * 4;    // one (category, child) box: 10 240 B
#ifndef PLF_TC_RING
#define PLF_TC_RING 4
#endif
constexpr int kRing = PLF_TC_RING;           // boxes per group ring (4 = 2 steps ahead)
// One 3xTF32 product as ONE accumulation chain of five MMAs: the A operand of a row is [hi(20) | lo(20)] (K = 40, five
// K = 8 steps, no padding), the B operand is the 40 x 40 block matrix
//            n < 20        n >= 20
//   k < 20   B_hi[n][k]    B_lo[n - 20][k]         D[:, 0:20]  = A_hi.B_hi + A_lo.B_hi
//   k >= 20  B_hi[n][k-20] 0                       D[:, 20:40] = A_hi.B_lo           (added on read-back)
// The price of a small MMA is per instruction, not per column (N is free up to 128: tools/microbench_mma.cu), so the
// obvious nine MMAs per product -- one per 3xTF32 term and K step -- cost nearly twice what these five do.
constexpr int kN = 40;                       // MMA N.  M = 128 with N % 8 == 0 is accepted by the hardware at cta_group::1 (CUTLASS's
                                             // static asserts want N % 16 == 0)
constexpr int kK = 40;                       // K (five K = 8 steps)
constexpr int kBMat = (kK / 4) * kN * 16;    // one B matrix in canonical K-major layout: 10 chunks x 40 rows x 16 B = 6400 B
constexpr int kNumB = 9;                     // P_left[4], P_right[4], EV
constexpr int kTraceCols = 13;               // PLF_TC_TRACE: counters per worker warp
constexpr int kThreads = 384;               // 8 worker warps, 2 producer warps, 2 MMA-issuer warps

// shared memory carve-up
constexpr size_t kOffRing = 0;                                         // [2 groups][kRing][kBoxBytes]
constexpr size_t kOffB = kOffRing + 2 * kRing * kBoxBytes;            // [kNumB][kBMat]
constexpr size_t kOffBar = kOffB + (size_t)kNumB * kBMat;             // barriers
constexpr size_t kOffOut = kOffBar + 512;             // after 26 mbarriers, the TMEM base address and (trace build) the clock stamps
constexpr size_t kSmemBytes = kOffOut + 2 * 4 * kBoxBytes;   // output staging: [2 groups][4 categories] boxes of {20 floats x 128 sites}

// TMEM columns of one group (the second group sits 256 columns further)
constexpr uint32_t kColAccA = 0, kColAccB = 40, kColAccX = 80;       // accumulators, 40 columns each
constexpr uint32_t kColA1 = 120, kColA2 = 160, kColP = 200;          // A operands [hi(20) | lo(20)]

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 40, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

__device__ __forceinline__ uint32_t rna_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = rna_tf32(x);
    lo = rna_tf32(x - __uint_as_float(hi));
}
// The split of the DATA path: hi = x truncated to tf32 (one LOP3), lo = x - hi (exact, <= 13 significant bits; the
// tensor core reads its leading 11).  cvt.rna.tf32.f32 is emulated on sm_100a (about five ALU instructions: ncu
// counted 2000 of the first version's 3300 instructions per tile and warp in it), and the two roundings to nearest buy
// nothing here: |x - (hi + tf32(lo))| <= 2^-21 |x| either way, against the 1e-5 tolerance of this mode.
__device__ __forceinline__ void split_trunc(float x, uint32_t &hi, uint32_t &lo)
{
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row x 16-byte core matrices,
// SBO = 128 B between 8-row groups, LBO = kN * 16 B between the two 16-byte K chunks of one K = 8 step.
__device__ __forceinline__ uint64_t b_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(((kN * 16) >> 4) & 0x3FFFu) << 16;      // leading byte offset (K direction)
    d |= (uint64_t)((128 >> 4) & 0x3FFFu) << 32;             // stride byte offset (N direction)
    d |= 1ull << 46;                                         // descriptor version (Blackwell)
    return d;                                                // base offset 0, layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// one lane of the (converged) warp, chosen by the hardware: the predicate nvcc recognises for uniform-datapath code
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float &a, float &b, float &c, float &d)
{
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    a = __uint_as_float(r0);
    b = __uint_as_float(r1);
    c = __uint_as_float(r2);
    d = __uint_as_float(r3);
}

// 20 consecutive columns of this thread's lane: one x16 and one x4 instruction
__device__ __forceinline__ void tmem_st20(uint32_t taddr, const uint32_t (&v)[20])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                    "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    tmem_st4(taddr + 16, v[16], v[17], v[18], v[19]);
}
__device__ __forceinline__ void tmem_ld20(uint32_t taddr, float (&f)[20])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]);
    tmem_ld4(taddr + 16, f[16], f[17], f[18], f[19]);
}

__device__ __forceinline__ void tma_box(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// shared -> global tensor copy of one {20 floats x 128 sites} box (SASS UTMASTG), bulk async-group of the issuing thread;
// rows past the end of the tensor are clipped by the TMA unit
__device__ __forceinline__ void tma_store_box(const CUtensorMap *map, int c0, int c1, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_global_v8(float *p, const float *v)
{
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// One matrix, split into hi and lo, as the 40 x 40 block matrix above in the canonical K-major layout:  B[n][k] at chunk
// (k / 4), row n, word (k % 4).  transpose == false: M[n][k] = src[n * 20 + k] (branch matrix P[kout][l]);  true:
// M[n][k] = src[k * 20 + n] (EV[k][l]).
__device__ __forceinline__ void stage_b(unsigned char *smem_b, int m, const float *__restrict__ src, bool transpose, int tid, int nthreads)
{
    float *dst = reinterpret_cast<float *>(smem_b + (size_t)m * kBMat);
    for (int idx = tid; idx < kS * kS; idx += nthreads) {
        const int n = idx / kS, k = idx - n * kS;
        const float v = __ldg(src + (transpose ? k * kS + n : n * kS + k));
        uint32_t h, l;
        split_tf32(v, h, l);
        auto word = [](int nn, int kk) { return (kk >> 2) * (kN * 4) + nn * 4 + (kk & 3); };
        dst[word(n, k)] = __uint_as_float(h);              // A_hi . B_hi
        dst[word(n, k + kS)] = __uint_as_float(h);         // A_lo . B_hi
        dst[word(n + kS, k)] = __uint_as_float(l);         // A_hi . B_lo
        dst[word(n + kS, k + kS)] = 0.0f;
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
plf_newview_aa_tc(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                  const __grid_constant__ CUtensorMap map3, const float *__restrict__ ev, const float *__restrict__ pl, const float *__restrict__ pr,
                  float *__restrict__ x3, unsigned char *__restrict__ scaler, const int *__restrict__ wgt, size_t n,
                  unsigned long long *__restrict__ scaler_sum, const int *__restrict__ cnt1, const int *__restrict__ cnt2,
                  int *__restrict__ cnt3, int flags, long long *__restrict__ trace)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ring = smem + kOffRing;
    unsigned char *smem_b = smem + kOffB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    uint64_t *full = bars;                    // [2][kRing]
    uint64_t *empty = bars + 2 * kRing;       // [2][kRing]
    uint64_t *mma_ab = bars + 4 * kRing;      // [2]  tensor core -> workers: a and b are in TMEM
    uint64_t *mma_x = mma_ab + 2;             // [2]  tensor core -> workers: x3 is in TMEM
    uint64_t *rdy_a = mma_x + 2;              // [2]  workers -> issuer: x1 operand is in TMEM (4 warp arrivals)
    uint64_t *rdy_b = rdy_a + 2;              // [2]  workers -> issuer: x2 operand is in TMEM
    uint64_t *rdy_p = rdy_b + 2;              // [2]  workers -> issuer: p operand is in TMEM
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(rdy_p + 2);
    // PLF_TC_TRACE only: clock stamps exchanged between workers and issuers of this CTA (same SM, same counter)
    volatile uint32_t *stamp_arrive_b = reinterpret_cast<volatile uint32_t *>(smem + kOffBar + 256);      // [2 groups][4 warps]
    volatile uint32_t *stamp_commit = stamp_arrive_b + 8;                                                 // [2 groups]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t n_tiles = (n + kTile - 1) / kTile;

    // ---- prologue: barriers, TMEM allocation, the nine matrices split and laid out for the tensor core ----
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * kRing; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 4);           // one arrive per worker warp of the group
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&mma_ab[i], 1);
            mbar_init(&mma_x[i], 1);
            mbar_init(&rdy_a[i], 4);
            mbar_init(&rdy_b[i], 4);
            mbar_init(&rdy_p[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = 0; j < 4; ++j) {
        stage_b(smem_b, j, pl + j * kS * kS, false, threadIdx.x, kThreads);
        stage_b(smem_b, 4 + j, pr + j * kS * kS, false, threadIdx.x, kThreads);
    }
    stage_b(smem_b, 8, ev, true, threadIdx.x, kThreads);
    fence_proxy_async_smem();                  // the matrices were written through the generic proxy; the MMA reads them through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    unsigned long long my_sum = 0;

    if (warp >= 10) {
        // ===== MMA issuers: one warp per group.  tcgen05.mma is issued by a single thread; a dedicated warp keeps the issue
        // off the workers' critical path.  The whole warp runs the loop and waits on the mbarriers, and one ELECTED lane
        // issues: with warp-uniform control flow the operands live in uniform registers and the five UTCHMMA of a product
        // sit back to back (170-230 cycles with the commit), where a branch on the lane number makes nvcc wrap every MMA
        // in an ELECT / R2UR.BROADCAST loop (~100 cycles each: profiles/r02_protein_tc.md). =====
        {
            const int g = __shfl_sync(0xffffffffu, warp, 0) - 10;
            const bool leader = elect_one();
            const uint32_t tcol = tmem_base + (uint32_t)g * 256u;
            const uint32_t b_base = smem_u32(smem_b);
            const size_t first = (size_t)blockIdx.x * 2 + g, stride = (size_t)gridDim.x * 2;
            const size_t my_tiles = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
            uint32_t ph = 0;                                   // all three ready barriers complete once per step
            uint32_t iacc[3] = {0u, 0u, 0u};
            auto product = [&](uint32_t acc, uint32_t a_op, int m) {      // acc = [A_hi | A_lo] . B[m]: one chain of five MMAs
                const uint32_t bm = b_base + (uint32_t)m * kBMat;
#pragma unroll
                for (int ks = 0; ks < kK / 8; ++ks) mma_ts(acc, a_op + 8 * ks, b_desc(bm + ks * 2 * kN * 16), ks > 0);
            };
            // the workers' order: ab(0); then per step s: x(s), ab(s + 1)
            const size_t steps = my_tiles * 4;
            for (size_t sidx = 0; sidx <= steps; ++sidx) {
                if (sidx > 0) {                                // x(sidx - 1) = p . EV
                    mbar_wait(&rdy_p[g], ph ^ 1u);             // completed in the PREVIOUS step's phase
                    tc_fence_after();
                    if (leader) {
                        product(tcol + kColAccX, tcol + kColP, 8);
                        mma_commit(&mma_x[g]);
                    }
                    __syncwarp();
                }
                if (sidx < steps) {                            // ab(sidx)
                    const int c = (int)(sidx & 3);
                    mbar_wait(&rdy_a[g], ph);
                    tc_fence_after();
                    if (leader) product(tcol + kColAccA, tcol + kColA1, c);
                    __syncwarp();
                    mbar_wait(&rdy_b[g], ph);
                    tc_fence_after();
                    if (leader) {
                        uint32_t t_wake = 0;
                        if constexpr (TRACE) {
                            t_wake = (uint32_t)clock();
                            uint32_t last = stamp_arrive_b[g * 4];
                            for (int q = 1; q < 4; ++q) {
                                const uint32_t v = stamp_arrive_b[g * 4 + q];
                                if ((int32_t)(v - last) > 0) last = v;
                            }
                            iacc[0] += t_wake - last;          // last worker arrive -> issuer awake
                        }
                        product(tcol + kColAccB, tcol + kColA2, 4 + c);
                        mma_commit(&mma_ab[g]);
                        if constexpr (TRACE) {
                            const uint32_t t_done = (uint32_t)clock();
                            iacc[1] += t_done - t_wake;        // five MMAs and the commit issued
                            stamp_commit[g] = t_done;
                            ++iacc[2];
                        }
                    }
                    __syncwarp();
                    ph ^= 1u;
                }
            }
            if constexpr (TRACE) {
                if (trace && leader) {
                    long long *row = trace + ((size_t)gridDim.x * 8 + (size_t)blockIdx.x * 2 + g) * kTraceCols;
                    row[0] = iacc[0];
                    row[1] = iacc[1];
                    row[2] = iacc[2];
                }
            }
        }
    } else if (warp >= 8) {
        // ===== producers: one lane per group =====
        if (lane == 0) {
            const int g = warp - 8;
            unsigned char *gring = ring + (size_t)g * kRing * kBoxBytes;
            uint32_t slot = 0, phase = 0;
            for (size_t tile = (size_t)blockIdx.x * 2 + g; tile < n_tiles; tile += (size_t)gridDim.x * 2) {
                const int row0 = (int)(tile * kTile);
                for (int c = 0; c < 4; ++c)
                    for (int child = 0; child < 2; ++child) {
                        mbar_wait(&empty[g * kRing + slot], phase ^ 1u);
                        mbar_arrive_expect_tx(&full[g * kRing + slot], kBoxBytes);
                        tma_box(gring + (size_t)slot * kBoxBytes, child ? &map2 : &map1, c * kS, row0, &full[g * kRing + slot]);
                        if (++slot == kRing) {
                            slot = 0;
                            phase ^= 1u;
                        }
                    }
            }
        }
    } else {
        // ===== worker groups =====
        const int g = warp >> 2;
        const int t = threadIdx.x & 127;                                  // site row of the tile = TMEM lane
        const unsigned char *gring = ring + (size_t)g * kRing * kBoxBytes;
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)g * 256u;    // this warp's lane quarter, this group's columns

        uint32_t slot = 0, phase = 0, ph_ab = 0, ph_x = 0;
        // PLF_TC_TRACE (debug build of the kernel): every worker thread adds up the cycles it spends in each kind of wait;
        // lane 0 of every worker warp writes its eight counters at the end (tools/tc_trace.py).  Unlike per-event stamps
        // this costs two clock reads per wait and perturbs nothing else.
        uint32_t wacc[kTraceCols] = {};
        uint32_t wacc_commit = 0;                              // commit issued by the issuer -> this worker awake
        const uint32_t t_begin = TRACE ? (uint32_t)clock() : 0u;
        auto wait_on = [&](int kind, uint64_t *bar, uint32_t parity) {
            if constexpr (TRACE) {
                const uint32_t t0 = (uint32_t)clock();
                mbar_wait(bar, parity);
                wacc[kind] += (uint32_t)clock() - t0;
            } else {
                mbar_wait(bar, parity);
            }
        };
        auto seg_begin = [&]() -> uint32_t { return TRACE ? (uint32_t)clock() : 0u; };
        auto seg_end = [&](int kind, uint32_t t0) {
            if constexpr (TRACE) wacc[kind] += (uint32_t)clock() - t0;
        };
        // convert: both children of the NEXT (tile, category) in ring order: ring -> registers -> hi/lo -> TMEM, then the
        // 10 branch MMAs  a = x1 . P_left[c]^T,  b = x2 . P_right[c]^T  (one chain of five per child, see kN)
        auto convert = [&]() {
#pragma unroll
            for (int child = 0; child < 2; ++child) {
                wait_on(child, &full[g * kRing + slot], phase);
                const float4 *row = reinterpret_cast<const float4 *>(gring + (size_t)slot * kBoxBytes + (size_t)t * (kS * 4));
                float4 v[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) v[q] = row[q];
                uint32_t hi[kS], lo[kS];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    split_trunc(v[q].x, hi[4 * q], lo[4 * q]);
                    split_trunc(v[q].y, hi[4 * q + 1], lo[4 * q + 1]);
                    split_trunc(v[q].z, hi[4 * q + 2], lo[4 * q + 2]);
                    split_trunc(v[q].w, hi[4 * q + 3], lo[4 * q + 3]);
                }
                // hand the slot back: every lane's reads are complete once their values have been used (the XOR below
                // depends on all of them), or, with the fenced release, ordered by fence.proxy.async
                unsigned dep = 0;
#pragma unroll
                for (int q = 0; q < 5; ++q) dep ^= hi[4 * q];
                if (flags & kFlagFencedRelease) fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (dep == 0x9E3779B9u) fence_proxy_async_smem();
                    mbar_arrive(&empty[g * kRing + slot]);
                }
                if (++slot == kRing) {
                    slot = 0;
                    phase ^= 1u;
                }
                tmem_st20(trow + (child ? kColA2 : kColA1), hi);
                tmem_st20(trow + (child ? kColA2 : kColA1) + kS, lo);
                // this warp's quarter of the operand is in TMEM: tell the issuer (it starts the five MMAs of this child
                // when all four warps have arrived, while the workers convert the other child)
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if constexpr (TRACE) {
                    if (lane == 0 && child) stamp_arrive_b[g * 4 + (warp & 3)] = (uint32_t)clock();
                }
                if (lane == 0) mbar_arrive(child ? &rdy_b[g] : &rdy_a[g]);
            }
        };
        // finish_ab: p = a * b in registers, split, back to TMEM as the A operand of the five EV MMAs
        auto finish_ab = [&]() {
            wait_on(2, &mma_ab[g], ph_ab);
            if constexpr (TRACE) wacc_commit += (uint32_t)clock() - stamp_commit[g];
            ph_ab ^= 1u;
            tc_fence_after();
            float a[kS], a2[kS], b[kS], b2[kS];               // the two column blocks of each accumulator
            tmem_ld20(trow + kColAccA, a);
            tmem_ld20(trow + kColAccA + kS, a2);
            tmem_ld20(trow + kColAccB, b);
            tmem_ld20(trow + kColAccB + kS, b2);
            tc_wait_ld();
            uint32_t hi[kS], lo[kS];
#pragma unroll
            for (int k = 0; k < kS; ++k) split_trunc((a[k] + a2[k]) * (b[k] + b2[k]), hi[k], lo[k]);
            tmem_st20(trow + kColP, hi);
            tmem_st20(trow + kColP + kS, lo);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&rdy_p[g]);
        };

        // Software pipeline over the (tile, category) steps of this group: the EV product of step s runs on the tensor
        // core while the CUDA cores convert the operands of step s + 1, and the branch products of step s + 1 run while
        // the results of step s are read back -- so of the two MMA round trips per step only a part of one is exposed.
        const size_t first = (size_t)blockIdx.x * 2 + g, stride = (size_t)gridDim.x * 2;
        if (first < n_tiles) convert();
        // Results.  A site's 320 bytes are contiguous in x3, but the lanes of a warp are 320 B apart: direct 256-bit stores
        // are 32 sector requests per instruction, 1280 per tile and group, and were a quarter of the time of the version that
        // used them (profiles/r02_protein_tc.md).  Instead every WARP stages its 32 sites in the global layout (32 rows of
        // 320 B, category by category as the results come out of TMEM) and its lane 0 hands them to the TMA unit as ONE
        // {80 floats x 32 sites} tensor store (SASS UTMASTG): asynchronous, rows past the end of the site range clipped
        // by the unit, and no synchronisation beyond the warp.  The row-per-lane writes at a 320 B pitch are 4-way bank
        // conflicts; they buy 32 requests of 320 B per warp and tile where category boxes {20 x 128} cost 128 of 80 B --
        // the TMA unit takes about one request per 3.7 cycles whatever its size (tools/microbench_tma.cu), and with
        // 80-byte rows in both directions that rate, not HBM, was the kernel's bound.
        const int wq = warp & 3;
        unsigned char *stage = smem + kOffOut + ((size_t)g * 4 + wq) * kBoxBytes;      // this warp's 32 rows x 320 B
        for (size_t tile = first; tile < n_tiles; tile += stride) {
            // running maximum of |x3| over the site's 80 results, as integer maxima of the magnitude bits in four
            // independent chains (a NaN or Inf has larger magnitude bits than any finite value, so "all 80 below 2^-32"
            // keeps the reference's meaning)
            uint32_t mx[4] = {0u, 0u, 0u, 0u};
            float out[4][kS];
            // the children's scaler counts and the site weight are needed at the end of the tile: ask for them now
            const size_t site = tile * kTile + (size_t)t;
            int cnt_in = 0, weight = 1;
            if (site < n) {
                if (cnt3) cnt_in = (cnt1 ? __ldg(cnt1 + site) : 0) + (cnt2 ? __ldg(cnt2 + site) : 0);
                if (wgt) weight = __ldg(wgt + site);
            }
            // Category c of this thread's site, times f, into its staging row.  Rows are 320 B apart, so the same 16-byte
            // chunk of eight consecutive rows falls on two bank groups (4-way conflicts).  Each pair of lanes therefore
            // writes its five chunks in an order rotated by (lane / 2) mod 4: instruction q stores chunk (q + rot) mod 5,
            // eight lanes hit eight bank groups (one 2-way collision in three of the five instructions).  The rotation
            // of the values is two levels of selects -- registers cannot be indexed by the lane.
            auto put_row = [&](int c, float f) {
                float4 v[5];
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    v[q] = make_float4(__fmul_rn(out[c][4 * q], f), __fmul_rn(out[c][4 * q + 1], f), __fmul_rn(out[c][4 * q + 2], f),
                                       __fmul_rn(out[c][4 * q + 3], f));
                const int rot = (lane >> 1) & 3;
                auto sel = [](bool take_b, const float4 &a, const float4 &b) {
                    return make_float4(take_b ? b.x : a.x, take_b ? b.y : a.y, take_b ? b.z : a.z, take_b ? b.w : a.w);
                };
                const bool r1 = rot & 1, r2 = rot & 2;
                float4 w[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) w[q] = sel(r1, v[q], v[(q + 1) % 5]);          // w[q] = v[(q + (rot & 1)) % 5]
#pragma unroll
                for (int q = 0; q < 5; ++q) v[q] = sel(r2, w[q], w[(q + 2) % 5]);          // v[q] = original chunk (q + rot) % 5
                float4 *row = reinterpret_cast<float4 *>(stage + (size_t)lane * (kSite * 4) + (size_t)c * (kS * 4));
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    int k = q + rot;
                    if (k >= 5) k -= 5;
                    row[k] = v[q];
                }
            };
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t ts = seg_begin();
                finish_ab();
                seg_end(9, ts);
                ts = seg_begin();
                if (c < 3 || tile + stride < n_tiles) convert();
                seg_end(8, ts);
                ts = seg_begin();
                wait_on(3, &mma_x[g], ph_x);
                ph_x ^= 1u;
                tc_fence_after();
                float o2[kS];
                tmem_ld20(trow + kColAccX, out[c]);
                tmem_ld20(trow + kColAccX + kS, o2);
                tc_wait_ld();
#pragma unroll
                for (int l = 0; l < kS; ++l) {
                    out[c][l] += o2[l];
                    mx[l & 3] = max(mx[l & 3], __float_as_uint(out[c][l]) & 0x7FFFFFFFu);
                }
                // staging runs one category behind: the previous tile's tensor stores have had a whole step more to read
                // the boxes before this tile's first write
                if (c == 1) {
                    const uint32_t tr = seg_begin();
                    if (lane == 0) bulk_wait_read_all();
                    __syncwarp();
                    seg_end(4, tr);
                }
                if (c >= 1) put_row(c - 1, 1.0f);
                seg_end(10, ts);
            }
            const uint32_t te = seg_begin();
            const bool small = max(max(mx[0], mx[1]), max(mx[2], mx[3])) < 0x2F800000u;      // bits of 2^-32
            put_row(3, small ? kTwoToThe32 : 1.0f);
            if (small) {                                       // the rows already staged, again, from the registers
                put_row(0, kTwoToThe32);
                put_row(1, kTwoToThe32);
                put_row(2, kTwoToThe32);
            }
            fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA unit
            __syncwarp();
            if (lane == 0) {
                tma_store_box(&map3, 0, (int)(tile * kTile) + wq * 32, stage);
                bulk_commit();
            }
            if (site < n) {
                if (scaler) scaler[site] = small ? 1 : 0;
                if (cnt3) cnt3[site] = cnt_in + (small ? 1 : 0);
                if (small) my_sum += (unsigned long long)(long long)weight;
            }
            seg_end(11, te);
        }
        if (lane == 0) bulk_wait_all();                        // every tensor store this lane has issued has completed
        if constexpr (TRACE) {
            if (trace && lane == 0) {
                wacc[6] = (uint32_t)clock() - t_begin;
                wacc[7] = (uint32_t)((n_tiles > first ? (n_tiles - first + stride - 1) / stride : 0));
                wacc[12] = wacc_commit;
                for (int k = 0; k < kTraceCols; ++k) trace[((size_t)blockIdx.x * 8 + warp) * kTraceCols + k] = (long long)wacc[k];
            }
        }
    }
    if (scaler_sum) block_add_u64<kThreads>(my_sum, scaler_sum);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// CLV as a 2-D tensor {80 floats, n sites}, row pitch 320 B; box = one category of 128 sites
static bool make_map(CUtensorMap *map, const float *x, size_t n, bool whole_rows = false)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kSite, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)kSite * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)(whole_rows ? kSite : kS), (cuuint32_t)(whole_rows ? 32 : kTile)};
    const cuuint32_t elem[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

int launch_newview_aa_tc(const float *x1, const float *x2, float *x3, unsigned char *scaler, const float *ev, const float *pl,
                         const float *pr, const int *wgt, size_t n, unsigned long long *scaler_sum, int flags, cudaStream_t stream,
                         const int *cnt1, const int *cnt2, int *cnt3)
{
    if (n == 0) return PLF_OK;
    if (n >= (1ull << 31)) return PLF_ERR_INVALID;                     // TMA coordinates are 32-bit
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return PLF_ERR_CUDA;
    CUtensorMap m1, m2, m3;
    if (!tc::make_map(&m1, x1, n) || !tc::make_map(&m2, x2, n) || !tc::make_map(&m3, x3, n, true)) return PLF_ERR_CUDA;
    const char *trace_path = getenv("PLF_TC_TRACE");
    auto kernel = trace_path ? tc::plf_newview_aa_tc<true> : tc::plf_newview_aa_tc<false>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes) != cudaSuccess)
        return PLF_ERR_CUDA;
    const size_t tiles = (n + tc::kTile - 1) / tc::kTile;
    size_t grid = (tiles + 1) / 2;
    if (grid > (size_t)sms) grid = sms;
    if (flags & kAaSingleCta) grid = 1;
    // PLF_TC_TRACE=<file> (debug): per worker warp, the cycles spent in each kind of wait, dumped as text:
    //   block warp  x1_box x2_box mma_ab mma_x staging_free staged total tiles | segments: convert finish_ab results tile_end |
    //               commit_to_awake          and per issuer lane (warp 10 + group):  arrive_to_awake  issue_of_b_chain  steps
    long long *d_trace = nullptr;
    const size_t kTraceWords = grid * 10 * tc::kTraceCols;      // 8 worker warps + 2 issuer lanes per CTA
    if (trace_path) {
        if (cudaMalloc(&d_trace, kTraceWords * sizeof(long long)) != cudaSuccess) return PLF_ERR_NOMEM;
        cudaMemsetAsync(d_trace, 0, kTraceWords * sizeof(long long), stream);
    }
    kernel<<<(int)grid, tc::kThreads, tc::kSmemBytes, stream>>>(m1, m2, m3, ev, pl, pr, x3, scaler, wgt, n, scaler_sum, cnt1, cnt2,
                                                                              cnt3, flags & kFlagFencedRelease, d_trace);
    count_launches(1);
    const bool ok = cudaGetLastError() == cudaSuccess;
    if (d_trace) {
        std::vector<long long> h(kTraceWords);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), d_trace, kTraceWords * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(d_trace);
        if (FILE *f = fopen(trace_path, "w")) {
            for (size_t w = 0; w < grid * 10; ++w) {        // worker rows, then the issuers' (warp 10 + group)
                if (w < grid * 8)
                    fprintf(f, "%zu %zu", w / 8, w % 8);
                else
                    fprintf(f, "%zu %zu", (w - grid * 8) / 2, 10 + (w - grid * 8) % 2);
                for (int k = 0; k < tc::kTraceCols; ++k) fprintf(f, " %lld", h[w * tc::kTraceCols + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return ok ? PLF_OK : PLF_ERR_CUDA;
}

int aa_tc_kernel_info(int *regs, int *block_threads, size_t *smem, int *tile_sites)
{
    cudaFuncAttributes attr;
    if (cudaFuncGetAttributes(&attr, tc::plf_newview_aa_tc<false>) != cudaSuccess) return PLF_ERR_CUDA;
    if (regs) *regs = attr.numRegs;
    if (block_threads) *block_threads = tc::kThreads;
    if (smem) *smem = tc::kSmemBytes;
    if (tile_sites) *tile_sites = tc::kTile;
    return PLF_OK;
}

}  // namespace plf
