// csrc/plf_kernels.cuh -- the fused PLF "newview" kernels for sm_100a (B200).
//
// One kernel replaces the reference's whole device pipeline
//   mm2sleft / mm2sright  (hls/src/mm2s{left,right}_memDNAwindowComb.cpp:16-100)  DDR -> 4 lanes
//   transpose             (hls/src/transpose.cpp:6-24)                            P -> P^T
//   mmul_branch x8        (aie/src/128x9DNAwindow8192Comb/kernels/mmul_branch.cpp:6-41)
//   combine x4            (.../kernels/combine.cpp:4-39)
//   ev x4                 (.../kernels/ev.cpp:4-27)
//   s2mm                  (hls/src/s2mm_memDNAwindowComb.cpp:20-101)              rescale + scaler byte
//   host scaler reduction (app/src/host_mem.cpp:384-388)
// and computes exactly what the CPU golden plf() computes (app/src/plf.cpp:19-65).
//
// Mapping: one thread per (site, rate category) -- the reference's "lane" (one 128-bit PLIO stream
// per category, mm2sleft_memDNAwindowComb.cpp:88-96).  A warp covers 8 consecutive sites per
// 128-bit load instruction, so every warp-level request is 512 contiguous bytes.  The category's
// two 4x4 P matrices and its EV matrix (48 floats) live in registers for the whole kernel.
// The kernel is HBM-bound (193 B and ~400 flop per site): tensor cores are deliberately unused.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace plf {

constexpr float kMinLikelihood = 0x1p-32f;   // plf.cpp:5-6 (exact in fp32)
constexpr float kTwoToThe32 = 0x1p+32f;      // plf.cpp:5

// Kernel `flags` argument (bit 0 is the old ev_per_category int, so existing callers keep working).
constexpr int kFlagEvPerCategory = 1;     // ev points to EV4[4][16], one matrix per category (gen mode)
constexpr int kFlagFencedRelease = 2;     // release ring slots with fence.proxy.async instead of the data dependency

// ---------------------------------------------------------------------------------------------
// Arithmetic policies.
// STRICT reproduces the reference's rounding: each product and each sum is rounded to fp32
// separately and sums run left to right (plf.cpp:32-39, 45-50).  __fmul_rn/__fadd_rn are never
// contracted into FMA by nvcc.  FMA is the contracted variant (<= 1e-5 relative).
//
// The reference starts every sum from +0.0f.  In round-to-nearest "+0 + q0" differs from q0 only
// for q0 == -0, and by induction the whole sum differs only when ALL four products are -0 (result
// -0 instead of +0).  A zero a[k] or b[k] with the wrong sign can only flip the sign of zero
// products further down, and those matter only when an output element is exactly zero -- in
// which case the reference output is +0 (a sum that started from +0 is never -0).  So the leading
// add is dropped from the two branch mat-vecs (dot4_inner) and kept in the final EV mat-vec
// (dot4_final), which canonicalises any -0 to +0: outputs stay bit-identical to plf() at 92
// instead of 100 fp32 operations per (site, category).
// ---------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100+: SASS FMUL2 / FFMA2).  Each half is an independent IEEE-754
// round-to-nearest fp32 operation, so results are bit-identical to two scalar instructions, at half
// the issue slots.  The register pairs are the natural ones: (x0,x1),(x2,x3) of a 128-bit load and
// adjacent matrix constants, so no moves are needed.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct MathStrict {
    // the four products of a dot product: two FMUL2, each product rounded on its own
    static __device__ __forceinline__ void prod4(float a0, float a1, float a2, float a3, float b0, float b1,
                                                 float b2, float b3, float &q0, float &q1, float &q2, float &q3)
    {
        f2_unpack(f2_mul(f2_pack(a0, a1), f2_pack(b0, b1)), q0, q1);
        f2_unpack(f2_mul(f2_pack(a2, a3), f2_pack(b2, b3)), q2, q3);
    }
    // p[k] = a[k] * b[k] for two k at once
    static __device__ __forceinline__ void mul_pair(float a0, float a1, float b0, float b1, float &p0, float &p1)
    {
        f2_unpack(f2_mul(f2_pack(a0, a1), f2_pack(b0, b1)), p0, p1);
    }
    // ((a0*b0 + a1*b1) + a2*b2) + a3*b3
    static __device__ __forceinline__ float dot4_inner(float a0, float a1, float a2, float a3,
                                                       float b0, float b1, float b2, float b3)
    {
        float q0, q1, q2, q3;
        prod4(a0, a1, a2, a3, b0, b1, b2, b3, q0, q1, q2, q3);
        float acc = __fadd_rn(q0, q1);
        acc = __fadd_rn(acc, q2);
        acc = __fadd_rn(acc, q3);
        return acc;
    }
    // ((((+0 + a0*b0) + a1*b1) + a2*b2) + a3*b3)
    static __device__ __forceinline__ float dot4_final(float a0, float a1, float a2, float a3,
                                                       float b0, float b1, float b2, float b3)
    {
        float q0, q1, q2, q3;
        prod4(a0, a1, a2, a3, b0, b1, b2, b3, q0, q1, q2, q3);
        float acc = __fadd_rn(0.0f, q0);
        acc = __fadd_rn(acc, q1);
        acc = __fadd_rn(acc, q2);
        acc = __fadd_rn(acc, q3);
        return acc;
    }
};

// The same arithmetic with scalar FMUL/FADD only (92 instructions per site-category).  Used by the
// batch/tree kernel, whose 96-register budget (17 warps are allocated as 20) has no room for the
// pair constraints of the packed version: with FMUL2 it spills 164 bytes and a traversal is 12 %
// slower (profiles/r01_tree.md).
struct MathStrictScalar {
    static __device__ __forceinline__ void mul_pair(float a0, float a1, float b0, float b1, float &p0, float &p1)
    {
        p0 = __fmul_rn(a0, b0);
        p1 = __fmul_rn(a1, b1);
    }
    static __device__ __forceinline__ float dot4_inner(float a0, float a1, float a2, float a3,
                                                       float b0, float b1, float b2, float b3)
    {
        float acc = __fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
        acc = __fadd_rn(acc, __fmul_rn(a2, b2));
        acc = __fadd_rn(acc, __fmul_rn(a3, b3));
        return acc;
    }
    static __device__ __forceinline__ float dot4_final(float a0, float a1, float a2, float a3,
                                                       float b0, float b1, float b2, float b3)
    {
        float acc = __fadd_rn(0.0f, __fmul_rn(a0, b0));
        acc = __fadd_rn(acc, __fmul_rn(a1, b1));
        acc = __fadd_rn(acc, __fmul_rn(a2, b2));
        acc = __fadd_rn(acc, __fmul_rn(a3, b3));
        return acc;
    }
};

struct MathFma {
    static __device__ __forceinline__ void mul_pair(float a0, float a1, float b0, float b1, float &p0, float &p1)
    {
        p0 = a0 * b0;
        p1 = a1 * b1;
    }
    static __device__ __forceinline__ float dot4_inner(float a0, float a1, float a2, float a3,
                                                       float b0, float b1, float b2, float b3)
    {
        float acc = a0 * b0;
        acc = __fmaf_rn(a1, b1, acc);
        acc = __fmaf_rn(a2, b2, acc);
        acc = __fmaf_rn(a3, b3, acc);
        return acc;
    }
    static __device__ __forceinline__ float dot4_final(float a0, float a1, float a2, float a3,
                                                       float b0, float b1, float b2, float b3)
    {
        return dot4_inner(a0, a1, a2, a3, b0, b1, b2, b3);
    }
};

// ---------------------------------------------------------------------------------------------
// Memory access policies: streaming 128-bit loads/stores.  CLVs are touched exactly once, so
// loads bypass L1 allocation and stores are marked streaming.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_plain(const float4 *p) { return *p; }
__device__ __forceinline__ void st_plain(float4 *p, const float4 &v) { *p = v; }

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still draining: its CTAs take over SMs
// as the predecessor's CTAs exit and run their prologue (barrier init, smem carve-up).  pdl_wait()
// blocks until the predecessor has completed and its memory is visible -- it must precede the first
// global access; pdl_launch_dependents() lets the NEXT kernel in the stream begin the same way.
// Without the launch attribute both are no-ops.  Hides 2-3 us of launch latency per launch, which
// matters for 1 Mi-site calls (30 us kernels) and the small upper levels of a tree.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// The 48 per-category constants: P_left[j], P_right[j] as [k][l]; EV (or EV4[j]) TRANSPOSED as [l][k], so that
// the four factors of every dot product sit in adjacent registers (pairs for the packed multiplies).
struct CatConst {
    float L[16];
    float R[16];
    float E[16];
};

__device__ __forceinline__ void load_cat_const(CatConst &c, const float *__restrict__ ev,
                                               const float *__restrict__ pl,
                                               const float *__restrict__ pr, int cat,
                                               int ev_per_category)
{
    const float4 *l4 = reinterpret_cast<const float4 *>(pl) + 4 * cat;
    const float4 *r4 = reinterpret_cast<const float4 *>(pr) + 4 * cat;
    const float4 *e4 = reinterpret_cast<const float4 *>(ev) + (ev_per_category ? 4 * cat : 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float4 a = __ldg(l4 + k), b = __ldg(r4 + k), e = __ldg(e4 + k);
        c.L[4 * k + 0] = a.x; c.L[4 * k + 1] = a.y; c.L[4 * k + 2] = a.z; c.L[4 * k + 3] = a.w;
        c.R[4 * k + 0] = b.x; c.R[4 * k + 1] = b.y; c.R[4 * k + 2] = b.z; c.R[4 * k + 3] = b.w;
        c.E[k] = e.x; c.E[4 + k] = e.y; c.E[8 + k] = e.z; c.E[12 + k] = e.w;      // transposed: E[4*l + k] = EV[k][l]
    }
}

// One (site, category): a = P_l x1, b = P_r x2 (plf.cpp:29-39), p = a.b (:41),
// x3[l] = sum_k p[k] EV[k][l] (:45-50).  Returns true when all four |x3| < 2^-32 (:53-56).
// Split in two so that a branch product can also come from a table (tip children, see the batch kernel):
//   category_branch : a[k] = sum_l x[l] P[k][l]          category_finish : p = a.b, x3 = p EV, threshold test
template <class M>
__device__ __forceinline__ float4 category_branch(const float (&P)[16], const float4 &x)
{
    return make_float4(M::dot4_inner(x.x, x.y, x.z, x.w, P[0], P[1], P[2], P[3]),
                       M::dot4_inner(x.x, x.y, x.z, x.w, P[4], P[5], P[6], P[7]),
                       M::dot4_inner(x.x, x.y, x.z, x.w, P[8], P[9], P[10], P[11]),
                       M::dot4_inner(x.x, x.y, x.z, x.w, P[12], P[13], P[14], P[15]));
}

template <class M>
__device__ __forceinline__ bool category_finish(const CatConst &c, const float4 &a, const float4 &b, float4 &o)
{
    float p[4];
    M::mul_pair(a.x, a.y, b.x, b.y, p[0], p[1]);
    M::mul_pair(a.z, a.w, b.z, b.w, p[2], p[3]);
    o.x = M::dot4_final(p[0], p[1], p[2], p[3], c.E[0], c.E[1], c.E[2], c.E[3]);
    o.y = M::dot4_final(p[0], p[1], p[2], p[3], c.E[4], c.E[5], c.E[6], c.E[7]);
    o.z = M::dot4_final(p[0], p[1], p[2], p[3], c.E[8], c.E[9], c.E[10], c.E[11]);
    o.w = M::dot4_final(p[0], p[1], p[2], p[3], c.E[12], c.E[13], c.E[14], c.E[15]);
    // NaN compares false, exactly like ABS(x) < minlikelihood on the CPU.
    return (fabsf(o.x) < kMinLikelihood) & (fabsf(o.y) < kMinLikelihood) &
           (fabsf(o.z) < kMinLikelihood) & (fabsf(o.w) < kMinLikelihood);
}

template <class M>
__device__ __forceinline__ bool category_newview(const CatConst &c, const float4 &u, const float4 &v,
                                                 float4 &o)
{
    return category_finish<M>(c, category_branch<M>(c.L, u), category_branch<M>(c.R, v), o);
}

__device__ __forceinline__ void rescale(float4 &o)
{
    o.x = __fmul_rn(o.x, kTwoToThe32);   // exact: power-of-two scale of a value < 2^-32
    o.y = __fmul_rn(o.y, kTwoToThe32);
    o.z = __fmul_rn(o.z, kTwoToThe32);
    o.w = __fmul_rn(o.w, kTwoToThe32);
}

// Site s of a warp tile occupies lanes 4s..4s+3: the site rescales when its whole nibble is set.
__device__ __forceinline__ bool nibble_all(unsigned ballot, int site_in_warp)
{
    return ((ballot >> (4 * site_in_warp)) & 0xFu) == 0xFu;
}

// Block-wide sum of per-thread counters -> one atomic per block (host_mem.cpp:384-388 fused).
template <int THREADS>
__device__ __forceinline__ void block_add_u64(unsigned long long v, unsigned long long *dst)
{
    __shared__ unsigned long long warp_sums[THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long t = threadIdx.x < THREADS / 32 ? warp_sums[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0 && t != 0ull) atomicAdd(dst, t);
    }
}

// ---------------------------------------------------------------------------------------------
// Variant 1 ("ldg"): persistent grid-stride kernel, register-staged.
//   U      : 128-bit loads per child per thread per tile; a warp tile is 8*U sites.
//   STREAM : use the streaming load/store policies above.
// Per tile each thread issues 2*U independent 128-bit loads before the first use (memory-level
// parallelism), computes U (site,category) results, votes per site with one ballot per load
// row, stores U 128-bit results and lanes 0..8U-1 store one scaler byte each (one contiguous
// 8U-byte run per warp).
// ---------------------------------------------------------------------------------------------
template <class M, int U, bool STREAM, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
plf_newview_ldg(const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                float4 *__restrict__ x3, unsigned char *__restrict__ scaler,
                const float *__restrict__ ev, const float *__restrict__ pl,
                const float *__restrict__ pr, const int *__restrict__ wgt, size_t n,
                unsigned long long *__restrict__ scaler_sum, int flags,
                unsigned long long * /*work: unused, static schedule*/)
{
    constexpr int TILE = 8 * U;                       // sites per warp tile
    const int lane = threadIdx.x & 31;
    const int cat = lane & 3;
    const int site_in_row = lane >> 2;                // 0..7

    pdl_wait();
    pdl_launch_dependents();
    CatConst c;
    load_cat_const(c, ev, pl, pr, cat, flags & kFlagEvPerCategory);

    const size_t warps_total = (size_t)gridDim.x * (THREADS / 32);
    const size_t warp_id = (size_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const size_t n_tiles = (n + TILE - 1) / TILE;
    unsigned long long my_sum = 0;

    for (size_t tile = warp_id; tile < n_tiles; tile += warps_total) {
        const size_t s0 = tile * TILE;
        const bool full = s0 + TILE <= n;
        float4 a[U], b[U], o[U];
        unsigned ballots[U];

        if (full) {
            const float4 *p1 = x1 + s0 * 4 + lane;
            const float4 *p2 = x2 + s0 * 4 + lane;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = STREAM ? ld_stream(p1 + 32 * u) : ld_plain(p1 + 32 * u);
                b[u] = STREAM ? ld_stream(p2 + 32 * u) : ld_plain(p2 + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool small = category_newview<M>(c, a[u], b[u], o[u]);
                ballots[u] = __ballot_sync(0xffffffffu, small);
                if (nibble_all(ballots[u], site_in_row)) rescale(o[u]);
            }
            float4 *p3 = x3 + s0 * 4 + lane;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (STREAM) st_stream(p3 + 32 * u, o[u]); else st_plain(p3 + 32 * u, o[u]);
            }
        } else {
            // ragged last tile: same code, predicated per load row
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t s = s0 + 8 * u + site_in_row;
                const bool live = s < n;
                a[u] = live ? x1[s * 4 + cat] : make_float4(1.f, 1.f, 1.f, 1.f);
                b[u] = live ? x2[s * 4 + cat] : make_float4(1.f, 1.f, 1.f, 1.f);
                bool small = category_newview<M>(c, a[u], b[u], o[u]);
                ballots[u] = __ballot_sync(0xffffffffu, small && live);
                if (nibble_all(ballots[u], site_in_row)) rescale(o[u]);
                if (live) x3[s * 4 + cat] = o[u];
            }
        }

        // scaler bytes + weighted count: lane L owns site s0+L of the tile
        if (lane < TILE) {
            unsigned bal = ballots[0];
#pragma unroll
            for (int u = 1; u < U; ++u) bal = (lane >> 3) == u ? ballots[u] : bal;
            const bool scaled = nibble_all(bal, lane & 7);
            const size_t s = s0 + lane;
            if (s < n) {
                if (scaler) scaler[s] = scaled ? 1 : 0;
                if (scaled) my_sum += wgt ? (unsigned long long)(long long)wgt[s] : 1ull;
            }
        }
    }
    if (scaler_sum) block_add_u64<THREADS>(my_sum, scaler_sum);
}

// ---------------------------------------------------------------------------------------------
// Variant 2 ("tma"): warp-specialised, shared-memory staged with the bulk-copy engine.
//
//   producer warp (1 elected lane): cp.async.bulk global -> shared (TMA bulk copy, SASS UBLKCP)
//       of one STAGE = WARPS*8*U consecutive sites of x1 and of x2, completion counted on an
//       mbarrier (complete_tx::bytes); DEPTH stages form a ring, so up to DEPTH*STAGE*128 bytes
//       are in flight per CTA without holding a single register.
//   consumer warps: wait on the stage's "full" mbarrier, read their float4 with conflict-free
//       LDS.128 (lane i reads bytes 16i..16i+15 of a 512-byte row), compute, vote, rescale and
//       store x3 / scaler bytes straight to global memory with 128-bit streaming stores, then
//       release the stage through its "empty" mbarrier.
// The ring decouples memory-level parallelism from occupancy: the 48 matrix constants per thread
// cost registers, not bytes in flight.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Releasing a ring slot.  The consumer's LDS reads of the slot must be COMPLETE before the producer is
// allowed to refill it with the bulk-copy engine: an mbarrier.arrive issued right after the LDS
// instructions can overtake them, and with a fast refill (data in L2, one busy CTA) the first bytes of
// the next stage then land in the slot while a warp is still reading it (found by the tree stress
// test, profiles/r01_tree.md).  Two ways to close the window:
//   * fence.proxy.async.shared::cta before the arrive (the CUTLASS consumer_release pattern for
//     ld.shared consumers).  ptxas emits MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S, which also waits for the
//     warp's outstanding GLOBAL stores -- invisible while DRAM-bound, but it caps a traffic-light
//     kernel (compressed tips) at ~2600 cycles per stage;
//   * make the arrive DATA-DEPENDENT on the loaded registers: a load whose value has been consumed
//     has been performed, and a performed read cannot be affected by a later write of any proxy.
// mbar_release_slot() does the second: `dep` is an XOR over one register of every LDS of the slot;
// the comparison can never be proven by the compiler, both branches arrive exactly once, and the
// (practically never taken) equal branch falls back to the fence.
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_release_slot(uint64_t *bar, int lane, unsigned dep)
{
    __syncwarp();
    if (lane == 0) {
        if (dep != 0x9E3779B9u) {
            mbar_arrive(bar);
        } else {
            fence_proxy_async_smem();
            mbar_arrive(bar);
        }
    }
}
// The release written against the PTX memory model: EVERY lane orders its own generic-proxy reads of the
// slot before later async-proxy writes (fence.proxy.async), the warp converges, one lane arrives.
__device__ __forceinline__ void mbar_release_slot_fenced(uint64_t *bar, int lane)
{
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
template <bool = true>
__device__ __forceinline__ void release_slot(uint64_t *bar, int lane, unsigned dep, int flags)
{
    if (flags & kFlagFencedRelease) mbar_release_slot_fenced(bar, lane);
    else mbar_release_slot(bar, lane, dep);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class M, int U, int WARPS, int DEPTH, int MINB>
__global__ void __launch_bounds__((WARPS + 1) * 32, MINB)
plf_newview_tma(const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                float4 *__restrict__ x3, unsigned char *__restrict__ scaler,
                const float *__restrict__ ev, const float *__restrict__ pl,
                const float *__restrict__ pr, const int *__restrict__ wgt, size_t n,
                unsigned long long *__restrict__ scaler_sum, int flags,
                unsigned long long * /*work: unused, static schedule*/)
{
    constexpr int THREADS = (WARPS + 1) * 32;
    constexpr int TILE = 8 * U;                 // sites per consumer warp per stage
    constexpr int STAGE = WARPS * TILE;         // sites per stage
    constexpr int STAGE_F4 = STAGE * 4;         // float4 per child per stage

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s1 = reinterpret_cast<float4 *>(smem_raw);                 // [DEPTH][STAGE_F4]
    float4 *s2 = s1 + (size_t)DEPTH * STAGE_F4;                        // [DEPTH][STAGE_F4]
    uint64_t *full = reinterpret_cast<uint64_t *>(s2 + (size_t)DEPTH * STAGE_F4);
    uint64_t *empty = full + DEPTH;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const size_t n_stages = (n + STAGE - 1) / STAGE;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            mbar_init(&full[d], 1);             // one arrive.expect_tx by the producer
            mbar_init(&empty[d], WARPS);        // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();                    // everything above overlapped the previous kernel's tail
    pdl_launch_dependents();

    unsigned long long my_sum = 0;

    if (warp == WARPS) {
        // ===== producer =====
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (size_t st = blockIdx.x; st < n_stages; st += gridDim.x) {
                mbar_wait(&empty[slot], phase ^ 1u);
                const size_t s0 = st * STAGE;
                const size_t left = n - s0;
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
        // ===== consumers =====
        const int cat = lane & 3;
        const int site_in_row = lane >> 2;
        CatConst c;
        load_cat_const(c, ev, pl, pr, cat, flags & kFlagEvPerCategory);

        const size_t last_full = n / STAGE;                 // stages [0, last_full) are complete
        const size_t stride_f4 = (size_t)gridDim.x * STAGE_F4;
        // this lane's output slot in the CTA's first stage; advanced by one grid stride per iteration
        float4 *out = x3 + ((size_t)blockIdx.x * STAGE + (size_t)warp * TILE) * 4 + lane;
        unsigned char *sc_out = scaler ? scaler + (size_t)blockIdx.x * STAGE + warp * TILE + lane : nullptr;
        const int *w_in = wgt ? wgt + (size_t)blockIdx.x * STAGE + warp * TILE + lane : nullptr;
        const uint32_t tile_off = warp * (TILE * 4) + lane;

        uint32_t slot = 0, phase = 0;
        for (size_t st = blockIdx.x; st < n_stages; st += gridDim.x) {
            const float4 *t1 = s1 + slot * STAGE_F4 + tile_off;
            const float4 *t2 = s2 + slot * STAGE_F4 + tile_off;
            mbar_wait(&full[slot], phase);
            float4 a[U], b[U], o[U];
            unsigned ballots[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = t1[32 * u];
                b[u] = t2[32 * u];
            }
            // release the slot as soon as this warp's reads of it are complete
            unsigned dep = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) dep ^= __float_as_uint(a[u].x) ^ __float_as_uint(b[u].w);
            release_slot(&empty[slot], lane, dep, flags);
            if (++slot == DEPTH) {
                slot = 0;
                phase ^= 1u;
            }

            if (st < last_full) {
                // complete stage: no bounds predicates anywhere
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bool small = category_newview<M>(c, a[u], b[u], o[u]);
                    ballots[u] = __ballot_sync(0xffffffffu, small);
                    if (nibble_all(ballots[u], site_in_row)) rescale(o[u]);
                    st_stream(out + 32 * u, o[u]);
                }
                if (lane < TILE) {
                    unsigned bal = ballots[0];
#pragma unroll
                    for (int u = 1; u < U; ++u) bal = (lane >> 3) == u ? ballots[u] : bal;
                    const bool scaled = nibble_all(bal, lane & 7);
                    if (sc_out) *sc_out = scaled ? 1 : 0;
                    if (scaled) my_sum += w_in ? (unsigned long long)(long long)*w_in : 1ull;
                }
            } else {
                // the single ragged stage at the end of the site range
                const size_t s0 = st * STAGE + (size_t)warp * TILE;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool live = s0 + 8 * u + site_in_row < n;
                    bool small = category_newview<M>(c, a[u], b[u], o[u]);
                    ballots[u] = __ballot_sync(0xffffffffu, small && live);
                    if (nibble_all(ballots[u], site_in_row)) rescale(o[u]);
                    if (live) st_stream(out + 32 * u, o[u]);
                }
                if (lane < TILE && s0 + lane < n) {
                    unsigned bal = ballots[0];
#pragma unroll
                    for (int u = 1; u < U; ++u) bal = (lane >> 3) == u ? ballots[u] : bal;
                    const bool scaled = nibble_all(bal, lane & 7);
                    if (sc_out) *sc_out = scaled ? 1 : 0;
                    if (scaled) my_sum += w_in ? (unsigned long long)(long long)*w_in : 1ull;
                }
            }
            out += stride_f4;
            if (sc_out) sc_out += (size_t)gridDim.x * STAGE;
            if (w_in) w_in += (size_t)gridDim.x * STAGE;
        }
    }
    if (scaler_sum) block_add_u64<THREADS>(my_sum, scaler_sum);
}

template <int U, int WARPS, int DEPTH>
constexpr size_t tma_smem_bytes()
{
    return (size_t)DEPTH * (WARPS * 8 * U) * 64 * 2 + (size_t)DEPTH * 2 * sizeof(uint64_t) + 64;
}

// ---------------------------------------------------------------------------------------------
// Variant 2b ("tma-dyn"): the same ring kernel with DYNAMIC stage scheduling.  With the static
// schedule (stage st to CTA st % grid) every SM is given the same share, so the launch lasts as long
// as its slowest SM: SMs do not get equal DRAM service (two dies, distance to the L2 slices), and
// the launch-to-launch spread of the static kernel shows it (8 Mi sites: mean 0.256 ms, min
// 0.240 ms).  Here the producer of each CTA takes the next stage index from a global counter
// (`work[0]`), publishes it to its consumers through shared memory together with the stage's
// data, and a sentinel index ends the loop.  The first DEPTH stages of a CTA are dealt statically and
// two counter fetches are kept in flight, so neither the ramp of a launch nor its steady state waits
// for an atomic round trip.  The last PRODUCER to run out of work (`work[1]` counts them, as soon as a
// producer has consumed its last fetch -- not behind a barrier at the end of the kernel) zeroes both
// words: the pair is clean for its next user without a memset.
// ---------------------------------------------------------------------------------------------
template <class M, int U, int WARPS, int DEPTH, int MINB>
__global__ void __launch_bounds__((WARPS + 1) * 32, MINB)
plf_newview_tma_dyn(const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                    float4 *__restrict__ x3, unsigned char *__restrict__ scaler,
                    const float *__restrict__ ev, const float *__restrict__ pl,
                    const float *__restrict__ pr, const int *__restrict__ wgt, size_t n,
                    unsigned long long *__restrict__ scaler_sum, int flags,
                    unsigned long long *work)
{
    constexpr int THREADS = (WARPS + 1) * 32;
    constexpr int TILE = 8 * U;
    constexpr int STAGE = WARPS * TILE;
    constexpr int STAGE_F4 = STAGE * 4;
    constexpr uint32_t kDone = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s1 = reinterpret_cast<float4 *>(smem_raw);
    float4 *s2 = s1 + (size_t)DEPTH * STAGE_F4;
    uint64_t *full = reinterpret_cast<uint64_t *>(s2 + (size_t)DEPTH * STAGE_F4);
    uint64_t *empty = full + DEPTH;
    volatile uint32_t *stage_of = reinterpret_cast<volatile uint32_t *>(empty + DEPTH);   // [DEPTH]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_stages = (n + STAGE - 1) / STAGE;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            mbar_init(&full[d], 1);
            mbar_init(&empty[d], WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();

    unsigned long long my_sum = 0;

    if (warp == WARPS) {
        // ===== producer =====
        if (lane == 0) {
            // The first DEPTH stages of every CTA are dealt statically (stage blockIdx.x + i*gridDim.x): the whole
            // ring is requested in the first microsecond of the launch instead of one stage per atomic round trip.
            // The counter hands out the stages from DEPTH*gridDim.x on; two fetches are kept in flight, so a
            // producer never waits for an atomic that it issued less than two stages ago.
            uint32_t slot = 0, phase = 0;
            const unsigned long long base = (unsigned long long)DEPTH * gridDim.x;
            unsigned long long next_a = base + atomicAdd(work, 1ull);
            unsigned long long next_b = base + atomicAdd(work, 1ull);
            for (unsigned long long it = 0;; ++it) {
                unsigned long long st;
                if (it < (unsigned long long)DEPTH) {
                    st = blockIdx.x + it * gridDim.x;         // >= n_stages only if every dynamic index is too
                } else {
                    st = next_a;
                    next_a = next_b;
                    next_b = base + atomicAdd(work, 1ull);    // needed two iterations from now
                }
                mbar_wait(&empty[slot], phase ^ 1u);
                if (st >= n_stages) {                     // out of work: tell the consumers and stop
                    stage_of[slot] = kDone;
                    mbar_arrive(&full[slot]);
                    break;
                }
                stage_of[slot] = (uint32_t)st;
                const size_t s0 = (size_t)st * STAGE;
                const size_t left = n - s0;
                const uint32_t bytes = (uint32_t)(left < (size_t)STAGE ? left : (size_t)STAGE) * 64u;
                mbar_arrive_expect_tx(&full[slot], 2u * bytes);
                bulk_g2s(s1 + slot * STAGE_F4, x1 + s0 * 4, bytes, &full[slot]);
                bulk_g2s(s2 + slot * STAGE_F4, x2 + s0 * 4, bytes, &full[slot]);
                if (++slot == DEPTH) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
            // Every fetch this producer issued has returned (the empty asm consumes both values), so this CTA is
            // done with the counter: report it NOW, while the consumers still work on the last ring stages, instead
            // of behind a block barrier at the very end of the kernel (an atomic round trip on every CTA's tail).
            // The last producer out zeroes the pair: clean for its next user without a memset.
            asm volatile("" :: "l"(next_a), "l"(next_b));
            const unsigned long long finished = atomicAdd(work + 1, 1ull);
            if (finished == gridDim.x - 1) {
                work[0] = 0ull;
                work[1] = 0ull;
                __threadfence();
            }
        }
    } else {
        // ===== consumers =====
        const int cat = lane & 3;
        const int site_in_row = lane >> 2;
        CatConst c;
        load_cat_const(c, ev, pl, pr, cat, flags & kFlagEvPerCategory);
        const size_t last_full = n / STAGE;
        const uint32_t tile_off = warp * (TILE * 4) + lane;

        uint32_t slot = 0, phase = 0;
        for (;;) {
            const float4 *t1 = s1 + slot * STAGE_F4 + tile_off;
            const float4 *t2 = s2 + slot * STAGE_F4 + tile_off;
            mbar_wait(&full[slot], phase);
            const uint32_t st = stage_of[slot];
            if (st == kDone) break;
            float4 a[U], b[U], o[U];
            unsigned ballots[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = t1[32 * u];
                b[u] = t2[32 * u];
            }
            unsigned dep = st;
#pragma unroll
            for (int u = 0; u < U; ++u) dep ^= __float_as_uint(a[u].x) ^ __float_as_uint(b[u].w);
            release_slot(&empty[slot], lane, dep, flags);
            if (++slot == DEPTH) {
                slot = 0;
                phase ^= 1u;
            }
            const size_t s0 = (size_t)st * STAGE + (size_t)warp * TILE;
            float4 *out = x3 + s0 * 4 + lane;
            const bool complete = st < last_full;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool live = complete || s0 + 8 * u + site_in_row < n;
                bool small = category_newview<M>(c, a[u], b[u], o[u]);
                ballots[u] = __ballot_sync(0xffffffffu, small && live);
                if (nibble_all(ballots[u], site_in_row)) rescale(o[u]);
                if (live) st_stream(out + 32 * u, o[u]);
            }
            if (lane < TILE && (complete || s0 + lane < n)) {
                unsigned bal = ballots[0];
#pragma unroll
                for (int u = 1; u < U; ++u) bal = (lane >> 3) == u ? ballots[u] : bal;
                const bool scaled = nibble_all(bal, lane & 7);
                if (scaler) scaler[s0 + lane] = scaled ? 1 : 0;
                if (scaled) my_sum += wgt ? (unsigned long long)(long long)wgt[s0 + lane] : 1ull;
            }
        }
    }
    if (scaler_sum) block_add_u64<THREADS>(my_sum, scaler_sum);
}

// ---------------------------------------------------------------------------------------------
// Variant 3 ("batch"): many independent newviews of the same site count in ONE launch, with
// per-site scaler-COUNT accumulation -- the building block of a chained newview over a tree
// (all inner nodes of one tree level are independent: BASELINE.json configs[4]).
//
// Same producer/consumer ring as plf_newview_tma.  The work list is the concatenation of the
// stages of all ops, dealt to the CTAs in chunks of consecutive stages, so a CTA changes op (and
// reloads the 48 matrix constants and the op's pointers) rarely.  Per site the op additionally
// reads the children's int32 scaler counts (NULL = tip = 0) and writes
//     cnt3[i] = cnt1[i] + cnt2[i] + (site i rescaled ? 1 : 0)
// which is what RAxML-style codes carry up the tree instead of the single byte of one call
// (up to 204 B/site/node: 192 + 3 x 4).
// The children's counts ride in the ring too: the producer adds one bulk copy per count vector
// (STAGE x 4 bytes) to the stage's mbarrier transaction.  Reading them with per-warp 64-byte
// loads instead cost 23 % of the traversal (profiles/r01_tree.md): the small requests arrive
// late under a saturated DRAM queue and stall the warp that issued them.
// Count vectors must be 16-byte aligned and padded to a multiple of 4 ints.
// ---------------------------------------------------------------------------------------------
struct cuda_true { static constexpr bool value = true; };
struct cuda_false { static constexpr bool value = false; };

struct BatchOp {
    const float4 *x1;
    const float4 *x2;
    float4 *x3;
    const int *cnt1;      // per-site scaler counts of the children; NULL for a tip
    const int *cnt2;
    int *cnt3;            // may be NULL
    unsigned char *scaler;  // this newview's own 0/1 byte per site; may be NULL
    const float *pl;      // P_left[64], P_right[64], EV[16] of this op
    const float *pr;
    const float *ev;
    // Compressed tips: when tip1 / tip2 is non-NULL that child is a tip stored
    // as one state code per site (0..15) and its CLV is x[i][j][l] = tipvec[code_i][l] for every
    // category j (RAxML's tipVector lookup); x1 / x2 is then ignored.  1 B/site instead of 64.
    const unsigned char *tip1;
    const unsigned char *tip2;
    const float *tipvec;  // [16][4], shared by all ops of a launch; NULL when no op has a compressed tip
    // Tip-tip nodes (both children compressed tips): the parent CLV of a site depends only on (code1, code2), so the
    // whole newview -- branch products, a.b, EV product, threshold test, x 2^32 -- is tabulated ONCE per op for the 256
    // code pairs by plf_tiptip_tables (16 KB per op, with the functions of the per-site path: same bits) and a site then
    // costs one L1-resident 128-bit table read per category instead of ~45 instructions.  RAxML tabulates the two branch
    // products for this case; with 4-bit codes the whole product fits.  NULL for every other kind of node.
    const float4 *tiptab;          // [256 code pairs][4 categories]: the final x3 of the pair, already rescaled if it rescales
    const unsigned char *tipflag;  // [256]: 1 when the pair rescales
};

// One block of 256 threads per op of the tree; blocks of ops without a tip-tip table return at once.
template <class M>
__global__ void __launch_bounds__(256)
plf_tiptip_tables(const BatchOp *__restrict__ ops, int n_ops)
{
    const int op = blockIdx.x;
    if (op >= n_ops) return;
    const BatchOp o = ops[op];
    if (!o.tiptab) return;
    const int c1 = threadIdx.x >> 4, c2 = threadIdx.x & 15;
    const float4 *tv = reinterpret_cast<const float4 *>(o.tipvec);
    const float4 t1 = __ldg(tv + c1), t2 = __ldg(tv + c2);
    float4 r[4];
    bool small = true;
#pragma unroll
    for (int cat = 0; cat < 4; ++cat) {
        CatConst c;
        load_cat_const(c, o.ev, o.pl, o.pr, cat, 0);
        small = category_finish<M>(c, category_branch<M>(c.L, t1), category_branch<M>(c.R, t2), r[cat]) && small;
    }
    float4 *tab = const_cast<float4 *>(o.tiptab) + threadIdx.x * 4;
#pragma unroll
    for (int cat = 0; cat < 4; ++cat) {
        if (small) rescale(r[cat]);
        tab[cat] = r[cat];
    }
    const_cast<unsigned char *>(o.tipflag)[threadIdx.x] = small ? 1 : 0;
}

template <int U, int WARPS, int DEPTH>
constexpr size_t batch_smem_bytes()
{
    return tma_smem_bytes<U, WARPS, DEPTH>() + (size_t)DEPTH * (WARPS * 8 * U) * (sizeof(int) * 2 + 2) + 256 +   // tma_smem_bytes already holds the barriers + stage indices
           (size_t)WARPS * 2 * 64 * sizeof(float4);                                                               // per-warp tip product tables
}

template <class M, int U, int WARPS, int DEPTH, int MINB>
__global__ void __launch_bounds__((WARPS + 1) * 32, MINB)
plf_newview_batch(const BatchOp *__restrict__ ops, int n_ops, size_t n,
                  const int *__restrict__ wgt, unsigned long long *__restrict__ scaler_sum,
                  unsigned chunk, unsigned long long *work, int flags)
{
    constexpr int THREADS = (WARPS + 1) * 32;
    constexpr int TILE = 8 * U;
    constexpr int STAGE = WARPS * TILE;
    constexpr int STAGE_F4 = STAGE * 4;
    constexpr uint32_t kDone = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s1 = reinterpret_cast<float4 *>(smem_raw);                 // [DEPTH][STAGE_F4]
    float4 *s2 = s1 + (size_t)DEPTH * STAGE_F4;                        // [DEPTH][STAGE_F4]
    int *c1 = reinterpret_cast<int *>(s2 + (size_t)DEPTH * STAGE_F4);  // [DEPTH][STAGE]
    int *c2 = c1 + (size_t)DEPTH * STAGE;                              // [DEPTH][STAGE]
    unsigned char *k1s = reinterpret_cast<unsigned char *>(c2 + (size_t)DEPTH * STAGE);   // [DEPTH][STAGE] tip codes
    unsigned char *k2s = k1s + (size_t)DEPTH * STAGE;
    float4 *tv = reinterpret_cast<float4 *>(k2s + (size_t)DEPTH * STAGE);                 // [16] tip vector table
    float4 *tabs = tv + 16;                                                               // [WARPS][2][64] tip product tables
    uint64_t *full = reinterpret_cast<uint64_t *>(tabs + (size_t)WARPS * 128);
    uint64_t *empty = full + DEPTH;
    volatile uint32_t *stage_of = reinterpret_cast<volatile uint32_t *>(empty + DEPTH);   // [DEPTH] global stage index

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // The work list: stage g of the launch is stage (g % spo) of op (g / spo).  It is dealt out in
    // chunks of `chunk` consecutive stages.  With a work counter the producers take chunks dynamically
    // (as plf_newview_tma_dyn takes stages; the fetch of the next chunk is issued a chunk ahead);
    // without one chunk q goes to CTA q % gridDim.x.  Either way the producer publishes the global
    // stage index of every slot to its consumers and ends with a sentinel.
    const uint32_t spo = (uint32_t)((n + STAGE - 1) / STAGE);          // stages per op
    const uint32_t full_stages = (uint32_t)(n / STAGE);                // stages [0, full_stages) are complete
    const uint32_t total = spo * (uint32_t)n_ops;
    const uint32_t n_chunks = (total + chunk - 1) / chunk;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            mbar_init(&full[d], 1);
            mbar_init(&empty[d], WARPS);
        }
        mbar_fence_init();
    }
    pdl_wait();                    // everything above overlapped the previous kernel's tail
    if (threadIdx.x < 16 && ops[0].tipvec) tv[threadIdx.x] = __ldg(reinterpret_cast<const float4 *>(ops[0].tipvec) + threadIdx.x);
    __syncthreads();
    pdl_launch_dependents();

    unsigned long long my_sum = 0;
    if (warp == WARPS) {
        // ===== producer =====
        if (lane == 0) {
            uint32_t slot = 0, phase = 0, cur_op = 0xffffffffu;
            const float4 *x1 = nullptr, *x2 = nullptr;
            const int *k1 = nullptr, *k2 = nullptr;
            const unsigned char *tp1 = nullptr, *tp2 = nullptr;
            unsigned long long q = work ? atomicAdd(work, 1ull) : (unsigned long long)blockIdx.x;
            while (q < n_chunks) {
                const unsigned long long q_next = work ? atomicAdd(work, 1ull) : q + gridDim.x;   // used a chunk from now
                const uint32_t g0 = (uint32_t)q * chunk;
                const uint32_t g1 = g0 + chunk < total ? g0 + chunk : total;
                uint32_t op = g0 / spo, st = g0 - op * spo;
                for (uint32_t g = g0; g < g1; ++g) {
                    if (op != cur_op) {          // global loads on the producer's critical path: only on op change
                        x1 = ops[op].x1;
                        x2 = ops[op].x2;
                        k1 = ops[op].cnt1;
                        k2 = ops[op].cnt2;
                        tp1 = ops[op].tip1;
                        tp2 = ops[op].tip2;
                        cur_op = op;
                    }
                    mbar_wait(&empty[slot], phase ^ 1u);
                    stage_of[slot] = g;
                    const size_t s0 = (size_t)st * STAGE;
                    const uint32_t sites = st < full_stages ? (uint32_t)STAGE : (uint32_t)(n - s0);
                    const uint32_t bytes = sites * 64u;
                    const uint32_t cbytes = ((sites * 4u) + 15u) & ~15u;     // count vectors are padded to 16 B
                    const uint32_t tbytes = (sites + 15u) & ~15u;            // so are tip code vectors
                    mbar_arrive_expect_tx(&full[slot], (tp1 ? tbytes : bytes) + (tp2 ? tbytes : bytes) +
                                                           (k1 ? cbytes : 0u) + (k2 ? cbytes : 0u));
                    if (tp1) bulk_g2s(k1s + slot * STAGE, tp1 + s0, tbytes, &full[slot]);
                    else bulk_g2s(s1 + slot * STAGE_F4, x1 + s0 * 4, bytes, &full[slot]);
                    if (tp2) bulk_g2s(k2s + slot * STAGE, tp2 + s0, tbytes, &full[slot]);
                    else bulk_g2s(s2 + slot * STAGE_F4, x2 + s0 * 4, bytes, &full[slot]);
                    if (k1) bulk_g2s(c1 + slot * STAGE, k1 + s0, cbytes, &full[slot]);
                    if (k2) bulk_g2s(c2 + slot * STAGE, k2 + s0, cbytes, &full[slot]);
                    if (++slot == DEPTH) {
                        slot = 0;
                        phase ^= 1u;
                    }
                    if (++st == spo) {
                        st = 0;
                        ++op;
                    }
                }
                q = q_next;
            }
            mbar_wait(&empty[slot], phase ^ 1u);          // out of work: tell the consumers and stop
            stage_of[slot] = kDone;
            mbar_arrive(&full[slot]);
            if (work) {                                   // every fetched chunk index has been consumed: done with the counter
                const unsigned long long finished = atomicAdd(work + 1, 1ull);
                if (finished == gridDim.x - 1) {          // last producer out: leave the pair clean for the next level
                    work[0] = 0ull;
                    work[1] = 0ull;
                    __threadfence();
                }
            }
        }
    } else {
        // ===== consumers =====
        const int cat = lane & 3;
        const int site_in_row = lane >> 2;
        const uint32_t tile_off = warp * (TILE * 4) + lane;
        float4 *my_tab = tabs + warp * 128;
        CatConst c;
        BatchOp o;
        uint32_t cur_op = 0xffffffffu, op_base = 0;
        uint32_t slot = 0, phase = 0;
        for (;;) {
            const float4 *t1 = s1 + slot * STAGE_F4 + tile_off;
            const float4 *t2 = s2 + slot * STAGE_F4 + tile_off;
            mbar_wait(&full[slot], phase);
            const uint32_t g = stage_of[slot];
            if (g == kDone) break;
            uint32_t st = g - op_base;
            if (cur_op == 0xffffffffu || st >= spo) {    // first stage, or a stage of another op: its pointers and 48 constants
                cur_op = g / spo;
                op_base = cur_op * spo;
                st = g - op_base;
                o = ops[cur_op];
                load_cat_const(c, o.ev, o.pl, o.pr, cat, 0);
                // Tip children: the branch product a[k] = sum_l tipvec[code][l] P[k][l] depends only on (code, category),
                // so it is computed ONCE per op for the 16 codes (RAxML's umpX tables) with the very function the
                // per-site path uses -- same inputs, same bits -- and a tip child then costs one table read per site
                // instead of a 16-byte CLV read and 28 fp32 operations.  Tables are private to the warp (no block
                // barrier); rows are stored at code ^ (code >> 1) so that the four nucleotide codes 1, 2, 4, 8 fall on
                // both halves of the banks.
                if (o.tip1 || o.tip2) {
                    __syncwarp();                              // every lane is done with the previous op's tables
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int code = 8 * h + site_in_row;
                        const int row = code ^ (code >> 1);
                        if (o.tip1) my_tab[row * 4 + cat] = category_branch<M>(c.L, tv[code]);
                        if (o.tip2) my_tab[64 + row * 4 + cat] = category_branch<M>(c.R, tv[code]);
                    }
                    __syncwarp();
                }
            }
            const size_t s0 = (size_t)st * STAGE + (size_t)warp * TILE;   // first site of this warp's tile
            const size_t s_lane = s0 + lane;                              // the site whose count this lane owns
            const bool complete = st < full_stages;
            const bool lane_live = lane < TILE && (complete || s_lane < n);
            // The rest of the stage, compiled once per (child 1 is a tip, child 2 is a tip): o.tip1 / o.tip2 are
            // warp-uniform run-time values, and left as such the compiler predicates both sides -- every dense stage
            // then also executes the tip path's byte and table reads, and the other way round.
            // tip-tip node with a table: codes -> table rows -> stores; no arithmetic, no vote
            auto tiptip_body = [&]() {
                const uint32_t code_off = slot * STAGE + warp * TILE + site_in_row;
                float4 r[U];
                unsigned dep = g;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned pair = ((k1s[code_off + 8 * u] & 15u) << 4) | (k2s[code_off + 8 * u] & 15u);
                    dep ^= pair << (u & 7);
                    r[u] = __ldg(o.tiptab + pair * 4 + cat);
                }
                unsigned my_pair = 0;
                if (lane_live) my_pair = ((k1s[slot * STAGE + warp * TILE + lane] & 15u) << 4) | (k2s[slot * STAGE + warp * TILE + lane] & 15u);
                dep ^= my_pair << 8;
                release_slot(&empty[slot], lane, dep, flags);
                float4 *out = o.x3 + s0 * 4 + lane;
                if (complete) {
#pragma unroll
                    for (int u = 0; u < U; ++u) st_stream(out + 32 * u, r[u]);
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (s0 + 8 * u + site_in_row < n) st_stream(out + 32 * u, r[u]);
                }
                if (lane_live) {
                    const bool scaled = __ldg(o.tipflag + my_pair) != 0;
                    if (o.scaler) o.scaler[s_lane] = scaled ? 1 : 0;
                    if (o.cnt3) o.cnt3[s_lane] = scaled ? 1 : 0;           // tips carry no counts
                    if (scaled) my_sum += wgt ? (unsigned long long)(long long)wgt[s_lane] : 1ull;
                }
            };
            auto stage_body = [&](auto tip1_c, auto tip2_c) {
                constexpr bool TIP1 = decltype(tip1_c)::value, TIP2 = decltype(tip2_c)::value;
                float4 a[U], b[U], r[U];
                unsigned ballots[U];
                const uint32_t code_off = slot * STAGE + warp * TILE + site_in_row;
                // a[u] / b[u]: the child's CLV slice, or for a tip child already its branch product from the table
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if constexpr (TIP1) {
                        const unsigned code = k1s[code_off + 8 * u] & 15u;
                        a[u] = my_tab[((code ^ (code >> 1)) << 2) + cat];
                    } else {
                        a[u] = t1[32 * u];
                    }
                    if constexpr (TIP2) {
                        const unsigned code = k2s[code_off + 8 * u] & 15u;
                        b[u] = my_tab[64 + ((code ^ (code >> 1)) << 2) + cat];
                    } else {
                        b[u] = t2[32 * u];
                    }
                }
                int cnt = 0;
                if (lane_live) {
                    if (o.cnt1) cnt = c1[slot * STAGE + warp * TILE + lane];
                    if (o.cnt2) cnt += c2[slot * STAGE + warp * TILE + lane];
                }
                unsigned dep = (unsigned)cnt ^ g;
#pragma unroll
                for (int u = 0; u < U; ++u) dep ^= __float_as_uint(a[u].x) ^ __float_as_uint(b[u].w);
                release_slot(&empty[slot], lane, dep, flags);
                float4 *out = o.x3 + s0 * 4 + lane;
                auto finish = [&](int u) {
                    float4 av, bv;
                    if constexpr (TIP1) av = a[u]; else av = category_branch<M>(c.L, a[u]);
                    if constexpr (TIP2) bv = b[u]; else bv = category_branch<M>(c.R, b[u]);
                    return category_finish<M>(c, av, bv, r[u]);
                };
                if (complete) {                    // no bounds predicates on the hot path
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const bool small = finish(u);
                        ballots[u] = __ballot_sync(0xffffffffu, small);
                        if (nibble_all(ballots[u], site_in_row)) rescale(r[u]);
                        st_stream(out + 32 * u, r[u]);
                    }
                } else {                           // the single ragged stage at the end of an op's site range
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const bool live = s0 + 8 * u + site_in_row < n;
                        const bool small = finish(u);
                        ballots[u] = __ballot_sync(0xffffffffu, small && live);
                        if (nibble_all(ballots[u], site_in_row)) rescale(r[u]);
                        if (live) st_stream(out + 32 * u, r[u]);
                    }
                }
                if (lane_live) {
                    unsigned bal = ballots[0];
#pragma unroll
                    for (int u = 1; u < U; ++u) bal = (lane >> 3) == u ? ballots[u] : bal;
                    const bool scaled = nibble_all(bal, lane & 7);
                    if (o.scaler) o.scaler[s_lane] = scaled ? 1 : 0;
                    if (o.cnt3) o.cnt3[s_lane] = cnt + (scaled ? 1 : 0);
                    if (scaled) my_sum += wgt ? (unsigned long long)(long long)wgt[s_lane] : 1ull;
                }
            };
            using yes = cuda_true;
            using no = cuda_false;
            if (o.tip1) {
                if (o.tip2 && o.tiptab) tiptip_body();
                else if (o.tip2) stage_body(yes{}, yes{});
                else stage_body(yes{}, no{});
            } else {
                if (o.tip2) stage_body(no{}, yes{});
                else stage_body(no{}, no{});
            }
            if (++slot == DEPTH) {
                slot = 0;
                phase ^= 1u;
            }
        }
    }
    if (scaler_sum) block_add_u64<THREADS>(my_sum, scaler_sum);
}

// ---------------------------------------------------------------------------------------------
// INPUT_SRC=gen analogue: no CLV is read.  Every lane holds its category's slice of the constant
// site pattern (mm2sleft_genDNAwindowComb.cpp:44-49 / mm2sright_...:45-50); an opaque register
// move per site keeps the compiler from hoisting the (site-invariant) arithmetic out of the loop,
// so the full per-site work is issued exactly as in the MEM kernel.
//   DISCARD = false : write CLV + scaler (cfg4a)      DISCARD = true : checksum only (cfg4b,
//   the sink of s2mm_genDNAwindowComb.cpp:15-53 reads the streams and drops them).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void opaque(float4 &v)
{
    asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w));
}

template <class M, int U, bool DISCARD, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
plf_newview_gen(float4 *__restrict__ x3, unsigned char *__restrict__ scaler,
                const float *__restrict__ pat_x1, const float *__restrict__ pat_x2,
                const float *__restrict__ ev4, const float *__restrict__ pl,
                const float *__restrict__ pr, size_t n,
                unsigned long long *__restrict__ scaler_sum, double *__restrict__ checksum)
{
    constexpr int TILE = 8 * U;
    const int lane = threadIdx.x & 31;
    const int cat = lane & 3;
    const int site_in_row = lane >> 2;

    CatConst c;
    load_cat_const(c, ev4, pl, pr, cat, 1);
    float4 g1 = __ldg(reinterpret_cast<const float4 *>(pat_x1) + cat);
    float4 g2 = __ldg(reinterpret_cast<const float4 *>(pat_x2) + cat);

    const size_t warps_total = (size_t)gridDim.x * (THREADS / 32);
    const size_t warp_id = (size_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const size_t n_tiles = (n + TILE - 1) / TILE;
    unsigned long long my_sum = 0;
    double my_check = 0.0;

    for (size_t tile = warp_id; tile < n_tiles; tile += warps_total) {
        const size_t s0 = tile * TILE;
        unsigned ballots[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t s = s0 + 8 * u + site_in_row;
            const bool live = s < n;
            float4 a = g1, b = g2, o;
            opaque(a);
            opaque(b);
            bool small = category_newview<M>(c, a, b, o);
            ballots[u] = __ballot_sync(0xffffffffu, small && live);
            if (nibble_all(ballots[u], site_in_row)) rescale(o);
            if (DISCARD) {
                if (live) my_check += (double)((o.x + o.y) + (o.z + o.w));
            } else if (live) {
                st_stream(x3 + s * 4 + cat, o);
            }
        }
        if (lane < TILE) {
            unsigned bal = ballots[0];
#pragma unroll
            for (int u = 1; u < U; ++u) bal = (lane >> 3) == u ? ballots[u] : bal;
            const bool scaled = nibble_all(bal, lane & 7);
            const size_t s = s0 + lane;
            if (s < n) {
                if (!DISCARD && scaler) scaler[s] = scaled ? 1 : 0;
                if (scaled) my_sum += 1ull;
            }
        }
    }
    if (scaler_sum) block_add_u64<THREADS>(my_sum, scaler_sum);
    if (DISCARD && checksum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_check += __shfl_xor_sync(0xffffffffu, my_check, o);
        if (lane == 0 && my_check != 0.0) atomicAdd(checksum, my_check);
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic stimulus on the device (bench / INPUT generation for GB-scale configs).
// Counter-based: element e of x1/x2 is a pure function of (seed, global element index), so any
// site range can be produced independently on any GPU and re-derived on the host for checks.
// Distribution follows host_mem.cpp:198-204: U(0,1), left CLV of every 4th site times 1e-12f.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// 24 random bits -> (0,1): (k + 0.5) / 2^24, exactly representable, never 0 or 1.
__host__ __device__ __forceinline__ float u01_from_bits(uint32_t r)
{
    return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
__host__ __device__ __forceinline__ void gen_pair(uint64_t seed, uint64_t elem, float &l, float &r)
{
    const uint64_t h = splitmix64(seed ^ splitmix64(elem));
    l = u01_from_bits((uint32_t)h);
    r = u01_from_bits((uint32_t)(h >> 32));
    if ((elem & 63u) < 16u) l = l * 1.0e-12f;     // j % 64 < 16  (host_mem.cpp:200-202)
}

static __global__ void __launch_bounds__(256)
plf_generate_kernel(float4 *__restrict__ x1, float4 *__restrict__ x2, uint64_t first_elem,
                    size_t n_vec4, uint64_t seed)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec4; v += stride) {
        float4 a, b;
        const uint64_t e = first_elem + 4 * (uint64_t)v;
        gen_pair(seed, e + 0, a.x, b.x);
        gen_pair(seed, e + 1, a.y, b.y);
        gen_pair(seed, e + 2, a.z, b.z);
        gen_pair(seed, e + 3, a.w, b.w);
        x1[v] = a;
        x2[v] = b;
    }
}

}  // namespace plf
