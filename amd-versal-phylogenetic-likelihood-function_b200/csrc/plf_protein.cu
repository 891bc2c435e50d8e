// csrc/plf_protein.cu -- 20-state (protein) newview, SURVEY.md section 8f.3: the reference's STATES knob.
//
// The reference implements DNA only ("could be implemented for any other type of data with more or fewer
// states", README.md:36; STATES knob README.md:67; open to-do "Implement protein-based PLF", README.md:202).
// This is plf() (app/src/plf.cpp:19-65) with the state count 4 -> 20 and nothing else changed:
//
//   a[k]      = sum_l x1[i,j,l] * left [j,k,l]        l ascending, fp32, each product and sum rounded
//   b[k]      = sum_l x2[i,j,l] * right[j,k,l]
//   p[k]      = a[k] * b[k]
//   x3[i,j,l] = sum_k p[k] * EV[k,l]                  k ascending
//   if all 80 |x3[i,*]| < 2^-32:  x3[i,*] *= 2^32, scaler byte = 1, scalerIncrement += wgt[i]
//
// Layouts: x1,x2,x3 float[n*80] [site][category j][state l]; left/right float[4*400] [j][k][l]; EV float[400] [k][l].
//
// Roofline.  961 algorithmic bytes and 4800 multiply-adds per site: 5.0 mul-add per byte, against a B200 balance
// of 116 mul-add/clk/SM * 148 SMs * 1.965 GHz / 7.1 TB/s = 4.8 (tools/microbench_sm.cu, profiles/r01_protein.md).
// Unlike the DNA kernel this one sits ON the ridge: the fp32 pipe and HBM saturate together at about 7 G sites/s.
// Tensor cores stay unused: tf32 inputs (10-bit mantissa) cannot meet the 1e-5 tolerance, and the 3xTF32 split
// would need the operands re-packed through shared memory twice per site.
//
// Shape.  Lane = (site s = lane>>2, category c = lane&3) as in the DNA kernels, T sites per lane, so a warp owns a
// tile of 8T consecutive sites and the rescale vote is one ballot per row.  The register tile is 20 outputs x T sites:
// every matrix word fetched from shared memory feeds T multiply-adds.  That ratio is what bounds the kernel: an
// LDS.128 always costs 4 wavefronts of the 128 B/clk shared-memory pipe (512 B of register write-back, broadcast or
// not -- ncu "L1 Wavefronts Shared Ideal"), so a site costs (300 + 10T)*4/(8T) wavefronts: 68 at T = 2, 42.5 at
// T = 4, against 41 clk of fp32 pipe.  T = 4 (232 registers, 8 warps) is the FMA default, T = 2 (12 warps) the strict one.
// Every warp runs its own pipeline on one tile buffer: lane 0 bulk-copies (TMA, cp.async.bulk + mbarrier) the x1
// half of the NEXT tile as soon as the left-branch products are done and the x2 half after the right-branch
// products -- no block-wide synchronisation after the prologue.  The three matrices live in shared memory (P
// transposed to [l][k], categories 404 floats apart so that the four categories of a quarter-warp hit disjoint bank
// groups).  Arithmetic is packed fp32x2 (FFMA2; FMUL2 + scalar FADD in strict mode): pairs
// run over k (branch products) and over l (back-transform), the per-site scalar enters as the broadcast operand.
#include "../../include/b200plf.h"
#include "plf_kernels.cuh"
#include "plf_registry.h"

namespace plf {

constexpr int kAaStates = 20;
constexpr int kAaSite = 4 * kAaStates;            // floats per site (320 B)
constexpr int kAaMat = kAaStates * kAaStates;     // 400
constexpr int kAaMatPitch = kAaMat + 4;           // 404 = 20 mod 32: categories land in disjoint 4-bank groups

// Arithmetic policies: acc <- acc + m * x on packed pairs.  Every sum starts from a zeroed accumulator, exactly like
// the reference (plf.cpp:25-27,32-33), so there is no special first term and the chunk loop of the two branch stages
// stays rolled: half the code per tile body (22 KB instead of 43 KB), which removed the instruction-fetch stalls of
// the T = 4 kernel (ncu no_instruction 0.42 warps per issue, +15 %).
//
// STRICT: every product and every sum rounded on its own, in the reference's order.  Products are packed (FMUL2);
// the sums are scalar __fadd_rn on the two halves, because ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one
// FFMA2 (seen in the SASS) and __fadd_rn forbids that.  Three instructions per pair instead of one: strict mode is
// bound by the fp32 pipe and the issue slots (tools/microbench_sm.cu: 97 multiply-adds per clk per SM for this mix
// against 116 for FFMA2).  Moving 4 of every 10 sums to prod * 1.0 + acc (run-time 1.0, FFMA2) kept the bits and
// changed nothing.
struct AaStrict {
    static __device__ __forceinline__ unsigned long long mac(unsigned long long m, unsigned long long xx, unsigned long long acc)
    {
        float a0, a1, p0, p1;
        f2_unpack(acc, a0, a1);
        f2_unpack(f2_mul(m, xx), p0, p1);
        return f2_pack(__fadd_rn(a0, p0), __fadd_rn(a1, p1));
    }
};
// FMA: one FFMA2 per pair; fma(m, x, +0) == m * x up to the sign of a zero, which the tolerance mode does not pin.
struct AaFma {
    static __device__ __forceinline__ unsigned long long mac(unsigned long long m, unsigned long long xx, unsigned long long acc)
    {
        return f2_fma(m, xx, acc);
    }
};

// acc[t][kp] = (a[2kp], a[2kp+1]) of site row t:  a[k] = sum_l x[l] * P[k][l], P^T in shared memory as [l][k].
// Returns an XOR over one register of every shared-memory read of the x tile: a value that exists only once all of
// those reads have been performed (see the refill in the kernel).
template <class M, int T>
__device__ __forceinline__ unsigned aa_branch(const float *__restrict__ xs, const float *__restrict__ mat_t, int s, int c,
                                              unsigned long long (&acc)[T][10])
{
    const float4 *m4 = reinterpret_cast<const float4 *>(mat_t);
    unsigned dep = 0;
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int kp = 0; kp < 10; ++kp) acc[t][kp] = 0ull;      // bits of (+0.0f, +0.0f)
#pragma unroll 1
    for (int q = 0; q < 5; ++q) {
        float4 xv[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            xv[t] = *reinterpret_cast<const float4 *>(xs + (8 * t + s) * kAaSite + c * kAaStates + 4 * q);
            dep ^= __float_as_uint(xv[t].x);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int l = 4 * q + j;
            unsigned long long m[10];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float4 v = m4[l * 5 + i];
                m[2 * i] = f2_pack(v.x, v.y);
                m[2 * i + 1] = f2_pack(v.z, v.w);
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const float x = j == 0 ? xv[t].x : j == 1 ? xv[t].y : j == 2 ? xv[t].z : xv[t].w;
                const unsigned long long xx = f2_pack(x, x);
#pragma unroll
                for (int kp = 0; kp < 10; ++kp) acc[t][kp] = M::mac(m[kp], xx, acc[t][kp]);
            }
        }
    }
    return dep;
}

// out[t][lp] = (x3[2lp], x3[2lp+1]):  x3[l] = sum_k p[k] * EV[k][l], EV in shared memory as [k][l].
template <class M, int T>
__device__ __forceinline__ void aa_backtransform(const float *__restrict__ ev_s, const unsigned long long (&p)[T][10],
                                                 unsigned long long (&out)[T][10])
{
    const float4 *e4 = reinterpret_cast<const float4 *>(ev_s);
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int lp = 0; lp < 10; ++lp) out[t][lp] = 0ull;
    // k runs in pairs (one packed p register each); unrolled: p is indexed by k and must stay in registers
#pragma unroll
    for (int k2 = 0; k2 < kAaStates / 2; ++k2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 2 * k2 + h;
            unsigned long long e[10];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float4 v = e4[k * 5 + i];
                e[2 * i] = f2_pack(v.x, v.y);
                e[2 * i + 1] = f2_pack(v.z, v.w);
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                float lo, hi;
                f2_unpack(p[t][k2], lo, hi);
                const float pk = h ? hi : lo;
                const unsigned long long pp = f2_pack(pk, pk);
#pragma unroll
                for (int lp = 0; lp < 10; ++lp) out[t][lp] = M::mac(e[lp], pp, out[t][lp]);
            }
        }
    }
}

// The three matrices are DEVICE arrays in the reference's layouts (EV [k][l], P [category][k][l]): the head of an
// instance's packed buffers, a tree's per-node matrices, or the stream's staging record for host-matrix callers.  Every
// block copies them into shared memory once (P transposed to [l][k] on the way): 14.4 KB per block, L2-resident.

__host__ __device__ constexpr size_t aa_bar_offset() { return (size_t)(8 * kAaMatPitch + kAaMat) * sizeof(float); }           // 14528
__host__ __device__ constexpr size_t aa_tile_offset(int warps) { return (aa_bar_offset() + (size_t)warps * 2 * sizeof(uint64_t) + 127) & ~(size_t)127; }
__host__ __device__ constexpr size_t aa_smem_bytes(int t, int warps)
{
    return aa_tile_offset(warps) + (size_t)warps * 2 /*children*/ * (size_t)(8 * t) * kAaSite * sizeof(float);
}

template <class M, int T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
plf_newview_aa(const float *__restrict__ ev, const float *__restrict__ pl, const float *__restrict__ pr,
               const float *__restrict__ x1, const float *__restrict__ x2,
               float *__restrict__ x3, unsigned char *__restrict__ scaler, const int *__restrict__ wgt, size_t n,
               unsigned long long *__restrict__ scaler_sum, int flags, const int *__restrict__ cnt1,
               const int *__restrict__ cnt2, int *__restrict__ cnt3)
{
    constexpr int TILE = 8 * T;                                   // sites per warp round
    constexpr int TILE_FLOATS = TILE * kAaSite;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_pl = reinterpret_cast<float *>(smem_raw);
    float *s_pr = s_pl + 4 * kAaMatPitch;
    float *s_ev = s_pr + 4 * kAaMatPitch;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + aa_bar_offset());
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = lane & 3, s = lane >> 2;
    float *my_x1 = reinterpret_cast<float *>(smem_raw + aa_tile_offset(WARPS)) + (size_t)warp * 2 * TILE_FLOATS;
    float *my_x2 = my_x1 + TILE_FLOATS;
    uint64_t *bar1 = bars + 2 * warp, *bar2 = bar1 + 1;

    // prologue: matrices from global into shared memory, P transposed to [l][k]; barriers
    for (int idx = threadIdx.x; idx < 4 * kAaMat; idx += WARPS * 32) {
        const int j = idx / kAaMat, rem = idx - j * kAaMat, k = rem / kAaStates, l = rem - k * kAaStates;
        s_pl[j * kAaMatPitch + l * kAaStates + k] = __ldg(pl + idx);
        s_pr[j * kAaMatPitch + l * kAaStates + k] = __ldg(pr + idx);
    }
    for (int idx = threadIdx.x; idx < kAaMat; idx += WARPS * 32) s_ev[idx] = __ldg(ev + idx);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * WARPS; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const size_t n_tiles = (n + TILE - 1) / TILE;
    const size_t stride = (size_t)gridDim.x * WARPS;
    size_t tile = (size_t)blockIdx.x * WARPS + warp;
    unsigned long long my_sum = 0;

    // One tile buffer per warp, refilled EARLY: the x1 half as soon as the left-branch products of the current tile are
    // done, the x2 half after the right-branch products, so each copy has two thirds of a tile time to land.  The
    // refill must not start before every read of the half has been performed; `dep` (an XOR over a register of
    // every such read) exists only then, and the never-true comparison makes the copy wait for it.
    // With kFlagFencedRelease (the default) every lane first orders its own reads of the half before later
    // async-proxy writes with fence.proxy.async -- the release as the PTX memory model words it.
    auto fetch = [&](const float *src, float *dst, uint64_t *bar, size_t tl, unsigned dep) {
        if (flags & kFlagFencedRelease) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            const size_t s0 = tl * TILE;
            const uint32_t sites = (uint32_t)(n - s0 < (size_t)TILE ? n - s0 : (size_t)TILE);
            const uint32_t bytes = sites * kAaSite * (uint32_t)sizeof(float);
            if (dep == 0x9E3779B9u) fence_proxy_async_smem();
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(dst, src + s0 * kAaSite, bytes, bar);
        }
    };
    if (tile < n_tiles) {
        fetch(x1, my_x1, bar1, tile, 0u);
        fetch(x2, my_x2, bar2, tile, 0u);
    }

    for (unsigned it = 0; tile < n_tiles; tile += stride, ++it) {
        const bool more = tile + stride < n_tiles;
        unsigned long long a[T][10], b[T][10];
        mbar_wait(bar1, it & 1u);
        const unsigned dep1 = aa_branch<M, T>(my_x1, s_pl + c * kAaMatPitch, s, c, a);
        if (more) fetch(x1, my_x1, bar1, tile + stride, dep1);
        mbar_wait(bar2, it & 1u);
        const unsigned dep2 = aa_branch<M, T>(my_x2, s_pr + c * kAaMatPitch, s, c, b);
        if (more) fetch(x2, my_x2, bar2, tile + stride, dep2);
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int kp = 0; kp < 10; ++kp) a[t][kp] = f2_mul(a[t][kp], b[t][kp]);
        aa_backtransform<M, T>(s_ev, a, b);         // b now holds x3

        const size_t s0 = tile * TILE;
        unsigned ballots[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const size_t site = s0 + 8 * t + s;
            const bool live = site < n;
            bool small = true;
            float o[kAaStates];
#pragma unroll
            for (int lp = 0; lp < 10; ++lp) {
                f2_unpack(b[t][lp], o[2 * lp], o[2 * lp + 1]);
                small = small && (fabsf(o[2 * lp]) < kMinLikelihood) && (fabsf(o[2 * lp + 1]) < kMinLikelihood);
            }
            ballots[t] = __ballot_sync(0xffffffffu, small && live);
            if (nibble_all(ballots[t], s)) {
#pragma unroll
                for (int l = 0; l < kAaStates; ++l) o[l] = __fmul_rn(o[l], kTwoToThe32);
            }
            if (live) {
                float4 *dst = reinterpret_cast<float4 *>(x3 + site * kAaSite + c * kAaStates);
#pragma unroll
                for (int q = 0; q < 5; ++q) st_stream(dst + q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
            }
        }
        if (lane < TILE) {
            unsigned bal = ballots[0];
#pragma unroll
            for (int t = 1; t < T; ++t) bal = (lane >> 3) == t ? ballots[t] : bal;
            const bool scaled = nibble_all(bal, lane & 7);
            const size_t site = s0 + lane;
            if (site < n) {
                if (scaler) scaler[site] = scaled ? 1 : 0;
                if (cnt3) cnt3[site] = (cnt1 ? __ldg(cnt1 + site) : 0) + (cnt2 ? __ldg(cnt2 + site) : 0) + (scaled ? 1 : 0);
                if (scaled) my_sum += wgt ? (unsigned long long)(long long)wgt[site] : 1ull;
            }
        }
    }
    if (scaler_sum) block_add_u64<WARPS * 32>(my_sum, scaler_sum);
}

// A second kernel family was built and measured, and removed again: lane = site, warp = category, the matrices
// read through the constant bank into uniform registers (FFMA2 R, R.F32, UR.F32x2, R), which takes the matrix
// traffic off the shared-memory pipe altogether.  In isolation the uniform path sustains 104-110 multiply-adds per
// clk per SM, but only with immediate constant offsets (UR-indexed LDCU collapses to 1-13 per clk) and 16-32 warps
// per SM to cover ~27 clk per LDCU per warp; a lane = site mapping then needs 640 B of staged CLV per site in flight
// (four warps sharing a tile, block-wide barriers, four copies of the code), and end to end it reached 3.1 G sites/s
// against 4.3 (now 5.0) for the register-tile kernel above.  tools/microbench_const.cu, profiles/r01_protein.md.

// ---- stimulus for S states: the recipe of host_mem.cpp:198-204 carried over (every 4th site has a tiny left child,
// so exactly ceil(n/4) sites rescale).  The 20-state sums are 25x larger than the 4-state ones, hence 1e-14 instead
// of the reference's 1e-12 (which is kept for S = 4). --------------------------------------------------------------
__host__ __device__ __forceinline__ void gen_pair_states(uint64_t seed, uint64_t elem, unsigned site_floats, float &l, float &r)
{
    const uint64_t h = splitmix64(seed ^ splitmix64(elem));
    l = u01_from_bits((uint32_t)h);
    r = u01_from_bits((uint32_t)(h >> 32));
    if (((elem / site_floats) & 3u) == 0u) l = l * (site_floats == 16u ? 1.0e-12f : 1.0e-14f);
}

static __global__ void __launch_bounds__(256)
plf_generate_states_kernel(float4 *__restrict__ x1, float4 *__restrict__ x2, uint64_t first_elem, size_t n_vec4,
                           unsigned site_floats, uint64_t seed)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec4; v += stride) {
        float4 a, b;
        const uint64_t e = first_elem + 4 * (uint64_t)v;
        gen_pair_states(seed, e + 0, site_floats, a.x, b.x);
        gen_pair_states(seed, e + 1, site_floats, a.y, b.y);
        gen_pair_states(seed, e + 2, site_floats, a.z, b.z);
        gen_pair_states(seed, e + 3, site_floats, a.w, b.w);
        x1[v] = a;
        x2[v] = b;
    }
}

namespace {

using AaFn = void (*)(const float *, const float *, const float *, const float *, const float *, float *, unsigned char *,
                      const int *, size_t, unsigned long long *, int, const int *, const int *, int *);
struct AaSel {
    AaFn fn = nullptr;
    int t = 0, warps = 0;
};

template <class M>
AaSel aa_pick(int t, int warps)
{
    if (t == 1 && warps == 16) return {plf_newview_aa<M, 1, 16>, 1, 16};
    if (t == 2 && warps == 8) return {plf_newview_aa<M, 2, 8>, 2, 8};
    if (t == 2 && warps == 12) return {plf_newview_aa<M, 2, 12>, 2, 12};
    if (t == 4 && warps == 4) return {plf_newview_aa<M, 4, 4>, 4, 4};
    if (t == 4 && warps == 8) return {plf_newview_aa<M, 4, 8>, 4, 8};
    return {};
}

// variant = sites per lane (1, 2, 4); 0 = the fastest measured shape for the arithmetic mode: 4 x 8 warps with FMA
// (255 registers), 2 x 12 warps in strict mode (the scalar adds need the registers and the issue slots).
AaSel aa_select(int math, int variant, int threads)
{
    const bool fma = math == PLF_MATH_FMA;
    const int t = variant ? variant : (fma ? 4 : 2);
    if (threads % 32) return {};
    const int warps = threads ? threads / 32 : (t == 1 ? 16 : t == 2 ? 12 : 8);
    return fma ? aa_pick<AaFma>(t, warps) : aa_pick<AaStrict>(t, warps);
}

}  // namespace

int aa_kernel_info(int math, int variant, int threads, int *regs, int *block_threads, size_t *smem, int *tile_sites)
{
    if (variant == kAaVariantTensorCore || (variant == 0 && threads == 0 && math == PLF_MATH_FMA))      // FMA default for long calls
        return math == PLF_MATH_FMA && (threads == 0 || threads == 384) ? aa_tc_kernel_info(regs, block_threads, smem, tile_sites)
                                                                        : PLF_ERR_INVALID;
    const AaSel k = aa_select(math, variant, threads);
    if (!k.fn) return PLF_ERR_INVALID;
    cudaFuncAttributes attr;
    if (cudaFuncGetAttributes(&attr, k.fn) != cudaSuccess) return PLF_ERR_CUDA;
    if (regs) *regs = attr.numRegs;
    if (block_threads) *block_threads = k.warps * 32;
    if (smem) *smem = aa_smem_bytes(k.t, k.warps);
    if (tile_sites) *tile_sites = 8 * k.t;
    return PLF_OK;
}

// ev / pl / pr are DEVICE arrays in the reference's layouts; cnt1 / cnt2 / cnt3 are optional per-site scaler counts of
// the children and of the result (chained newview over a tree), NULL otherwise.
int launch_newview_aa(const float *x1, const float *x2, float *x3, unsigned char *scaler, const float *ev,
                      const float *pl, const float *pr, const int *wgt, size_t n, unsigned long long *scaler_sum,
                      int math, int variant, int threads, int flags, cudaStream_t stream, const int *cnt1, const int *cnt2,
                      int *cnt3)
{
    // variant 9: the tensor-core kernel.  3xTF32 is fp32-class but not the reference's rounding sequence: FMA mode only.
    // It is also what variant 0 means in FMA mode from kAaTensorCoreMinSites sites on (5.4 against 4.7 G sites/s for the
    // CUDA-core FMA kernel on B200); shorter calls keep the CUDA-core kernel, whose prologue is lighter.
    if (variant == kAaVariantTensorCore || (variant == 0 && threads == 0 && math == PLF_MATH_FMA && n >= kAaTensorCoreMinSites)) {
        if (math != PLF_MATH_FMA || (threads != 0 && threads != 384)) return PLF_ERR_INVALID;
        // slot release: the tensor-core kernel is not DRAM-bound and the fence costs it 3 % (5.70 against 5.88 G sites/s),
        // so its own default is the data-dependency release; PLF_SAFE_RELEASE / plf_set_release_mode / opts->flags override
        if (flags & kAaReleaseUnset) flags = (flags & ~(kAaReleaseUnset | kFlagFencedRelease)) | (fenced_release(false) ? kFlagFencedRelease : 0);
        return launch_newview_aa_tc(x1, x2, x3, scaler, ev, pl, pr, wgt, n, scaler_sum, flags, stream, cnt1, cnt2, cnt3);
    }
    if (flags & kAaReleaseUnset) flags = (flags & ~(kAaReleaseUnset | kFlagFencedRelease)) | (fenced_release(true) ? kFlagFencedRelease : 0);
    const AaSel k = aa_select(math, variant, threads);
    if (!k.fn) return PLF_ERR_INVALID;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return PLF_ERR_CUDA;
    if (n == 0) return PLF_OK;
    const size_t smem = aa_smem_bytes(k.t, k.warps);
    if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PLF_ERR_CUDA;
    const size_t tiles = (n + 8 * k.t - 1) / (8 * k.t);
    size_t grid = (tiles + k.warps - 1) / k.warps;
    if (grid > (size_t)sms) grid = sms;
    if (flags & kAaSingleCta) grid = 1;
    k.fn<<<(int)grid, k.warps * 32, smem, stream>>>(ev, pl, pr, x1, x2, x3, scaler, wgt, n, scaler_sum, flags & kFlagFencedRelease,
                                                    cnt1, cnt2, cnt3);
    count_launches(1);
    return cudaGetLastError() == cudaSuccess ? PLF_OK : PLF_ERR_CUDA;
}

int launch_generate_states(int states, float *x1, float *x2, size_t first_site, size_t n, uint64_t seed, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return PLF_ERR_CUDA;
    const unsigned site_floats = 4u * (unsigned)states;
    const size_t n_vec4 = n * site_floats / 4;
    size_t grid = (n_vec4 + 255) / 256;
    if (grid > (size_t)sms * 8) grid = (size_t)sms * 8;
    if (grid == 0) return PLF_OK;
    plf_generate_states_kernel<<<(int)grid, 256, 0, stream>>>(reinterpret_cast<float4 *>(x1), reinterpret_cast<float4 *>(x2),
                                                             (uint64_t)first_site * site_floats, n_vec4, site_floats, seed);
    count_launches(1);
    return cudaGetLastError() == cudaSuccess ? PLF_OK : PLF_ERR_CUDA;
}

void generate_states_host(int states, float *x1, float *x2, size_t first_site, size_t n, uint64_t seed)
{
    const unsigned site_floats = 4u * (unsigned)states;
    for (size_t e = 0; e < n * site_floats; ++e)
        gen_pair_states(seed, (uint64_t)first_site * site_floats + e, site_floats, x1[e], x2[e]);
}

}  // namespace plf
