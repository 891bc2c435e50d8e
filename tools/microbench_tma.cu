// tools/microbench_tma.cu -- how fast can one SM's TMA unit FEED a shared-memory ring, by request shape (developer tool).
//
// The tcgen05 20-state kernel loads {20 floats x 128 sites} boxes: 128 rows of 80 B, 320 B apart in global memory.  Is the
// row size what limits its operand feed?  This program runs the kernel's ring alone -- one producer lane, consumers that
// only wait for a slot and hand it back -- over a [n sites x 80 floats] array, with three ways to fill a 10 KB slot:
//
//   box80    cp.async.bulk.tensor.2d, box {20 floats, 128 sites}: 128 requests of 80 B        (what the kernel does)
//   box320   cp.async.bulk.tensor.2d, box {80 floats, 32 sites}:   32 requests of 320 B
//   bulk     cp.async.bulk, 10 240 contiguous bytes                                          (what the 4-state kernel does)
//   rows320  32 cp.async.bulk of 320 B each, into rows 336 B apart (a conflict-free pitch for row-per-lane reads)
//
// and prints GB/s over all SMs for ring depths 4, 8, 16 (slots in flight per CTA).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/microbench_tma tools/microbench_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}

constexpr int kSlot = 10752;          // 32 rows x 336 B >= 10 240

template <int MODE>
__global__ void __launch_bounds__(160, 1) feed(const __grid_constant__ CUtensorMap map80, const __grid_constant__ CUtensorMap map320,
                                               const float *x, size_t n_slots, int depth, unsigned long long *sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem), *empty = full + 16;
    unsigned char *ring = smem + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4) {
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (size_t s = blockIdx.x; s < n_slots; s += gridDim.x) {
                mbar_wait(&empty[slot], phase ^ 1u);
                unsigned char *dst = ring + (size_t)slot * kSlot;
                mbar_expect(&full[slot], 10240);
                // slot s = 10 240 bytes of the array.  box80: tile s/4, category s%4 (128 sites x 80 B);  others: 32 whole sites
                if (MODE == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map80)), "r"((int)(s & 3) * 20), "r"((int)(s >> 2) * 128),
                                    "r"(smem_u32(&full[slot])) : "memory");
                } else if (MODE == 1) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map320)), "r"(0), "r"((int)s * 32), "r"(smem_u32(&full[slot])) : "memory");
                } else if (MODE == 2) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem_u32(dst)), "l"(x + s * 2560), "r"(10240), "r"(smem_u32(&full[slot])) : "memory");
                } else {
                    for (int r = 0; r < 32; ++r)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"(smem_u32(dst + r * 336)), "l"(x + s * 2560 + r * 80), "r"(320), "r"(smem_u32(&full[slot])) : "memory");
                }
                if (++slot == (uint32_t)depth) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        uint32_t slot = 0, phase = 0;
        unsigned long long acc = 0;
        for (size_t s = blockIdx.x; s < n_slots; s += gridDim.x) {
            mbar_wait(&full[slot], phase);
            acc += *reinterpret_cast<const uint32_t *>(ring + (size_t)slot * kSlot + threadIdx.x * 16);      // touch the data
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (++slot == (uint32_t)depth) {
                slot = 0;
                phase ^= 1u;
            }
        }
        if (acc == 0x123456789ull) *sink = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const size_t n = (size_t)4 << 20;                       // 4 Mi sites x 320 B = 1.34 GB
    float *x = nullptr;
    unsigned long long *sink = nullptr;
    cudaMalloc(&x, n * 320);
    cudaMalloc(&sink, 8);
    cudaMemset(x, 0, n * 320);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)p;
    CUtensorMap m80, m320;
    const cuuint64_t dims[2] = {80, n}, strides[1] = {320};
    const cuuint32_t b80[2] = {20, 128}, b320[2] = {80, 32}, el[2] = {1, 1};
    if (enc(&m80, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, b80, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&m320, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, b320, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("tensor map encode failed\n");
        return 1;
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t n_slots = n / 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[4] = {"box80  (tensor, 128 rows of 80 B)", "box320 (tensor, 32 rows of 320 B)", "bulk   (10 240 contiguous bytes)", "rows320 (32 bulk copies of 320 B)"};
    printf("# %d SMs, %zu slots of 10 240 B, read-only feed of a shared-memory ring, GB/s\n", sms, n_slots);
    for (int mode = 0; mode < 4; ++mode)
        for (int depth : {4, 8, 16}) {
            const int smem = 1024 + depth * kSlot;
            auto k = mode == 0 ? feed<0> : mode == 1 ? feed<1> : mode == 2 ? feed<2> : feed<3>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                k<<<sms, 160, smem>>>(m80, m320, x, n_slots, depth, sink);
                cudaEventRecord(e1);
                if (cudaDeviceSynchronize() != cudaSuccess) {
                    printf("%s depth %d: %s\n", names[mode], depth, cudaGetErrorString(cudaGetLastError()));
                    return 1;
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("%-36s ring %2d slots (%3d KB)  %7.1f GB/s\n", names[mode], depth, depth * 10, n * 320 / best / 1e6);
        }
    return 0;
}
