// tools/microbench_sm.cu -- developer microbenchmark (not part of the product): per-SM issue rates that decide
// the shape of the 20-state kernel.  One CTA per SM.  LDS.128 patterns are timed over the whole kernel with CUDA
// events (and every loaded component is consumed: ptxas narrows the load otherwise -- an earlier version of this
// file measured LDS.32 that way); the fp32 mixes likewise.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_sm tools/microbench_sm.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// mode 0: uniform address; 1: 4 distinct 16 B addresses by lane&3 in distinct banks (pitch 404 words);
// 2: 4 distinct addresses, same banks (pitch 400 words = 16 mod 32 -> 2-way); 3: all lanes distinct, conflict-free
__global__ void lds_kernel(int mode, int iters, long long *cycles, float *sink)
{
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int base;
    if (mode == 0) base = 0;
    else if (mode == 1) base = (lane & 3) * 101;
    else if (mode == 2) base = (lane & 3) * 100;
    else if (mode == 3) base = lane;
    else if (mode == 4) base = lane & 3;            // 4 addresses inside one 64-byte segment
    else if (mode == 5) base = lane & 7;            // 8 addresses inside one 128-byte row
    else base = (lane & 1) * 101;                   // 2 addresses, distinct banks
    unsigned x = 0;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float4 v;
            const unsigned addr = sbase + (unsigned)(base + ((it + j) & 15) * 32) * 16u;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            x ^= __float_as_uint(v.x) ^ __float_as_uint(v.y) ^ __float_as_uint(v.z) ^ __float_as_uint(v.w);   // all four: ptxas narrows the load otherwise
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (x == 0x12345678u) sink[0] = 1.0f;
}

// MODE 0: FFMA scalar (16 chains); 1: FFMA2 (16 chains of pairs); 2: FMUL2 + packed add; 3: FMUL2 + 2 scalar FADD;
// 4: FFMA2 with one uniform LDS.128 per 4 FFMA2; 5: FMUL2 + adds split 1/3 packed, 2/3 scalar; 6: FFMA2 with scalar broadcast operand
template <int MODE>
__global__ void fma_kernel(int iters, long long *cycles, float *sink, const float *__restrict__ init)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(init[0], init[1], init[2], init[3]);
    __syncthreads();
    float a[16], b[16];
    unsigned long long p[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = init[i]; b[i] = init[16 + i];
        p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(b[i]);
        q[i] = ((unsigned long long)__float_as_uint(init[32 + i]) << 32) | __float_as_uint(init[48 + i]);
    }
    const float m = init[64];
    const unsigned long long mm = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(init[65]);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, b[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = f2_fma(p[i], mm, q[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = f2_add(p[i], f2_mul(q[i], mm));
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                unsigned long long r = f2_mul(q[i], mm);
                a[i] = __fadd_rn(a[i], __uint_as_float((unsigned)r));
                b[i] = __fadd_rn(b[i], __uint_as_float((unsigned)(r >> 32)));
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float4 v = sm[(it + i) & 63];
                unsigned long long m0 = ((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
                unsigned long long m1 = ((unsigned long long)__float_as_uint(v.w) << 32) | __float_as_uint(v.z);
                p[i] = f2_fma(m0, mm, p[i]);
                p[i + 1] = f2_fma(m1, mm, p[i + 1]);
                p[i + 2] = f2_fma(m0, q[i], p[i + 2]);
                p[i + 3] = f2_fma(m1, q[i + 1], p[i + 3]);
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                unsigned long long r = f2_mul(q[i], mm);
                if (i % 3 == 0) {
                    p[i] = f2_add(p[i], r);
                } else {
                    a[i] = __fadd_rn(a[i], __uint_as_float((unsigned)r));
                    b[i] = __fadd_rn(b[i], __uint_as_float((unsigned)(r >> 32)));
                }
            }
        } else {
            const unsigned long long xx = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = f2_fma(q[i], xx, p[i]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + b[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    if (s == 12345.678f) sink[0] = s;
}

template <int MODE>
void run_fma(const char *name, int warps, int iters, long long *cyc, float *sink, const float *init)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<MODE><<<148, warps * 32>>>(iters, cyc, sink, init);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
    }
    // whole-kernel time: one warp's clock64 span overstates the rate when the arbiter favours that warp
    printf("%-28s warps=%2d : %.2f mul-add/clk/SM (events, 1.965 GHz)\n", name, warps,
           (double)iters * 16 * warps * 32 * (MODE == 0 ? 1 : 2) / (ms * 1e-3 * 1.965e9));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main()
{
    long long *cyc;
    float *sink;
    CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
    CK(cudaMalloc(&sink, 16));
    long long h[1024];
    const int iters = 16384;       // long enough that launch overhead is < 1 %
    CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const char *lname[] = {"uniform", "4 addr distinct banks", "4 addr same banks", "32 distinct", "4 addr in one 64 B", "8 addr in one 128 B", "2 addr distinct banks"};
    // whole-kernel time (CUDA events): a single warp's clock64 span flatters whatever the arbiter favours
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int warps : {8, 16}) {
        for (int mode = 0; mode < 7; ++mode) {
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                lds_kernel<<<148, warps * 32, 65536>>>(mode, iters, cyc, sink);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("LDS.128 %-24s warps=%2d : %.3f clk per warp-instruction per SM (events, 1.965 GHz)\n", lname[mode], warps,
                   ms * 1e-3 * 1.965e9 / ((double)iters * 16 * warps));
        }
    }
    float hinit[128];
    for (int i = 0; i < 128; ++i) hinit[i] = 1.0f + 1e-3f * i;
    float *init;
    CK(cudaMalloc(&init, sizeof hinit));
    CK(cudaMemcpy(init, hinit, sizeof hinit, cudaMemcpyHostToDevice));
    for (int warps : {4, 8, 12, 16}) {
        run_fma<0>("FFMA (scalar)", warps, iters, cyc, sink, init);
        run_fma<1>("FFMA2", warps, iters, cyc, sink, init);
        run_fma<6>("FFMA2 (scalar bcast operand)", warps, iters, cyc, sink, init);
        run_fma<2>("FMUL2 + packed add", warps, iters, cyc, sink, init);
        run_fma<3>("FMUL2 + 2 FADD", warps, iters, cyc, sink, init);
        run_fma<5>("FMUL2 + 1/3 packed 2/3 FADD", warps, iters, cyc, sink, init);
        run_fma<4>("FFMA2 + LDS.128 per 4", warps, iters, cyc, sink, init);
    }
    return 0;
}
