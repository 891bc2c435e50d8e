// tools/microbench_const.cu -- developer microbenchmark: fp32 multiply-add rate when the matrix operand comes from the
// constant bank (warp-uniform, compile-time offsets) while streaming through a 14.4 KB footprint, as a 20-state
// newview with lane = site would.  Whole-kernel time by CUDA events (not one warp's clock).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__constant__ float cm[2 * 9 * 400];

__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long pk(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

template <int MODE, int T, int THREADS>   // MODE 0: FFMA2 with constant pair; 1: scalar FFMA with constant.  T sites per lane share each pair.
__global__ void __launch_bounds__(THREADS) k(int iters, float *out, const float *__restrict__ in, int coff)
{
    const float *cmo = cm + coff;   // runtime-uniform offset: LDCU c[3][UR+imm] when coff is not a compile-time constant
    float x[T][20];
    float acc[T][20];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int l = 0; l < 20; ++l) { x[t][l] = in[(threadIdx.x + l + t) & 63]; acc[t][l] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int st = 0; st < 9; ++st) {
#pragma unroll
            for (int l = 0; l < 20; ++l) {
                if (MODE == 0) {
#pragma unroll
                    for (int kp = 0; kp < 10; ++kp) {
                        const unsigned long long m = pk(cmo[st * 400 + l * 20 + 2 * kp], cmo[st * 400 + l * 20 + 2 * kp + 1]);
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            unsigned long long a = pk(acc[t][2 * kp], acc[t][2 * kp + 1]);
                            a = f2_fma(m, pk(x[t][l], x[t][l]), a);
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[t][2 * kp]), "=f"(acc[t][2 * kp + 1]) : "l"(a));
                        }
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < 20; ++kk)
#pragma unroll
                        for (int t = 0; t < T; ++t) acc[t][kk] = fmaf(cmo[st * 400 + l * 20 + kk], x[t][l], acc[t][kk]);
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
                for (int l = 0; l < 20; ++l) x[t][l] = acc[t][l] * 1e-3f;     // next stage consumes this one (like p -> x3)
        }
    }
    float s = 0;
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int i = 0; i < 20; ++i) s += acc[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int T, int THREADS>
void run(int iters, float *out, const float *in, int coff = 0)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE, T, THREADS><<<148, THREADS>>>(iters, out, in, coff);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)iters * 9 * 400 * T * THREADS * 148;
    printf("%s coff=%d T=%d threads=%4d: %.3f ms  %.2f T mul-add/s  (%.1f per clk per SM at 1.965 GHz)\n", MODE ? "FFMA  UR" : "FFMA2 UR", coff, T,
           THREADS, ms, fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    float h[7200];
    for (int i = 0; i < 7200; ++i) h[i] = 1e-2f * (i % 97);
    CK(cudaMemcpyToSymbol(cm, h, sizeof h));
    float *out, *in;
    CK(cudaMalloc(&out, 148 * 1024 * 4));
    CK(cudaMalloc(&in, 256));
    CK(cudaMemset(in, 0, 256));
    const int iters = 100;
    run<0, 2, 512>(iters, out, in, 3600);
    run<0, 1, 1024>(iters, out, in, 3600);
    run<0, 1, 256>(iters, out, in);
    run<0, 1, 512>(iters, out, in);
    run<0, 1, 768>(iters, out, in);
    run<0, 1, 1024>(iters, out, in);
    run<0, 2, 256>(iters, out, in);
    run<0, 2, 384>(iters, out, in);
    run<0, 2, 512>(iters, out, in);
    run<0, 2, 640>(iters, out, in);
    run<0, 4, 128>(iters, out, in);
    run<0, 4, 256>(iters, out, in);
    run<1, 1, 1024>(iters, out, in);
    run<1, 2, 512>(iters, out, in);
    return 0;
}
