// Issue-rate microbenchmarks for the instruction mix of the nearest-hit scan (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
// Each test: every warp runs ITERS iterations of an unrolled body of independent chains; reports warp-instructions
// per clock per SM (4 schedulers => 4.0 is the issue limit).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

__device__ __forceinline__ unsigned long long pack2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo2(unsigned long long v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int KIND>
__global__ void __launch_bounds__(256) bench(float *out, long long *cycles, float seed, const float4 *gsrc)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed * 2, seed * 3, seed * 4);
    __syncthreads();
    float a[CHAINS], b = seed, c = seed * 0.5f;
    unsigned long long p[CHAINS], pb = pack2(seed, seed), pc = pack2(c, c);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + i; p[i] = pack2(seed + i, seed - i); }
    const float4 *gen = (seed > 100.f) ? gsrc : sm;       // generic pointer (really shared)
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
            if (KIND == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (KIND == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (KIND == 4) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                             asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); }
            if (KIND == 5) { float4 v = sm[(it + i) & 63]; a[i] += v.x + v.w; }                 // LDS.128 broadcast + 2 FADD
            if (KIND == 6) { float4 v = gen[(it + i) & 63]; a[i] += v.x + v.w; }                // LD.E.128 generic + 2 FADD
            if (KIND == 7) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                             asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); }
            if (KIND == 8) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (KIND == 9) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (KIND == 10) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (KIND == 11) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + lo2(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, int per_iter, int ctas_per_sm, float *out, long long *cyc, const float4 *g)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    bench<KIND><<<grid, 256>>>(out, cyc, 1.0f, g);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KIND><<<grid, 256>>>(out, cyc, 1.0f, g);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double winst = (double)ITERS * CHAINS * per_iter * 8 /*warps per CTA*/ * ctas_per_sm;
    printf("%-34s ctas/SM=%d  warp-instr/clk/SM = %.3f   (%.3f ms, %s)\n", name, ctas_per_sm, winst / avg, ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out; long long *cyc; float4 *g;
    cudaMalloc(&out, 4 * 256 * 2048); cudaMalloc(&cyc, 8 * 2048); cudaMalloc(&g, 64 * 16);
    for (int c = 1; c <= 4; c *= 2) {
        run<0>("FFMA", 1, c, out, cyc, g);
        run<1>("FFMA2 (fma.rn.f32x2)", 1, c, out, cyc, g);
        run<8>("FMUL", 1, c, out, cyc, g);
        run<11>("FMUL2", 1, c, out, cyc, g);
        run<9>("FADD", 1, c, out, cyc, g);
        run<10>("FADD2", 1, c, out, cyc, g);
        run<2>("FMNMX", 1, c, out, cyc, g);
        run<3>("FMNMX3", 1, c, out, cyc, g);
        run<4>("FFMA + FMNMX", 2, c, out, cyc, g);
        run<7>("FFMA2 + FMNMX", 2, c, out, cyc, g);
        run<5>("LDS.128 bcast + 2 FADD", 3, c, out, cyc, g);
        run<6>("LD.E.128 generic(smem) + 2 FADD", 3, c, out, cyc, g);
    }
    return 0;
}
