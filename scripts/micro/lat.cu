// Dev tool: dependent-chain latencies (cycles) of the instruction kinds on the critical path of one PDHG phase.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k(double* out, long long* t, double a, double b, int m, const double* g)
{
    __shared__ double sm[1024];
    __shared__ int si[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; si[i] = (i * 7 + 1) & 1023; }
    __syncthreads();
    double x = a; long long c0, c1; int j = threadIdx.x & 1023;
    // 0: DFMA
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    c1 = clock64(); if (threadIdx.x == 0) t[0] = c1 - c0;
    // 1: DADD
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x + b;
    c1 = clock64(); if (threadIdx.x == 0) t[1] = c1 - c0;
    // 2: fmax
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fmax(x * b, a);
    c1 = clock64(); if (threadIdx.x == 0) t[2] = c1 - c0;   // DMUL + DMNMX
    // 3: shfl double + add
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
    c1 = clock64(); if (threadIdx.x == 0) t[3] = c1 - c0;
    // 4: dependent LDS (index chain)
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) j = si[j];
    c1 = clock64(); if (threadIdx.x == 0) t[4] = c1 - c0;
    // 5: int modulo
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) j = (j + 12345) % m;
    c1 = clock64(); if (threadIdx.x == 0) t[5] = c1 - c0;
    // 6: __syncthreads
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) __syncthreads();
    c1 = clock64(); if (threadIdx.x == 0) t[6] = c1 - c0;
    // 7: globaltimer
    unsigned long long gt = 0;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { unsigned long long q; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(q)); gt += q; }
    c1 = clock64(); if (threadIdx.x == 0) t[7] = c1 - c0;
    // 8: dependent global load through L1 (ld.ca) : pointer chase in a small array
    const double* p = g; double acc = 0;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) { double v = __ldca(p + (j & 1023)); j = (int)v; acc += v; }
    c1 = clock64(); if (threadIdx.x == 0) t[8] = c1 - c0;
    // 9: dependent global load through L2 (ld.cg)
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) { double v = __ldcg(p + (j & 1023)); j = (int)v; acc += v; }
    c1 = clock64(); if (threadIdx.x == 0) t[9] = c1 - c0;
    // 10: store then load same address through L2 (st + ld.cg) dependent
    double* w = const_cast<double*>(g) + 2048 + threadIdx.x;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) { *w = acc; acc = __ldcg(w) + 1.0; }
    c1 = clock64(); if (threadIdx.x == 0) t[10] = c1 - c0;
    // 11: DFMA with 4 independent chains (throughput-ish per warp)
    double y0 = a, y1 = b, y2 = a + 1, y3 = b + 1;
    c0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { y0 = fma(y0, b, a); y1 = fma(y1, b, a); y2 = fma(y2, b, a); y3 = fma(y3, b, a); }
    c1 = clock64(); if (threadIdx.x == 0) t[11] = c1 - c0;
    out[threadIdx.x] = x + j + gt + acc + y0 + y1 + y2 + y3;
}
int main()
{
    double* out; long long* t; double* g;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&t, 8 * 16); cudaMalloc(&g, 8 * 8192);
    double h[8192]; for (int i = 0; i < 8192; ++i) h[i] = (double)((i * 13 + 5) & 1023);
    cudaMemcpy(g, h, sizeof(h), cudaMemcpyHostToDevice);
    const char* names[] = {"DFMA", "DADD", "DMUL+fmax", "shfl64+DADD", "LDS chain", "int %", "__syncthreads", "globaltimer", "LDG.ca chain", "LDG.cg chain", "ST+LDG.cg", "DFMA x4 indep"};
    for (int threads : {32, 128, 1024}) {
        k<<<1, threads>>>(out, t, 1.0000001, 0.9999999, 977, g);
        k<<<1, threads>>>(out, t, 1.0000001, 0.9999999, 977, g);
        cudaDeviceSynchronize();
        long long ht[16]; cudaMemcpy(ht, t, sizeof(ht), cudaMemcpyDeviceToHost);
        printf("threads %d:", threads);
        for (int i = 0; i < 12; ++i) printf("  %s %.1f", names[i], (double)ht[i] / N);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
