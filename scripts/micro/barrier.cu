// Dev tool: cost of one grid-wide phase boundary on a cooperative grid of one 1024-thread CTA per SM, for several
// barrier constructions.  Every CTA stores a few doubles (like a PDHG phase does), meets, and reads a neighbour's
// doubles (checks that the barrier orders the stores).  Prints ns per barrier.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o barrier barrier.cu && ./barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void v_counter(unsigned* c, unsigned& target, int sleep_ns)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(1u) : "memory");
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if ((int)(v - target) >= 0) break;
            if (sleep_ns) __nanosleep(sleep_ns);
        }
    }
    __syncthreads();
}
// fence + relaxed arrive, relaxed poll, fence after
__device__ __forceinline__ void v_relaxed(unsigned* c, unsigned& target, int sleep_ns)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(1u) : "memory");
        unsigned v;
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if ((int)(v - target) >= 0) break;
            if (sleep_ns) __nanosleep(sleep_ns);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
// one flag per CTA (distinct words, 148 words = 5 lines); warp 0 polls all flags
__device__ __forceinline__ void v_flags(unsigned* flags, unsigned& epoch, int sleep_ns)
{
    __syncthreads();
    ++epoch;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
        }
        const int G = gridDim.x;
        for (;;) {
            bool ok = true;
            for (int g = threadIdx.x; g < G; g += 32) {
                unsigned v;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + g) : "memory");
                ok &= (int)(v - epoch) >= 0;
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (sleep_ns) __nanosleep(sleep_ns);
        }
        if (threadIdx.x == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
// counter sharded over S lines: CTA g arrives on shard g % S, thread 0..S-1 poll one shard each
template <int S>
__device__ __forceinline__ void v_sharded(unsigned* c, unsigned& epoch, int sleep_ns)
{
    __syncthreads();
    ++epoch;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(c + 32 * (blockIdx.x % S)), "r"(1u) : "memory");
        }
        const unsigned G = gridDim.x;
        // shard s receives ceil((G - s) / S) arrivals per barrier
        const unsigned want = threadIdx.x < S ? ((G - threadIdx.x + S - 1) / S) * epoch : 0u;
        for (;;) {
            unsigned v = want;
            if (threadIdx.x < S) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c + 32 * threadIdx.x) : "memory");
            if (__all_sync(0xffffffffu, (int)(v - want) >= 0)) break;
            if (sleep_ns) __nanosleep(sleep_ns);
        }
        if (threadIdx.x == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
// clusters of C CTAs: hardware cluster barrier inside, the rank-0 CTA of every cluster meets the others on the counter
__device__ __forceinline__ void v_cluster(unsigned* c, unsigned& target, int sleep_ns, unsigned nclusters, unsigned crank)
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (crank == 0 && threadIdx.x == 0) {
        target += nclusters;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(1u) : "memory");
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if ((int)(v - target) >= 0) break;
            if (sleep_ns) __nanosleep(sleep_ns);
        }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int V>
__global__ void __launch_bounds__(1024, 1) k(unsigned* sync, double* data, int iters, int sleep_ns, int nstores, int* bad)
{
    unsigned target = 0, epoch = 0;
    unsigned crank = 0, nclusters = gridDim.x;
    if (V == 4) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        unsigned cs;
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
        nclusters = gridDim.x / cs;
    }
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x;
    int errs = 0;
    for (int it = 1; it <= iters; ++it) {
        if ((int)threadIdx.x < nstores) data[(size_t)blockIdx.x * 64 + threadIdx.x] = (double)it;
        if (V == 0) v_counter(sync, target, sleep_ns);
        if (V == 1) v_relaxed(sync, target, sleep_ns);
        if (V == 2) v_flags(sync, epoch, sleep_ns);
        if (V == 3) v_sharded<8>(sync, epoch, sleep_ns);
        if (V == 4) v_cluster(sync, target, sleep_ns, nclusters, crank);
        if (V == 5) grid.sync();
        if (V == 6) v_sharded<2>(sync, epoch, sleep_ns);
        if ((int)threadIdx.x < nstores) {
            const double v = __ldcg(data + (size_t)((blockIdx.x + 37) % G) * 64 + threadIdx.x);
            if (v != (double)it && v != (double)(it + 1)) ++errs;
        }
    }
    if (errs) atomicAdd(bad, errs);
}

template <int V>
static void run(const char* name, int cluster, int sleep_ns, int nstores)
{
    unsigned* sync; double* data; int* bad;
    cudaMalloc(&sync, 4096 * 4); cudaMemset(sync, 0, 4096 * 4);
    cudaMalloc(&data, 148 * 64 * 8 * 2); cudaMemset(data, 0, 148 * 64 * 8 * 2);
    cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    int iters = 4000;
    void* args[] = {&sync, &data, &iters, &sleep_ns, &nstores, &bad};
    cudaLaunchConfig_t cfg = {};
    int G = 148;
    if (cluster > 1) G = (148 / cluster) * cluster;
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension; at[1].val.clusterDim.x = cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = cluster > 1 ? 2 : 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 3 && err == cudaSuccess; ++rep) {
        cudaMemset(sync, 0, 4096 * 4);
        cudaEventRecord(e0);
        err = cudaLaunchKernelExC(&cfg, (const void*)k<V>, args);
        cudaEventRecord(e1);
        if (err == cudaSuccess) err = cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int hb = 0; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) { printf("%-44s cluster %d: %s\n", name, cluster, cudaGetErrorString(err)); cudaGetLastError(); }
    else printf("%-44s cluster %2d sleep %3d stores %2d grid %3d: %7.1f ns / barrier  (order errors %d)\n", name, cluster, sleep_ns, nstores, G, best * 1e6 / iters, hb);
    cudaFree(sync); cudaFree(data); cudaFree(bad);
}

int main()
{
    for (int nstores = 0; nstores <= 32; nstores += 32) {
        run<0>("counter red.release / ld.acquire", 1, 40, nstores);
        run<0>("counter red.release / ld.acquire", 1, 0, nstores);
        run<0>("counter red.release / ld.acquire", 1, 100, nstores);
        run<1>("counter fence + relaxed red / relaxed poll", 1, 40, nstores);
        run<1>("counter fence + relaxed red / relaxed poll", 1, 0, nstores);
        run<2>("one flag per CTA, warp polls all", 1, 40, nstores);
        run<2>("one flag per CTA, warp polls all", 1, 0, nstores);
        run<3>("counter sharded over 8 lines", 1, 40, nstores);
        run<3>("counter sharded over 8 lines", 1, 0, nstores);
        run<6>("counter sharded over 2 lines", 1, 0, nstores);
        run<4>("cluster barrier + counter between clusters", 2, 40, nstores);
        run<4>("cluster barrier + counter between clusters", 2, 0, nstores);
        run<4>("cluster barrier + counter between clusters", 4, 40, nstores);
        run<4>("cluster barrier + counter between clusters", 4, 0, nstores);
        run<4>("cluster barrier + counter between clusters", 8, 0, nstores);
        run<5>("cooperative_groups grid.sync()", 1, 0, nstores);
    }
    return 0;
}
