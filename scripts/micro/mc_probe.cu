// Dev tool: can this box do NVSwitch multicast (cuMulticast*) and does a 16-byte multimem.st reach every GPU's copy?
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o mc_probe mc_probe.cu -lcuda && ./mc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("FAILED %s: %s\n", #x, s_); return 1; } } while (0)

__global__ void k_mc_store(unsigned long long* mc, int n, unsigned long long tag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long a = 0x1000ull + (unsigned long long)i, b = a ^ tag;
    unsigned lo0 = (unsigned)a, hi0 = (unsigned)(a >> 32), lo1 = (unsigned)b, hi1 = (unsigned)(b >> 32);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 2 * (size_t)i), "r"(lo0), "r"(hi0), "r"(lo1), "r"(hi1) : "memory");
}
__global__ void k_check(const unsigned long long* uc, int n, unsigned long long tag, int* bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long a = uc[2 * (size_t)i], b = uc[2 * (size_t)i + 1];
    if (a != 0x1000ull + (unsigned long long)i || (a ^ b) != tag) atomicAdd(bad, 1);
}

int main()
{
    CU(cuInit(0));
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    printf("devices: %d\n", ndev);
    if (ndev < 2) { printf("need 2 devices\n"); return 0; }
    const int N = ndev > 8 ? 8 : ndev;
    for (int d = 0; d < N; ++d) {
        int sup = 0;
        CUdevice dev; CU(cuDeviceGet(&dev, d));
        CU(cuDeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
        printf("device %d multicast supported: %d\n", d, sup);
        if (!sup) return 0;
    }
    CUmulticastObjectProp mp = {};
    mp.numDevices = N; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR; mp.flags = 0;
    size_t gran = 0;
    mp.size = 1 << 20;
    CU(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    printf("granularity %zu\n", gran);
    mp.size = ((size_t)(1 << 20) + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle mc;
    CU(cuMulticastCreate(&mc, &mp));
    int fd = -1;
    CU(cuMemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    printf("multicast object created, exported fd %d\n", fd);
    for (int d = 0; d < N; ++d) { CUdevice dev; CU(cuDeviceGet(&dev, d)); CU(cuMulticastAddDevice(mc, dev)); }
    std::vector<CUmemGenericAllocationHandle> mem(N);
    std::vector<CUdeviceptr> uc(N), mcva(N);
    for (int d = 0; d < N; ++d) {
        cudaSetDevice(d); cudaFree(0);
        CUmemAllocationProp ap = {};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED; ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ap.location.id = d;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        CU(cuMemGetAllocationGranularity(&mg, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        CU(cuMemCreate(&mem[d], mp.size, &ap, 0));
        CU(cuMulticastBindMem(mc, 0, mem[d], 0, mp.size, 0));
        CUmemAccessDesc ad = {}; ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = d; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CU(cuMemAddressReserve(&uc[d], mp.size, gran, 0, 0)); CU(cuMemMap(uc[d], mp.size, 0, mem[d], 0)); CU(cuMemSetAccess(uc[d], mp.size, &ad, 1));
        CU(cuMemAddressReserve(&mcva[d], mp.size, gran, 0, 0)); CU(cuMemMap(mcva[d], mp.size, 0, mc, 0)); CU(cuMemSetAccess(mcva[d], mp.size, &ad, 1));
        cudaMemset((void*)uc[d], 0, mp.size);
    }
    for (int d = 0; d < N; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); }
    const int n = 4096;
    cudaSetDevice(0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mc_store<<<n / 256, 256>>>((unsigned long long*)mcva[0], n, 0xabcdefull);
    cudaError_t ce = cudaDeviceSynchronize();
    printf("multimem.st kernel: %s\n", cudaGetErrorString(ce));
    for (int d = 0; d < N; ++d) {
        cudaSetDevice(d);
        int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
        k_check<<<n / 256, 256>>>((const unsigned long long*)uc[d], n, 0xabcdefull, bad);
        int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
        printf("device %d: %d of %d words wrong\n", d, hb, n);
    }
    printf("MC_PROBE_DONE\n");
    return 0;
}
