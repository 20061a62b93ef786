// multicast.cu -- NVSwitch multicast mailbox for the row-partitioned exchange (host side; driver API through
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda and still loads on a box without a GPU).
//
// The in-kernel exchange of k_pdhg_rowpart stores every new dual value into the mailbox of EVERY peer.  With peer pointers
// that is N - 1 small NVLink writes per value (measured: the wait for the peers' words grows 1.7 -> 3.0 -> 6.8 us on osa-60
// and 2.6 -> 5.2 -> 9.9 us on ken-18 from 2 to 4 to 8 GPUs: 16-byte writes use a fraction of the link).  Here the mailboxes
// of all ranks are bound into ONE multicast object: a single `multimem.st` on the multicast address is replicated by the
// switch into every GPU's mailbox.  The mailbox itself becomes a cuMemCreate allocation (multicast needs the VMM API), mapped
// twice per rank: unicast (the local polls of unpack_mail) and multicast (the stores).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/mllp_b200.h"
#include "pdhg_host.h"

namespace mllp {
void set_last_error(const std::string& msg);

namespace {
struct DriverApi {
    bool ok = false;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
};

template <class F>
bool entry(const char* name, F* fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

DriverApi* driver()
{
    static DriverApi d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        d.ok = entry("cuGetErrorString", &d.GetErrorString) && entry("cuDeviceGet", &d.DeviceGet) &&
               entry("cuDeviceGetAttribute", &d.DeviceGetAttribute) && entry("cuMulticastCreate", &d.MulticastCreate) &&
               entry("cuMulticastAddDevice", &d.MulticastAddDevice) && entry("cuMulticastBindMem", &d.MulticastBindMem) &&
               entry("cuMulticastUnbind", &d.MulticastUnbind) && entry("cuMulticastGetGranularity", &d.MulticastGetGranularity) &&
               entry("cuMemCreate", &d.MemCreate) && entry("cuMemRelease", &d.MemRelease) &&
               entry("cuMemExportToShareableHandle", &d.MemExportToShareableHandle) &&
               entry("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle) &&
               entry("cuMemAddressReserve", &d.MemAddressReserve) && entry("cuMemAddressFree", &d.MemAddressFree) &&
               entry("cuMemMap", &d.MemMap) && entry("cuMemUnmap", &d.MemUnmap) && entry("cuMemSetAccess", &d.MemSetAccess);
    }
    return d.ok ? &d : nullptr;
}

int cu_fail(DriverApi* D, CUresult r, const char* what)
{
    const char* s = nullptr;
    if (D && D->GetErrorString) D->GetErrorString(r, &s);
    set_last_error(std::string(what) + ": " + (s ? s : "driver error") + " (" + std::to_string((int)r) + ")");
    return 3000 + (int)r;
}
#define CU_OK(call)                                         \
    do {                                                    \
        CUresult r_ = (call);                               \
        if (r_ != CUDA_SUCCESS) return cu_fail(D, r_, #call); \
    } while (0)

CUmulticastObjectProp mc_prop(int nranks, size_t size)
{
    CUmulticastObjectProp p = {};
    p.numDevices = (unsigned)nranks;
    p.size = size;
    p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    p.flags = 0;
    return p;
}
}  // namespace

int mc_supported(int device, int* out)
{
    *out = 0;
    DriverApi* D = driver();
    if (!D) return 0;
    CUdevice dev;
    CU_OK(D->DeviceGet(&dev, device));
    int sup = 0;
    CU_OK(D->DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
    *out = sup;
    return 0;
}

// size of the multicast object / of every rank's mailbox allocation for `bytes` of payload (the same on every rank)
static int mc_size(DriverApi* D, int nranks, size_t bytes, size_t* size, size_t* gran)
{
    CUmulticastObjectProp p = mc_prop(nranks, bytes);
    CU_OK(D->MulticastGetGranularity(gran, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    *size = (bytes + *gran - 1) / *gran * *gran;
    return 0;
}

int mc_create(McState* S, int nranks, size_t bytes, int* fd)
{
    DriverApi* D = driver();
    if (!D) { set_last_error("multicast: the driver entry points are not available"); return MLLP_E_STATE; }
    int rc = mc_size(D, nranks, bytes, &S->size, &S->gran);
    if (rc) return rc;
    CUmulticastObjectProp p = mc_prop(nranks, S->size);
    CUmemGenericAllocationHandle h;
    CU_OK(D->MulticastCreate(&h, &p));
    S->mc = (unsigned long long)h; S->have_mc = true;
    int f = -1;
    CU_OK(D->MemExportToShareableHandle(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd = f;
    return 0;
}

int mc_import(McState* S, int nranks, size_t bytes, int fd)
{
    DriverApi* D = driver();
    if (!D) { set_last_error("multicast: the driver entry points are not available"); return MLLP_E_STATE; }
    int rc = mc_size(D, nranks, bytes, &S->size, &S->gran);
    if (rc) return rc;
    CUmemGenericAllocationHandle h;
    CU_OK(D->MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    S->mc = (unsigned long long)h; S->have_mc = true;
    return 0;
}

int mc_add_device(McState* S, int device)
{
    DriverApi* D = driver();
    CUdevice dev;
    CU_OK(D->DeviceGet(&dev, device));
    CU_OK(D->MulticastAddDevice((CUmemGenericAllocationHandle)S->mc, dev));
    S->device = device;
    return 0;
}

// after EVERY rank has added its device: this rank's mailbox memory, bound into the object and mapped twice
int mc_bind_map(McState* S)
{
    DriverApi* D = driver();
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = S->device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle mem;
    CU_OK(D->MemCreate(&mem, S->size, &ap, 0));
    S->mem = (unsigned long long)mem; S->have_mem = true;
    CU_OK(D->MulticastBindMem((CUmemGenericAllocationHandle)S->mc, 0, mem, 0, S->size, 0));
    S->bound = true;
    CUmemAccessDesc ad = {};
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = S->device;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUdeviceptr uc = 0, mv = 0;
    CU_OK(D->MemAddressReserve(&uc, S->size, S->gran, 0, 0));
    S->uc = (unsigned long long)uc;
    CU_OK(D->MemMap(uc, S->size, 0, mem, 0));
    S->uc_mapped = true;
    CU_OK(D->MemSetAccess(uc, S->size, &ad, 1));
    CU_OK(D->MemAddressReserve(&mv, S->size, S->gran, 0, 0));
    S->mcva = (unsigned long long)mv;
    CU_OK(D->MemMap(mv, S->size, 0, (CUmemGenericAllocationHandle)S->mc, 0));
    S->mc_mapped = true;
    CU_OK(D->MemSetAccess(mv, S->size, &ad, 1));
    if (cudaMemset((void*)uc, 0, S->size) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        set_last_error("multicast: clearing the mailbox failed");
        return MLLP_E_STATE;
    }
    return 0;
}

void mc_destroy(McState* S)
{
    DriverApi* D = driver();
    if (!D || !S) return;
    if (S->mc_mapped) D->MemUnmap((CUdeviceptr)S->mcva, S->size);
    if (S->mcva) D->MemAddressFree((CUdeviceptr)S->mcva, S->size);
    if (S->uc_mapped) D->MemUnmap((CUdeviceptr)S->uc, S->size);
    if (S->uc) D->MemAddressFree((CUdeviceptr)S->uc, S->size);
    if (S->bound) {
        CUdevice dev;
        if (D->DeviceGet(&dev, S->device) == CUDA_SUCCESS) D->MulticastUnbind((CUmemGenericAllocationHandle)S->mc, dev, 0, S->size);
    }
    if (S->have_mem) D->MemRelease((CUmemGenericAllocationHandle)S->mem);
    if (S->have_mc) D->MemRelease((CUmemGenericAllocationHandle)S->mc);
    *S = McState();
}

}  // namespace mllp
