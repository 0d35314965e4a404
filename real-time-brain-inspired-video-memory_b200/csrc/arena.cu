// arena.cu -- growable device ranges for a store that must grow WITHOUT reallocation.
//
// The reference's store only ever grows (one SET c.embedding per new chunk, neo4j_handler.py:221-253).  A resident
// store that doubles by allocate-and-copy needs twice its size in HBM at the moment of growth and stops the world for
// a device-to-device copy; a 76.8 GB shard could not grow at all.  Here a store reserves a VIRTUAL address range for
// its maximum capacity once (cuMemAddressReserve: no memory behind it) and maps physical chunks (cuMemCreate +
// cuMemMap) at its end as rows arrive.  Base addresses never change, so resident rows are never copied, cached TMA
// descriptors of the resident prefix stay valid, and the HBM in use is the rows actually stored, rounded up to one chunk.
// The driver entry points are resolved through the runtime (cudaGetDriverEntryPoint): the library keeps no link-time
// dependency on libcuda and still loads on hosts without a driver.
#include "common.cuh"

#include <cuda.h>
#include <vector>

namespace vm {

namespace {
struct DriverApi {
    bool ok = false;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
};
DriverApi g_drv;

template <typename F> bool resolve(const char *name, F *out)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    *out = reinterpret_cast<F>(p);
    return true;
}

int driver_load()
{
    if (g_drv.ok) return VM_OK;
    bool ok = resolve("cuMemAddressReserve", &g_drv.MemAddressReserve) && resolve("cuMemAddressFree", &g_drv.MemAddressFree) &&
              resolve("cuMemCreate", &g_drv.MemCreate) && resolve("cuMemRelease", &g_drv.MemRelease) &&
              resolve("cuMemMap", &g_drv.MemMap) && resolve("cuMemUnmap", &g_drv.MemUnmap) &&
              resolve("cuMemSetAccess", &g_drv.MemSetAccess) &&
              resolve("cuMemGetAllocationGranularity", &g_drv.MemGetAllocationGranularity);
    resolve("cuGetErrorString", &g_drv.GetErrorString);
    VM_REQUIRE(ok, VM_ERR_CUDA, "the CUDA driver lacks the virtual memory management entry points");
    g_drv.ok = true;
    return VM_OK;
}

const char *drv_err(CUresult r)
{
    const char *s = nullptr;
    if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "unknown driver error";
}
}  // namespace

#define VM_DRV_CHECK(expr)                                                                   \
    do {                                                                                     \
        CUresult _r = (expr);                                                                \
        if (_r != CUDA_SUCCESS) {                                                            \
            set_error("%s failed: %s (%s:%d)", #expr, drv_err(_r), __FILE__, __LINE__);      \
            return _r == CUDA_ERROR_OUT_OF_MEMORY ? VM_ERR_OOM : VM_ERR_CUDA;                \
        }                                                                                    \
    } while (0)

// One reserved virtual range, physically backed from its start up to `mapped` bytes.
struct Arena {
    int device = 0;
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, gran = 0, chunk = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    std::vector<size_t> sizes;
};

static CUmemAllocationProp arena_prop(int device)
{
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    return p;
}

int arena_create(Arena **out, int device, size_t max_bytes)
{
    int rc = driver_load();
    if (rc != VM_OK) return rc;
    VM_CUDA_CHECK(cudaFree(nullptr));  // make sure the primary context exists and is current
    Arena *a = new (std::nothrow) Arena();
    VM_REQUIRE(a, VM_ERR_OOM, "host allocation failed");
    a->device = device;
    const CUmemAllocationProp prop = arena_prop(device);
    CUresult r = g_drv.MemGetAllocationGranularity(&a->gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || a->gran == 0) { delete a; set_error("cuMemGetAllocationGranularity failed: %s", drv_err(r)); return VM_ERR_CUDA; }
    // physical chunks of >= 64 MB (few map calls per GB), always a multiple of the granularity
    a->chunk = ((size_t(64) << 20) + a->gran - 1) / a->gran * a->gran;
    a->reserved = (max_bytes + a->chunk - 1) / a->chunk * a->chunk;
    if (a->reserved == 0) a->reserved = a->chunk;
    r = g_drv.MemAddressReserve(&a->base, a->reserved, 0, 0, 0);
    if (r != CUDA_SUCCESS) { set_error("cuMemAddressReserve(%zu) failed: %s", a->reserved, drv_err(r)); delete a; return VM_ERR_OOM; }
    *out = a;
    return VM_OK;
}

// Back [0, bytes) with physical memory (no-op when already backed).  Existing mappings are untouched.
int arena_grow(Arena *a, size_t bytes)
{
    VM_REQUIRE(bytes <= a->reserved, VM_ERR_OVERFLOW, "arena: %zu bytes exceed the reserved range of %zu", bytes, a->reserved);
    const CUmemAllocationProp prop = arena_prop(a->device);
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    while (a->mapped < bytes) {
        size_t want = bytes - a->mapped;
        // grow geometrically up to 1 GB per chunk so that a large store is a few dozen mappings
        size_t step = a->chunk;
        while (step < want && step < (size_t(1) << 30) && step < a->mapped) step *= 2;
        if (a->mapped + step > a->reserved) step = a->reserved - a->mapped;
        CUmemGenericAllocationHandle h;
        VM_DRV_CHECK(g_drv.MemCreate(&h, step, &prop, 0));
        CUresult r = g_drv.MemMap(a->base + a->mapped, step, 0, h, 0);
        if (r == CUDA_SUCCESS) r = g_drv.MemSetAccess(a->base + a->mapped, step, &acc, 1);
        if (r != CUDA_SUCCESS) {
            g_drv.MemUnmap(a->base + a->mapped, step);
            g_drv.MemRelease(h);
            set_error("mapping %zu bytes at offset %zu failed: %s", step, a->mapped, drv_err(r));
            return VM_ERR_CUDA;
        }
        a->handles.push_back(h);
        a->sizes.push_back(step);
        a->mapped += step;
    }
    return VM_OK;
}

void arena_destroy(Arena *a)
{
    if (!a) return;
    size_t off = 0;
    for (size_t i = 0; i < a->handles.size(); ++i) {
        g_drv.MemUnmap(a->base + off, a->sizes[i]);
        g_drv.MemRelease(a->handles[i]);
        off += a->sizes[i];
    }
    if (a->base) g_drv.MemAddressFree(a->base, a->reserved);
    delete a;
}

void *arena_base(const Arena *a) { return reinterpret_cast<void *>(a->base); }
size_t arena_mapped(const Arena *a) { return a->mapped; }
size_t arena_reserved(const Arena *a) { return a->reserved; }

}  // namespace vm
