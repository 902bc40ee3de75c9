// Runtime entry points of libsdrgpu: device selection, memory helpers, error string, launch counter.
#include "common.cuh"

namespace sdrgpu {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
int g_tuning[SDRGPU_TUNE_COUNT] = {0};
}  // namespace sdrgpu

using namespace sdrgpu;

extern "C" {

const char *sdrgpu_last_error(void) { return g_last_error.c_str(); }
const char *sdrgpu_version(void) { return "sdrgpu 0.1 (sm_100a)"; }
uint64_t sdrgpu_launch_count(void) { return g_launches.load(); }

sdrgpu_status sdrgpu_set_tuning(int knob, int value)
{
    if (knob < 0 || knob >= SDRGPU_TUNE_COUNT) return fail(SDRGPU_ERR_INVALID_ARG, "unknown tuning knob %d", knob);
    if (value < 0) return fail(SDRGPU_ERR_INVALID_ARG, "tuning values are >= 0");
    g_tuning[knob] = value;
    return SDRGPU_OK;
}
int sdrgpu_get_tuning(int knob) { return (knob < 0 || knob >= SDRGPU_TUNE_COUNT) ? -1 : g_tuning[knob]; }

sdrgpu_status sdrgpu_device_count(int *count)
{
    if (!count) return fail(SDRGPU_ERR_INVALID_ARG, "count is NULL");
    *count = 0;
    SDRGPU_CUDA(cudaGetDeviceCount(count));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SDRGPU_ERR_CUDA, "no CUDA device available (%s); libsdrgpu has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SDRGPU_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, n);
    SDRGPU_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SDRGPU_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SDRGPU_ERR_CUDA, "device %d is sm_%d%d; libsdrgpu is built for sm_100a only", device, prop.major,
                    prop.minor);
    SDRGPU_CUDA(cudaFree(0));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_alloc_pinned(void **ptr, size_t bytes)
{
    if (!ptr) return fail(SDRGPU_ERR_INVALID_ARG, "ptr is NULL");
    SDRGPU_CUDA(cudaMallocHost(ptr, bytes));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_free_pinned(void *ptr)
{
    SDRGPU_CUDA(cudaFreeHost(ptr));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_device_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(SDRGPU_ERR_INVALID_ARG, "ptr is NULL");
    SDRGPU_CUDA(cudaMalloc(ptr, bytes));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_device_free(void *ptr)
{
    SDRGPU_CUDA(cudaFree(ptr));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_memcpy(void *dst, const void *src, size_t bytes, int dst_mem, int src_mem)
{
    cudaMemcpyKind kind = cudaMemcpyDefault;
    (void)dst_mem;
    (void)src_mem;
    SDRGPU_CUDA(cudaMemcpy(dst, src, bytes, kind));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_device_synchronize(void)
{
    SDRGPU_CUDA(cudaDeviceSynchronize());
    return SDRGPU_OK;
}

}  // extern "C"
