// abi_core.cu -- error plumbing, launch accounting and host<->device helpers of the C-ABI.
#include "zr_common.cuh"
#include <cstdarg>
#include <cstdio>
#include <atomic>

namespace zr
{
    static thread_local char g_err[512] = { 0 };
    static std::atomic<uint64_t> g_launches{ 0 };

    void set_error(const char* fmt, ...)
    {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g_err, sizeof(g_err), fmt, ap);
        va_end(ap);
    }

    zr_status cuda_fail(cudaError_t e, const char* what)
    {
        set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
        return ZR_ERR_CUDA;
    }

    void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}

extern "C"
{
    const char* zr_last_error(void) { return zr::g_err; }
    uint32_t zr_abi_version(void) { return (1u << 16) | 0u; }
    uint64_t zr_kernel_launch_count(void) { return zr::g_launches.load(); }

    zr_status zr_device_malloc(void** d_ptr, size_t bytes)
    {
        if (!d_ptr) { zr::set_error("zr_device_malloc: null out pointer"); return ZR_ERR_INVALID_ARG; }
        cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
        if (e != cudaSuccess) { zr::cuda_fail(e, "cudaMalloc"); return e == cudaErrorMemoryAllocation ? ZR_ERR_OUT_OF_MEMORY : ZR_ERR_CUDA; }
        return ZR_OK;
    }
    void zr_device_free(void* d_ptr) { if (d_ptr) cudaFree(d_ptr); }
    zr_status zr_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream)
    {
        ZR_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        return ZR_OK;
    }
    zr_status zr_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream)
    {
        ZR_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        return ZR_OK;
    }
    zr_status zr_memset_d(void* d_dst, int value, size_t bytes, void* stream)
    {
        ZR_CUDA(cudaMemsetAsync(d_dst, value, bytes, (cudaStream_t)stream));
        return ZR_OK;
    }
    zr_status zr_stream_synchronize(void* stream)
    {
        ZR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        return ZR_OK;
    }
}
