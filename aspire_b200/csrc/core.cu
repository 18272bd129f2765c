// Error plumbing, version and device check of the C-ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace asp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return ASP_ERR_CUDA;
    }
    return ASP_OK;
}

int require_device() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_rc = ASP_ERR_DEVICE;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("no CUDA device: %s (aspire_b200 has no CPU fallback)", cudaGetErrorString(e));
        return ASP_ERR_DEVICE;
    }
    if (dev == cached_dev) return cached_rc;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess || major != 10) {
        cudaGetLastError();
        set_error("device %d is not compute capability 10.x (sm_100a kernels only, no fallback)", dev);
        cached_dev = dev;
        cached_rc = ASP_ERR_DEVICE;
        return ASP_ERR_DEVICE;
    }
    cached_dev = dev;
    cached_rc = ASP_OK;
    return ASP_OK;
}

}  // namespace asp

extern "C" {

int asp_version(void) { return ASP_VERSION; }

const char* asp_last_error(void) { return asp::g_err; }

int asp_device_check(void) { return asp::require_device(); }

size_t asp_reduce_workspace_bytes(void) { return sizeof(asp::ReduceWs); }

}  // extern "C"
