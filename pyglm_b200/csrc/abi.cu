// Library-level C ABI: last-error string, version, device probe.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void pyglm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* pyglm_last_error(void) { return g_err; }

extern "C" int pyglm_abi_version(void) { return 1; }

// 0 when the current device can run this library (compute capability 10.x), else an error.
extern "C" int pyglm_device_check(void) {
    int dev = 0;
    PYGLM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    PYGLM_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        pyglm_set_error("pyglm_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", dev, p.major, p.minor, p.name);
        return PYGLM_ERR_UNSUPPORTED;
    }
    return PYGLM_OK;
}
