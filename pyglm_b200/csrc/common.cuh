// Shared device/host helpers for the pyglm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pyglm_b200 kernels are written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------- error plumbing (C ABI: int status)
enum {
    PYGLM_OK = 0,
    PYGLM_ERR_INVALID = 1,   // bad argument (shape / alignment / null pointer)
    PYGLM_ERR_CUDA = 2,      // a CUDA runtime call or launch failed
    PYGLM_ERR_UNSUPPORTED = 3
};

void pyglm_set_error(const char* fmt, ...);

#define PYGLM_CHECK_ARG(cond, ...)                      \
    do {                                                \
        if (!(cond)) {                                  \
            pyglm_set_error(__VA_ARGS__);               \
            return PYGLM_ERR_INVALID;                   \
        }                                               \
    } while (0)

#define PYGLM_CUDA(call)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            pyglm_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,               \
                            cudaGetErrorString(e__));                                          \
            return PYGLM_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

#define PYGLM_LAUNCH_CHECK() PYGLM_CUDA(cudaGetLastError())

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device primitives
#ifdef __CUDACC__

// D(8x8) += A(8x4, row) * B(4x8, col), all FP64: the native DMMA.8x8x4 of sm_100a.
// lane = 4*g + q:  A element (row g, col q); B element (row q, col g); C elements (row g, cols 2q, 2q+1).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 16-byte async global->shared copy; src_bytes == 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a fixed order (deterministic); result valid in thread 0. scratch: >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < nw) ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

#endif  // __CUDACC__
