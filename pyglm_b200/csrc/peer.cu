// Exchange steps of the multi-GPU sweep as kernels over peer-mapped (symmetric) memory instead of NCCL calls
// (VERDICT r1 item 4).  The buffers are allocated and exchanged by the host (torch symmetric memory: one
// allocation per rank, every rank holds the peers' base pointers in a device array); a device-side barrier of the same
// facility orders the pushes against the reads.  See also gram_tc_finalize_peers_kernel (gram_tc.cu), the fused
// reduce-scatter + finalize of the time-sharded Gram.
#include "common.cuh"

namespace {

// src (nwords x 16 bytes) -> the same offset of every peer's buffer: the all-gather of the new state rows
// [a | W | b | status] (models.py:169-171 leaves each regression's state with its owner; every rank needs all rows for
// the next psi and for the host network step) as plain NVLink stores, one kernel.
__global__ void __launch_bounds__(256)
peer_push_kernel(const uint4* __restrict__ src, long long nwords, void* const* __restrict__ peers, int world,
                 long long dst_off_bytes) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        const uint4 v = src[i];
        for (int r = 0; r < world; ++r)
            reinterpret_cast<uint4*>(static_cast<char*>(peers[r]) + dst_off_bytes)[i] = v;
    }
}

}  // namespace

// Copy nbytes (a multiple of 16, src 16-byte aligned) from src to offset dst_off_bytes of each of the `world` buffers
// whose base pointers are in the DEVICE array peers (this rank's own buffer included).
extern "C" int pyglm_peer_push(const void* src, long long nbytes, void* const* peers, int world, long long dst_off_bytes,
                               cudaStream_t stream) {
    PYGLM_CHECK_ARG(src && peers && world >= 1 && nbytes >= 0 && nbytes % 16 == 0 && dst_off_bytes >= 0 &&
                    dst_off_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                    "pyglm_peer_push: sizes and offsets must be multiples of 16 bytes");
    if (nbytes == 0) return PYGLM_OK;
    const long long nwords = nbytes / 16;
    const int blocks = (int)((nwords + 255) / 256 < 592 ? (nwords + 255) / 256 : 592);
    peer_push_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint4*>(src), nwords, peers, world, dst_off_bytes);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}
