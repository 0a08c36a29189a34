// Philox4x32-10 counter-based generator and the stream convention shared with the CPU oracle
// (oracle/pg_devroye.c): key = (seed_lo, seed_hi), counter = (elem_lo, elem_hi, call_id, block#).
// One independent stream per (seed, call_id, element); a stream hands out 32-bit words four at a
// time.  Results do not depend on grid shape or on how neurons / time bins are sharded over GPUs.
#pragma once
#include <stdint.h>

// one out-of-line copy of the 10-round block function (inlined at every unif() call site it dominated the
// instruction footprint of the Polya-gamma kernel); operands and result travel in registers
static __device__ __noinline__ uint4 philox4x32_10_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct PhiloxStream {
    uint32_t c0, c1, c2, c3, k0, k1;
    uint32_t o0, o1, o2, o3;   // scalars, not an array: dynamic indexing would spill to local memory
    int have;

    __device__ __forceinline__ void seed(uint64_t seed_, uint32_t call_id, uint64_t elem) {
        k0 = (uint32_t)seed_; k1 = (uint32_t)(seed_ >> 32);
        c0 = (uint32_t)elem;  c1 = (uint32_t)(elem >> 32);
        c2 = call_id; c3 = 0; have = 0;
    }
    __device__ __forceinline__ void refill() {
        const uint4 o = philox4x32_10_block(c0, c1, c2, c3, k0, k1);
        o0 = o.x; o1 = o.y; o2 = o.z; o3 = o.w;
        c3 += 1; have = 4;
    }
    __device__ __forceinline__ uint32_t u32() {
        if (have == 0) refill();
        uint32_t v = (have == 4) ? o0 : (have == 3) ? o1 : (have == 2) ? o2 : o3;
        --have;
        return v;
    }
    // uniform on [0,1) from 53 random bits (27 from the first word, 26 from the second)
    __device__ __forceinline__ double unif() {
        uint32_t a = u32(), b = u32();
        return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
    }
    static __device__ __forceinline__ double to_unif(uint32_t a, uint32_t b) {
        return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
    }
    // The next two uniforms of the stream (the same values as two unif() calls) with ONE unconditional block
    // evaluation: between uniforms `have` is 0 or 2, so the pair is either a whole new block or the last two words
    // of the current block and the first two of the next.  Lanes of a warp whose streams are at different offsets
    // still evaluate the block function together.  have += 2 afterwards hands the second uniform back (which can
    // leave a whole block unread, have == 4: the one case that needs no evaluation).
    __device__ __forceinline__ void unif2(double& ua, double& ub) {
        const bool half = (have == 2);
        const uint32_t p2 = o2, p3 = o3;
        if (have != 4) refill();
        ua = half ? to_unif(p2, p3) : to_unif(o0, o1);
        ub = half ? to_unif(o0, o1) : to_unif(o2, o3);
        have = half ? 2 : 0;
    }
    __device__ __forceinline__ void skip_block() { c3 += 1; have = 0; }
    __device__ __forceinline__ double expon() { return -log1p(-unif()); }
    // square of a standard normal (Box-Muller, cosine branch)
    __device__ __forceinline__ double norm_sq() {
        double u1 = unif(), u2 = unif();
        double c = cospi(2.0 * u2);   // == cos(2 pi u2) without the large-argument slow path
        return -2.0 * log1p(-u1) * c * c;
    }
    // standard normal (Box-Muller, cosine branch)
    __device__ __forceinline__ double norm() {
        double u1 = unif(), u2 = unif();
        return sqrt(-2.0 * log1p(-u1)) * cospi(2.0 * u2);
    }
};
