// (2) Polya-gamma augmentation: omega[t, j] ~ PG(1, psi[t, j]) for every time bin and local neuron.
//
// Replaces pypolyagamma.pgdrawvpar (reference call site regression.py:501-508; third-party C++,
// not vendored).  Same exact sampler -- Devroye's alternating-series rejection method for J*(1, z)
// with truncation point t = 0.64 (Polson, Scott & Windle 2013) -- restated for one thread per draw
// with a private counter-based Philox4x32-10 stream keyed by (seed, call_id, global element index),
// so draws do not depend on grid shape or on how neurons / time are sharded across GPUs, and match
// the CPU oracle (oracle/pg_devroye.c, rng_kind 0) element by element.
// Divergence: the outer loop accepts on its first proposal ~99.9% of the time and the series test ends
// after one or two terms, but the two proposal samplers differ (3 uniforms and no loop against >= 6
// uniforms and two nested rejection loops): one thread per draw (pg_draw_kernel) runs at 9 of 32 lanes.
// The default path is therefore the branch-compacted two-pass form further down (pg_pick_kernel,
// pg_ig_small_kernel, pg_ig_kernel), which produces the same draws.
// Algorithmic traffic: 8 B psi read + 8 B omega write per draw (HBM roofline); the kernels
// are bound by instruction issue (~900 thread-instructions per exact draw), not by bytes.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr double PG_T = 0.64;
constexpr double PG_T_RECIP = 1.5625;
constexpr double PG_PI = 3.141592653589793238462643383279502884;
constexpr double PG_LOG_HALF_PI = 0.45158270528945486472619522989488;   // log(pi/2)

// One out-of-line copy of each double-precision transcendental.  Inlined at every call site the kernel was
// ~77 KB of SASS and ncu showed `no_instruction` (instruction-cache miss) as the dominant stall; divergent
// lanes of a warp sit in different libm bodies, so footprint matters more than call overhead.
__device__ __noinline__ double d_exp(double x) { return exp(x); }
__device__ __noinline__ double d_log(double x) { return log(x); }
__device__ __noinline__ double d_erfc(double x) { return erfc(x); }

struct PgRng {
    PhiloxStream s;
    __device__ __forceinline__ double unif() { return s.unif(); }
    // Exp(1): 1 - u is exactly representable, so this equals -log1p(-u) up to the rounding of log
    __device__ __forceinline__ double expon() { return -d_log(1.0 - s.unif()); }
    __device__ __forceinline__ double norm_sq() {
        const double u1 = s.unif(), u2 = s.unif();
        const double c = cospi(2.0 * u2);
        return -2.0 * d_log(1.0 - u1) * c * c;
    }
};

// a_n(x); lx = log(x) is only read when x <= t
__device__ __forceinline__ double pg_a(int n, double x, double lx) {
    const double K = (n + 0.5) * PG_PI;
    if (x > PG_T) return K * d_exp(-0.5 * K * K * x);
    if (x > 0.0) return d_exp(-1.5 * (PG_LOG_HALF_PI + lx) + d_log(K) - 2.0 * (n + 0.5) * (n + 0.5) / x);
    return 0.0;
}

// p/(p+q), the probability of proposing from the exponential tail.  q/p = (4/pi) fz e^{fz t} [e^{-z} Phi(b) +
// e^{z} Phi(a)] evaluated directly (the oracle's log form guards an overflow that cannot occur for z < 12).
// For z >= 12 the mass is below 2^-53, the resolution of the uniform it is compared with, and is taken as 0.
__device__ __forceinline__ double pg_mass_texpon(double z, double fz) {
    if (z >= 12.0) return 0.0;
    const double rt = 1.25;                                     // sqrt(1/t)
    const double b = rt * (PG_T * z - 1.0), a = -rt * (PG_T * z + 1.0);
    const double phib = 0.5 * d_erfc(-b * 0.70710678118654752440);
    const double phia = 0.5 * d_erfc(-a * 0.70710678118654752440);
    const double qdivp = 4.0 / PG_PI * fz * (d_exp(fz * PG_T - z) * phib + d_exp(fz * PG_T + z) * phia);
    return 1.0 / (1.0 + qdivp);
}

// ---- single-precision screening -------------------------------------------------------------------------------
// Every accept / reject decision of the sampler compares a uniform (or a cheap quantity) with a transcendental
// expression.  Only the VALUE that is returned (X) has to be double precision; a comparison does not, unless the
// two sides are close.  Each test is therefore evaluated in FP32 first (MUFU-based expf / logf / erfcf, error
// < 1e-5 relative on these argument ranges) and decided there when the sides differ by more than 2e-4 relative;
// otherwise -- and for NaN / overflow -- the FP64 expression of the oracle decides.  Decisions, the stream of
// uniforms consumed and the returned values are exactly the oracle's; ~10 of the ~12 FP64 transcendentals per draw
// become FP32.
constexpr float PG_GUARD = 2e-4f;

// +1: a < b for sure, 0: a > b for sure, -1: too close to call in single precision (or not finite)
__device__ __forceinline__ int pg_less32(float a, float b, float scale) {
    const float d = a - b;
    if (!(fabsf(d) > PG_GUARD * scale)) return -1;
    return d < 0.0f ? 1 : 0;
}

// -log(1 - u) in single precision with full relative accuracy at both ends of (0, 1)
__device__ __forceinline__ float pg_expon32(double u) {
    return (u < 0.5) ? -log1pf(-(float)u) : -__logf((float)(1.0 - u));
}

__device__ __forceinline__ float pg_a32(int n, float x, float lx) {
    const float K = (n + 0.5f) * 3.14159265358979f;
    if (x > 0.64f) return K * __expf(-0.5f * K * K * x);
    return __expf(-1.5f * (0.451582705f + lx) + __logf(K) - 2.0f * (n + 0.5f) * (n + 0.5f) / x);
}

// the alternating-series test of pg1_draw in FP64 (the oracle's arithmetic): 1 accept, 0 reject
__device__ __noinline__ int pg_series64(double X, double u) {
    const double lx = (X <= PG_T) ? d_log(X) : 0.0;
    double S = pg_a(0, X, lx);
    const double Y = u * S;
    int n = 0;
    for (;;) {
        ++n;
        if (n & 1) { S -= pg_a(n, X, lx); if (Y <= S) return 1; }
        else       { S += pg_a(n, X, lx); if (Y > S) return 0; }
    }
}

__device__ __forceinline__ int pg_series(double X, double u) {
    const float xf = (float)X, uf = (float)u;
    // too close to the truncation point to know which branch of a_n the FP64 code takes, or tiny x (underflow)
    if (fabsf(xf - 0.64f) < 1e-4f || !(xf > 1e-3f)) return pg_series64(X, u);
    const float lx = __logf(xf);
    const float a0 = pg_a32(0, xf, lx);
    float S = a0;
    const float Y = uf * a0;
    for (int n = 1; n <= 6; ++n) {
        const float an = pg_a32(n, xf, lx);
        if (n & 1) {
            S -= an;
            const int c = pg_less32(S, Y, a0);            // Y <= S  <=>  not (S < Y)
            if (c < 0) break;
            if (c == 0) return 1;
        } else {
            S += an;
            const int c = pg_less32(S, Y, a0);            // Y > S   <=>  S < Y
            if (c < 0) break;
            if (c == 1) return 0;
        }
    }
    return pg_series64(X, u);
}

// u < mass_texpon(z)?  (the branch choice of pg1_draw)
__device__ __forceinline__ bool pg_pick_texpon(double u, double z, double fz) {
    if (z >= 12.0) return false;                              // mass taken as 0, see pg_mass_texpon
    const float zf = (float)z, fzf = (float)fz;
    const float b = 1.25f * (0.64f * zf - 1.0f), a = -1.25f * (0.64f * zf + 1.0f);
    const float phib = 0.5f * erfcf(-b * 0.70710678f), phia = 0.5f * erfcf(-a * 0.70710678f);
    const float q = 1.27323954f * fzf * (__expf(fzf * 0.64f - zf) * phib + __expf(fzf * 0.64f + zf) * phia);
    const float m = 1.0f / (1.0f + q);
    const int c = pg_less32((float)u, m, 1.0f);
    if (c >= 0) return c == 1;
    return u < pg_mass_texpon(z, fz);
}

__device__ __forceinline__ double pg_rtigauss(double z, PgRng& r) {
    const double t = PG_T;
    double X = t + 1.0;
    if (PG_T_RECIP > z) {
        bool first = true;
        for (;;) {
            const double ua = r.unif();
            if (!first) {                                     // while (unif > alpha), alpha = exp(-z^2 X / 2)
                const float al = __expf(-0.5f * (float)(z * z * X));
                const int c = pg_less32(al, (float)ua, 1.0f); // alpha < u: continue the loop
                const bool again = (c >= 0) ? (c == 1) : (ua > d_exp(-0.5 * z * z * X));
                if (!again) break;
            }
            first = false;
            double u1, u2;
            for (;;) {                                        // until E1^2 <= 2 E2 / t
                u1 = r.unif(); u2 = r.unif();
                const float e1 = pg_expon32(u1), e2 = pg_expon32(u2);
                const float lhs = e1 * e1, rhs = 3.125f * e2;
                const int c = pg_less32(rhs, lhs, fmaxf(lhs, rhs));   // rhs < lhs: reject the pair
                bool reject;
                if (c >= 0) reject = (c == 1);
                else {
                    const double E1 = -d_log(1.0 - u1), E2 = -d_log(1.0 - u2);
                    reject = E1 * E1 > 2.0 * E2 / t;
                }
                if (!reject) break;
            }
            const double E1 = -d_log(1.0 - u1);
            X = 1.0 + E1 * t;
            X = t / (X * X);
        }
    } else {
        const double mu = 1.0 / z;
        while (X > t) {
            const double Y = r.norm_sq();
            const double half_mu = 0.5 * mu, mu_Y = mu * Y;
            X = mu + half_mu * mu_Y - half_mu * sqrt(4.0 * mu_Y + mu_Y * mu_Y);
            if (r.unif() > mu / (mu + X)) X = mu * mu / X;
        }
    }
    return X;
}

__device__ double pg1_draw(double psi, PgRng& r) {
    const double z = fabs(psi) * 0.5;
    const double fz = 0.125 * PG_PI * PG_PI + 0.5 * z * z;
    for (;;) {
        double X;
        if (pg_pick_texpon(r.unif(), z, fz)) X = PG_T + r.expon() / fz;
        else X = pg_rtigauss(z, r);
        if (pg_series(X, r.unif())) return 0.25 * X;
    }
}

// grid-stride over the T x n_valid valid entries; element id = (t_off + t) * n_total + (n_off + j)
__global__ void __launch_bounds__(256, 4)
pg_draw_kernel(const double* __restrict__ psi, int ldpsi, long long T, int n_valid, int ld_out,
               unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
               double* __restrict__ omega) {
    const long long total = T * (long long)n_valid;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long t = idx / n_valid;
        const int j = (int)(idx - t * n_valid);
        PgRng r;
        r.s.seed(seed, call_id, (unsigned long long)((t_off + t) * (long long)n_total + n_off + j));
        omega[t * ld_out + j] = pg1_draw(psi[t * ldpsi + j], r);
    }
}

// ---- branch-compacted two-pass form ("warp-coherent rejection") -----------------------------------------------
// One thread per draw keeps only ~9 of 32 lanes busy: about half of the lanes propose from the exponential tail (3
// uniforms, no loop), the others from the truncated inverse Gaussian (two nested rejection loops, >= 6 uniforms), and
// a warp pays for both.  Because an element's stream is a pure function of (seed, call, element), a draw can be
// resumed by ANY thread.  Pass 1 makes the proposal choice for every element with all lanes converged, finishes the
// exponential-tail draws on the spot and appends the others to a list (small z from the front, large z -- the
// other inverse-Gaussian sampler -- from the back; warp-aggregated atomics).  Pass 2 walks each list with every
// lane of a warp inside the same sampler, re-deriving the element's stream (one extra Philox block).  Decisions,
// uniforms consumed and values are those of pg1_draw, so the oracle stream test applies unchanged.
__global__ void __launch_bounds__(256, 4)
pg_pick_kernel(const double* __restrict__ psi, int ldpsi, long long T, int n_valid, int ld_out,
               unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
               double* __restrict__ omega, unsigned* __restrict__ counts, unsigned* __restrict__ list,
               unsigned capacity) {
    // List space is claimed once per CTA iteration (256 elements), not once per warp row: both counters live in
    // one L2 sector, and same-address atomics retire about one per L2 clock -- at two per warp row (1.25 M at
    // cfg3) that serialisation alone was most of this kernel's time (ncu r01m: long-scoreboard 69 % of the stall
    // cycles, issue slots 40 % busy).  Buffers alternate with the iteration parity, so two barriers per iteration do.
    __shared__ unsigned s_cnt[2][2][8], s_base[2][2];
    const unsigned total = (unsigned)(T * (long long)n_valid);            // < 2^32 - 1 on this path
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned par = 0;
    // same trip count for every thread of a CTA (barriers and full-mask ballots below); psi is fetched one
    // iteration ahead
    unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x;
    double p_next = 0.0;
    if (base + threadIdx.x < total) {
        const unsigned i0 = (unsigned)base + threadIdx.x, t0 = i0 / (unsigned)n_valid;
        p_next = psi[(long long)t0 * ldpsi + (i0 - t0 * (unsigned)n_valid)];
    }
    for (; base < total; base += stride, par ^= 1u) {
        const unsigned long long idx = base + threadIdx.x;
        const double p = p_next;
        if (idx + stride < total) {
            const unsigned i1 = (unsigned)(idx + stride), t1 = i1 / (unsigned)n_valid;
            p_next = psi[(long long)t1 * ldpsi + (i1 - t1 * (unsigned)n_valid)];
        }
        int kind = 0;                                         // 0: done here, 1: small-z IG, 2: large-z IG
        if (idx < total) {
            const long long t = (unsigned)idx / (unsigned)n_valid;
            const int j = (int)((unsigned)idx - (unsigned)t * (unsigned)n_valid);
            PgRng r;
            r.s.seed(seed, call_id, (unsigned long long)((t_off + t) * (long long)n_total + n_off + j));
            const double z = fabs(p) * 0.5;
            const double fz = 0.125 * PG_PI * PG_PI + 0.5 * z * z;
            if (pg_pick_texpon(r.unif(), z, fz)) {
                const double X = PG_T + r.expon() / fz;
                omega[t * ld_out + j] = pg_series(X, r.unif()) ? 0.25 * X : pg1_draw(p, r);
            } else {
                kind = (PG_T_RECIP > z) ? 1 : 2;
            }
        }
        const unsigned m1 = __ballot_sync(0xffffffffu, kind == 1), m2 = __ballot_sync(0xffffffffu, kind == 2);
        if (lane == 0) { s_cnt[par][0][warp] = __popc(m1); s_cnt[par][1][warp] = __popc(m2); }
        __syncthreads();
        if (threadIdx.x < 2) {
            unsigned tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_cnt[par][threadIdx.x][w];
            s_base[par][threadIdx.x] = tot ? atomicAdd(&counts[threadIdx.x], tot) : 0u;
        }
        __syncthreads();
        if (kind) {
            unsigned off = s_base[par][kind - 1];
            for (unsigned w = 0; w < warp; ++w) off += s_cnt[par][kind - 1][w];
            if (kind == 1) list[off + __popc(m1 & lt)] = (unsigned)idx;
            else list[capacity - 1u - (off + __popc(m2 & lt))] = (unsigned)idx;
        }
    }
}

// Pass 2, small z (z < 1/t): the truncated inverse Gaussian by rejection from the tail of 1/E^2 -- two nested
// rejection loops in pg_rtigauss.  With one list entry per thread a warp waits for its slowest lane (the maximum of
// 32 geometric trip counts).  Here the loops are flattened into one step per iteration -- a pair trial (E1, E2),
// and, when it is accepted, the alpha test and the series test -- and a lane that finishes a draw takes the next
// entry of its warp's contiguous chunk of the list at once (ballot + popc, no atomics), so every lane does a pair
// trial in every iteration until the chunk runs dry.  Each step's two uniforms come from unif2(): one block
// evaluation per step, executed by all lanes together whatever their stream offsets.  Block 0 of a listed
// element's stream (the proposal-choice uniform and the first, never tested, alpha uniform) is skipped, not computed.
__global__ void __launch_bounds__(256, 4)
pg_ig_small_kernel(const double* __restrict__ psi, int ldpsi, int n_valid, int ld_out,
                   unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
                   double* __restrict__ omega, const unsigned* __restrict__ counts,
                   const unsigned* __restrict__ list) {
    const unsigned long long count = counts[0];
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = (gridDim.x * (unsigned long long)blockDim.x) >> 5;
    const unsigned long long per = (count + n_warps - 1) / n_warps;
    unsigned long long next = warp * per < count ? warp * per : count;
    const unsigned long long end = next + per < count ? next + per : count;
    bool active = false;
    PgRng r;
    double z = 0.0, p = 0.0;
    long long out = 0;
    for (;;) {
        const unsigned want = __ballot_sync(0xffffffffu, !active);
        if (want) {
            const unsigned long long mine = next + __popc(want & lt);
            if (!active && mine < end) {
                const unsigned idx = list[mine];
                const long long t = idx / (unsigned)n_valid;
                const int j = (int)(idx - (unsigned)t * (unsigned)n_valid);
                p = psi[t * ldpsi + j];
                z = fabs(p) * 0.5;
                out = t * ld_out + j;
                r.s.seed(seed, call_id, (unsigned long long)((t_off + t) * (long long)n_total + n_off + j));
                r.s.skip_block();
                active = true;
            }
            next += __popc(want);
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            double u1, u2;
            r.s.unif2(u1, u2);
            const float e1 = pg_expon32(u1), e2 = pg_expon32(u2);
            const float lhs = e1 * e1, rhs = 3.125f * e2;
            int c = pg_less32(rhs, lhs, fmaxf(lhs, rhs));      // rhs < lhs: reject the pair
            bool reject;
            if (c >= 0) reject = (c == 1);
            else {
                const double E1 = -d_log(1.0 - u1), E2 = -d_log(1.0 - u2);
                reject = E1 * E1 > 2.0 * E2 / PG_T;
            }
            if (!reject) {
                const double E1 = -d_log(1.0 - u1);
                double X = 1.0 + E1 * PG_T;
                X = PG_T / (X * X);
                double ua, us;
                r.s.unif2(ua, us);
                const float al = __expf(-0.5f * (float)(z * z * X));
                c = pg_less32(al, (float)ua, 1.0f);            // alpha < u: propose again
                const bool again = (c >= 0) ? (c == 1) : (ua > d_exp(-0.5 * z * z * X));
                if (again) {
                    r.s.have += 2;                              // the series uniform was not drawn
                } else {
                    omega[out] = pg_series(X, us) ? 0.25 * X : pg1_draw(p, r);
                    active = false;
                }
            }
        }
    }
}

template <bool LARGE>
__global__ void __launch_bounds__(256, 4)
pg_ig_kernel(const double* __restrict__ psi, int ldpsi, int n_valid, int ld_out,
             unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
             double* __restrict__ omega, const unsigned* __restrict__ counts, const unsigned* __restrict__ list,
             unsigned capacity) {
    const unsigned count = counts[LARGE ? 1 : 0];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const unsigned idx = LARGE ? list[capacity - 1u - i] : list[i];
        const long long t = idx / (unsigned)n_valid;
        const int j = (int)(idx - (unsigned)t * (unsigned)n_valid);
        const double p = psi[t * ldpsi + j];
        PgRng r;
        r.s.seed(seed, call_id, (unsigned long long)((t_off + t) * (long long)n_total + n_off + j));
        (void)r.unif();                                       // the proposal-choice uniform, spent in pass 1
        const double X = pg_rtigauss(fabs(p) * 0.5, r);
        omega[t * ld_out + j] = pg_series(X, r.unif()) ? 0.25 * X : pg1_draw(p, r);
    }
}

__global__ void philox_unif_kernel(unsigned long long seed, unsigned call_id, unsigned long long elem0, int n_elem,
                                   int count, double* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    PhiloxStream r;
    r.seed(seed, call_id, elem0 + e);
    for (int i = 0; i < count; ++i) out[(size_t)e * count + i] = r.unif();
}

}  // namespace

// omega[t*ld_out + j] ~ PG(1, psi[t*ldpsi + j]) for t < T, j < n_valid.  Columns j >= n_valid of omega are
// left untouched (callers keep them zero).  (t_off, n_off, n_total) locate this shard in the global
// (time x neuron) grid so that the random stream of an element is independent of the sharding.
extern "C" int pyglm_pg_draw(const double* psi, int ldpsi, long long T, int n_valid, double* omega, int ld_out,
                             unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
                             cudaStream_t stream) {
    PYGLM_CHECK_ARG(psi && omega, "pyglm_pg_draw: null pointer");
    PYGLM_CHECK_ARG(T > 0 && n_valid > 0 && ldpsi >= n_valid && ld_out >= n_valid && n_total >= n_off + n_valid,
                    "pyglm_pg_draw: bad shape (T=%lld n=%d ldpsi=%d ld_out=%d n_off=%d n_total=%d)", T, n_valid, ldpsi, ld_out, n_off, n_total);
    long long total = T * (long long)n_valid;
    int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
    pg_draw_kernel<<<blocks, 256, 0, stream>>>(psi, ldpsi, T, n_valid, ld_out, seed, call_id, t_off, n_off, n_total, omega);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

// The same draws through the branch-compacted two-pass kernels.  workspace: pyglm_pg_draw_ws_bytes(T, n_valid) bytes
// of device memory (16 B of counters + one 32-bit element index per draw), contents irrelevant on entry.  Falls back
// to the one-pass kernel when the element indices do not fit 32 bits or the workspace is too small.
extern "C" size_t pyglm_pg_draw_ws_bytes(long long T, int n_valid) {
    return 16 + 4 * (size_t)(T > 0 ? T : 0) * (size_t)(n_valid > 0 ? n_valid : 0);
}

extern "C" int pyglm_pg_draw_ws(const double* psi, int ldpsi, long long T, int n_valid, double* omega, int ld_out,
                                unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    PYGLM_CHECK_ARG(psi && omega, "pyglm_pg_draw_ws: null pointer");
    PYGLM_CHECK_ARG(T > 0 && n_valid > 0 && ldpsi >= n_valid && ld_out >= n_valid && n_total >= n_off + n_valid,
                    "pyglm_pg_draw_ws: bad shape (T=%lld n=%d ldpsi=%d ld_out=%d n_off=%d n_total=%d)", T, n_valid, ldpsi, ld_out, n_off, n_total);
    const long long total = T * (long long)n_valid;
    if (!workspace || total >= 0xffffffffLL || workspace_bytes < pyglm_pg_draw_ws_bytes(T, n_valid))
        return pyglm_pg_draw(psi, ldpsi, T, n_valid, omega, ld_out, seed, call_id, t_off, n_off, n_total, stream);
    unsigned* counts = (unsigned*)workspace;
    unsigned* list = counts + 4;
    PYGLM_CUDA(cudaMemsetAsync(counts, 0, 16, stream));
    const int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
    pg_pick_kernel<<<blocks, 256, 0, stream>>>(psi, ldpsi, T, n_valid, ld_out, seed, call_id, t_off, n_off, n_total,
                                               omega, counts, list, (unsigned)total);
    PYGLM_LAUNCH_CHECK();
    // one resident wave (4 CTAs per SM): each warp then owns a long chunk of the list, and the idle tail at the end of
    // a chunk -- lanes waiting for the warp's last draw -- is paid once per warp, not once per 8 entries per lane
    const int small_blocks = blocks < 148 * 4 ? blocks : 148 * 4;
    pg_ig_small_kernel<<<small_blocks, 256, 0, stream>>>(psi, ldpsi, n_valid, ld_out, seed, call_id, t_off, n_off, n_total,
                                                   omega, counts, list);
    PYGLM_LAUNCH_CHECK();
    pg_ig_kernel<true><<<blocks, 256, 0, stream>>>(psi, ldpsi, n_valid, ld_out, seed, call_id, t_off, n_off, n_total,
                                                   omega, counts, list, (unsigned)total);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

// Test hook: out[e*count + i] = i-th uniform of stream (seed, call_id, elem0 + e).
extern "C" int pyglm_philox_uniforms(unsigned long long seed, unsigned call_id, unsigned long long elem0, int n_elem,
                                     int count, double* out, cudaStream_t stream) {
    PYGLM_CHECK_ARG(out && n_elem > 0 && count > 0, "pyglm_philox_uniforms: bad arguments");
    philox_unif_kernel<<<ceil_div_i(n_elem, 128), 128, 0, stream>>>(seed, call_id, elem0, n_elem, count, out);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}
