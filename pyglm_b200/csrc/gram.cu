// (3) Per-postsynaptic-neuron sufficient statistics  J_n = X~^T diag(omega_n) X~  for ALL local
// neurons in one launch, lower triangle only, on the FP64 tensor pipe (DMMA.8x8x4).
//
// Replaces regression.py:251-256 (XO = X*omega[:,None]; XO.T.dot(X); XO.sum(0); omega.sum()) executed
// once per neuron by models.py:169-171.  The reference forms an 8*T*NB-byte temporary and a dense
// dgemm per neuron; here the contraction is restated as ONE GEMM over all neurons,
//
//     J[(i,j), n] = sum_t  Z[t,(i,j)] * omega[t,n],      Z[t,(i,j)] = X~[t,i] * X~[t,j]   (i >= j)
//
// i.e. (pairs x T) . (T x neurons): the Khatri-Rao operand Z is never materialised -- each warp builds
// its 8x4 DMMA A-fragments in registers as the product of two shared-memory reads -- and only pairs on or
// below the diagonal are enumerated, which halves the flops against the reference's full dgemm.  The bias
// row/column of J (X^T omega and sum omega) is not a side reduction: X~ carries the ones column.
//
// Work decomposition
//   tile   = 8 rows i (one per warp) x 32 columns j (4 DMMA m-tiles per warp) of the (i,j) plane,
//            enumerated on the host for j0 <= i0 + 7 (pyglm_gram_tiles)
//   CTA    = (tile, block of 8*NT neurons, time slab); 8 warps; accumulators 4 x NT x (8x8) FP64 in registers
//   stream = time, in chunks of 32 bins: per bin the CTA stages 32 + 8 columns of X~ and 8*NT columns of
//            omega with 16-byte cp.async into a 4-stage ring (row pitch == 4 mod 16 doubles, which makes
//            every DMMA fragment load bank-conflict free)
//   reuse  = 2048*NT MACs per (40 + 8*NT) staged doubles  (NT=5: 16 MAC per byte) -> FP64-pipe bound.
// Time slabs (gridDim.z > 1) exist only to fill the 148 SMs when (tiles x neuron blocks) is small; their
// partial sums are combined in a fixed order by gram_reduce_kernel (deterministic, no atomics).
//
// The same kernel yields h = X~^T (y - 1/2) (regression.py:259-260): with omega := kappa the bias ROW of J
// is exactly h, so pyglm_xt_kappa runs it over the tiles of that one row.
#include "common.cuh"
#include <type_traits>

namespace {

constexpr int GR_KC = 32;      // time bins per stage
constexpr int GR_STAGES = 4;
constexpr int GR_TJ = 32;      // columns per tile
constexpr int GR_TI = 8;       // rows per tile (one per warp)
constexpr int GR_OFF_I = GR_TJ;
constexpr int GR_OFF_W = GR_TJ + GR_TI;

template <int NT>
struct GramCfg {
    static constexpr int LD = ((GR_OFF_W + 8 * NT + 15) / 16) * 16 + 4;   // == 4 (mod 16)
    static constexpr int CH = (GR_OFF_W + 8 * NT) / 2;                    // 16-byte chunks per staged row
    static constexpr size_t SMEM = (size_t)GR_STAGES * GR_KC * LD * sizeof(double);
};

template <int NT>
__global__ void __launch_bounds__(256, 2)
gram_kernel(const double* __restrict__ Xp, int ldx, const double* __restrict__ Om, int ldo, int T, int t_per_slab,
            const int2* __restrict__ tiles, int n_valid, double* __restrict__ J, long long stride_n, int ldj,
            int i_base, long long slab_stride) {
    using Cfg = GramCfg<NT>;
    extern __shared__ __align__(16) double gsm[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int2 tile = tiles[blockIdx.x];
    const int i0 = tile.x, j0 = tile.y;
    const int n0 = blockIdx.y * 8 * NT;
    // n-tiles of this neuron block that hold real neurons (CTA-uniform); the last block may be ragged
    const int ntcount = min(NT, (n_valid - n0 + 7) / 8);
    const int t_begin = blockIdx.z * t_per_slab;
    const int t_end = min(T, t_begin + t_per_slab);
    // m-tiles of this CTA that touch the lower triangle (CTA-uniform): columns j0+8m <= i0+7
    const int mcount = min(4, (i0 + 7 - j0) / 8 + 1);

    double acc[4][NT][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;

    const int nchunks = (t_end - t_begin + GR_KC - 1) / GR_KC;

    auto load_chunk = [&](int c) {
        if (c < nchunks) {
            double* dst = gsm + (size_t)(c % GR_STAGES) * GR_KC * Cfg::LD;
            const int tc = t_begin + c * GR_KC;
            for (int x = tid; x < GR_KC * Cfg::CH; x += 256) {
                const int r = x / Cfg::CH, cc = x - r * Cfg::CH;
                const int t = tc + r;
                int ok = (t < t_end) ? 16 : 0;                   // rows past the slab end are zero-filled
                const size_t tr = (size_t)min(t, t_end - 1);
                const double* src;
                if (cc < GR_TJ / 2) src = Xp + tr * ldx + j0 + 2 * cc;
                else if (cc < GR_OFF_W / 2) src = Xp + tr * ldx + i0 + 2 * (cc - GR_TJ / 2);
                else {
                    int col = n0 + 2 * (cc - GR_OFF_W / 2);      // ragged last block: columns past ldo are zero-filled
                    if (col >= ldo) { col = 0; ok = 0; }
                    src = Om + tr * ldo + col;
                }
                cp_async16(dst + r * Cfg::LD + 2 * cc, src, ok);
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < GR_STAGES - 1; ++s) load_chunk(s);

    // The hot loop exists twice: CTAs whose neuron block is full and whose tile lies wholly below the diagonal take
    // the predicate-free copy (runtime predicates around the DMMAs cost ~10% when measured); the ragged last
    // neuron block and the diagonal tiles take the predicated one.
    auto consume = [&](auto full_tag, int c) {
        constexpr bool FULL = decltype(full_tag)::value;
        const double* row = gsm + (size_t)(c % GR_STAGES) * GR_KC * Cfg::LD + q * Cfg::LD;
#pragma unroll 2
        for (int kk = 0; kk < GR_KC; kk += 4) {
            const double* rk = row + kk * Cfg::LD;
            const double xi = rk[GR_OFF_I + warp];
            double a[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) a[m] = xi * rk[8 * m + g];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (FULL || j < ntcount) {
                    const double b = rk[GR_OFF_W + 8 * j + g];
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        if (FULL || m < mcount) dmma884(acc[m][j][0], acc[m][j][1], a[m], b);
                }
            }
        }
    };
    const bool full = (ntcount == NT) && (mcount == 4);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<GR_STAGES - 2>();
        __syncthreads();
        load_chunk(c + GR_STAGES - 1);
        if (full) consume(std::true_type{}, c);
        else consume(std::false_type{}, c);
    }
    cp_async_wait<0>();

    // epilogue: this thread owns pair (i0+warp, j0+8m+g) x neurons n0+8j+2q+{0,1}
    const int i = i0 + warp;
    double* Jout = J + (size_t)blockIdx.z * slab_stride + (size_t)(i - i_base) * ldj;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (m >= mcount) continue;
        const int jc = j0 + 8 * m + g;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + 8 * j + 2 * q + e;
                if (n < n_valid) Jout[(size_t)n * stride_n + jc] = acc[m][j][e];
            }
        }
    }
}

// J[n][i][j] = sum_s part[s][n][i][j] over the written entries of every tile, slab order fixed.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ part, int nslabs, long long slab_stride, const int2* __restrict__ tiles,
                   double* __restrict__ J, long long stride_n, int ldj, int i_base) {
    const int2 tile = tiles[blockIdx.x];
    const int n = blockIdx.y;
    const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
    const int mcount = min(4, (tile.x + 7 - tile.y) / 8 + 1);
    if (c >= 8 * mcount) return;
    const size_t off = (size_t)n * stride_n + (size_t)(tile.x + r - i_base) * ldj + tile.y + c;
    double s = 0.0;
    for (int k = 0; k < nslabs; ++k) s += part[(size_t)k * slab_stride + off];
    J[off] = s;
}

// n-tiles (groups of 8 neurons) per CTA.  At most 5: gram_kernel<5> needs 123 registers, so two CTAs share an
// SM (16 warps keep the DMMA pipe ~85% busy; ncu profiles/r01); wider tiles fall to one CTA per SM and measure
// slower.  Blocks are balanced (n8 = 13 -> 5,5,3 handled by the runtime ntcount), so nothing is padded.
int pick_nt(int n8) {
    const int nblocks = (n8 + 4) / 5;
    return (n8 + nblocks - 1) / nblocks;
}

template <int NT>
int launch_gram(dim3 grid, const double* Xp, int ldx, const double* Om, int ldo, int T, int t_per_slab,
                const int2* tiles, int n_valid, double* out, long long stride_n, int ldj, int i_base,
                long long slab_stride, cudaStream_t stream) {
    PYGLM_CUDA(cudaFuncSetAttribute(gram_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GramCfg<NT>::SMEM));
    gram_kernel<NT><<<grid, 256, GramCfg<NT>::SMEM, stream>>>(Xp, ldx, Om, ldo, T, t_per_slab, tiles, n_valid, out,
                                                              stride_n, ldj, i_base, slab_stride);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

}  // namespace

// Host-side enumeration of the (i0, j0) tiles: mode 0 = every tile touching the lower triangle of a D x D
// matrix; mode 1 = only the tiles of the row block holding row D-1 (the bias row, used for h).
// Writes 2*count ints (i0, j0 interleaved) when tiles != NULL and capacity suffices; returns count.
extern "C" int pyglm_gram_tiles(int D, int mode, int* tiles, int capacity) {
    int count = 0;
    const int i_first = (mode == 1) ? ((D - 1) / GR_TI) * GR_TI : 0;
    for (int i0 = i_first; i0 < D; i0 += GR_TI)
        for (int j0 = 0; j0 < D && j0 <= i0 + GR_TI - 1; j0 += GR_TJ) {
            if (tiles && count < capacity) { tiles[2 * count] = i0; tiles[2 * count + 1] = j0; }
            ++count;
        }
    return count;
}

// Number of time slabs the Gram launch should use for (ntiles, n_valid, T): 1 when the tile x neuron-block grid
// already gives the 148 SMs several waves, more when it does not (each slab >= 1024 bins).
extern "C" int pyglm_gram_slabs(int ntiles, int n_valid, int T) {
    const int n8 = (n_valid + 7) / 8;
    const int nt = pick_nt(n8);
    const long long ctas = (long long)ntiles * ((n8 + nt - 1) / nt);
    long long s = (148LL * 8 + ctas - 1) / ctas;
    long long smax = T / 1024 > 0 ? T / 1024 : 1;
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    return (int)s;
}

// J[n*stride_n + (i - i_base)*ldj + j] = sum_t Xp[t,i] * Xp[t,j] * Om[t,n]   for the listed tiles, n < n_valid.
//   Xp (T x ldx) padded design matrix; Om (T x ldo) weights (omega, or kappa for h); tiles: device int2[ntiles].
//   nslabs > 1 needs workspace of nslabs * n_valid * stride_n doubles.  Entries outside the listed tiles are
//   left untouched (in particular the strict upper triangle beyond the diagonal tiles).
extern "C" int pyglm_weighted_gram(const double* Xp, int ldx, int T, const double* Om, int ldo, int n_valid,
                                   const int* tiles, int ntiles, int nslabs, double* J, long long stride_n, int ldj,
                                   int i_base, double* workspace, cudaStream_t stream) {
    PYGLM_CHECK_ARG(Xp && Om && tiles && J, "pyglm_weighted_gram: null pointer");
    PYGLM_CHECK_ARG(T > 0 && n_valid > 0 && ntiles > 0 && nslabs > 0, "pyglm_weighted_gram: T, n_valid, ntiles, nslabs must be positive");
    PYGLM_CHECK_ARG(ldx % 32 == 0 && ldo % 2 == 0 && ldo >= ((n_valid + 63) / 64) * 64,
                    "pyglm_weighted_gram: ldx=%d must be a multiple of 32, ldo=%d even and >= n rounded up to 64", ldx, ldo);
    PYGLM_CHECK_ARG(((uintptr_t)Xp & 15) == 0 && ((uintptr_t)Om & 15) == 0, "pyglm_weighted_gram: operands must be 16-byte aligned");
    PYGLM_CHECK_ARG(nslabs == 1 || workspace, "pyglm_weighted_gram: nslabs > 1 needs a workspace");
    const int n8 = (n_valid + 7) / 8;
    const int NT = pick_nt(n8);
    int t_per_slab = ((T + nslabs - 1) / nslabs + GR_KC - 1) / GR_KC * GR_KC;
    nslabs = (T + t_per_slab - 1) / t_per_slab;
    dim3 grid(ntiles, (n8 + NT - 1) / NT, nslabs);
    PYGLM_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "pyglm_weighted_gram: grid too large");
    double* out = (nslabs == 1) ? J : workspace;
    const long long slab_stride = (long long)n_valid * stride_n;
    const int2* t2 = reinterpret_cast<const int2*>(tiles);
    int rc = PYGLM_ERR_INVALID;
#define GR_CASE(NTV) \
    case NTV: rc = launch_gram<NTV>(grid, Xp, ldx, Om, ldo, T, t_per_slab, t2, n_valid, out, stride_n, ldj, i_base, slab_stride, stream); break;
    switch (NT) { GR_CASE(1) GR_CASE(2) GR_CASE(3) GR_CASE(4) GR_CASE(5) }
#undef GR_CASE
    if (rc) return rc;
    if (nslabs > 1) {
        dim3 rgrid(ntiles, n_valid);
        gram_reduce_kernel<<<rgrid, 256, 0, stream>>>(workspace, nslabs, slab_stride, t2, J, stride_n, ldj, i_base);
        PYGLM_LAUNCH_CHECK();
    }
    return PYGLM_OK;
}
