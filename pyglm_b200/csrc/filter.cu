// (1) Filtered-spike-train build: X[t,n,b] = sum_{l=1..L} basis[l-1,b] * S[t-l,n]  (strictly causal,
// zero history before t = 0), written straight into the padded design-matrix layout
//   Xp (T x ldx) row-major:  columns [0, N*B) = X in n-major / b-minor order (regression.py:173-180),
//                            column N*B = 1.0 (the affine / bias column), columns above = 0.
// Replaces pyglm/utils/basis.py:5-34 (per-b scipy.signal.fftconvolve, truncated to T, clipped at 0
// when basis >= 0 and S >= 0).  A direct L-tap sum: L = 100 taps makes the FFT pointless on a GPU.
// HBM-bound: 8*T*N bytes read + 8*T*ldx bytes written; every S element is staged once per CTA in
// shared memory and reused L*B times from there.
#include "common.cuh"

namespace {

constexpr int FILT_NX = 32;   // neurons per CTA (one per lane: coalesced S reads)
constexpr int FILT_NY = 8;    // time rows per CTA pass
constexpr int FILT_TT = 128;  // output time bins per CTA
constexpr int FILT_RB = 4;    // consecutive time bins per thread (register block)
constexpr int FILT_BB = 4;    // basis functions handled by the register-blocked path

__global__ void __launch_bounds__(FILT_NX * FILT_NY)
filter_kernel(const double* __restrict__ S, const double* __restrict__ basis, int T, int N, int L, int B,
              int clip, double* __restrict__ Xp, int ldx) {
    extern __shared__ double smem[];
    double* Ss = smem;                                   // (FILT_TT + L) x FILT_NX history window
    double* bs = smem + (size_t)(FILT_TT + L) * FILT_NX; // L x B basis
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int n = blockIdx.y * FILT_NX + tx;
    const int t0 = blockIdx.x * FILT_TT;   // time on grid.x: T/128 may exceed the 65535 limit of grid.y

    for (int r = ty; r < FILT_TT + L; r += FILT_NY) {
        int t = t0 - L + r;
        Ss[r * FILT_NX + tx] = (t >= 0 && t < T && n < N) ? S[(size_t)t * N + n] : 0.0;
    }
    for (int i = ty * FILT_NX + tx; i < L * B; i += FILT_NX * FILT_NY) bs[i] = basis[i];
    __syncthreads();
    if (n >= N) return;

    // Register blocking: a thread produces FILT_RB consecutive time bins of its neuron for up to FILT_BB basis functions
    // at once.  Going from lag l to l+1 the window of a bin is the previous bin's, so one new spike value and B
    // (warp-broadcast) basis values are loaded per 4*B FMAs -- the first version loaded two operands per FMA and ran
    // at the shared-memory bandwidth.  Every output still accumulates its taps in the order l = 1..L.
    if (B <= FILT_BB) {
        for (int r = ty * FILT_RB; r < FILT_TT; r += FILT_NY * FILT_RB) {
            if (t0 + r >= T) break;
            double acc[FILT_RB][FILT_BB];
#pragma unroll
            for (int k = 0; k < FILT_RB; ++k)
#pragma unroll
                for (int b = 0; b < FILT_BB; ++b) acc[k][b] = 0.0;
            const double* win = Ss + (size_t)(r + L) * FILT_NX + tx;      // win[(k - l) * NX] = S[t0 + r + k - l, n]
            double w[FILT_RB];                                             // w[k] = S[t0 + r + k - l, n], lag l = 1
#pragma unroll
            for (int k = 0; k < FILT_RB; ++k) w[k] = win[(k - 1) * FILT_NX];
#pragma unroll 4
            for (int l = 1; l <= L; ++l) {
#pragma unroll
                for (int b = 0; b < FILT_BB; ++b) {
                    if (b < B) {
                        const double bv = bs[(l - 1) * B + b];
#pragma unroll
                        for (int k = 0; k < FILT_RB; ++k) acc[k][b] = fma(bv, w[k], acc[k][b]);
                    }
                }
#pragma unroll
                for (int k = FILT_RB - 1; k >= 1; --k) w[k] = w[k - 1];   // lag l+1: bin k sees what bin k-1 saw
                if (l < L) w[0] = win[-(l + 1) * FILT_NX];
            }
#pragma unroll
            for (int k = 0; k < FILT_RB; ++k) {
                const int t = t0 + r + k;
                if (t < T) {
#pragma unroll
                    for (int b = 0; b < FILT_BB; ++b)
                        if (b < B) Xp[(size_t)t * ldx + (size_t)n * B + b] = clip ? fmax(acc[k][b], 0.0) : acc[k][b];
                }
            }
        }
        return;
    }
    for (int r = ty; r < FILT_TT; r += FILT_NY) {
        int t = t0 + r;
        if (t >= T) break;
        const double* win = Ss + (size_t)(r + L) * FILT_NX + tx;   // win[-l*NX] = S[t-l, n]
        for (int b = 0; b < B; ++b) {
            double acc = 0.0;
            for (int l = 1; l <= L; ++l) acc = fma(bs[(l - 1) * B + b], win[-l * FILT_NX], acc);
            if (clip) acc = fmax(acc, 0.0);
            Xp[(size_t)t * ldx + (size_t)n * B + b] = acc;
        }
    }
}

// bias column = 1, padding columns = 0
__global__ void pad_columns_kernel(double* __restrict__ Xp, int T, int NB, int ldx) {
    int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= T) return;
    for (int c = NB + threadIdx.x; c < ldx; c += blockDim.x) Xp[(size_t)t * ldx + c] = (c == NB) ? 1.0 : 0.0;
}

// dense (T x NB) -> padded (T x ldx)
__global__ void pack_design_kernel(const double* __restrict__ X, int T, int NB, double* __restrict__ Xp, int ldx) {
    int t = blockIdx.x;
    for (int c = threadIdx.x; c < ldx; c += blockDim.x)
        Xp[(size_t)t * ldx + c] = (c < NB) ? X[(size_t)t * NB + c] : ((c == NB) ? 1.0 : 0.0);
}

// padded (T x ldx) -> dense (T x NB)
__global__ void unpack_design_kernel(const double* __restrict__ Xp, int T, int NB, int ldx, double* __restrict__ X) {
    int t = blockIdx.x;
    for (int c = threadIdx.x; c < NB; c += blockDim.x) X[(size_t)t * NB + c] = Xp[(size_t)t * ldx + c];
}

}  // namespace

extern "C" int pyglm_filter_spikes(const double* S, const double* basis, int T, int N, int L, int B, int clip,
                                   double* Xp, int ldx, cudaStream_t stream) {
    PYGLM_CHECK_ARG(S && basis && Xp, "pyglm_filter_spikes: null pointer");
    PYGLM_CHECK_ARG(T > 0 && N > 0 && L > 0 && B > 0, "pyglm_filter_spikes: T,N,L,B must be positive");
    PYGLM_CHECK_ARG(ldx >= N * B + 1, "pyglm_filter_spikes: ldx=%d < N*B+1=%d", ldx, N * B + 1);
    size_t smem = ((size_t)(FILT_TT + L) * FILT_NX + (size_t)L * B) * sizeof(double);
    PYGLM_CHECK_ARG(smem <= 227 * 1024, "pyglm_filter_spikes: basis too long (L=%d, B=%d) for the shared-memory window", L, B);
    PYGLM_CUDA(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div_i(T, FILT_TT), ceil_div_i(N, FILT_NX)), block(FILT_NX, FILT_NY);
    filter_kernel<<<grid, block, smem, stream>>>(S, basis, T, N, L, B, clip, Xp, ldx);
    PYGLM_LAUNCH_CHECK();
    dim3 pblock(32, 8);
    pad_columns_kernel<<<ceil_div_i(T, 8), pblock, 0, stream>>>(Xp, T, N * B, ldx);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

extern "C" int pyglm_pack_design(const double* X, int T, int NB, double* Xp, int ldx, cudaStream_t stream) {
    PYGLM_CHECK_ARG(X && Xp && T > 0 && NB > 0 && ldx >= NB + 1, "pyglm_pack_design: bad arguments");
    pack_design_kernel<<<T, 128, 0, stream>>>(X, T, NB, Xp, ldx);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

extern "C" int pyglm_unpack_design(const double* Xp, int T, int NB, int ldx, double* X, cudaStream_t stream) {
    PYGLM_CHECK_ARG(X && Xp && T > 0 && NB > 0 && ldx >= NB + 1, "pyglm_unpack_design: bad arguments");
    unpack_design_kernel<<<T, 128, 0, stream>>>(Xp, T, NB, ldx, X);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}
