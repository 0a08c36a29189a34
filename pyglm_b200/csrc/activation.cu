// (5) psi = X~ . W~ for ALL local postsynaptic neurons at once, with the consumer fused into the
// epilogue: store psi | Bernoulli log-likelihood reduction | logistic mean.
//
// Replaces N separate dgemv's of regression.py:195-201 (activation), :491-494 (log_likelihood,
// looped per neuron at models.py:93-94) and :524-526 (mean, models.py:159-161) by one FP64 GEMM
//   psi (T x n) = Xp (T x K) . Wt (K x n),   K = N*B+1 rounded up to 4 (bias handled as the ones column)
// on the FP64 tensor pipe (DMMA.8x8x4).  CTA tile: 128 time bins x (8*NT) neurons, 8 warps, each
// warp 16 bins x 8*NT neurons; K is streamed in 16-column slabs with double-buffered cp.async.
// Strict FP64: 2*T*K*n flop against 8*T*K bytes -> bound by the FP64 pipe, not HBM.
#include "common.cuh"

namespace {

constexpr int ACT_BM = 128;  // time bins per CTA
constexpr int ACT_KC = 16;   // K slab
constexpr int ACT_LDA = 20;  // smem row pitch of the X slab: == 4 (mod 16) -> conflict-free DMMA fragment loads

enum { ACT_STORE_PSI = 0, ACT_LOGLIK = 1, ACT_MEAN = 2 };

__device__ __forceinline__ double softplus(double x) {
    // log(1 + e^x) without the overflow of regression.py:494 for x > 709 (value-preserving elsewhere)
    return fmax(x, 0.0) + log1p(exp(-fabs(x)));
}

template <int NT, int MODE>
__global__ void __launch_bounds__(256)
activation_kernel(const double* __restrict__ Xp, int ldx, const double* __restrict__ Wt, int ldw, int T, int K,
                  int n_valid, const double* __restrict__ Y, int ldy, int y_col0, double* __restrict__ out,
                  int ldo, double* __restrict__ partials) {
    constexpr int LDW = 8 * NT + 4;
    extern __shared__ __align__(16) double act_smem[];
    double (*Xs)[ACT_BM * ACT_LDA] = reinterpret_cast<double (*)[ACT_BM * ACT_LDA]>(act_smem);
    double (*Ws)[ACT_KC * LDW] = reinterpret_cast<double (*)[ACT_KC * LDW]>(act_smem + 2 * ACT_BM * ACT_LDA);
    __shared__ double red[32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int t0 = blockIdx.x * ACT_BM;
    const int n0 = blockIdx.y * 8 * NT;

    double acc[2][NT][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;

    const int nslab = (K + ACT_KC - 1) / ACT_KC;

    auto load_slab = [&](int s, int buf) {
        const int k0 = s * ACT_KC;
        // X slab: 128 rows x 16 doubles = 1024 16-byte chunks, 4 per thread
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int c = tid + it * 256;
            int r = c >> 3, cc = (c & 7) * 2;
            int t = min(t0 + r, T - 1);                  // clamp: rows beyond T are never stored
            cp_async16(&Xs[buf][r * ACT_LDA + cc], Xp + (size_t)t * ldx + k0 + cc, 16);
        }
        // W slab: 16 rows x 8*NT doubles
        for (int c = tid; c < ACT_KC * 4 * NT; c += 256) {
            int r = c / (4 * NT), cc = (c % (4 * NT)) * 2;
            cp_async16(&Ws[buf][r * LDW + cc], Wt + (size_t)(k0 + r) * ldw + n0 + cc, 16);
        }
        cp_async_commit();
    };

    load_slab(0, 0);
    for (int s = 0; s < nslab; ++s) {
        const int buf = s & 1;
        if (s + 1 < nslab) { load_slab(s + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* xs = &Xs[buf][(warp * 16) * ACT_LDA];
        const double* ws = &Ws[buf][0];
#pragma unroll
        for (int kk = 0; kk < ACT_KC; kk += 4) {
            double a0 = xs[g * ACT_LDA + kk + q];
            double a1 = xs[(g + 8) * ACT_LDA + kk + q];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                double b = ws[(kk + q) * LDW + 8 * j + g];
                dmma884(acc[0][j][0], acc[0][j][1], a0, b);
                dmma884(acc[1][j][0], acc[1][j][1], a1, b);
            }
        }
        __syncthreads();
    }

    // epilogue: thread owns rows t0 + warp*16 + {g, g+8}, neurons n0 + 8j + 2q + {0,1}
    double ll = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int t = t0 + warp * 16 + g + 8 * m;
        if (t >= T) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + 8 * j + 2 * q + e;
                const double psi = acc[m][j][e];
                if (MODE == ACT_STORE_PSI) {
                    out[(size_t)t * ldo + n] = psi;          // padded neurons hold psi = 0
                } else if (n < n_valid) {
                    if (MODE == ACT_LOGLIK) {
                        const double y = Y[(size_t)t * ldy + y_col0 + n];
                        ll += y * psi - softplus(psi);
                    } else {
                        out[(size_t)t * ldo + n] = 1.0 / (1.0 + exp(-psi));
                    }
                }
            }
        }
    }
    if (MODE == ACT_LOGLIK) {
        double tot = block_sum(ll, red);
        if (tid == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

// fixed-order (deterministic) sum of per-CTA partials
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ p, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += p[i];
    double tot = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = tot;
}

// Neuron tile (1..8 groups of 8 neurons).  Every neuron tile streams the whole X slab of its 128 bins again, so the cost
// of a tiling is (number of tiles) x (tile width + the slab traffic, worth about two groups of DMMA work), not the
// padding alone: 13 groups (n = 100) run as 2 tiles of 7, not as 13 tiles of 1 (measured 0.75 -> 0.36 ms at T = 1e5, D = 401).
// Ties go to the wider tile.  The accumulation order over K of an element does not depend on the tile width.
int pick_nt(int n8) {
    int best = 1, best_cost = 1 << 30;
    for (int nt = 1; nt <= 8; ++nt) {
        int cost = ((n8 + nt - 1) / nt) * (nt + 2);
        if (cost <= best_cost) { best_cost = cost; best = nt; }
    }
    return best;
}

template <int NT, int MODE>
int launch_activation_nt(dim3 grid, const double* Xp, int ldx, const double* Wt, int ldw, int T, int K, int n_valid,
                         const double* Y, int ldy, int y_col0, double* out, int ldo, double* partials,
                         cudaStream_t stream) {
    size_t smem = (size_t)(2 * ACT_BM * ACT_LDA + 2 * ACT_KC * (8 * NT + 4)) * sizeof(double);
    PYGLM_CUDA(cudaFuncSetAttribute(activation_kernel<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    activation_kernel<NT, MODE><<<grid, 256, smem, stream>>>(Xp, ldx, Wt, ldw, T, K, n_valid, Y, ldy, y_col0, out, ldo, partials);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

template <int MODE>
int launch_activation(const double* Xp, int ldx, const double* Wt, int ldw, int T, int K, int n_valid,
                      const double* Y, int ldy, int y_col0, double* out, int ldo, double* partials,
                      int* nparts, cudaStream_t stream) {
    const int n8 = (n_valid + 7) / 8;
    const int NT = pick_nt(n8);
    dim3 grid(ceil_div_i(T, ACT_BM), ceil_div_i(n8, NT));
    if (nparts) *nparts = grid.x * grid.y;
#define ACT_CASE(NTV) \
    case NTV: return launch_activation_nt<NTV, MODE>(grid, Xp, ldx, Wt, ldw, T, K, n_valid, Y, ldy, y_col0, out, ldo, partials, stream);
    switch (NT) {
        ACT_CASE(1) ACT_CASE(2) ACT_CASE(3) ACT_CASE(4) ACT_CASE(5) ACT_CASE(6) ACT_CASE(7) ACT_CASE(8)
    }
#undef ACT_CASE
    return PYGLM_ERR_INVALID;
}

int check_activation_args(const char* who, const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n) {
    PYGLM_CHECK_ARG(Xp && Wt, "%s: null pointer", who);
    PYGLM_CHECK_ARG(T > 0 && D > 0 && n > 0, "%s: T, D, n must be positive", who);
    PYGLM_CHECK_ARG(ldx % 16 == 0 && ldx >= ((D + 15) / 16) * 16, "%s: ldx=%d must be a multiple of 16 and >= D=%d rounded up to 16", who, ldx, D);
    PYGLM_CHECK_ARG(ldw % 8 == 0 && ldw >= ((n + 63) / 64) * 64, "%s: ldw=%d must be a multiple of 8 and >= n=%d rounded up to 64", who, ldw, n);
    PYGLM_CHECK_ARG(((uintptr_t)Xp & 15) == 0 && ((uintptr_t)Wt & 15) == 0, "%s: pointers must be 16-byte aligned", who);
    return PYGLM_OK;
}

}  // namespace

// psi[t, j] = sum_d Xp[t, d] * Wt[d, j]  for j < n (columns n..ldo-1 of psi receive 0 up to the tile edge).
// Wt is (ldx x ldw): row d < N*B holds a[j,d/B]*W[j,d/B,d%B], row N*B the bias, other rows 0.
extern "C" int pyglm_activation(const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n,
                                double* psi, int ldpsi, cudaStream_t stream) {
    int rc = check_activation_args("pyglm_activation", Xp, ldx, Wt, ldw, T, D, n);
    if (rc) return rc;
    PYGLM_CHECK_ARG(psi && ldpsi % 2 == 0 && ldpsi >= ((n + 63) / 64) * 64, "pyglm_activation: ldpsi=%d must be even and >= n rounded up to 64", ldpsi);
    return launch_activation<ACT_STORE_PSI>(Xp, ldx, Wt, ldw, T, ((D + 15) / 16) * 16, n, nullptr, 0, 0, psi, ldpsi, nullptr, nullptr, stream);
}

// ll[0] = sum_{t, j<n} Y[t, y_col0+j] * psi[t,j] - log(1 + exp(psi[t,j]));  workspace: >= ceil(T/128)*ceil(n/8) doubles.
extern "C" int pyglm_loglik(const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n,
                            const double* Y, int ldy, int y_col0, double* ll, double* workspace,
                            cudaStream_t stream) {
    int rc = check_activation_args("pyglm_loglik", Xp, ldx, Wt, ldw, T, D, n);
    if (rc) return rc;
    PYGLM_CHECK_ARG(Y && ll && workspace, "pyglm_loglik: null pointer");
    int nparts = 0;
    rc = launch_activation<ACT_LOGLIK>(Xp, ldx, Wt, ldw, T, ((D + 15) / 16) * 16, n, Y, ldy, y_col0, nullptr, 0, workspace, &nparts, stream);
    if (rc) return rc;
    reduce_partials_kernel<<<1, 1024, 0, stream>>>(workspace, nparts, ll);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

// mu[t, j] = logistic(psi[t, j]) for j < n, written with row pitch ldmu.
extern "C" int pyglm_means(const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n,
                           double* mu, int ldmu, cudaStream_t stream) {
    int rc = check_activation_args("pyglm_means", Xp, ldx, Wt, ldw, T, D, n);
    if (rc) return rc;
    PYGLM_CHECK_ARG(mu && ldmu >= n, "pyglm_means: ldmu=%d < n=%d", ldmu, n);
    return launch_activation<ACT_MEAN>(Xp, ldx, Wt, ldw, T, ((D + 15) / 16) * 16, n, nullptr, 0, 0, mu, ldmu, nullptr, nullptr, stream);
}
