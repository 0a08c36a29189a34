// Forward simulation of the Bernoulli network GLM on the device.
//
// Replaces the Python loop of pyglm/models.py:98-151 (generate) with pyglm/regression.py:528-541 (rvs):
//     for t:  X[t]   = Y[t-L:t]^T flipud(basis)                (zero history before t = 0)
//             psi[t] = W X[t] + b                              (W = model.weights reshaped (N, N*B), models.py:124)
//             Y[t]   = rand(N) < logistic(psi[t])
// The recursion is sequential in time, so the whole simulation is ONE persistent thread-block cluster: the
// postsynaptic neurons are split over the CTAs of the cluster; each CTA keeps its rows of W in shared memory (or
// streams them from L2 when they do not fit), a ring of the last L spikes of its own neurons, and a double-buffered
// copy of the full regressor vector x[t].  Per step a CTA
//   A. computes psi for its neurons (one warp per neuron, fixed butterfly order) and draws the spikes
//      (Philox stream keyed by (seed, call_id, t*N + n): independent of the cluster shape),
//   B. forms its slice of x[t+1] with the SAME fma order as filter_kernel (filter.cu) -- so the X returned equals
//      convolve_with_basis(Y) bit for bit -- writes it to the padded design layout in HBM and into every CTA's
//      x buffer through distributed shared memory,
//   C. meets the other CTAs at ONE cluster barrier.
// Latency-bound by construction (T dependent steps); HBM traffic is the 8*T*(ldx + N) bytes it writes.
#include "common.cuh"
#include "philox.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int GEN_THREADS = 512;

struct GenArgs {
    const double* W;       // (N, N*B) row-major: row n = postsynaptic neuron
    const double* bias;    // (N)
    const double* basis;   // (L, B): row 0 = lag 1
    int N, B, L, npc;      // npc = neurons per CTA
    long long T;
    unsigned long long seed;
    unsigned call_id;
    double* Xp; int ldx;   // (T, ldx): columns [0, N*B) written here; bias / padding columns by the caller
    double* Y;             // (T, N)
    double* U;             // optional (T, N): the uniforms (Bernoulli) or standard normals (Gaussian) used (test hook)
    int w_in_smem;
    double gauss_sd;       // < 0: Bernoulli spikes; >= 0: Gaussian observations y = psi + gauss_sd * z (regression.py:417)
};

__global__ void __launch_bounds__(GEN_THREADS, 1)
generate_kernel(const GenArgs A) {
    extern __shared__ __align__(16) double gsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int N = A.N, B = A.B, L = A.L, NB = A.N * A.B;
    const int NBpad = (NB + 1) & ~1;
    const int n0 = min(N, rank * A.npc), n1 = min(N, n0 + A.npc);
    const int nown = n1 - n0, ownB = nown * B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = GEN_THREADS / 32;

    double* xfull = gsm;                                   // [2][NBpad]
    double* bs = xfull + 2 * NBpad;                        // [L*B]
    double* Ws = bs + ((L * B + 1) & ~1);                  // [npc][NB] when w_in_smem
    const bool gauss = A.gauss_sd >= 0.0;
    double* yd = Ws + (A.w_in_smem ? (size_t)A.npc * NB : 0);           // Gaussian: [L][npc] ring of the last L values
    unsigned char* yh = reinterpret_cast<unsigned char*>(yd + (gauss ? (size_t)L * A.npc : 0));   // Bernoulli: [L][npc] ring

    for (int i = tid; i < 2 * NBpad; i += GEN_THREADS) xfull[i] = 0.0;
    for (int i = tid; i < L * B; i += GEN_THREADS) bs[i] = A.basis[i];
    for (int i = tid; i < L * A.npc; i += GEN_THREADS) { if (gauss) yd[i] = 0.0; else yh[i] = 0; }
    if (A.w_in_smem)
        for (int i = tid; i < nown * NB; i += GEN_THREADS) Ws[i] = A.W[(size_t)n0 * NB + i];
    cluster.sync();

    for (long long t = 0; t < A.T; ++t) {
        const double* x = xfull + (t & 1) * NBpad;
        // A. psi and the spike draw for the own neurons
        for (int nl = warp; nl < nown; nl += NW) {
            const double* wr = A.w_in_smem ? Ws + (size_t)nl * NB : A.W + (size_t)(n0 + nl) * NB;
            double acc = 0.0;
            for (int k = lane; k < NB; k += 32) acc = fma(wr[k], x[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const int n = n0 + nl;
                const double psi = acc + A.bias[n];
                PhiloxStream r;
                r.seed(A.seed, A.call_id, (unsigned long long)t * (unsigned long long)N + (unsigned long long)n);
                if (gauss) {
                    const double z = r.norm();
                    const double y = psi + A.gauss_sd * z;
                    yd[(int)(t % L) * A.npc + nl] = y;
                    A.Y[(size_t)t * N + n] = y;
                    if (A.U) A.U[(size_t)t * N + n] = z;
                } else {
                    const double u = r.unif();
                    const double p = 1.0 / (1.0 + exp(-psi));
                    const int y = u < p;
                    yh[(int)(t % L) * A.npc + nl] = (unsigned char)y;
                    A.Y[(size_t)t * N + n] = (double)y;
                    if (A.U) A.U[(size_t)t * N + n] = u;
                }
            }
        }
        __syncthreads();
        // B. own slice of x[t+1]: sum_{l=1..L} basis[l-1,b] * Y[t+1-l, n] in filter_kernel's order (lag 1 first);
        //    zero spikes are skipped -- fma(c, 0, acc) == acc exactly, so the result is bit-identical
        if (t + 1 < A.T) {
            double* xn = xfull + ((t + 1) & 1) * NBpad;
            for (int j = tid; j < ownB; j += GEN_THREADS) {
                const int nl = j / B, b = j - nl * B;
                double acc = 0.0;
                int slot = (int)(t % L);                   // slot of time t+1-l for l = 1
                const int lmax = (int)min((long long)L, t + 1);
                if (gauss) {
                    for (int l = 1; l <= lmax; ++l) {
                        acc = fma(bs[(l - 1) * B + b], yd[slot * A.npc + nl], acc);
                        slot = (slot == 0) ? L - 1 : slot - 1;
                    }
                } else {
                    for (int l = 1; l <= lmax; ++l) {
                        if (yh[slot * A.npc + nl]) acc = fma(bs[(l - 1) * B + b], 1.0, acc);
                        slot = (slot == 0) ? L - 1 : slot - 1;
                    }
                }
                A.Xp[(size_t)(t + 1) * A.ldx + (size_t)n0 * B + j] = acc;
                for (int c = 0; c < C; ++c) cluster.map_shared_rank(xn, c)[n0 * B + j] = acc;
            }
        }
        // C. every CTA has x[t+1] (and nobody still reads x[t-1]'s buffer) after this barrier
        cluster.sync();
    }
}

}  // namespace

// Simulate T bins.  Wm (N x N*B), bias (N), basis (L x B) device doubles; Xp (T x ldx) with ldx >= N*B+1: the kernel
// writes columns [0, N*B) of rows 1..T-1 (row 0 is the zero-history row: the caller provides Xp zeroed with the bias
// column set); Y (T x N) receives 0/1; U (T x N) or NULL receives the uniforms.  gauss_sd >= 0 selects Gaussian
// observations y = psi + gauss_sd * z (SparseGaussianRegression.rvs, regression.py:406-417; U then receives z).
// pyglm/models.py:98-151.
extern "C" int pyglm_generate(const double* Wm, const double* bias, const double* basis, int N, int B, int L, long long T,
                              unsigned long long seed, unsigned call_id, double gauss_sd, double* Xp, int ldx, double* Y,
                              double* U, cudaStream_t stream) {
    PYGLM_CHECK_ARG(Wm && bias && basis && Xp && Y, "pyglm_generate: null pointer");
    PYGLM_CHECK_ARG(N > 0 && B > 0 && L > 0 && T > 0 && ldx >= N * B + 1, "pyglm_generate: bad shape");
    int C = 1;
    while (C < 8 && N >= 4 * C) C *= 2;                   // at least two neurons per CTA
    GenArgs A;
    A.W = Wm; A.bias = bias; A.basis = basis; A.N = N; A.B = B; A.L = L; A.T = T; A.seed = seed; A.call_id = call_id;
    A.Xp = Xp; A.ldx = ldx; A.Y = Y; A.U = U; A.gauss_sd = gauss_sd;
    A.npc = (N + C - 1) / C;
    const int NB = N * B, NBpad = (NB + 1) & ~1;
    const size_t base = ((size_t)2 * NBpad + ((L * B + 1) & ~1)) * sizeof(double) + (size_t)L * A.npc + 16 +
                        (gauss_sd >= 0.0 ? (size_t)L * A.npc * sizeof(double) : 0);
    const size_t wbytes = (size_t)A.npc * NB * sizeof(double);
    PYGLM_CHECK_ARG(base <= 200 * 1024, "pyglm_generate: N*B=%d, L=%d too large for the shared-memory state", NB, L);
    A.w_in_smem = (base + wbytes <= 220 * 1024) ? 1 : 0;
    const size_t smem = base + (A.w_in_smem ? wbytes : 0);
    PYGLM_CUDA(cudaFuncSetAttribute(generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, 1, 1);
    cfg.blockDim = dim3(GEN_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PYGLM_CUDA(cudaLaunchKernelEx(&cfg, generate_kernel, A));
    return PYGLM_OK;
}
