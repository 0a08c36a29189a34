// Pieces shared by the two spike-and-slab kernels (spike_slab.cu: one CTA per neuron, P in L2; spike_slab_dsm.cu: one
// thread-block cluster per neuron, P in distributed shared memory).
#pragma once
#include "common.cuh"

namespace pyglm_ss {

constexpr int SS_BMAX = 16;

struct SpikeSlabArgs {
    int N, B, D, n_loc;
    const double* J; long long stride_n; int ldj;     // likelihood J (lower triangle valid), per local neuron
    const double* h; int ldh;                          // likelihood h
    const double* J0w;                                 // (n_loc, N, B, B) prior precision blocks
    const double* h0w;                                 // (n_loc, N, B)
    const double* J0b; const double* h0b;              // (n_loc,)
    const double* cprior;                              // (n_loc, N)   1/2 log|J0_m| - 1/2 h0_m^T J0_m^-1 h0_m
    const double* logit_rho;                           // (n_loc, N)   log rho - log(1 - rho)
    const int* perm;                                   // (n_loc, N)
    const double* us;                                  // (n_loc, N)
    const double* z;                                   // (n_loc, ldz) standard normals keyed by coordinate
    int ldz;
    const unsigned char* do_scan;                      // (n_loc,) 0 -> keep a as given (deterministic sparsity)
    unsigned char* a;                                  // (n_loc, N) in/out
    double* W;                                         // (n_loc, N, B) out
    double* bias;                                      // (n_loc,) out
    double* P;                                         // workspace (n_loc, D, D)
    double* logodds;                                   // optional (n_loc, N): log-odds per scan step
    double* ml;                                        // optional (n_loc,): marginal likelihood of the final a
    int* status;                                       // (n_loc,) 0 ok, 1 = a Schur complement lost positive definiteness
    int debug;                                         // PYGLM_SS_DEBUG=1: CTA 0 prints its cycles per phase (profiling aid)
    int la_G;                                          // slots of the scan's lookahead table (0 = off), set by the launcher
};

template <int B>
struct SmallSolve {
    double L[B][B], il[B], y[B];      // Cholesky factor, 1/diag, L^-1 r
    double G[B][B], gr[B], xm[B];     // (L L^T)^-1, G r, L^-T z
    double dpost;
};

// Cholesky of the lower triangle of S (BS x BS) and the quadratic form: dpost = sgn 1/2 log|S| + 1/2 r^T S^-1 r
// (NaN when S is not positive definite).
// small_factor_parts leaves the transcendental to the caller: det = |S| (<= 0 or NaN when S is not positive definite),
// q = r^T S^-1 r; small_factor is the complete form.
template <int B, int BS>
__device__ __forceinline__ bool small_factor_parts(SmallSolve<B>& w, const double (&S)[B][B], const double (&r)[B],
                                                   double& det_out, double& q_out) {
    double det = 1.0, q = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < BS; ++j) {
        double d = S[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= w.L[j][k] * w.L[j][k];
        ok = ok && (d > 0.0);
        det *= d;
        w.il[j] = rsqrt(d);
        w.L[j][j] = d * w.il[j];
#pragma unroll
        for (int i = j + 1; i < BS; ++i) {
            double v = S[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= w.L[i][k] * w.L[j][k];
            w.L[i][j] = v * w.il[j];
        }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i) {                        // y = L^-1 r
        double v = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= w.L[i][k] * w.y[k];
        w.y[i] = v * w.il[i];
        q += w.y[i] * w.y[i];
    }
    det_out = det;
    q_out = q;
    return ok;
}

template <int B, int BS>
__device__ __forceinline__ void small_factor(SmallSolve<B>& w, const double (&S)[B][B], const double (&r)[B], double sgn) {
    double det = 1.0, q = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < BS; ++j) {
        double d = S[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= w.L[j][k] * w.L[j][k];
        ok = ok && (d > 0.0);
        det *= d;
        w.il[j] = rsqrt(d);
        w.L[j][j] = d * w.il[j];
#pragma unroll
        for (int i = j + 1; i < BS; ++i) {
            double v = S[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= w.L[i][k] * w.L[j][k];
            w.L[i][j] = v * w.il[j];
        }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i) {                        // y = L^-1 r
        double v = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= w.L[i][k] * w.y[k];
        w.y[i] = v * w.il[i];
        q += w.y[i] * w.y[i];
    }
    w.dpost = ok ? sgn * 0.5 * log(det) + 0.5 * q : nan("");
}

// What a committed step needs on top of small_factor: G = S^-1, gr = G r, and xm = L^-T z for the draws.
template <int B, int BS>
__device__ __forceinline__ void small_finish(SmallSolve<B>& w, const double (&r)[B], const double* zc, int coord0) {
#pragma unroll
    for (int c = 0; c < BS; ++c) {
        double u[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) v -= w.L[i][k] * u[k];
            u[i] = v * w.il[i];
        }
#pragma unroll
        for (int i = BS - 1; i >= 0; --i) {
            double v = u[i];
#pragma unroll
            for (int k = i + 1; k < BS; ++k) v -= w.L[k][i] * w.G[k][c];
            w.G[i][c] = v * w.il[i];
        }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < BS; ++k) v += w.G[i][k] * r[k];
        w.gr[i] = v;
        w.xm[i] = 0.0;
    }
    if (zc) {
#pragma unroll
        for (int i = BS - 1; i >= 0; --i) {               // xm = L^-T z_m
            double v = zc[coord0 + i];
#pragma unroll
            for (int k = i + 1; k < BS; ++k) v -= w.L[k][i] * w.xm[k];
            w.xm[i] = v * w.il[i];
        }
    }
}


// Cluster kernel (spike_slab_dsm.cu).  csize = 0: pick the smallest cluster whose shared memory holds the neuron's
// state; 2 / 4 / 8: that size.  PYGLM_ERR_UNSUPPORTED (nothing launched, no error text) when the state does not fit.
int spike_slab_dsm_launch(const SpikeSlabArgs& A, int csize, cudaStream_t stream);
// true when the cluster kernel applies and is expected to beat one CTA per neuron (see spike_slab_dsm.cu)
bool spike_slab_dsm_preferred(const SpikeSlabArgs& A);

}  // namespace pyglm_ss
