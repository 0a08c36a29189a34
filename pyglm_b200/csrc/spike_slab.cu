// (4) Spike-and-slab block update of (a_n, W_n, b_n) for every local postsynaptic neuron: one CTA per neuron.
//
// Replaces regression.py:265-340: _collapsed_resample_a (:282-320; N sequential flips, each calling
// _marginal_likelihood :343-378 = two dense Choleskys + two dpotrs of size <= D) and _resample_W (:323-340;
// sample_gaussian(J=,h=) = Cholesky + triangular solves).
//
// Same conditional distributions, different linear algebra.  Let S be the active coordinate set (the B-blocks
// with a=1 plus the bias).  The kernel keeps P = (Jp_SS)^-1 and mu = P hp_S in compact form and uses the
// block-bordering identities
//     add block m:     t = P c,  c = Jp[S,m],  Sm = Jp[m,m] - c^T t,  r = hp[m] - c^T mu
//                      ml(S+m) - ml(S) = -1/2 log|Sm| + 1/2 r^T Sm^-1 r + prior_m
//     remove block m:  the same quantity read off P:  +1/2 log|P_mm| + 1/2 mu_m^T P_mm^-1 mu_m + prior_m
// so one flip costs O(K^2 B) (a mat-vec against P and a rank-B update) instead of O(K^3), and every step is a
// parallel reduction rather than a sequential factorisation.  The prior J0 is block diagonal
// (regression.py:218), so its contribution is the closed-form prior_m = 1/2 log|J0_m| - 1/2 h0_m^T J0_m^-1 h0_m.
// The log-odds of a_m = 1 is  ml(1) - ml(0) + log rho_m - log(1 - rho_m); the flip is decided exactly like
// sample_discrete_from_log(lps) with uniform u: a_m = (u > p0), p0 = 1 / (1 + exp(logodds)).
//
// The Gaussian draw replays the final active set by bordering in ASCENDING coordinate order with the bias
// last (the order np.ix_(mask, mask) gives the reference): appending block m with Sm = L L^T,
//     x_m = L^-T z_m,   x_prev -= t x_m,
// which is, step for step, the back-substitution  x = chol(Jp_SS)^-T z  of sample_gaussian; mu + x is the draw.
// The same replay accumulates  -1/2 log|Jp_SS| + 1/2 hp^T Jp^-1 hp, i.e. _marginal_likelihood, for parity tests.
//
// Randomness (permutation, uniforms, normals) is read from device buffers so tests can inject the reference's
// own draws; in production those buffers are filled by pyglm_scan_randomness (Philox, below).
#include "spike_slab_common.cuh"
#include <cooperative_groups.h>
#include "philox.cuh"
#include <stdlib.h>

using namespace pyglm_ss;

namespace {

template <int SS_THREADS>
struct Ctx {
    static constexpr int SS_WARPS = SS_THREADS / 32;
    static constexpr int SS_ROWS = 4;          // rows of P per warp per batch in the passes over P
    // problem
    int N, B, D, NB, ldj, ldp;
    const double* Jn; const double* hn; const double* J0w; const double* h0w; double J0b, h0b;
    double* P;
    // shared state
    double *mu, *xs, *cb, *tb, *gb, *S, *G, *r, *gr, *xm, *misc;
    int *cidx, *slot;
    int K;
    int tid, lane, warp;

    __device__ __forceinline__ double Jp(int i, int j) const {
        const int hi = max(i, j), lo = min(i, j);
        double v = Jn[(size_t)hi * ldj + lo];
        if (hi < NB) {
            const int m = hi / B;
            if (lo / B == m) v += J0w[(size_t)m * B * B + (hi - m * B) * B + (lo - m * B)];
        } else if (lo == hi) {
            v += J0b;
        }
        return v;
    }
    __device__ __forceinline__ double hp(int d) const { return hn[d] + (d < NB ? h0w[d] : h0b); }

    // thread 0: in-place lower Cholesky of the bs x bs matrix in S (row pitch SS_BMAX); returns log|S| or NaN
    __device__ double chol_small(int bs) {
        double logdet = 0.0;
        for (int j = 0; j < bs; ++j) {
            double d = S[j * SS_BMAX + j];
            for (int k = 0; k < j; ++k) d -= S[j * SS_BMAX + k] * S[j * SS_BMAX + k];
            if (!(d > 0.0)) return nan("");
            const double l = sqrt(d);
            S[j * SS_BMAX + j] = l;
            logdet += 2.0 * log(l);
            for (int i = j + 1; i < bs; ++i) {
                double s = S[i * SS_BMAX + j];
                for (int k = 0; k < j; ++k) s -= S[i * SS_BMAX + k] * S[j * SS_BMAX + k];
                S[i * SS_BMAX + j] = s / l;
            }
        }
        return logdet;
    }
    // thread 0: G = (L L^T)^-1 with L in S
    __device__ void inverse_from_chol(int bs) {
        for (int c = 0; c < bs; ++c) {
            double y[SS_BMAX];
            for (int i = 0; i < bs; ++i) {               // L y = e_c
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = 0; k < i; ++k) s -= S[i * SS_BMAX + k] * y[k];
                y[i] = s / S[i * SS_BMAX + i];
            }
            for (int i = bs - 1; i >= 0; --i) {          // L^T g = y
                double s = y[i];
                for (int k = i + 1; k < bs; ++k) s -= S[k * SS_BMAX + i] * G[k * SS_BMAX + c];
                G[i * SS_BMAX + c] = s / S[i * SS_BMAX + i];
            }
        }
    }
    // thread 0: |L^-1 v|^2
    __device__ double quad_from_chol(const double* v, int bs) {
        double y[SS_BMAX], q = 0.0;
        for (int i = 0; i < bs; ++i) {
            double s = v[i];
            for (int k = 0; k < i; ++k) s -= S[i * SS_BMAX + k] * y[k];
            y[i] = s / S[i * SS_BMAX + i];
            q += y[i] * y[i];
        }
        return q;
    }

    // Evaluate appending coordinates [coord0, coord0+bs) to the active set.  On return (after the trailing
    // barrier) thread-0 results sit in shared memory: misc[0] = -1/2 log|Sm| + 1/2 r^T Sm^-1 r (NaN on failure),
    // S = chol(Sm), G = Sm^-1, r, gr = G r, and xm = L^-T z_m when zc != nullptr.
    __device__ void eval_add(int coord0, int bs, const double* zc) {
        for (int e = tid; e < K * bs; e += SS_THREADS) {
            const int c1 = e / bs, bb = e - c1 * bs;
            cb[c1 * B + bb] = Jp(cidx[c1], coord0 + bb);
        }
        __syncthreads();
        // t = P c: SS_ROWS rows per warp at a time, four right-hand sides per pass.  All 4*SS_ROWS loads of a lane
        // are issued before the first FMA, so a pass over P costs a handful of L2 latencies instead of one per row
        // (P lives in L2: K*K doubles do not fit in shared memory).
        for (int r0 = warp * SS_ROWS; r0 < K; r0 += SS_WARPS * SS_ROWS) {
            for (int b0 = 0; b0 < bs; b0 += 4) {
                const int nb = min(4, bs - b0);
                double s[SS_ROWS][4];
#pragma unroll
                for (int rr = 0; rr < SS_ROWS; ++rr) s[rr][0] = s[rr][1] = s[rr][2] = s[rr][3] = 0.0;
                for (int c0 = lane; c0 < K; c0 += 128) {
                    double pv[SS_ROWS][4];
#pragma unroll
                    for (int rr = 0; rr < SS_ROWS; ++rr) {
                        const double* prow = P + (size_t)min(r0 + rr, K - 1) * ldp;
#pragma unroll
                        for (int u = 0; u < 4; ++u) pv[rr][u] = (c0 + 32 * u < K) ? prow[c0 + 32 * u] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double* cr = cb + min(c0 + 32 * u, K - 1) * B + b0;
                        const double v0 = cr[0];
                        const double v1 = (nb > 1) ? cr[1] : 0.0;
                        const double v2 = (nb > 2) ? cr[2] : 0.0;
                        const double v3 = (nb > 3) ? cr[3] : 0.0;
#pragma unroll
                        for (int rr = 0; rr < SS_ROWS; ++rr) {
                            s[rr][0] += pv[rr][u] * v0;
                            s[rr][1] += pv[rr][u] * v1;
                            if (nb > 2) { s[rr][2] += pv[rr][u] * v2; s[rr][3] += pv[rr][u] * v3; }
                        }
                    }
                }
#pragma unroll
                for (int rr = 0; rr < SS_ROWS; ++rr) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < nb) {
                            const double v = warp_sum(s[rr][k]);
                            if (lane == 0 && r0 + rr < K) tb[(r0 + rr) * B + b0 + k] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // Sm = Jp[m,m] - c^T t  and  r = hp[m] - c^T mu : bs*bs + bs dot products of length K
        for (int p = warp; p < bs * bs + bs; p += SS_WARPS) {
            double s = 0.0;
            if (p < bs * bs) {
                const int bb = p / bs, b2 = p - bb * bs;
                for (int c1 = lane; c1 < K; c1 += 32) s += cb[c1 * B + bb] * tb[c1 * B + b2];
                s = warp_sum(s);
                if (lane == 0) S[bb * SS_BMAX + b2] = Jp(coord0 + bb, coord0 + b2) - s;
            } else {
                const int bb = p - bs * bs;
                for (int c1 = lane; c1 < K; c1 += 32) s += cb[c1 * B + bb] * mu[c1];
                s = warp_sum(s);
                if (lane == 0) r[bb] = hp(coord0 + bb) - s;
            }
        }
        __syncthreads();
        if (tid == 0) {
            const double logdet = chol_small(bs);
            if (logdet == logdet) {
                misc[0] = -0.5 * logdet + 0.5 * quad_from_chol(r, bs);
                inverse_from_chol(bs);
                for (int i = 0; i < bs; ++i) {
                    double s = 0.0;
                    for (int k = 0; k < bs; ++k) s += G[i * SS_BMAX + k] * r[k];
                    gr[i] = s;
                }
                if (zc) {                                 // xm = L^-T z_m
                    for (int i = bs - 1; i >= 0; --i) {
                        double s = zc[coord0 + i];
                        for (int k = i + 1; k < bs; ++k) s -= S[k * SS_BMAX + i] * xm[k];
                        xm[i] = s / S[i * SS_BMAX + i];
                    }
                }
            } else {
                misc[0] = nan("");
            }
        }
        __syncthreads();
    }

    // P[c1][c2] += sign * sum_k gb[c1][k] tb[c2][k] over the K x K active block, SS_ROWS rows per warp with every
    // load of the batch in flight before the first store.
    __device__ __forceinline__ void rank_update(int bs, double sign) {
        for (int r0 = warp * SS_ROWS; r0 < K; r0 += SS_WARPS * SS_ROWS) {
            for (int c0 = lane; c0 < K; c0 += 128) {
                double pv[SS_ROWS][4];
#pragma unroll
                for (int rr = 0; rr < SS_ROWS; ++rr) {
                    const double* prow = P + (size_t)min(r0 + rr, K - 1) * ldp;
#pragma unroll
                    for (int u = 0; u < 4; ++u) pv[rr][u] = (c0 + 32 * u < K) ? prow[c0 + 32 * u] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c2 = c0 + 32 * u;
                    if (c2 < K) {
                        const double* t2 = tb + c2 * B;
#pragma unroll
                        for (int rr = 0; rr < SS_ROWS; ++rr) {
                            if (r0 + rr < K) {
                                const double* g1 = gb + (r0 + rr) * B;
                                double acc = 0.0;
                                for (int k = 0; k < bs; ++k) acc += g1[k] * t2[k];
                                P[(size_t)(r0 + rr) * ldp + c2] = pv[rr][u] + sign * acc;
                            }
                        }
                    }
                }
            }
        }
    }

    // Append the block evaluated by the last eval_add.
    __device__ void commit_add(int coord0, int bs, bool draw) {
        for (int e = tid; e < K * bs; e += SS_THREADS) {          // gb = t G
            const int c1 = e / bs, bb = e - c1 * bs;
            double s = 0.0;
            for (int k = 0; k < bs; ++k) s += tb[c1 * B + k] * G[k * SS_BMAX + bb];
            gb[c1 * B + bb] = s;
        }
        __syncthreads();
        rank_update(bs, 1.0);                                      // P += gb t^T
        for (int e = tid; e < K * bs; e += SS_THREADS) {          // new border rows / columns
            const int c1 = e / bs, bb = e - c1 * bs;
            const double v = -gb[c1 * B + bb];
            P[(size_t)c1 * ldp + K + bb] = v;
            P[(size_t)(K + bb) * ldp + c1] = v;
        }
        for (int e = tid; e < bs * bs; e += SS_THREADS) {
            const int bb = e / bs, b2 = e - bb * bs;
            P[(size_t)(K + bb) * ldp + K + b2] = G[bb * SS_BMAX + b2];
        }
        for (int c1 = tid; c1 < K; c1 += SS_THREADS) {
            double s = 0.0, sx = 0.0;
            for (int k = 0; k < bs; ++k) { s += gb[c1 * B + k] * r[k]; sx += tb[c1 * B + k] * xm[k]; }
            mu[c1] -= s;
            if (draw) xs[c1] -= sx;
        }
        if (tid < bs) {
            mu[K + tid] = gr[tid];
            if (draw) xs[K + tid] = xm[tid];
            cidx[K + tid] = coord0 + tid;
        }
        __syncthreads();
        K += bs;
    }

    // Evaluate removing the B-block stored at compact position pos: misc[0] = ml(with) - ml(without), posterior part.
    __device__ void eval_remove(int pos) {
        if (tid == 0) {
            for (int i = 0; i < B; ++i) {
                for (int k = 0; k < B; ++k) S[i * SS_BMAX + k] = P[(size_t)(pos + i) * ldp + pos + k];
                r[i] = mu[pos + i];
            }
            const double logdet = chol_small(B);
            if (logdet == logdet) {
                misc[0] = 0.5 * logdet + 0.5 * quad_from_chol(r, B);
                inverse_from_chol(B);                     // G = P_mm^-1
            } else {
                misc[0] = nan("");
            }
        }
        __syncthreads();
    }

    // Remove the block at pos (after eval_remove) and move the last block into its place.
    __device__ void commit_remove(int pos) {
        for (int e = tid; e < K * B; e += SS_THREADS) {           // tb = P[:, pos block] (read as rows: P symmetric)
            const int bb = e / K, c2 = e - bb * K;
            tb[c2 * B + bb] = P[(size_t)(pos + bb) * ldp + c2];
        }
        __syncthreads();
        for (int e = tid; e < K * B; e += SS_THREADS) {           // gb = tb P_mm^-1
            const int c1 = e / B, bb = e - c1 * B;
            double s = 0.0;
            for (int k = 0; k < B; ++k) s += tb[c1 * B + k] * G[k * SS_BMAX + bb];
            gb[c1 * B + bb] = s;
        }
        __syncthreads();
        rank_update(B, -1.0);                                      // P -= gb tb^T
        for (int c1 = tid; c1 < K; c1 += SS_THREADS) {            // mu_R -= P_Rm P_mm^-1 mu_m   (r holds mu_m)
            double s = 0.0;
            for (int k = 0; k < B; ++k) s += gb[c1 * B + k] * r[k];
            mu[c1] -= s;
        }
        __syncthreads();
        const int last = K - B;
        if (pos != last) {
            for (int e = tid; e < K * B; e += SS_THREADS) {       // rows of the last block -> rows at pos
                const int bb = e / K, c2 = e - bb * K;
                P[(size_t)(pos + bb) * ldp + c2] = P[(size_t)(last + bb) * ldp + c2];
            }
            __syncthreads();
            for (int e = tid; e < last * B; e += SS_THREADS) {    // columns of the last block -> columns at pos
                const int c1 = e / B, bb = e - c1 * B;
                P[(size_t)c1 * ldp + pos + bb] = P[(size_t)c1 * ldp + last + bb];
            }
            if (tid < B) {
                mu[pos + tid] = mu[last + tid];
                cidx[pos + tid] = cidx[last + tid];
            }
            if (tid == 0) slot[cidx[last] / B] = pos;
        }
        __syncthreads();
        K -= B;
    }
};

template <int SS_THREADS, int MINB>
__global__ void __launch_bounds__(SS_THREADS, MINB)
spike_slab_kernel(SpikeSlabArgs A) {
    extern __shared__ __align__(16) double ssm[];
    const int ln = blockIdx.x;
    const int N = A.N, B = A.B, D = A.D;
    const int Dpad = (D + 1) & ~1;

    Ctx<SS_THREADS> c;
    c.N = N; c.B = B; c.D = D; c.NB = N * B; c.ldj = A.ldj; c.ldp = D;
    c.Jn = A.J + (size_t)ln * A.stride_n;
    c.hn = A.h + (size_t)ln * A.ldh;
    c.J0w = A.J0w + (size_t)ln * N * B * B;
    c.h0w = A.h0w + (size_t)ln * N * B;
    c.J0b = A.J0b[ln]; c.h0b = A.h0b[ln];
    c.P = A.P + (size_t)ln * D * D;
    double* p = ssm;
    c.mu = p; p += Dpad;
    c.xs = p; p += Dpad;
    c.cb = p; p += (size_t)Dpad * B;
    c.tb = p; p += (size_t)Dpad * B;
    c.gb = p; p += (size_t)Dpad * B;
    c.S = p; p += SS_BMAX * SS_BMAX;
    c.G = p; p += SS_BMAX * SS_BMAX;
    c.r = p; p += SS_BMAX;
    c.gr = p; p += SS_BMAX;
    c.xm = p; p += SS_BMAX;
    c.misc = p; p += 8;
    c.cidx = reinterpret_cast<int*>(p);
    c.slot = c.cidx + Dpad;
    c.tid = threadIdx.x; c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5;
    c.K = 0;
    const int tid = threadIdx.x;

    unsigned char* a = A.a + (size_t)ln * N;
    const double* cprior = A.cprior + (size_t)ln * N;
    const double* lrho = A.logit_rho + (size_t)ln * N;
    int fail = 0;

    if (tid < SS_BMAX) c.xm[tid] = 0.0;
    for (int m = tid; m < N; m += SS_THREADS) c.slot[m] = -1;
    __syncthreads();

    if (A.do_scan[ln]) {
        // ---- build P, mu for the current active set: bias first, then the active blocks
        c.eval_add(D - 1, 1, nullptr);
        if (!(c.misc[0] == c.misc[0])) fail = 1;
        c.commit_add(D - 1, 1, false);
        for (int m = 0; m < N && !fail; ++m) {
            if (!a[m]) continue;
            if (tid == 0) c.slot[m] = c.K;
            c.eval_add(m * B, B, nullptr);
            if (!(c.misc[0] == c.misc[0])) { fail = 1; break; }
            c.commit_add(m * B, B, false);
        }
        // ---- the collapsed scan (regression.py:286-320)
        const int* perm = A.perm + (size_t)ln * N;
        const double* us = A.us + (size_t)ln * N;
        for (int step = 0; step < N && !fail; ++step) {
            const int m = perm[step];
            const int pos = c.slot[m];
            if (pos < 0) c.eval_add(m * B, B, nullptr);
            else c.eval_remove(pos);
            const double dpost = c.misc[0];
            if (!(dpost == dpost)) { fail = 1; break; }
            const double lo = dpost + cprior[m] + lrho[m];
            const double p0 = 1.0 / (1.0 + exp(lo));
            const int v = us[step] > p0;
            if (A.logodds && tid == 0) A.logodds[(size_t)ln * N + step] = lo;
            if (pos < 0 && v) {
                if (tid == 0) { c.slot[m] = c.K; a[m] = 1; }
                c.commit_add(m * B, B, false);
            } else if (pos >= 0 && !v) {
                if (tid == 0) { c.slot[m] = -1; a[m] = 0; }
                c.commit_remove(pos);
            } else {
                __syncthreads();
            }
        }
    }
    __syncthreads();

    // ---- Gaussian draw of [W_active; b] by bordering in ascending coordinate order, bias last
    c.K = 0;
    double ml = 0.0;
    const double* zc = A.z + (size_t)ln * A.ldz;
    for (int m = 0; m <= N && !fail; ++m) {
        const bool is_bias = (m == N);
        if (!is_bias && !a[m]) continue;
        const int coord0 = is_bias ? D - 1 : m * B, bs = is_bias ? 1 : B;
        c.eval_add(coord0, bs, zc);
        const double dpost = c.misc[0];
        if (!(dpost == dpost)) { fail = 1; break; }
        ml += dpost + (is_bias ? 0.5 * log(c.J0b) - 0.5 * c.h0b * c.h0b / c.J0b : cprior[m]);
        c.commit_add(coord0, bs, true);
    }
    __syncthreads();
    double* Wn = A.W + (size_t)ln * N * B;
    for (int e = tid; e < N * B; e += SS_THREADS) Wn[e] = 0.0;
    __syncthreads();
    if (!fail) {
        for (int k = tid; k < c.K; k += SS_THREADS) {
            const int d = c.cidx[k];
            const double v = c.mu[k] + c.xs[k];
            if (d < N * B) Wn[d] = v; else A.bias[ln] = v;
        }
    }
    if (tid == 0) {
        if (A.ml) A.ml[ln] = fail ? nan("") : ml;
        A.status[ln] = fail;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path, B <= 4 (every configuration in BASELINE.json has B in {1,2,3}).  Same algorithm and the same order of
// updates as the generic kernel above; what changes is how a step is executed:
//   * B is a template parameter: block loops unroll, index divisions fold, and the three phases (build, scan, draw)
//     run through ONE loop so each step routine is instantiated once (the generic kernel is 390 KB of SASS and
//     stalls on instruction fetch);
//   * the small dense algebra of a step (Cholesky of the B x B Schur complement, log-determinant, quadratic form,
//     the flip decision, and -- only when the step is committed -- the inverse) is evaluated REDUNDANTLY by every
//     thread in registers from block-wide partial sums, instead of by thread 0 with 511 threads parked at a barrier;
//   * S_m = J_mm - c^T P c and r = h_m - c^T mu are accumulated inside the pass over P (each warp multiplies the
//     rows of t it has just produced), so an evaluation is two barriers, a rejected removal none;
//   * passes over P keep 16 loads per lane in flight (4 rows x 4 column chunks): P lives in L2 (K x K doubles do
//     not fit in shared memory), so memory-level parallelism sets the time of a pass;
//   * t G (the scaled border) is recomputed per row from registers in the rank update: no staging pass.
// Cluster variant (C = 2): the C CTAs of a cluster work on ONE neuron.  Both execute the same control flow on their own
// copies of the small state (mu, cidx, panels ...); the K x K passes over P -- t = P c, the rank update, the trailing
// updates of the blocked factorisations -- are split by row batches, the rows of t and the per-warp partial sums of an
// evaluation are written into both CTAs' shared memory (DSMEM), and every block barrier becomes a cluster barrier
// (release / acquire at cluster scope, which also orders the P rows the peer wrote to L2).  P is read with ld.global.cg
// (L2 only) so no CTA ever sees a stale L1 line of a row its peer updated.
template <int C>
__device__ __forceinline__ void csync() {
    if (C == 1) __syncthreads();
    else cooperative_groups::this_cluster().sync();
}
template <int C>
__device__ __forceinline__ unsigned crank_of() {
    return (C == 1) ? 0u : cooperative_groups::this_cluster().block_rank();
}
// store to the same shared-memory location in every CTA of the cluster
template <int C>
__device__ __forceinline__ void store_all(double* p, double v) {
    if (C == 1) { *p = v; return; }
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
#pragma unroll
    for (int r = 0; r < C; ++r) *cl.map_shared_rank(p, r) = v;
}
__device__ __forceinline__ double ldP(const double* p) { return __ldcg(p); }

template <int B, int NTHR, int C>
struct FastCtx {
    static constexpr int NWARP = NTHR / 32;
    static constexpr int ROWS = 4;              // rows of P per warp per batch
    static constexpr int PB = B * B + B;        // partial sums per warp: c^T t (B x B) and c^T mu (B)
    int N, D, NB, ldj, ldp;
    const double* Jn; const double* hn; const double* J0w; const double* h0w; double J0b, h0b;
    double* P;
    double *mu, *xs, *cb, *tb, *part;
    int *cidx, *slot;
    int K, tid, lane, warp;
    int crank;                                  // rank of this CTA in the cluster working on the neuron
    // Lookahead table of the scan: for the next la_G INACTIVE neurons in scan order, c_j = Jp[S, j-block] and
    // t_j = P c_j are formed together in ONE pass over P (a K x K by K x (G B) product on the FP64 tensor cores) and
    // kept current under every flip in between by rank-B corrections (shared memory only), so the scan makes one
    // pass over P per la_G add-evaluations instead of one per evaluation.
    static constexpr int LA_MAX = 8;
    static constexpr int LA_COLS = 16;          // G * B <= 16: two 8-column DMMA tiles
    double *Cb, *Tb;                            // [la_G][la_ld]
    double *la_part, *la_E;                     // [NWARP][B][16] partial products; [LA_MAX][2][B][B] reduced E and D
    int* cand;                                  // [LA_MAX] presynaptic neuron per slot, -1 = free
    int la_G, la_ld;

    __device__ __forceinline__ double Jp(int i, int j) const {
        const int hi = max(i, j), lo = min(i, j);
        double v = Jn[(size_t)hi * ldj + lo];
        if (hi < NB) {
            const int m = hi / B;
            if (lo / B == m) v += J0w[(size_t)m * B * B + (hi - m * B) * B + (lo - m * B)];
        } else if (lo == hi) {
            v += J0b;
        }
        return v;
    }
    __device__ __forceinline__ double hp(int d) const { return hn[d] + (d < NB ? h0w[d] : h0b); }

    __device__ __forceinline__ int la_find(int m) const {
        int g = -1;
#pragma unroll
        for (int i = 0; i < LA_MAX; ++i) g = (i < la_G && cand[i] == m) ? i : g;
        return g;
    }

    // A (row gr, col q) of t^T or P, B (row q, col gr) of the c vectors, as dmma884 wants them (lane = 4 gr + q)
    __device__ __forceinline__ double la_cval(int col, int n) const {
        const int g = n / B, bb = n - g * B;
        return (col < K && g < la_G) ? Cb[g * la_ld + col * B + bb] : 0.0;
    }

    // Fill the table with the next la_G inactive neurons of the scan (the current one first).
    __device__ __forceinline__ void la_refill(const int* perm, int cursor) {
        __syncthreads();
        if (tid == 0) {
            int g = 0;
            for (int i = cursor; i < N && g < la_G; ++i) {
                const int m = perm[i];
                if (slot[m] < 0) cand[g++] = m;
            }
            for (; g < LA_MAX; ++g) cand[g] = -1;
        }
        __syncthreads();
        for (int e = tid; e < la_G * K * B; e += NTHR) {
            const int g = e / (K * B), rem = e - g * K * B, c1 = rem / B, bb = rem - c1 * B;
            const int m = cand[g];
            Cb[g * la_ld + c1 * B + bb] = (m >= 0) ? Jp(cidx[c1], m * B + bb) : 0.0;
        }
        __syncthreads();
        const int gr = lane >> 2, q = lane & 3;
        for (int rt = warp; rt * 8 < K; rt += NWARP) {      // T = P C, one 8-row tile per warp, all K columns
            const int row = rt * 8 + gr;
            const double* prow = P + (size_t)min(row, K - 1) * ldp;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int k0 = 0; k0 < K; k0 += 64) {
                double av[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int col = k0 + 4 * u + q;
                    av[u] = (row < K && col < K) ? ldP(prow + col) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int kc = k0 + 4 * u;
                    if (kc < K) {
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) dmma884(acc[nt][0], acc[nt][1], av[u], la_cval(kc + q, nt * 8 + gr));
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = nt * 8 + 2 * q + h, g = n / B, bb = n - g * B;
                    if (row < K && g < la_G) Tb[g * la_ld + row * B + bb] = acc[nt][h];
                }
        }
        __syncthreads();
    }

    // S = J_jj - c_j^T t_j, r = h_j - c_j^T mu from the table; t_j is copied to tb for the commit.
    __device__ __forceinline__ void la_eval(int g, int coord0, double (&S)[B][B], double (&r)[B]) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
#pragma unroll
            for (int b2 = 0; b2 <= b; ++b2) S[b][b2] = Jp(coord0 + b, coord0 + b2);
            r[b] = hp(coord0 + b);
        }
        const double* cg = Cb + g * la_ld;
        const double* tg = Tb + g * la_ld;
        double ps[B][B], pr[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            pr[b] = 0.0;
#pragma unroll
            for (int b2 = 0; b2 < B; ++b2) ps[b][b2] = 0.0;
        }
        __syncthreads();                                   // `part` and tb of the previous step are no longer read
        for (int row = tid; row < K; row += NTHR) {
            const double m_r = mu[row];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const double cvb = cg[row * B + b];
                tb[row * B + b] = tg[row * B + b];
                pr[b] += cvb * m_r;
#pragma unroll
                for (int b2 = 0; b2 < B; ++b2) ps[b][b2] += cvb * tg[row * B + b2];
            }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            pr[b] = warp_sum(pr[b]);
#pragma unroll
            for (int b2 = 0; b2 < B; ++b2) ps[b][b2] = warp_sum(ps[b][b2]);
        }
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                part[warp * PB + B * B + b] = pr[b];
#pragma unroll
                for (int b2 = 0; b2 < B; ++b2) part[warp * PB + b * B + b2] = ps[b][b2];
            }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < B; ++b) {
            r[b] -= warp_sum(lane < NWARP ? part[lane * PB + B * B + b] : 0.0);
#pragma unroll
            for (int b2 = 0; b2 <= b; ++b2) S[b][b2] -= warp_sum(lane < NWARP ? part[lane * PB + b * B + b2] : 0.0);
        }
    }

    // Slot g leaves the table (its neuron has been visited).
    __device__ __forceinline__ void la_release(int g) {
        __syncthreads();
        if (tid == 0) cand[g] = -1;
        __syncthreads();
    }

    // The block of slot g (coordinates coord0..) is being appended with t = tb, G = S^-1 (w.G): for every other pending
    // candidate j,  E = t^T c_j - D,  D = Jp[new block, j-block];  t_j += t (G E),  new rows of t_j = -G E,  new rows of
    // c_j = D.  Call before commit_add (K is still the old size).
    __device__ __forceinline__ void la_after_add(int g, const SmallSolve<B>& w, int coord0) {
        const int gr = lane >> 2, q = lane & 3;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int kc = warp * 4; kc < K; kc += NWARP * 4) {  // t^T C, the K range split over the warps
            const int col = kc + q;
            const double av = (gr < B && col < K) ? tb[col * B + gr] : 0.0;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) dmma884(acc[nt][0], acc[nt][1], av, la_cval(col, nt * 8 + gr));
        }
        if (gr < B) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) la_part[(warp * B + gr) * LA_COLS + nt * 8 + 2 * q + h] = acc[nt][h];
        }
        __syncthreads();
        if (tid < la_G * B * B) {                          // E and D of slot g2, entry (b, b2)
            const int g2 = tid / (B * B), rem = tid - g2 * B * B, bb = rem / B, b2 = rem - bb * B;
            const int m2 = cand[g2];
            double e = 0.0, d = 0.0;
            if (m2 >= 0 && g2 != g) {
                for (int wv = 0; wv < NWARP; ++wv) e += la_part[(wv * B + bb) * LA_COLS + g2 * B + b2];
                d = Jp(coord0 + bb, m2 * B + b2);
                e -= d;
            }
            la_E[(g2 * 2 + 0) * B * B + rem] = e;
            la_E[(g2 * 2 + 1) * B * B + rem] = d;
        }
        __syncthreads();
        for (int g2 = 0; g2 < la_G; ++g2) {
            if (g2 == g || cand[g2] < 0) continue;
            double GE[B][B];
#pragma unroll
            for (int i = 0; i < B; ++i)
#pragma unroll
                for (int b2 = 0; b2 < B; ++b2) {
                    double v = 0.0;
#pragma unroll
                    for (int kk = 0; kk < B; ++kk) v += w.G[i][kk] * la_E[(g2 * 2 + 0) * B * B + kk * B + b2];
                    GE[i][b2] = v;
                }
            double* tg = Tb + g2 * la_ld;
            double* cg = Cb + g2 * la_ld;
            for (int row = tid; row < K; row += NTHR) {
#pragma unroll
                for (int b2 = 0; b2 < B; ++b2) {
                    double v = tg[row * B + b2];
#pragma unroll
                    for (int i = 0; i < B; ++i) v += tb[row * B + i] * GE[i][b2];
                    tg[row * B + b2] = v;
                }
            }
            if (tid < B * B) {
                const int i = tid / B, b2 = tid - i * B;
                double sel = 0.0;
#pragma unroll
                for (int ii = 0; ii < B; ++ii)
#pragma unroll
                    for (int jj = 0; jj < B; ++jj) sel = (ii == i && jj == b2) ? GE[ii][jj] : sel;
                tg[(K + i) * B + b2] = -sel;
                cg[(K + i) * B + b2] = la_E[(g2 * 2 + 1) * B * B + i * B + b2];
            }
        }
        if (tid == 0) cand[g] = -1;
        __syncthreads();
    }

    // The block at pos is being removed (tb = P[:, pos block], w.G = P_mm^-1): t_j -= tb (G t_j[pos block]) on the
    // remaining rows.  Call after tb has been gathered, before anything moves.
    __device__ __forceinline__ void la_after_remove(const SmallSolve<B>& w, int pos) {
        for (int g2 = 0; g2 < la_G; ++g2) {
            if (cand[g2] < 0) continue;
            double* tg = Tb + g2 * la_ld;
            double v[B][B];
#pragma unroll
            for (int i = 0; i < B; ++i)
#pragma unroll
                for (int b2 = 0; b2 < B; ++b2) {
                    double s = 0.0;
#pragma unroll
                    for (int kk = 0; kk < B; ++kk) s += w.G[i][kk] * tg[(pos + kk) * B + b2];
                    v[i][b2] = s;
                }
            for (int row = tid; row < K; row += NTHR) {
                if (row >= pos && row < pos + B) continue;
#pragma unroll
                for (int b2 = 0; b2 < B; ++b2) {
                    double x = tg[row * B + b2];
#pragma unroll
                    for (int i = 0; i < B; ++i) x -= tb[row * B + i] * v[i][b2];
                    tg[row * B + b2] = x;
                }
            }
        }
    }

    // the rows of the last block take the place of the removed block in every table entry
    __device__ __forceinline__ void la_move(int pos, int last) {
        if (tid < la_G * B * B) {
            const int g2 = tid / (B * B), rem = tid - g2 * B * B, i = rem / B, b2 = rem - i * B;
            if (cand[g2] >= 0) {
                Tb[g2 * la_ld + (pos + i) * B + b2] = Tb[g2 * la_ld + (last + i) * B + b2];
                Cb[g2 * la_ld + (pos + i) * B + b2] = Cb[g2 * la_ld + (last + i) * B + b2];
            }
        }
    }

    // Evaluate appending coordinates [coord0, coord0+BS): t = P c goes to tb; S (lower triangle of the Schur
    // complement) and r are returned in every thread.
    template <int BS>
    __device__ __forceinline__ void eval_add(int coord0, double (&S)[B][B], double (&r)[B]) {
#pragma unroll
        for (int b = 0; b < BS; ++b) {                    // issued early: latency hides behind the pass over P
#pragma unroll
            for (int b2 = 0; b2 <= b; ++b2) S[b][b2] = Jp(coord0 + b, coord0 + b2);
            r[b] = hp(coord0 + b);
        }
        for (int e = tid; e < K * BS; e += NTHR) {
            const int c1 = e / BS, bb = e - c1 * BS;
            cb[c1 * B + bb] = Jp(cidx[c1], coord0 + bb);
        }
        csync<C>();
        double ps[BS][BS], pr[BS];
#pragma unroll
        for (int b = 0; b < BS; ++b) {
            pr[b] = 0.0;
#pragma unroll
            for (int b2 = 0; b2 < BS; ++b2) ps[b][b2] = 0.0;
        }
        for (int r0 = (crank * NWARP + warp) * ROWS; r0 < K; r0 += C * NWARP * ROWS) {
            double s[ROWS][BS];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                for (int b = 0; b < BS; ++b) s[rr][b] = 0.0;
            for (int c0 = lane; c0 < K; c0 += 128) {
                double pv[ROWS][4];
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr) {
                    const double* prow = P + (size_t)min(r0 + rr, K - 1) * ldp;
#pragma unroll
                    for (int u = 0; u < 4; ++u) pv[rr][u] = (c0 + 32 * u < K) ? ldP(prow + c0 + 32 * u) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double* cr = cb + min(c0 + 32 * u, K - 1) * B;
                    double cv[BS];
#pragma unroll
                    for (int b = 0; b < BS; ++b) cv[b] = cr[b];
#pragma unroll
                    for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                        for (int b = 0; b < BS; ++b) s[rr][b] += pv[rr][u] * cv[b];
                }
            }
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                if (r0 + rr < K) {
                    const int row = r0 + rr;
#pragma unroll
                    for (int b = 0; b < BS; ++b) s[rr][b] = warp_sum(s[rr][b]);
                    const double m_r = mu[row];
#pragma unroll
                    for (int b = 0; b < BS; ++b) {
                        if (lane == 0) store_all<C>(tb + row * B + b, s[rr][b]);
                        const double cvb = cb[row * B + b];
                        pr[b] += cvb * m_r;
#pragma unroll
                        for (int b2 = 0; b2 < BS; ++b2) ps[b][b2] += cvb * s[rr][b2];
                    }
                }
            }
        }
        if (lane == 0) {
            const int slotw = crank * NWARP + warp;
#pragma unroll
            for (int b = 0; b < BS; ++b) {
                store_all<C>(part + slotw * PB + B * B + b, pr[b]);
#pragma unroll
                for (int b2 = 0; b2 < BS; ++b2) store_all<C>(part + slotw * PB + b * B + b2, ps[b][b2]);
            }
        }
        csync<C>();
        // block-wide sums of the per-warp partials: lane w holds warp w's, butterfly in a fixed order (deterministic)
#pragma unroll
        for (int b = 0; b < BS; ++b) {
            r[b] -= warp_sum(lane < C * NWARP ? part[lane * PB + B * B + b] : 0.0);
#pragma unroll
            for (int b2 = 0; b2 <= b; ++b2) S[b][b2] -= warp_sum(lane < C * NWARP ? part[lane * PB + b * B + b2] : 0.0);
        }
    }

    // P[r][c] += sgn * (tb[r] G) . tb[c] over the K x K active block; mu (and xs) follow.
    template <int BS>
    __device__ __forceinline__ void rank_update(const SmallSolve<B>& w, double sgn, bool draw) {
        for (int r0 = (crank * NWARP + warp) * ROWS; r0 < K; r0 += C * NWARP * ROWS) {
            double g[ROWS][BS];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                const double* tr = tb + min(r0 + rr, K - 1) * B;
#pragma unroll
                for (int k = 0; k < BS; ++k) {
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j < BS; ++j) v += tr[j] * w.G[j][k];
                    g[rr][k] = sgn * v;
                }
            }
            for (int c0 = lane; c0 < K; c0 += 128) {
                double pv[ROWS][4];
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr) {
                    const double* prow = P + (size_t)min(r0 + rr, K - 1) * ldp;
#pragma unroll
                    for (int u = 0; u < 4; ++u) pv[rr][u] = (c0 + 32 * u < K) ? ldP(prow + c0 + 32 * u) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c2 = c0 + 32 * u;
                    if (c2 < K) {
                        double tv[BS];
#pragma unroll
                        for (int k = 0; k < BS; ++k) tv[k] = tb[c2 * B + k];
#pragma unroll
                        for (int rr = 0; rr < ROWS; ++rr) {
                            if (r0 + rr < K) {
                                double acc = pv[rr][u];
#pragma unroll
                                for (int k = 0; k < BS; ++k) acc += g[rr][k] * tv[k];
                                P[(size_t)(r0 + rr) * ldp + c2] = acc;
                            }
                        }
                    }
                }
            }
        }
        for (int c1 = tid; c1 < K; c1 += NTHR) {
            double sm = 0.0, sx = 0.0;
#pragma unroll
            for (int k = 0; k < BS; ++k) { sm += tb[c1 * B + k] * w.gr[k]; sx += tb[c1 * B + k] * w.xm[k]; }
            mu[c1] -= sm;
            if (draw) xs[c1] -= sx;
        }
    }

    // Append the block evaluated by the last eval_add<BS>.
    template <int BS>
    __device__ __forceinline__ void commit_add(const SmallSolve<B>& w, int coord0, bool draw) {
        rank_update<BS>(w, 1.0, draw);
        for (int e = tid; e < K * BS; e += NTHR) {        // new border rows / columns: -t G
            const int c1 = e / BS, bb = e - c1 * BS;
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < BS; ++j) {
                double gsel = 0.0;
#pragma unroll
                for (int k = 0; k < BS; ++k) gsel = (k == bb) ? w.G[j][k] : gsel;
                v -= tb[c1 * B + j] * gsel;
            }
            P[(size_t)c1 * ldp + K + bb] = v;
            P[(size_t)(K + bb) * ldp + c1] = v;
        }
#pragma unroll
        for (int b = 0; b < BS; ++b) {
#pragma unroll
            for (int b2 = 0; b2 < BS; ++b2)
                if (tid == 32 + b * BS + b2) P[(size_t)(K + b) * ldp + K + b2] = w.G[b][b2];
            if (tid == b) {
                mu[K + b] = w.gr[b];
                if (draw) xs[K + b] = w.xm[b];
                cidx[K + b] = coord0 + b;
            }
        }
        csync<C>();
        K += BS;
    }

    // Remove the block at pos (G = P_mm^-1, gr = P_mm^-1 mu_m in w) and move the last block into its place.
    __device__ __forceinline__ void commit_remove(const SmallSolve<B>& w, int pos) {
        for (int e = tid; e < K * B; e += NTHR) {         // tb = P[:, pos block] (read as rows: P symmetric)
            const int bb = e / K, c2 = e - bb * K;
            tb[c2 * B + bb] = ldP(P + (size_t)(pos + bb) * ldp + c2);
        }
        csync<C>();
        if (la_G > 0) la_after_remove(w, pos);
        rank_update<B>(w, -1.0, false);                   // P -= t P_mm^-1 t^T,  mu_R -= P_Rm P_mm^-1 mu_m
        csync<C>();
        const int last = K - B;
        if (pos != last) {
            if (la_G > 0) la_move(pos, last);
            for (int e = tid; e < K * B; e += NTHR) {     // rows of the last block -> rows at pos
                const int bb = e / K, c2 = e - bb * K;
                P[(size_t)(pos + bb) * ldp + c2] = ldP(P + (size_t)(last + bb) * ldp + c2);
            }
            csync<C>();
            for (int e = tid; e < last * B; e += NTHR) {  // columns of the last block -> columns at pos
                const int c1 = e / B, bb = e - c1 * B;
                P[(size_t)c1 * ldp + pos + bb] = ldP(P + (size_t)c1 * ldp + last + bb);
            }
            if (tid < B) {
                mu[pos + tid] = mu[last + tid];
                cidx[pos + tid] = cidx[last + tid];
            }
            if (tid == 0) slot[cidx[last] / B] = pos;
        }
        csync<C>();
        K -= B;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Blocked dense factorisations for the two phases of the update that are NOT inherently sequential.
//   BUILD  P = (Jp_SS)^-1 for the active set the scan starts from.  Bordering one block at a time (as the scan must)
//          costs two latency-bound passes over P per block; here the matrix is gathered once and inverted in place by
//          the symmetric sweep operator, GW columns per step:  D = A_pp^-1,  T = A_op D,  A_oo -= T A_op^T (lower
//          triangle only),  A_op = T,  A_pp = -D;  after all blocks A = -(Jp_SS)^-1.  K/GW passes instead of 2K/B.
//   DRAW   [W_S; b] = L^-T (L^-1 hp + z) with Jp_SS = L L^T in ascending coordinate order, bias last -- literally
//          sample_gaussian(J=, h=) (regression.py:334): right-looking blocked Cholesky in place, the forward solve of
//          hp carried along the panel updates, one blocked back-substitution.  -sum log L_ii + 1/2 |L^-1 hp|^2 is the
//          posterior part of _marginal_likelihood (regression.py:369-376).
// The panels of a step live in shared memory ([b][row] layout: conflict-free for lanes that walk along a row of P);
// the pivot block is factorised by one warp in shared memory.  Measured on cfg3 (K ~ 265): BUILD 5.3 M -> see
// profiles/, DRAW 5.0 M cycles before.
constexpr int GW = 8;

template <int NTHR, int C>
struct Blocked {
    static constexpr int NWARP = NTHR / 32;
    double* P; int ldp;
    double *Ft, *Tt;           // [GW][ldt]
    double* M8;                // [GW*GW] pivot block, M8[GW*GW] = positive-definite flag
    int ldt, tid, lane, warp, crank;

    // P[i][j] -= sum_b T[b][i] * F[b][j]   for lo <= j <= i < K, rows and columns in [s0, s1) excluded
    __device__ __forceinline__ void tri_update(int K, int lo, int s0, int s1, const double* T, const double* F) {
        for (int i0 = lo + (crank * NWARP + warp) * 2; i0 < K; i0 += C * NWARP * 2) {
            const int i1 = i0 + 1;
            const bool ok0 = !(i0 >= s0 && i0 < s1), ok1 = (i1 < K) && !(i1 >= s0 && i1 < s1);
            if (!ok0 && !ok1) continue;
            double t0[GW], t1[GW];
#pragma unroll
            for (int b = 0; b < GW; ++b) { t0[b] = T[b * ldt + i0]; t1[b] = T[b * ldt + min(i1, K - 1)]; }
            double* row0 = P + (size_t)i0 * ldp;
            double* row1 = P + (size_t)min(i1, K - 1) * ldp;
            const int jend = ok1 ? i1 : i0;
            for (int j0 = lo + lane; j0 <= jend; j0 += 128) {
                double p0[4], p1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u;
                    p0[u] = (ok0 && j <= i0) ? ldP(row0 + j) : 0.0;
                    p1[u] = (ok1 && j <= i1) ? ldP(row1 + j) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u;
                    if (j <= jend && !(j >= s0 && j < s1)) {
                        double a0 = p0[u], a1 = p1[u];
#pragma unroll
                        for (int b = 0; b < GW; ++b) {
                            const double f = F[b * ldt + j];
                            a0 -= t0[b] * f;
                            a1 -= t1[b] * f;
                        }
                        if (ok0 && j <= i0) row0[j] = a0;
                        if (ok1 && j <= i1) row1[j] = a1;
                    }
                }
            }
        }
    }

    // pivot block (lower triangle of P at [k0, k0+w)) -> M8: full symmetric (sym) or lower with zero upper, identity padded
    __device__ __forceinline__ void load_pivot(int k0, int w, bool sym) {
        if (tid < GW * GW) {
            const int r = tid / GW, c = tid - r * GW;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < w && c < w) v = (c <= r) ? ldP(P + (size_t)(k0 + r) * ldp + k0 + c) : (sym ? ldP(P + (size_t)(k0 + c) * ldp + k0 + r) : 0.0);
            M8[tid] = v;
        }
    }

    // warp 0: M8 <- M8^-1 (Gauss-Jordan, no pivoting: the block is a Schur complement of an SPD matrix)
    __device__ __forceinline__ bool inv8(int w) {
        bool ok = true;
        const int e0 = lane, e1 = lane + 32;
        const int r0 = e0 / GW, c0 = e0 - r0 * GW, r1 = e1 / GW, c1 = e1 - r1 * GW;
        for (int k = 0; k < w; ++k) {
            const double d = M8[k * GW + k];
            ok = ok && (d > 0.0);
            const double id = 1.0 / d;
            const double ark0 = M8[r0 * GW + k], akc0 = M8[k * GW + c0], a0 = M8[e0];
            const double ark1 = M8[r1 * GW + k], akc1 = M8[k * GW + c1], a1 = M8[e1];
            const double n0 = (r0 == k) ? ((c0 == k) ? id : akc0 * id) : ((c0 == k) ? -ark0 * id : a0 - ark0 * akc0 * id);
            const double n1 = (r1 == k) ? ((c1 == k) ? id : akc1 * id) : ((c1 == k) ? -ark1 * id : a1 - ark1 * akc1 * id);
            __syncwarp();
            M8[e0] = n0; M8[e1] = n1;
            __syncwarp();
        }
        return ok;
    }

    // warp 0: lower triangle of M8 <- its Cholesky factor
    __device__ __forceinline__ bool chol8(int w) {
        bool ok = true;
        const int e0 = lane, e1 = lane + 32;
        const int r0 = e0 / GW, c0 = e0 - r0 * GW, r1 = e1 / GW, c1 = e1 - r1 * GW;
        for (int k = 0; k < w; ++k) {
            const double d = M8[k * GW + k];
            ok = ok && (d > 0.0);
            const double il = rsqrt(d);
            const double ark0 = M8[r0 * GW + k], ack0 = M8[c0 * GW + k], a0 = M8[e0];
            const double ark1 = M8[r1 * GW + k], ack1 = M8[c1 * GW + k], a1 = M8[e1];
            // column k: a_ik / sqrt(d) (diagonal: sqrt(d)); trailing lower triangle: a_ij - a_ik a_jk / d
            const double n0 = (c0 == k) ? ((r0 == k) ? d * il : ark0 * il) : a0 - ark0 * ack0 * (il * il);
            const double n1 = (c1 == k) ? ((r1 == k) ? d * il : ark1 * il) : a1 - ark1 * ack1 * (il * il);
            __syncwarp();
            if (r0 >= k && c0 >= k && c0 <= r0) M8[e0] = n0;
            if (r1 >= k && c1 >= k && c1 <= r1) M8[e1] = n1;
            __syncwarp();
        }
        return ok;
    }

    // Lower triangle of P (K x K, SPD) -> full symmetric inverse, in place.  false: not positive definite.
    __device__ __forceinline__ bool invert(int K) {
        for (int k0 = 0; k0 < K; k0 += GW) {
            const int w = min(GW, K - k0), k1 = k0 + w;
            load_pivot(k0, w, true);
            csync<C>();
            if (warp == 0) {
                const bool ok = inv8(w);
                if (lane == 0) M8[GW * GW] = ok ? 1.0 : 0.0;
            }
            csync<C>();
            if (M8[GW * GW] == 0.0) return false;
            for (int j = tid; j < K; j += NTHR) {
                double f[GW];
                const bool inp = (j >= k0 && j < k1);
#pragma unroll
                for (int b = 0; b < GW; ++b)
                    f[b] = (b < w && !inp) ? ((j < k0) ? ldP(P + (size_t)(k0 + b) * ldp + j) : ldP(P + (size_t)j * ldp + k0 + b)) : 0.0;
#pragma unroll
                for (int b = 0; b < GW; ++b) {
                    double t = 0.0;
#pragma unroll
                    for (int c = 0; c < GW; ++c) t += f[c] * M8[c * GW + b];
                    Ft[b * ldt + j] = f[b];
                    Tt[b * ldt + j] = inp ? 0.0 : t;
                }
            }
            csync<C>();
            tri_update(K, 0, k0, k1, Tt, Ft);
            for (int j = tid; j < K; j += NTHR) {
                if (j >= k0 && j < k1) continue;
                for (int b = 0; b < w; ++b) {
                    if (j < k0) P[(size_t)(k0 + b) * ldp + j] = Tt[b * ldt + j];
                    else P[(size_t)j * ldp + k0 + b] = Tt[b * ldt + j];
                }
            }
            if (tid < GW * GW) {
                const int r = tid / GW, c = tid - r * GW;
                if (c <= r && r < w) P[(size_t)(k0 + r) * ldp + k0 + c] = -M8[tid];
            }
            csync<C>();
        }
        for (int i = crank * NWARP + warp; i < K; i += C * NWARP) {   // P = -A, mirrored into the upper triangle
            for (int j = lane; j <= i; j += 32) {
                const double v = -ldP(P + (size_t)i * ldp + j);
                P[(size_t)i * ldp + j] = v;
                P[(size_t)j * ldp + i] = v;
            }
        }
        csync<C>();
        return true;
    }

    // Lower triangle of P (K x K, SPD) -> its Cholesky factor L in place; hv (K, shared) -> L^-1 hv; *half_logdet =
    // sum_i log L_ii.  false: not positive definite.
    __device__ __forceinline__ bool cholesky(int K, double* hv, double* half_logdet) {
        double ld = 0.0;
        for (int k0 = 0; k0 < K; k0 += GW) {
            const int w = min(GW, K - k0), k1 = k0 + w;
            load_pivot(k0, w, false);
            csync<C>();
            if (warp == 0) {
                const bool ok = chol8(w);
                if (lane == 0) M8[GW * GW] = ok ? 1.0 : 0.0;
            }
            csync<C>();
            if (M8[GW * GW] == 0.0) return false;
            double y[GW];
#pragma unroll
            for (int b = 0; b < GW; ++b) {
                double v = (b < w) ? hv[k0 + b] : 0.0;
#pragma unroll
                for (int c = 0; c < b; ++c) v -= M8[b * GW + c] * y[c];
                y[b] = v / M8[b * GW + b];
                if (b < w) ld += log(M8[b * GW + b]);
            }
            if (tid < GW * GW) {
                const int r = tid / GW, c = tid - r * GW;
                if (c <= r && r < w) P[(size_t)(k0 + r) * ldp + k0 + c] = M8[tid];
            }
            // panel L[j][p] = C[j][p] L_pp^-T and the forward solve; rows split over the cluster (read-modify-write of
            // P), the panel and the updated right-hand side go to every CTA
            for (int j = k1 + crank * NTHR + tid; j < K; j += C * NTHR) {
                double v[GW];
#pragma unroll
                for (int b = 0; b < GW; ++b) v[b] = (b < w) ? ldP(P + (size_t)j * ldp + k0 + b) : 0.0;
                double hj = hv[j];
#pragma unroll
                for (int b = 0; b < GW; ++b) {
                    double sacc = v[b];
#pragma unroll
                    for (int c = 0; c < b; ++c) sacc -= v[c] * M8[b * GW + c];
                    v[b] = sacc / M8[b * GW + b];
                    hj -= v[b] * y[b];
                    store_all<C>(Tt + b * ldt + j, v[b]);
                    if (b < w) P[(size_t)j * ldp + k0 + b] = v[b];
                }
                store_all<C>(hv + j, hj);
            }
            csync<C>();
#pragma unroll
            for (int b = 0; b < GW; ++b)
                if (tid == b && b < w) hv[k0 + b] = y[b];
            tri_update(K, k1, -1, -1, Tt, Tt);
            csync<C>();
        }
        *half_logdet = ld;
        return true;
    }

    // v (K, shared) <- L^-T v with L in the lower triangle of P; s: K doubles of shared scratch.
    __device__ __forceinline__ void backsolve(int K, double* v, double* s) {
        for (int j = tid; j < K; j += NTHR) s[j] = 0.0;
        csync<C>();
        for (int k0 = ((K - 1) / GW) * GW; k0 >= 0; k0 -= GW) {
            const int w = min(GW, K - k0);
            load_pivot(k0, w, false);
            csync<C>();
            double o[GW];
#pragma unroll
            for (int b = GW - 1; b >= 0; --b) {
                double t = (b < w) ? v[k0 + b] - s[k0 + b] : 0.0;
#pragma unroll
                for (int c = b + 1; c < GW; ++c) t -= M8[c * GW + b] * o[c];
                o[b] = t / M8[b * GW + b];
            }
            for (int j = tid; j < k0; j += NTHR) {
                double acc = s[j];
                for (int b = 0; b < w; ++b) acc += ldP(P + (size_t)(k0 + b) * ldp + j) * o[b];
                s[j] = acc;
            }
            csync<C>();
#pragma unroll
            for (int b = 0; b < GW; ++b)
                if (tid == b && b < w) v[k0 + b] = o[b];
        }
        csync<C>();
    }
};

// DRAW, blocked: [W_S; b] = L^-T (L^-1 hp + z), Jp_SS = L L^T in ascending coordinate order with the bias last
// (np.ix_(mask, mask), regression.py:350-353; sample_gaussian(J=, h=), :334).  On return c.cidx / c.mu / c.xs / c.K
// describe the draw (value of coordinate cidx[k] = mu[k] + xs[k]) and ml holds _marginal_likelihood of the final a
// (the cprior terms are added by thread 0 only).  false: Jp_SS is not positive definite.
template <int B, int NTHR, int C>
__device__ __forceinline__ bool blocked_draw(FastCtx<B, NTHR, C>& c, Blocked<NTHR, C>& blk, const unsigned char* a,
                                             const double* zc, const double* cprior, bool want_ml, int* ksh, double& ml) {
    const int N = c.N, D = c.D, tid = c.tid;
    csync<C>();
    if (tid == 0) {
        int K = 0;
        for (int m = 0; m < N; ++m)
            if (a[m])
                for (int b = 0; b < B; ++b) c.cidx[K++] = m * B + b;
        c.cidx[K++] = D - 1;
        *ksh = K;
    }
    csync<C>();
    const int K = *ksh;
    for (int i = c.warp; i < K; i += NTHR / 32) {
        const int ci = c.cidx[i];
        for (int j = c.lane; j <= i; j += 32) c.P[(size_t)i * c.ldp + j] = c.Jp(ci, c.cidx[j]);
    }
    for (int j = tid; j < K; j += NTHR) c.mu[j] = c.hp(c.cidx[j]);
    csync<C>();
    double half_logdet = 0.0;
    if (!blk.cholesky(K, c.mu, &half_logdet)) return false;
    double quad = 0.0;                                    // every thread, same order: |L^-1 hp|^2
    for (int j = 0; j < K; ++j) quad += c.mu[j] * c.mu[j];
    for (int j = tid; j < K; j += NTHR) c.xs[j] = c.mu[j] + zc[c.cidx[j]];
    csync<C>();
    blk.backsolve(K, c.xs, c.tb);                         // xs = L^-T (L^-1 hp + z) = Jp^-1 hp + L^-T z
    for (int j = tid; j < K; j += NTHR) c.mu[j] = 0.0;
    ml = -half_logdet + 0.5 * quad + 0.5 * log(c.J0b) - 0.5 * c.h0b * c.h0b / c.J0b;
    if (want_ml && tid == 0)
        for (int m = 0; m < N; ++m)
            if (a[m]) ml += cprior[m];
    c.K = K;
    return true;
}

template <int B, int NTHR>
size_t fast_smem_bytes(int N) {
    const int D = N * B + 1, Dpad = (D + 1) & ~1;
    return ((size_t)2 * Dpad + (size_t)2 * Dpad * B + (size_t)(NTHR / 32) * (B * B + B) +
            (size_t)2 * GW * Dpad + GW * GW + 2) * sizeof(double) +
           ((size_t)Dpad + N + 2 + 8) * sizeof(int);
}

template <int B, int NTHR, int MINB, bool BLK, int C, bool DBG>
__global__ void __launch_bounds__(NTHR, MINB)
spike_slab_fast_kernel(SpikeSlabArgs A) {
    extern __shared__ __align__(16) double ssm[];
    const int ln = blockIdx.x / C;
    const int N = A.N, D = A.D;
    const int Dpad = (D + 1) & ~1;

    FastCtx<B, NTHR, C> c;
    c.N = N; c.D = D; c.NB = N * B; c.ldj = A.ldj; c.ldp = D;
    c.Jn = A.J + (size_t)ln * A.stride_n;
    c.hn = A.h + (size_t)ln * A.ldh;
    c.J0w = A.J0w + (size_t)ln * N * B * B;
    c.h0w = A.h0w + (size_t)ln * N * B;
    c.J0b = A.J0b[ln]; c.h0b = A.h0b[ln];
    c.P = A.P + (size_t)ln * D * D;
    double* p = ssm;
    c.mu = p; p += Dpad;
    c.xs = p; p += Dpad;
    c.cb = p; p += (size_t)Dpad * B;
    c.tb = p; p += (size_t)Dpad * B;
    c.part = p; p += (size_t)C * (NTHR / 32) * (B * B + B);
    Blocked<NTHR, C> blk;
    blk.P = c.P; blk.ldp = c.ldp; blk.ldt = Dpad;
    blk.Ft = p; p += (size_t)GW * Dpad;
    blk.Tt = p; p += (size_t)GW * Dpad;
    blk.M8 = p; p += GW * GW + 2;
    c.la_G = A.la_G; c.la_ld = Dpad * B;
    c.Cb = p; p += (size_t)c.la_G * c.la_ld;
    c.Tb = p; p += (size_t)c.la_G * c.la_ld;
    c.la_part = p; p += (size_t)(NTHR / 32) * B * 16;
    c.la_E = p; p += 16 * B * B;
    blk.tid = threadIdx.x; blk.lane = threadIdx.x & 31; blk.warp = threadIdx.x >> 5;
    blk.crank = (int)crank_of<C>();
    c.cidx = reinterpret_cast<int*>(p);
    c.slot = c.cidx + Dpad;
    int* ksh = c.slot + N;                                  // block-wide scalar: size of an index list
    c.cand = ksh + 2;
    if (threadIdx.x < 8) c.cand[threadIdx.x] = -1;
    c.tid = threadIdx.x; c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5;
    c.crank = (int)crank_of<C>();
    c.K = 0;
    const int tid = threadIdx.x;

    unsigned char* a = A.a + (size_t)ln * N;
    const double* cprior = A.cprior + (size_t)ln * N;
    const double* lrho = A.logit_rho + (size_t)ln * N;
    const int* perm = A.perm + (size_t)ln * N;
    const double* us = A.us + (size_t)ln * N;
    const double* zc = A.z + (size_t)ln * A.ldz;

    for (int m = tid; m < N; m += NTHR) c.slot[m] = -1;
    csync<C>();

    // One loop over all steps of the three phases:
    //   BUILD  P, mu for the current active set: bias first, then the active blocks in ascending order
    //   SCAN   the collapsed scan over a random permutation (regression.py:286-320)
    //   DRAW   restart from the empty set and border in ascending coordinate order, bias last, with the normals z:
    //          step for step the back-substitution x = chol(Jp_SS)^-T z of sample_gaussian; mu + x is the draw and
    //          the accumulated dpost is _marginal_likelihood (regression.py:343-378).
    enum { PH_BUILD_BIAS, PH_BUILD, PH_SCAN, PH_DRAW, PH_DONE };
    int phase = A.do_scan[ln] ? PH_BUILD_BIAS : PH_DRAW;
    int cursor = 0, fail = 0;
    double ml = 0.0;
    SmallSolve<B> w;
    long long clk0 = DBG ? clock64() : 0, clk_build = 0, clk_scan = 0;
    int k_build = 0, k_scan = 0, n_eval = 0, n_flip = 0;
    long long dbg_rm = 0, dbg_ev = 0, dbg_ca = 0, dbg_cr = 0, dbg_no = 0;
    if (BLK && phase == PH_BUILD_BIAS) {
        // BUILD, blocked: index list (bias first, then the active blocks in ascending order -- the order the bordering
        // build produces), gather Jp_SS, invert in place, mu = P hp_S
        if (tid == 0) {
            int K = 0;
            c.cidx[K++] = D - 1;
            for (int m = 0; m < N; ++m)
                if (a[m]) {
                    c.slot[m] = K;
                    for (int b = 0; b < B; ++b) c.cidx[K++] = m * B + b;
                }
            *ksh = K;
        }
        csync<C>();
        const int K = *ksh;
        for (int i = c.warp; i < K; i += NTHR / 32) {
            const int ci = c.cidx[i];
            for (int j = c.lane; j <= i; j += 32) c.P[(size_t)i * c.ldp + j] = c.Jp(ci, c.cidx[j]);
        }
        for (int j = tid; j < K; j += NTHR) c.tb[j] = c.hp(c.cidx[j]);
        csync<C>();
        if (!blk.invert(K)) {
            fail = 1;
            phase = PH_DONE;
        } else {
            for (int i = c.warp; i < K; i += NTHR / 32) {
                double acc = 0.0;
                for (int j = c.lane; j < K; j += 32) acc += ldP(c.P + (size_t)i * c.ldp + j) * c.tb[j];
                acc = warp_sum(acc);
                if (c.lane == 0) c.mu[i] = acc;
            }
            csync<C>();
            c.K = K;
            phase = PH_SCAN;
            if (DBG) { clk_build = clock64(); k_build = K; }
        }
    }
    while (phase != PH_DONE && !(BLK && phase == PH_DRAW)) {
        int m = -1;
        bool is_bias = false;
        if (phase == PH_BUILD_BIAS) {
            is_bias = true;
        } else if (phase == PH_BUILD) {
            while (cursor < N && !a[cursor]) ++cursor;
            if (cursor == N) { phase = PH_SCAN; cursor = 0; if (DBG) { clk_build = clock64(); k_build = c.K; } continue; }
            m = cursor++;
        } else if (phase == PH_SCAN) {
            if (cursor == N) {
                csync<C>(); if (DBG) { clk_scan = clock64(); k_scan = c.K; } phase = PH_DRAW; cursor = 0; c.K = 0;
                if (BLK) break;
                continue;
            }
            m = perm[cursor];
        } else {
            while (cursor < N && !a[cursor]) ++cursor;
            if (cursor == N) is_bias = true; else m = cursor++;
        }
        const bool scan = (phase == PH_SCAN), draw = (phase == PH_DRAW);
        const int pos = scan ? c.slot[m] : -1;
        const int coord0 = is_bias ? D - 1 : m * B;
        double S[B][B], r[B];
        int la_slot = -1;
        const long long t_a = DBG ? clock64() : 0;
        if (pos >= 0) {
            // removal: ml(with) - ml(without) read off P and mu, no pass over P, no barrier
#pragma unroll
            for (int i = 0; i < B; ++i) {
#pragma unroll
                for (int k = 0; k <= i; ++k) S[i][k] = ldP(c.P + (size_t)(pos + i) * c.ldp + pos + k);
                r[i] = c.mu[pos + i];
            }
            small_factor<B, B>(w, S, r, 1.0);
        } else if (is_bias) {
            c.template eval_add<1>(coord0, S, r);
            small_factor<B, 1>(w, S, r, -1.0);
        } else {
            if (scan && c.la_G > 0) {
                la_slot = c.la_find(m);
                if (la_slot < 0) { c.la_refill(perm, cursor); la_slot = c.la_find(m); }
                c.la_eval(la_slot, coord0, S, r);
            } else {
                c.template eval_add<B>(coord0, S, r);
            }
            small_factor<B, B>(w, S, r, -1.0);
        }
        const long long t_b = DBG ? clock64() : 0;
        if (DBG && scan) { if (pos >= 0) { dbg_rm += t_b - t_a; } else { dbg_ev += t_b - t_a; } }
        if (!(w.dpost == w.dpost)) { fail = 1; break; }
        bool do_add = true, do_remove = false;
        if (scan) {
            const double lo = w.dpost + cprior[m] + lrho[m];
            const double p0 = 1.0 / (1.0 + exp(lo));
            const int v = us[cursor] > p0;
            if (A.logodds && tid == 0) A.logodds[(size_t)ln * N + cursor] = lo;
            do_add = (pos < 0) && v;
            do_remove = (pos >= 0) && !v;
            if (DBG) { n_eval += (pos < 0); n_flip += (do_add || do_remove); }
            ++cursor;
        } else if (draw) {
            ml += w.dpost + (is_bias ? 0.5 * log(c.J0b) - 0.5 * c.h0b * c.h0b / c.J0b : cprior[m]);
        }
        if (do_add) {
            if (!is_bias && tid == 0) { c.slot[m] = c.K; a[m] = 1; }
            if (is_bias) {
                small_finish<B, 1>(w, r, draw ? zc : nullptr, coord0);
                c.template commit_add<1>(w, coord0, draw);
            } else {
                small_finish<B, B>(w, r, draw ? zc : nullptr, coord0);
                if (la_slot >= 0) c.la_after_add(la_slot, w, coord0);
                c.template commit_add<B>(w, coord0, draw);
            }
        } else if (do_remove) {
            small_finish<B, B>(w, r, nullptr, 0);
            csync<C>();                              // every thread has read slot[m] before it changes
            if (tid == 0) { c.slot[m] = -1; a[m] = 0; }
            c.commit_remove(w, pos);
        } else if (la_slot >= 0) {
            c.la_release(la_slot);                        // evaluated and left inactive
        }
        if (DBG && scan) { if (do_add) dbg_ca += clock64() - t_b; else if (do_remove) dbg_cr += clock64() - t_b; else dbg_no += clock64() - t_b; }
        if (phase == PH_BUILD_BIAS) phase = PH_BUILD;
        else if (draw && is_bias) phase = PH_DONE;
    }
    if (BLK && !fail && phase == PH_DRAW) {
        if (!blocked_draw<B, NTHR, C>(c, blk, a, zc, cprior, A.ml != nullptr, ksh, ml)) fail = 1;
    }
    csync<C>();
    double* Wn = A.W + (size_t)ln * N * B;
    for (int e = tid; e < N * B; e += NTHR) Wn[e] = 0.0;
    csync<C>();
    if (!fail) {
        for (int k = tid; k < c.K; k += NTHR) {
            const int d = c.cidx[k];
            const double v = c.mu[k] + c.xs[k];
            if (d < N * B) Wn[d] = v; else A.bias[ln] = v;
        }
    }
    if (tid == 0) {
        if (A.ml) A.ml[ln] = fail ? nan("") : ml;
        A.status[ln] = fail;
        if (DBG && A.debug && ln == 0 && c.crank == 0)
            printf("spike_slab cta0: build %lld cyc (K=%d)  scan %lld cyc (K=%d, %d add-evals, %d flips)  draw %lld cyc\n",
                   clk_build - clk0, k_build, clk_scan - clk_build, k_scan, n_eval, n_flip, clock64() - clk_scan);
        if (DBG && A.debug && ln == 0 && c.crank == 0)
            printf("   scan split: removal-evals %lld  add-evals %lld  commit-add %lld  commit-remove %lld  no-flip tail %lld\n",
                   dbg_rm, dbg_ev, dbg_ca, dbg_cr, dbg_no);
    }
}

template <int B, int NTHR, int MINB, bool BLK, int C>
int launch_fast(const SpikeSlabArgs& A, cudaStream_t stream) {
    const int Dpad_ = (A.N * B + 2) & ~1;
    size_t smem = fast_smem_bytes<B, NTHR>(A.N) + (size_t)(C - 1) * (NTHR / 32) * (B * B + B) * sizeof(double) +
                  ((size_t)(NTHR / 32) * B * 16 + 16 * B * B) * sizeof(double) + 8 * sizeof(int);
    // lookahead table of the scan: as many slots as shared memory allows (<= 8, G * B <= 16); off below two slots
    static int la_env = -1;
    if (la_env < 0) { const char* e = getenv("PYGLM_SS_LOOKAHEAD"); la_env = e ? atoi(e) : 8; }
    const size_t per_slot = (size_t)2 * Dpad_ * B * sizeof(double);
    // MINB = 2: two CTAs (two neurons) share an SM -- half of its shared memory each, so a shorter lookahead table
    const size_t budget = (MINB >= 2 ? 113 : 227) * 1024;
    int G = (smem < budget) ? (int)((budget - smem) / per_slot) : 0;
    if (G > 16 / B) G = 16 / B;
    if (G > 8) G = 8;
    if (G > la_env) G = la_env;
    if (G < 2 || C != 1) G = 0;
    smem += (size_t)G * per_slot;
    SpikeSlabArgs A2 = A;
    A2.la_G = G;
    if (smem > budget) {
        pyglm_set_error("pyglm_spike_slab_update: N*B=%d too large for the shared-memory state (%zu bytes)", A.N * B, smem);
        return PYGLM_ERR_INVALID;
    }
    if (A.debug) {                                        // PYGLM_SS_DEBUG=1: the instrumented instantiation
        PYGLM_CUDA(cudaFuncSetAttribute(spike_slab_fast_kernel<B, NTHR, MINB, BLK, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        PYGLM_CUDA(cudaFuncSetAttribute(spike_slab_fast_kernel<B, NTHR, MINB, BLK, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(A.n_loc * C, 1, 1);
    cfg.blockDim = dim3(NTHR, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (A.debug) PYGLM_CUDA(cudaLaunchKernelEx(&cfg, spike_slab_fast_kernel<B, NTHR, MINB, BLK, C, true>, A2));
    else PYGLM_CUDA(cudaLaunchKernelEx(&cfg, spike_slab_fast_kernel<B, NTHR, MINB, BLK, C, false>, A2));
    return PYGLM_OK;
}

template <int B>
int launch_fast_variant(const SpikeSlabArgs& A, int variant, cudaStream_t stream) {
    switch (variant) {
        case 2: return launch_fast<B, 512, 1, true, 2>(A, stream);     // 2-CTA cluster per neuron: measured 15 % faster per
                                                                       // neuron on twice the SMs (the K x K passes are not what
                                                                       // bounds a step) -- kept for measurement only
        case 3: return launch_fast<B, 512, 1, false, 1>(A, stream);    // bordering build / draw (the first version; A/B runs)
        // two neurons per SM, 256 threads each (N = 200 then fits ONE wave of 148 SMs instead of 148 + 52, at the
        // price of a 2-slot lookahead table).  Measured at cfg3 (profiles/r01n_scan_variants.log): 4.5-5.2 ms
        // against 3.6-4.4 ms -- the passes over P are L2-throughput bound, co-residency does not hide them; with 512
        // threads and 64 registers (spills) 7-7.8 ms.  Kept for measurement only.  Falls back to one CTA per SM when
        // the state of a neuron does not fit half an SM's shared memory.
        case 4: if constexpr (B == 2) { if (launch_fast<B, 256, 2, true, 1>(A, stream) == PYGLM_OK) return PYGLM_OK; } break;
        default: break;
    }
    return launch_fast<B, 512, 1, true, 1>(A, stream);                 // one CTA per neuron, one per SM
}

// perm: Fisher-Yates permutation of 0..N-1 per local neuron; us: N uniforms; z: D normals keyed by coordinate.
// Streams are keyed by the GLOBAL neuron index so that sharding does not change the draws.
__global__ void scan_randomness_kernel(int N, int D, int n_loc, int n_off, unsigned long long seed, unsigned call_id,
                                       int* __restrict__ perm, double* __restrict__ us, double* __restrict__ z, int ldz) {
    const int ln = blockIdx.x;
    const unsigned long long base = (unsigned long long)(n_off + ln) << 32;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        PhiloxStream r; r.seed(seed, call_id, base + 2ull * i);
        us[(size_t)ln * N + i] = r.unif();
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        PhiloxStream r; r.seed(seed, call_id, base + 2ull * d + 1ull);
        z[(size_t)ln * ldz + d] = r.norm();
    }
    if (threadIdx.x == 0) {
        int* pm = perm + (size_t)ln * N;
        for (int i = 0; i < N; ++i) pm[i] = i;
        PhiloxStream r; r.seed(seed, call_id, base + 0xFFFFFFFFull);
        for (int i = N - 1; i > 0; --i) {
            const int j = (int)(r.unif() * (i + 1));
            const int tmp = pm[i]; pm[i] = pm[j]; pm[j] = tmp;
        }
    }
}

}  // namespace

extern "C" size_t pyglm_spike_slab_workspace_doubles(int N, int B, int n_loc) {
    const size_t D = (size_t)N * B + 1;
    return (size_t)n_loc * D * D;
}

extern "C" int pyglm_scan_randomness(int N, int B, int n_loc, int n_off, unsigned long long seed, unsigned call_id,
                                     int* perm, double* us, double* z, int ldz, cudaStream_t stream) {
    PYGLM_CHECK_ARG(perm && us && z && N > 0 && B > 0 && n_loc > 0 && ldz >= N * B + 1, "pyglm_scan_randomness: bad arguments");
    scan_randomness_kernel<<<n_loc, 256, 0, stream>>>(N, N * B + 1, n_loc, n_off, seed, call_id, perm, us, z, ldz);
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}

// One spike-and-slab update of (a, W, b) for n_loc postsynaptic neurons.  See SpikeSlabArgs for shapes.
extern "C" int pyglm_spike_slab_update(int N, int B, int n_loc,
                                       const double* J, long long stride_n, int ldj, const double* h, int ldh,
                                       const double* J0w, const double* h0w, const double* J0b, const double* h0b,
                                       const double* cprior, const double* logit_rho,
                                       const int* perm, const double* us, const double* z, int ldz,
                                       const unsigned char* do_scan, unsigned char* a, double* W, double* bias,
                                       double* P_workspace, double* logodds, double* ml, int* status,
                                       cudaStream_t stream) {
    PYGLM_CHECK_ARG(N > 0 && B > 0 && n_loc > 0, "pyglm_spike_slab_update: N, B, n_loc must be positive");
    PYGLM_CHECK_ARG(B <= SS_BMAX, "pyglm_spike_slab_update: B=%d exceeds the supported block size %d", B, SS_BMAX);
    PYGLM_CHECK_ARG(J && h && J0w && h0w && J0b && h0b && cprior && logit_rho && perm && us && z && do_scan && a && W && bias && P_workspace && status,
                    "pyglm_spike_slab_update: null pointer");
    const int D = N * B + 1;
    PYGLM_CHECK_ARG(ldj >= D && ldh >= D && ldz >= D, "pyglm_spike_slab_update: leading dimensions must be >= D=%d", D);
    SpikeSlabArgs A;
    A.N = N; A.B = B; A.D = D; A.n_loc = n_loc;
    A.J = J; A.stride_n = stride_n; A.ldj = ldj; A.h = h; A.ldh = ldh;
    A.J0w = J0w; A.h0w = h0w; A.J0b = J0b; A.h0b = h0b; A.cprior = cprior; A.logit_rho = logit_rho;
    A.perm = perm; A.us = us; A.z = z; A.ldz = ldz; A.do_scan = do_scan; A.a = a; A.W = W; A.bias = bias;
    A.P = P_workspace; A.logodds = logodds; A.ml = ml; A.status = status;
    static int debug = -1;
    if (debug < 0) { const char* e = getenv("PYGLM_SS_DEBUG"); debug = e ? atoi(e) : 0; }
    A.debug = debug;
    const int Dpad = (D + 1) & ~1;
    size_t smem = ((size_t)2 * Dpad + (size_t)3 * Dpad * B + 2 * SS_BMAX * SS_BMAX + 3 * SS_BMAX + 8) * sizeof(double)
                + ((size_t)Dpad + N) * sizeof(int);
    PYGLM_CHECK_ARG(smem <= 227 * 1024, "pyglm_spike_slab_update: N*B=%d too large for the shared-memory state (%zu bytes)", N * B, smem);
    const char* venv = getenv("PYGLM_SS_VARIANT");         // read per call: tests switch kernels inside one process
    const int variant = venv ? atoi(venv) : 0;
    // Cluster kernel with P in distributed shared memory (spike_slab_dsm.cu): PYGLM_SS_VARIANT=82 / 84 / 88 force it
    // with 2 / 4 / 8 CTAs per neuron, 80 lets it pick the cluster size; by default it runs when it applies and the
    // neurons would leave SMs idle (spike_slab_dsm_preferred); 1 = never.
    if (B <= 4 && (variant == 80 || variant == 82 || variant == 84 || variant == 88 ||
                   (variant == 0 && spike_slab_dsm_preferred(A)))) {
        const int rc = spike_slab_dsm_launch(A, variant >= 80 ? variant - 80 : 0, stream);
        if (rc != PYGLM_ERR_UNSUPPORTED) return rc;
        PYGLM_CHECK_ARG(variant == 0, "pyglm_spike_slab_update: PYGLM_SS_VARIANT=%d: the state of a neuron (N*B=%d) does not "
                        "fit the shared memory of such a cluster", variant, N * B);
    }
    if (B <= 4 && (variant < 10 || variant >= 80)) {     // variant 10..79: force the generic kernel (tests, B > 4)
        switch (B) {
            case 1: return launch_fast_variant<1>(A, variant, stream);
            case 2: return launch_fast_variant<2>(A, variant, stream);
            case 3: return launch_fast_variant<3>(A, variant, stream);
            default: return launch_fast_variant<4>(A, variant, stream);
        }
    }
#define SS_LAUNCH(THR, MINB)                                                                                              \
    do {                                                                                                                   \
        PYGLM_CUDA(cudaFuncSetAttribute(spike_slab_kernel<THR, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        spike_slab_kernel<THR, MINB><<<n_loc, THR, smem, stream>>>(A);                                                      \
    } while (0)
    SS_LAUNCH(512, 1);
#undef SS_LAUNCH
    PYGLM_LAUNCH_CHECK();
    return PYGLM_OK;
}
