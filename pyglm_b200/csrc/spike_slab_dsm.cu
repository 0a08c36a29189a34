// (4b) Spike-and-slab update of (a_n, W_n, b_n) with ONE THREAD-BLOCK CLUSTER per postsynaptic neuron and the active-set
// inverse P = (Jp_SS)^-1 resident in DISTRIBUTED SHARED MEMORY (regression.py:265-340: _collapsed_resample_a :282-320,
// _marginal_likelihood :343-378, _resample_W :323-340).
//
// Why: with one CTA per neuron (spike_slab.cu) P (K x K doubles, 640 KB at cfg3) lives in L2 and every committed flip,
// every refill of the lookahead table and every panel step of the blocked inverse / Cholesky is an L2-latency-bound
// pass over it: ~2.7 ms per neuron however many SMs are idle -- the floor of the multi-GPU sweep.
// Here the C CTAs of a cluster own the rows of P cyclically (row p -> CTA p mod C, local row p / C, full rows), all
// small state is REPLICATED in every CTA and updated redundantly by identical instruction streams (so every CTA takes
// bit-identical decisions), and what crosses CTAs goes through DSMEM stores followed by one cluster barrier:
//   BUILD  symmetric sweep inverse, 8 columns per step: the column panel A[:, k0:k0+8] is all-gathered (each CTA stores
//          its rows' entries into every CTA), every CTA inverts the 8 x 8 pivot itself and updates its own rows.
//   SCAN   evaluations need NO communication: removals read the replicated diagonal blocks of P and mu; additions read
//          a lookahead table whose small matrices M = T^T C (all slot pairs) and r_g = hp_g - c_g^T mu are kept current
//          under every flip by O(B^3) updates (algebra: profiles/proto_scan_dsm.py, checked against the oracle).  A
//          committed flip all-gathers ONE K x B block (t_g for an addition, P[:, pos] for a removal), then every CTA
//          rank-B-updates its own rows of P and of the table from shared memory.  A removed block leaves a zeroed
//          tombstone that the next addition reuses (nothing moves between CTAs).  A table refill is one local
//          (K/C x K)(K x G B) product plus an all-reduce of (G B)^2 + G B partial sums.
//   DRAW   right-looking Cholesky of Jp_SS (ascending coordinates, bias last, np.ix_(mask, mask) of regression.py:350-353)
//          with the 8-column panel all-gathered per step, the forward solve of hp carried along, and a back-substitution
//          whose partial sums stay private per CTA until the pivot block needs them: sample_gaussian(J=, h=) (:334).
// Same decisions and draws as spike_slab.cu on the same randomness (tests: adjacency exact, log-odds / W / b / ml to 1e-8
// of the oracle).
#include "spike_slab_common.cuh"
#include <cooperative_groups.h>
#include <initializer_list>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace pyglm_ss {
namespace {

constexpr int NTHR = 512;
constexpr int NWARP = NTHR / 32;
constexpr int GW = 8;
constexpr int LA_MAXCOLS = 16;

template <int C>
__device__ __forceinline__ void store_all(cg::cluster_group& cl, double* p, double v) {
#pragma unroll
    for (int r = 0; r < C; ++r) *cl.map_shared_rank(p, r) = v;
}

// warp 0: M8 (8 x 8, row-major, identity padded beyond w) <- its inverse (Gauss-Jordan without pivoting: SPD input)
__device__ __forceinline__ bool inv8(double* M8, int lane, int w) {
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
__device__ __forceinline__ bool chol8(double* M8, int lane, int w) {
    bool ok = true;
    const int e0 = lane, e1 = lane + 32;
    const int r0 = e0 / GW, c0 = e0 - r0 * GW, r1 = e1 / GW, c1 = e1 - r1 * GW;
    for (int k = 0; k < w; ++k) {
        const double d = M8[k * GW + k];
        ok = ok && (d > 0.0);
        const double il = rsqrt(d);
        const double ark0 = M8[r0 * GW + k], ack0 = M8[c0 * GW + k], a0 = M8[e0];
        const double ark1 = M8[r1 * GW + k], ack1 = M8[c1 * GW + k], a1 = M8[e1];
        const double n0 = (c0 == k) ? ((r0 == k) ? d * il : ark0 * il) : a0 - ark0 * ack0 * (il * il);
        const double n1 = (c1 == k) ? ((r1 == k) ? d * il : ark1 * il) : a1 - ark1 * ack1 * (il * il);
        __syncwarp();
        if (r0 >= k && c0 >= k && c0 <= r0) M8[e0] = n0;
        if (r1 >= k && c1 >= k && c1 <= r1) M8[e1] = n1;
        __syncwarp();
    }
    return ok;
}

// P[lr][j] += sum_{k < KD} a(lr, k) b(k, j) on my rows (FP64 tensor cores, DMMA.8x8x4): row tiles of 8 local rows, column
// tiles [ct0, ct1) of 8, row tile rt stopping after column tile ctlast(rt) (lower-triangular storage: the rows of a tile
// end at different columns, rowlen() bounds each lane's stores).  Work items are (row tile, chunk of U column tiles),
// dealt round-robin to the warps over the FLATTENED list, so that a triangular update is balanced too; the U tiles of an
// item are independent DMMA chains.  a(lr, k) must return 0 for rows that are not to be touched; columns beyond the
// matrix receive garbage that nobody reads.
template <int KD, class RP, class RN, class AF, class BF, class CL>
__device__ __forceinline__ void tile_update(RP rowptr, RN rowlen, int RL, int nrt, int ct0, int ct1, int warp, int lane,
                                            AF afrag, BF bfrag, CL ctlast) {
    const int g = lane >> 2, q = lane & 3;
    constexpr int U = 4;                                  // column tiles in flight per item
    int base = 0;                                         // items before row tile rt in the flattened list
    for (int rt = 0; rt < nrt; ++rt) {
        const int cend = min(ct1, ctlast(rt) + 1);
        const int nch = (cend > ct0) ? (cend - ct0 + U - 1) / U : 0;
        int ch = (warp - base % NWARP + NWARP) % NWARP;   // my first chunk of this row tile
        base += nch;
        if (ch >= nch) continue;
        const int lr = rt * 8 + g;
        const double a0 = afrag(lr, q);
        const double a1 = (KD == 8) ? afrag(lr, 4 + q) : 0.0;
        const bool rowok = lr < RL;
        double* rowp = rowptr(rowok ? lr : 0) + 2 * q;
        const int rlen = rowlen(rowok ? lr : 0);           // allocated entries of the row (full rows: the pitch)
        for (; ch < nch; ch += NWARP) {
            const int ct = ct0 + ch * U;
            double2 c[U];
            double b0[U], b1[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = (ct + u) * 8;
                const bool in = ct + u < cend;
                ok[u] = in && rowok && (j + 2 * q + 1 < rlen);
                c[u] = ok[u] ? *reinterpret_cast<double2*>(rowp + j) : make_double2(0.0, 0.0);
                b0[u] = in ? bfrag(q, j + g) : 0.0;
                b1[u] = (KD == 8 && in) ? bfrag(4 + q, j + g) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) dmma884(c[u].x, c[u].y, a0, b0[u]);
            if (KD == 8) {
#pragma unroll
                for (int u = 0; u < U; ++u) dmma884(c[u].x, c[u].y, a1, b1[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) *reinterpret_cast<double2*>(rowp + (ct + u) * 8) = c[u];
        }
    }
}

struct DsmGeom {
    int RL, ldp, LAG, XSZ, FPN, PSZ, TG;   // TG: thresholds / prior terms of the scan steps in global scratch, not smem     // local rows, row pitch of Ploc, table slots, doubles per exchange buffer, doubles of Fp
    size_t smem;
};

// Shared-memory plan for (N, B, C); LAG = 0 when even one table slot does not fit.
inline DsmGeom dsm_geom(int N, int B, int C, bool sym) {
    DsmGeom g;
    const int D = N * B + 1, Dp = (D + 1) & ~1;
    g.RL = (D + C - 1) / C;
    g.ldp = ((D + 11) / 16) * 16 + 4;                 // == 4 (mod 16), >= D
    if (g.ldp < D) g.ldp += 16;
    g.LAG = 0;
    g.smem = 0;
    for (int G = (LA_MAXCOLS / B < 8 ? LA_MAXCOLS / B : 8); G >= 1; --G) {
        const int LAC = G * B, LACp = (LAC + 7) & ~7;
        int XSZ = max(B * g.ldp + B * LAC, C * (LAC * LAC + LAC));
        XSZ = max(XSZ, (4 * Dp - Dp * B + 1) / 2);      // the DRAW's four K-vectors alias xbuf + Pd
        int FPN = max(GW, LACp) * g.ldp + 8;
        if (sym) FPN = max(FPN, max(C * Dp, C * g.RL * LAC));   // mu partials / transposed-part exchange of the refill
        // P: full rows of pitch ldp, or (sym) the lower triangle: row lr of CTA c holds lr C + ((c + 2) & ~1) entries
        size_t psz = (size_t)g.RL * g.ldp;
        if (sym) {
            psz = 0;
            for (int c = 0; c < C; ++c) {
                const size_t nr = (D - c + C - 1) / C, l0 = (c + 2) & ~1;        // rows of CTA c, its first row's length
                psz = max(psz, (size_t)(C / 2) * nr * (nr - 1) + nr * l0);
            }
        }
        if (sym && ((D + 7) / 8) * (LACp / 8) > 8 * NWARP) continue;     // register-held partial tiles of the SYM refill
        bool found = false;
        for (int tg = 0; tg < 2 && !found; ++tg) {
            // tg = 1: the per-step thresholds and prior terms (2 N doubles) stay in global scratch (the unused P
            // workspace): one more table slot is worth more than their shared-memory latency
            size_t dbl = psz + FPN + Dp /*mu*/ + 2 * (size_t)XSZ + (size_t)Dp * B /*Pd*/ +
                         2 * (size_t)g.RL * LAC + LAC * LAC + LAC + LAC * B + 3 * B * LAC + B * B + B + (2 * GW * GW + 2 + GW) +
                         2 * C * GW + NWARP + (tg ? 0 : 2 * N) /*thr, cpl*/;
            size_t bytes = dbl * sizeof(double) + ((size_t)Dp + 3 * N + 8 + 8 + 2 * NWARP) * sizeof(int) + (size_t)((N + 7) & ~7);
            if (bytes <= 227 * 1024) {
                g.LAG = G; g.XSZ = XSZ; g.FPN = FPN; g.PSZ = (int)psz; g.TG = tg; g.smem = bytes;
                found = true;
            }
        }
        if (found) break;
    }
    return g;
}

template <int B, int C, bool SYM>
__global__ void __launch_bounds__(NTHR, 1)
spike_slab_dsm_kernel(SpikeSlabArgs A, int RL, int ldp, int LAG, int XSZ, int FPN, int PSZ, int TG) {
    extern __shared__ __align__(16) double sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int crank = (int)cl.block_rank();
    const int ln = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = A.N, D = A.D, NB = N * B;
    const int Dp = (D + 1) & ~1;
    const int LAC = LAG * B, LACp = (LAC + 7) & ~7;

    const double* Jn = A.J + (size_t)ln * A.stride_n;
    const double* hn = A.h + (size_t)ln * A.ldh;
    const double* J0w = A.J0w + (size_t)ln * N * B * B;
    const double* h0w = A.h0w + (size_t)ln * N * B;
    const double J0b = A.J0b[ln], h0b = A.h0b[ln];
    const int ldj = A.ldj;
    auto Jp = [&](int i, int j) -> double {
        const int hi = max(i, j), lo = min(i, j);
        double v = Jn[(size_t)hi * ldj + lo];
        if (hi < NB) {
            const int m = hi / B;
            if (lo / B == m) v += J0w[(size_t)m * B * B + (hi - m * B) * B + (lo - m * B)];
        } else if (lo == hi) {
            v += J0b;
        }
        return v;
    };
    auto hp = [&](int d) -> double { return hn[d] + (d < NB ? h0w[d] : h0b); };

    // ---- shared memory
    double* p = sm;
    double* Ploc = p; p += PSZ;                        // my rows of P (slot space): full rows, or (SYM) their lower triangle
    // SYM: row lr (global row i = lr C + crank) keeps its entries j <= i, padded to an even count; row offsets in closed form
    const int L0 = (crank + 2) & ~1;
    auto PR = [&](int lr) -> double* {
        return Ploc + (SYM ? (size_t)(C / 2) * lr * (lr - 1) + (size_t)lr * L0 : (size_t)lr * ldp);
    };
    auto PLEN = [&](int lr) -> int { return SYM ? lr * C + L0 : ldp; };
    double* Fp = p; p += FPN;                          // BUILD / DRAW: column panel F[b][j]; SCAN refill: C[col][j] (transient)
    double* mu = p; p += Dp;                           // replicated
    double* xbuf = p; p += 2 * (size_t)XSZ;            // two exchange buffers written by every CTA of the cluster
    double* Pd = p; p += (size_t)Dp * B;               // replicated diagonal blocks: Pd[(pos+i)*B + k] = P[pos+i][pos+k]
    double* Cloc = p; p += (size_t)RL * LAC;           // my rows of the table's c_g ...
    double* Tloc = p; p += (size_t)RL * LAC;           // ... and t_g = P c_g
    double* Mt = p; p += LAC * LAC;                    // replicated M[c1][c2] = t_c1^T c_c2
    double* rv = p; p += LAC;                          // replicated r_g = hp_g - c_g^T mu
    double* Jgg = p; p += LAC * B;                     // Jp diagonal block of each candidate
    double* Esh = p; p += B * LAC;
    double* Dsh = p; p += B * LAC;
    double* GEsh = p; p += B * LAC;
    double* Gsh = p; p += B * B;
    double* grsh = p; p += B;
    double* M8 = p; p += 2 * GW * GW + 2;              // two pivot buffers + flag
    double* rd8 = p; p += GW;                          // reciprocals of the pivot's diagonal
    double* spart = p; p += 2 * C * GW;
    double* blo = p; p += NWARP;                       // log-odds of the batch of speculative evaluations
    // thresholds log((1 - u) / u) of the scan steps and cprior + logit rho per neuron: shared memory, or (TG) this
    // neuron's slice of the global P workspace, which the cluster kernel does not otherwise use
    double* us_s = TG ? A.P + (size_t)ln * D * D : p;
    double* cpl_s = us_s + N;
    if (!TG) p += 2 * N;
    int* cidx = reinterpret_cast<int*>(p);
    int* slot = cidx + Dp;
    int* freel = slot + N;
    int* perm_s = freel + N;
    int* cand = perm_s + N;
    int* scal = cand + 8;
    int* bst = scal + 8;                               // per warp: outcome of its speculative evaluation, table slot
    int* bgs = bst + NWARP;
    unsigned char* a_s = reinterpret_cast<unsigned char*>(bgs + NWARP);
    // DRAW vectors alias the scan's exchange buffers and diagonal blocks (dead by then; xbuf and Pd are adjacent)
    double* hv = xbuf;
    double* xs = xbuf + Dp;
    double* sp = xbuf + 2 * Dp;
    double* rdg = xbuf + 3 * Dp;                       // 1 / L_kk, replicated (xbuf + Pd hold >= 3 Dp B + ... doubles)
    double* flag = M8 + 2 * GW * GW;
    double* M8base = M8;

    unsigned char* a_g = A.a + (size_t)ln * N;
    const double* zc = A.z + (size_t)ln * A.ldz;
    for (int m = tid; m < N; m += NTHR) {
        slot[m] = -1;
        a_s[m] = a_g[m];
        perm_s[m] = A.perm[(size_t)ln * N + m];
        {
            // a_m = 1  <=>  u > 1 / (1 + exp(logodds))  <=>  logodds > log((1 - u) / u): the per-step exp becomes a
            // threshold computed once, in parallel (sample_discrete_from_log with the uniform u, regression.py:315)
            const double u = A.us[(size_t)ln * N + m];
            if (!TG || crank == 0) us_s[m] = (u > 0.0) ? log((1.0 - u) / u) : 1e300;
        }
        if (!TG || crank == 0) cpl_s[m] = A.cprior[(size_t)ln * N + m] + A.logit_rho[(size_t)ln * N + m];
    }
    if (tid < 8) cand[tid] = -1;
    __syncthreads();
    cl.sync();                                          // every CTA of the cluster is resident before any DSMEM store

    int fail = 0;
    long long clk0 = 0, clk1 = 0, clk2 = 0;
    long long dbg[20] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // PYGLM_SS_DEBUG: cycles per sub-phase, counts
    const bool dbgon = A.debug != 0;
#define DBG_T0 const long long t0__ = dbgon ? clock64() : 0
#define DBG_ACC(k) do { if (dbgon) dbg[k] += clock64() - t0__; } while (0)
    if (A.debug) clk0 = clock64();
    auto no_limit = [](int) -> int { return 1 << 30; };

    if (A.do_scan[ln]) {
        // =============================================================== BUILD: P = (Jp_SS)^-1, mu = P hp_S
        if (tid == 0) {
            int K = 0;
            cidx[K++] = D - 1;                          // bias first, then the active blocks in ascending order
            for (int m = 0; m < N; ++m)
                if (a_s[m]) {
                    slot[m] = K;
                    for (int b = 0; b < B; ++b) cidx[K++] = m * B + b;
                }
            scal[0] = K;
        }
        __syncthreads();
        int Ks = scal[0];
        int nfree = 0;
        {
            const int K = Ks;
            const int nrt = ((K - crank + C - 1) / C + 7) / 8;     // row tiles that hold rows of mine
            for (int lr = warp; lr * C + crank < K; lr += NWARP) {
                const int i = lr * C + crank, ci = cidx[i];
                double* row = PR(lr);
                const int jend = SYM ? i + 1 : K;
                for (int j = lane; j < jend; j += 32) row[j] = Jp(ci, cidx[j]);
            }
            for (int j = tid; j < K; j += NTHR) hv[j] = hp(cidx[j]);
            __syncthreads();
            auto row_limit = [&](int rt) -> int { return SYM ? ((rt * 8 + 7) * C + crank) / 8 : (1 << 30); };
            for (int k0 = 0; k0 < K; k0 += GW) {
                const int w = min(GW, K - k0), k1 = k0 + w;
                long long tb0 = dbgon ? clock64() : 0;
                // all-gather the column panel F[b][j] = A[j][k0 + b] (= A[k0 + b][j]: A is symmetric)
                if (!SYM) {
                    for (int e = tid; e < RL * GW; e += NTHR) {
                        const int lr = e / GW, b = e - lr * GW, j = lr * C + crank;
                        if (j < K) store_all<C>(cl, &Fp[b * ldp + j], b < w ? PR(lr)[k0 + b] : 0.0);
                    }
                } else {
                    // rows j >= k1 hold A[j][k0 + b] themselves; for j < k0 it sits in pivot row k0 + b, whose owner sends
                    // the row's left part; the pivot block's lower triangle goes to the second pivot buffer
                    for (int e = tid; e < RL * GW; e += NTHR) {
                        const int lr = e / GW, b = e - lr * GW, j = lr * C + crank;
                        if (j < K && j >= k0) store_all<C>(cl, &Fp[b * ldp + j], (j >= k1 && b < w) ? PR(lr)[k0 + b] : 0.0);
                    }
#pragma unroll
                    for (int b = 0; b < GW; ++b) {
                        if (b < w && (k0 + b) % C == crank) {
                            const double* row = PR((k0 + b) / C);
                            for (int j = tid; j < k0; j += NTHR) store_all<C>(cl, &Fp[b * ldp + j], row[j]);
                            if (tid <= b) store_all<C>(cl, &M8[GW * GW + b * GW + tid], row[k0 + tid]);
                        } else if (b >= w) {
                            if (crank == 0) for (int j = tid; j < k0; j += NTHR) store_all<C>(cl, &Fp[b * ldp + j], 0.0);
                        }
                    }
                }
                if (dbgon) { const long long t = clock64(); dbg[0] += t - tb0; tb0 = t; }
                cl.sync();
                if (dbgon) { const long long t = clock64(); dbg[1] += t - tb0; tb0 = t; }
                if (tid < GW * GW) {
                    const int r = tid / GW, c = tid - r * GW;
                    double v = (r == c) ? 1.0 : 0.0;
                    if (r < w && c < w) v = SYM ? M8[GW * GW + max(r, c) * GW + min(r, c)] : Fp[c * ldp + k0 + r];
                    M8[tid] = v;
                }
                __syncthreads();
                if (warp == 0) {
                    const bool ok = inv8(M8, lane, w);
                    if (lane == 0) *flag = ok ? 1.0 : 0.0;
                }
                __syncthreads();
                if (*flag == 0.0) { fail = 1; break; }
                if (dbgon) { const long long t = clock64(); dbg[2] += t - tb0; tb0 = t; }
                // sweep step on my rows: T = A_op D;  A_oo -= T A_po;  A_op = T, A_po = T^T;  A_pp = -D.
                // A_oo -= T F on the tensor cores (everything, pivot rows and columns included: they are rewritten below)
                tile_update<8>(PR, PLEN, RL, nrt, 0, (K + 7) / 8, warp, lane,
                    [&](int lr, int k) -> double {
                        const int i = lr * C + crank;
                        if (i >= K || (i >= k0 && i < k1)) return 0.0;
                        double v = 0.0;
#pragma unroll
                        for (int c = 0; c < GW; ++c) v -= Fp[c * ldp + i] * M8[c * GW + k];
                        return v;
                    },
                    [&](int k, int j) -> double { return Fp[k * ldp + j]; }, row_limit);
                __syncthreads();
                for (int e = tid; e < RL * GW; e += NTHR) {      // pivot columns of my other rows: A_op = T
                    const int lr = e / GW, b = e - lr * GW, i = lr * C + crank;
                    if (i < K && !(i >= k0 && i < k1) && b < w && (!SYM || i >= k1)) {
                        double v = 0.0;
#pragma unroll
                        for (int c = 0; c < GW; ++c) v += Fp[c * ldp + i] * M8[c * GW + b];
                        PR(lr)[k0 + b] = v;
                    }
                }
#pragma unroll
                for (int r = 0; r < GW; ++r) {                    // pivot rows of mine: A_po = T^T, A_pp = -D
                    const int i = k0 + r;
                    if (r < w && i % C == crank) {
                        double* row = PR(i / C);
                        const int jend = SYM ? i + 1 : K;
                        for (int j = tid; j < jend; j += NTHR) {
                            double v;
                            if (j >= k0 && j < k1) {
                                v = -M8[r * GW + (j - k0)];
                            } else {
                                v = 0.0;
#pragma unroll
                                for (int c = 0; c < GW; ++c) v += Fp[c * ldp + j] * M8[c * GW + r];
                            }
                            row[j] = v;
                        }
                    }
                }
                if (dbgon) { const long long t = clock64(); dbg[3] += t - tb0; tb0 = t; }
                cl.sync();                              // every CTA is done with Fp before the next panel arrives
                if (dbgon) { const long long t = clock64(); dbg[1] += t - tb0; tb0 = t; }
            }
            if (!fail) {
                // the sweeps leave -(Jp_SS)^-1: negate, mu = P hp_S, and the replicated diagonal blocks
                if (!SYM) {
                    for (int lr = warp; lr * C + crank < K; lr += NWARP) {
                        const int i = lr * C + crank;
                        double* row = PR(lr);
                        double acc = 0.0;
                        for (int j = lane; j < K; j += 32) {
                            const double v = -row[j];
                            row[j] = v;
                            acc += v * hv[j];
                        }
                        acc = warp_sum(acc);
                        __syncwarp();
                        if (lane == 0) store_all<C>(cl, &mu[i], acc);
                        const int bs = (i == 0) ? 0 : i - ((i - 1) % B);
                        const int nb = (i == 0) ? 1 : B;
                        if (lane < nb) store_all<C>(cl, &Pd[i * B + lane], row[bs + lane]);
                    }
                    cl.sync();
                } else {
                    // lower triangle only: mu_j = sum_{i >= j} P_ij h_i (down my part of column j) + sum_{j' < j} P_jj' h_j'
                    // (along row j if it is mine); the CTAs' partial vectors are summed in rank order by everyone
                    double* pm = xbuf + Dp;                              // K partial sums (xs's place: free until the DRAW)
                    for (int lr = warp; lr * C + crank < K; lr += NWARP) {
                        const int i = lr * C + crank;
                        double* row = PR(lr);
                        for (int j = lane; j <= i; j += 32) row[j] = -row[j];
                    }
                    __syncthreads();
                    for (int j = tid; j < K; j += NTHR) {
                        double acc = 0.0;
                        for (int lr = (j - crank + C - 1 >= 0) ? max(0, (j - crank + C - 1) / C) : 0; lr * C + crank < K; ++lr)
                            acc += PR(lr)[j] * hv[lr * C + crank];
                        pm[j] = acc;
                    }
                    __syncthreads();
                    for (int lr = warp; lr * C + crank < K; lr += NWARP) {
                        const int i = lr * C + crank;
                        const double* row = PR(lr);
                        double acc = 0.0;
                        for (int j = lane; j < i; j += 32) acc += row[j] * hv[j];
                        acc = warp_sum(acc);
                        if (lane == 0) pm[i] += acc;
                        const int bs = (i == 0) ? 0 : i - ((i - 1) % B);
                        if (lane <= i - bs) store_all<C>(cl, &Pd[i * B + lane], row[bs + lane]);
                    }
                    __syncthreads();
                    for (int j = tid; j < K; j += NTHR) store_all<C>(cl, &Fp[crank * Dp + j], pm[j]);
                    cl.sync();
                    for (int j = tid; j < K; j += NTHR) {
                        double acc = 0.0;
#pragma unroll
                        for (int rr = 0; rr < C; ++rr) acc += Fp[rr * Dp + j];
                        mu[j] = acc;
                    }
                    cl.sync();                                          // Fp is read by everyone before it is reused
                }
            }
        }
        if (A.debug) clk1 = clock64();

        // =============================================================== SCAN (regression.py:286-320)
        int par = 0;
        unsigned live = 0u;
        // Evaluations are SPECULATIVELY BATCHED: warp w evaluates scan step `step + w` on the current state, i.e. assuming
        // that no flip is committed before it (true for most steps: ~40 of 200 flip at cfg3).  The decisions are then
        // taken in scan order; the first flip is committed and the evaluations behind it are discarded and redone on
        // the new state.  The serial chain of a step (B x B factorisation, logarithm, comparison) is thus paid once per
        // flip and once per 16 quiet steps instead of once per step, and it is no longer issued by 16 warps at once.
        enum { ST_NOFLIP = 0, ST_FLIP = 1, ST_REFILL = 2, ST_FAIL = 3, ST_END = 4 };
        int step = 0;
        while (step < N && !fail) {
            SmallSolve<B> w;
            double r[B];
            {
                const int s_w = step + warp;
                int st = ST_END, g_w = -1;
                double lo_w = 0.0;
                if (s_w < N) {
                    const int m_w = perm_s[s_w];
                    const int pos_w = slot[m_w];
                    const double cplv = cpl_s[m_w], thr = us_s[s_w];        // issued first: they may live in global memory
                    double S[B][B], sgn = -1.0;
                    bool have = true;
                    if (pos_w >= 0) {
                        // removal: ml(with) - ml(without) read off the replicated diagonal block and mu
#pragma unroll
                        for (int i = 0; i < B; ++i) {
#pragma unroll
                            for (int k = 0; k <= i; ++k) S[i][k] = Pd[(pos_w + i) * B + k];
                            r[i] = mu[pos_w + i];
                        }
                        sgn = 1.0;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) g_w = (i < LAG && ((live >> i) & 1u) && cand[i] == m_w) ? i : g_w;
                        have = g_w >= 0;
#pragma unroll
                        for (int b = 0; b < B; ++b) {
#pragma unroll
                            for (int b2 = 0; b2 <= b; ++b2)
                                S[b][b2] = have ? Jgg[(g_w * B + b) * B + b2] - Mt[(g_w * B + b) * LAC + g_w * B + b2] : 1.0;
                            r[b] = have ? rv[g_w * B + b] : 0.0;
                        }
                    }
                    if (!have) {
                        st = ST_REFILL;
                    } else {
                        // logodds = sgn 1/2 log|S| + 1/2 r^T S^-1 r + prior terms against the threshold log((1-u)/u) of the
                        // step; the logarithm in single precision first, in double precision only when the comparison is
                        // closer than 1e-4 or the log-odds are recorded
                        double det, qf;
                        const bool ok = small_factor_parts<B, B>(w, S, r, det, qf);
                        const double rest = 0.5 * qf + cplv;
                        double lo = sgn * 0.5 * (double)__logf((float)det) + rest;
                        if (A.logodds != nullptr || !(fabs(lo - thr) > 1e-4) || det < 1e-30 || det > 1e30)
                            lo = sgn * 0.5 * log(det) + rest;
                        lo_w = lo;
                        const int v = lo > thr;
                        st = (!ok || !(lo == lo)) ? ST_FAIL : (((pos_w < 0) == (v != 0)) ? ST_FLIP : ST_NOFLIP);
                    }
                }
                if (lane == 0) { bst[warp] = st; bgs[warp] = g_w; blo[warp] = lo_w; }
            }
            __syncthreads();
            int first = NWARP;
#pragma unroll
            for (int i = NWARP - 1; i >= 0; --i) first = (bst[i] != ST_NOFLIP) ? i : first;
            const int stf = (first < NWARP) ? bst[first] : ST_NOFLIP;
            // steps [step, step + first) are evaluated and left as they are; a flip at `first` is decided too
            const int ndec = first + ((stf == ST_FLIP) ? 1 : 0);
#pragma unroll
            for (int i = 0; i < NWARP; ++i)
                if (i < first && bgs[i] >= 0) live &= ~(1u << bgs[i]);
            if (A.logodds && crank == 0 && tid < ndec) A.logodds[(size_t)ln * N + step + tid] = blo[tid];
            if (stf == ST_FAIL) { fail = 1; break; }
            step += first;
            if (stf == ST_NOFLIP || stf == ST_END) { __syncthreads(); continue; }
            if (stf == ST_REFILL) {
            // ---- refill the table with the next LAG inactive neurons of the scan (this one first)
            DBG_T0;
            __syncthreads();
            if (tid == 0) {
                int q = 0;
                for (int i = step; i < N && q < LAG; ++i)
                    if (slot[perm_s[i]] < 0) cand[q++] = perm_s[i];
                scal[1] = q;
                for (; q < 8; ++q) cand[q] = -1;
            }
            __syncthreads();
            live = (1u << scal[1]) - 1u;
            const int Ksp = (Ks + 7) & ~7;
            for (int e0 = tid; e0 < Ksp * LACp; e0 += 4 * NTHR) {   // C[col][i] = Jp[S_i, candidate coordinate], 0 padded
                // four elements per thread with their global loads issued together (one L2 round trip, not four)
                double v[4];
                int dst[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * NTHR;
                    v[u] = 0.0;
                    dst[u] = -1;
                    if (e < Ksp * LACp) {
                        const int col = e / Ksp, i = e - col * Ksp;
                        dst[u] = col * ldp + i;
                        if (col < LAC && i < Ks) {
                            const int gq = col / B, mq = cand[gq], ci = cidx[i];
                            if (mq >= 0 && ci >= 0) {
                                const int cj = mq * B + (col - gq * B);      // candidate coordinates are inactive: ci != cj
                                const int hi = max(ci, cj), lo = min(ci, cj);
                                v[u] = Jn[(size_t)hi * ldj + lo];            // off-block entries carry no prior term
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (dst[u] >= 0) Fp[dst[u]] = v[u];
            }
            if (tid < LAC * B) {
                const int col = tid / B, b2 = tid - col * B, gq = col / B, mq = cand[gq];
                Jgg[tid] = (mq >= 0) ? Jp(mq * B + (col - gq * B), mq * B + b2) : 1.0;
            }
            __syncthreads();
            if (dbgon) dbg[11] += clock64() - t0__;
            {
                // my rows of T = P C on the tensor cores: one warp per (row tile, 8 table columns), two independent
                // accumulator pairs over the even / odd k-steps
                const int nrt = ((Ks - crank + C - 1) / C + 7) / 8, nctl = LACp / 8;
                const int gq = lane >> 2, q = lane & 3;
                for (int task = warp; task < nrt * nctl; task += NWARP) {
                    const int rt = task / nctl, ctl = task - rt * nctl;
                    const int lr = rt * 8 + gq, i = lr * C + crank;
                    const bool rowok = i < Ks;
                    const double* prow = PR(rowok ? lr : 0);
                    const double* crow = Fp + (size_t)(ctl * 8 + gq) * ldp;
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                    const int jlim = SYM ? min(Ks, i + 1) : Ks;          // SYM: the stored part of the row (j <= i)
                    // the trip count must be the same for all lanes (mma.sync): run to the longest row of the tile
                    const int jmax = SYM ? min(Ks, (rt * 8 + 7) * C + crank + 1) : Ks;
                    for (int j0 = 0; j0 < jmax; j0 += 8) {
                        const double a0 = (rowok && j0 + q < jlim) ? prow[j0 + q] : 0.0;
                        const double a1 = (rowok && j0 + 4 + q < jlim) ? prow[j0 + 4 + q] : 0.0;
                        dmma884(c0, c1, a0, crow[j0 + q]);
                        dmma884(d0, d1, a1, crow[j0 + 4 + q]);
                    }
                    c0 += d0; c1 += d1;
                    const int col = ctl * 8 + 2 * q;
                    if (rowok) {
                        if (col < LAC) Tloc[lr * LAC + col] = c0;
                        if (col + 1 < LAC) Tloc[lr * LAC + col + 1] = c1;
                    }
                }
                for (int e = tid; e < RL * LAC; e += NTHR) {
                    const int lr = e / LAC, c2 = e - lr * LAC, i = lr * C + crank;
                    if (i < Ks) Cloc[e] = Fp[c2 * ldp + i];
                }
                if (SYM) {
                    // the other half of the symmetric product: T[i] += sum over MY rows j > i of P[j][i] C[j], for ALL i.
                    // One warp per (8 output rows, 8 table columns); k runs over my local rows.  The partial rows go to
                    // their owners (row i -> CTA i mod C) once every CTA has finished reading its copy of C.
                    constexpr int MAXT = 8;
                    const int nit = (Ks + 7) / 8, ntask = nit * nctl;
                    double r0[MAXT], r1[MAXT];
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        r0[t] = r1[t] = 0.0;
                        const int task = warp + t * NWARP;
                        if (task < ntask) {
                            const int it = task / nctl, ctl = task - it * nctl;
                            const int i = it * 8 + gq;
                            const double* crow = Fp + (size_t)(ctl * 8 + gq) * ldp;
                            double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                            int lr0 = (it * 8 - crank) / C;                 // first local row that can exceed the tile's rows
                            if (lr0 < 0) lr0 = 0;
                            for (; lr0 * C + crank < Ks; lr0 += 8) {
                                const int lA = lr0 + q, jA = lA * C + crank, lB = lr0 + 4 + q, jB = lB * C + crank;
                                const double a0 = (jA < Ks && jA > i && i < Ks) ? PR(lA)[i] : 0.0;
                                const double a1 = (jB < Ks && jB > i && i < Ks) ? PR(lB)[i] : 0.0;
                                const double b0 = (jA < Ks) ? crow[jA] : 0.0;
                                const double b1 = (jB < Ks) ? crow[jB] : 0.0;
                                dmma884(c0, c1, a0, b0);
                                dmma884(d0, d1, a1, b1);
                            }
                            r0[t] = c0 + d0;
                            r1[t] = c1 + d1;
                        }
                    }
                    cl.sync();                                              // every CTA is done with C (Fp)
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        const int task = warp + t * NWARP;
                        if (task < ntask) {
                            const int it = task / nctl, ctl = task - it * nctl;
                            const int i = it * 8 + gq, col = ctl * 8 + 2 * q;
                            if (i < Ks) {
                                double* dst = cl.map_shared_rank(Fp + ((size_t)crank * RL + i / C) * LAC, (unsigned)(i % C));
                                if (col < LAC) dst[col] = r0[t];
                                if (col + 1 < LAC) dst[col + 1] = r1[t];
                            }
                        }
                    }
                    cl.sync();
                    for (int e = tid; e < RL * LAC; e += NTHR) {
                        const int lr = e / LAC, i = lr * C + crank;
                        if (i < Ks) {
                            double acc = Tloc[e];
#pragma unroll
                            for (int rr = 0; rr < C; ++rr) acc += Fp[(size_t)rr * RL * LAC + e];
                            Tloc[e] = acc;
                        }
                    }
                }
            }
            __syncthreads();
            if (dbgon) dbg[12] += clock64() - t0__;
            // partial sums over my rows of M = T^T C and C^T mu -> every CTA; then summed in rank order
            double* xb = xbuf + (size_t)par * XSZ;
            if (tid < LAC * LAC + LAC) {
                double s0 = 0.0, s1 = 0.0;
                if (tid < LAC * LAC) {
                    const int c1 = tid / LAC, c2 = tid - c1 * LAC;
                    int lr = 0;
                    for (; (lr + 1) * C + crank < Ks; lr += 2) {
                        s0 += Tloc[lr * LAC + c1] * Cloc[lr * LAC + c2];
                        s1 += Tloc[(lr + 1) * LAC + c1] * Cloc[(lr + 1) * LAC + c2];
                    }
                    if (lr * C + crank < Ks) s0 += Tloc[lr * LAC + c1] * Cloc[lr * LAC + c2];
                } else {
                    const int c2 = tid - LAC * LAC;
                    for (int lr = 0; lr * C + crank < Ks; ++lr) s0 += Cloc[lr * LAC + c2] * mu[lr * C + crank];
                }
                store_all<C>(cl, &xb[crank * (LAC * LAC + LAC) + tid], s0 + s1);
            }
            if (dbgon) dbg[13] += clock64() - t0__;
            cl.sync();
            if (dbgon) dbg[14] += clock64() - t0__;
            if (tid < LAC * LAC + LAC) {
                double s = 0.0;
#pragma unroll
                for (int rr = 0; rr < C; ++rr) s += xb[rr * (LAC * LAC + LAC) + tid];
                if (tid < LAC * LAC) {
                    Mt[tid] = s;
                } else {
                    const int c2 = tid - LAC * LAC, gq = c2 / B, mq = cand[gq];
                    rv[c2] = (mq >= 0) ? hp(mq * B + (c2 - gq * B)) - s : 0.0;
                }
            }
            __syncthreads();
            par ^= 1;
            DBG_ACC(4);
                if (dbgon) ++dbg[8];
                __syncthreads();
                continue;
            }
            // ---- a flip of scan step `step`, evaluated by warp `first` (whose registers hold the factorisation)
            const int m = perm_s[step];
            const int pos = slot[m];
            const int g = bgs[first];
            const bool fw = (warp == first);
            const int v = (pos < 0) ? 1 : 0;
            ++step;

            if (pos < 0 && v) {
                // ---- commit the addition of slot g (block m)
                DBG_T0;
                small_finish<B, B>(w, r, nullptr, 0);
                const bool grow = (nfree == 0);
                const int pn = grow ? Ks : freel[nfree - 1];
                const int Kn = grow ? Ks + B : Ks;
                double* tb = xbuf + (size_t)par * XSZ;               // t[b][i], pitch ldp
                for (int e = tid; e < RL * B; e += NTHR) {           // all-gather t = T[:, slot g]
                    const int lr = e / B, b = e - lr * B, i = lr * C + crank;
                    if (i < Ks) store_all<C>(cl, &tb[b * ldp + i], Tloc[lr * LAC + g * B + b]);
                }
                if (tid < B * LAC) {                                  // E = t^T c_j - D,  D = Jp[new block, j-block]
                    const int b = tid / LAC, c2 = tid - b * LAC, g2 = c2 / B;
                    const bool lv = ((live >> g2) & 1u) && g2 != g;
                    const double d = lv ? Jp(m * B + b, cand[g2] * B + (c2 - g2 * B)) : 0.0;
                    Dsh[tid] = d;
                    Esh[tid] = lv ? Mt[(g * B + b) * LAC + c2] - d : 0.0;
                }
                if (fw && lane == 0) {
#pragma unroll
                    for (int i = 0; i < B; ++i) {
                        grsh[i] = w.gr[i];
#pragma unroll
                        for (int k = 0; k < B; ++k) Gsh[i * B + k] = w.G[i][k];
                    }
                }
                if (dbgon) dbg[15] += clock64() - t0__;
                cl.sync();
                if (dbgon) dbg[16] += clock64() - t0__;
                if (grow && tid < B * B) tb[(tid / B) * ldp + Ks + (tid % B)] = 0.0;   // rows that do not exist yet
                if (tid < B * LAC) {
                    const int b = tid / LAC, c2 = tid - b * LAC;
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < B; ++k) s += Gsh[b * B + k] * Esh[k * LAC + c2];
                    GEsh[tid] = s;
                }
                __syncthreads();
                {
                    // P += (t G) t^T on my old rows (rank B on the tensor cores), then the new border rows / columns
                    const int nrt = ((Kn - crank + C - 1) / C + 7) / 8;
                    tile_update<4>(PR, PLEN, RL, nrt, 0, (Kn + 7) / 8, warp, lane,
                        [&](int lr, int k) -> double {
                            const int i = lr * C + crank;
                            if (k >= B || i >= Ks || (i >= pn && i < pn + B)) return 0.0;
                            double v = 0.0;
#pragma unroll
                            for (int j2 = 0; j2 < B; ++j2) v += tb[j2 * ldp + i] * Gsh[j2 * B + k];
                            return v;
                        },
                        [&](int k, int j) -> double { return (k < B) ? tb[k * ldp + j] : 0.0; },
                        [&](int rt) -> int { return SYM ? ((rt * 8 + 7) * C + crank) / 8 : (1 << 30); });
                    __syncthreads();
                    for (int e = tid; e < RL * B; e += NTHR) {       // border columns of my old rows: -t G
                        const int lr = e / B, b = e - lr * B, i = lr * C + crank;
                        if (i < Ks && !(i >= pn && i < pn + B) && (!SYM || i >= pn + B)) {
                            double v = 0.0;
#pragma unroll
                            for (int j2 = 0; j2 < B; ++j2) v -= tb[j2 * ldp + i] * Gsh[j2 * B + b];
                            PR(lr)[pn + b] = v;
                        }
                    }
#pragma unroll
                    for (int bi = 0; bi < B; ++bi) {                  // the new rows
                        const int i = pn + bi;
                        if (i % C == crank) {
                            double* row = PR(i / C);
                            const int jend = SYM ? i + 1 : Kn;
                            for (int j = tid; j < jend; j += NTHR) {
                                double x;
                                if (j >= pn && j < pn + B) {
                                    x = Gsh[bi * B + (j - pn)];
                                } else {
                                    x = 0.0;
#pragma unroll
                                    for (int k = 0; k < B; ++k) x -= tb[k * ldp + j] * Gsh[k * B + bi];
                                }
                                row[j] = x;
                            }
                        }
                    }
                }
                if (dbgon) dbg[17] += clock64() - t0__;
                for (int e = tid; e < RL * LAC; e += NTHR) {                        // my rows of the table
                    const int lr = e / LAC, c2 = e - lr * LAC, i = lr * C + crank, g2 = c2 / B;
                    if (i >= Kn) continue;
                    const bool lv = ((live >> g2) & 1u) && g2 != g;
                    if (i >= pn && i < pn + B) {
                        const int bi = i - pn;
                        Tloc[e] = lv ? -GEsh[bi * LAC + c2] : 0.0;
                        Cloc[e] = lv ? Dsh[bi * LAC + c2] : 0.0;
                    } else if (lv) {
                        double x = Tloc[e];
#pragma unroll
                        for (int b = 0; b < B; ++b) x += tb[b * ldp + i] * GEsh[b * LAC + c2];
                        Tloc[e] = x;
                    }
                }
                for (int i = tid; i < Kn; i += NTHR) {                              // replicated mu and diagonal blocks
                    if (i >= pn && i < pn + B) {
                        mu[i] = grsh[i - pn];
#pragma unroll
                        for (int k = 0; k < B; ++k) Pd[i * B + k] = Gsh[(i - pn) * B + k];
                    } else {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < B; ++k) s += tb[k * ldp + i] * grsh[k];
                        mu[i] -= s;
                        const int bs = (i == 0) ? 0 : i - ((i - 1) % B);
                        const int nb = (i == 0) ? 1 : B;
                        double gi[B];
#pragma unroll
                        for (int k2 = 0; k2 < B; ++k2) {
                            double gik = 0.0;
#pragma unroll
                            for (int j2 = 0; j2 < B; ++j2) gik += tb[j2 * ldp + i] * Gsh[j2 * B + k2];
                            gi[k2] = gik;
                        }
#pragma unroll
                        for (int k = 0; k < B; ++k) {
                            if (k < nb) {
                                double x = 0.0;
#pragma unroll
                                for (int k2 = 0; k2 < B; ++k2) x += gi[k2] * tb[k2 * ldp + bs + k];
                                Pd[i * B + k] += x;
                            }
                        }
                    }
                }
                if (tid < LAC * LAC) {                                               // M += E^T G E,  r += E^T G r_m
                    const int c1 = tid / LAC, c2 = tid - c1 * LAC;
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < B; ++b) s += Esh[b * LAC + c1] * GEsh[b * LAC + c2];
                    Mt[tid] += s;
                } else if (tid < LAC * LAC + LAC) {
                    const int c2 = tid - LAC * LAC;
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < B; ++b) s += Esh[b * LAC + c2] * grsh[b];
                    rv[c2] += s;
                }
                if (tid == NTHR - 1) {
#pragma unroll
                    for (int b = 0; b < B; ++b) cidx[pn + b] = m * B + b;
                    slot[m] = pn;
                    a_s[m] = 1;
                }
                if (grow) Ks += B; else --nfree;
                live &= ~(1u << g);
                par ^= 1;
                __syncthreads();
                DBG_ACC(5);
                if (dbgon) ++dbg[9];
            } else if (pos >= 0 && !v) {
                // ---- commit the removal of the block at pos
                DBG_T0;
                small_finish<B, B>(w, r, nullptr, 0);                  // G = P_mm^-1, gr = P_mm^-1 mu_m
                double* tb = xbuf + (size_t)par * XSZ;                 // P[pos + b][:], pitch ldp
                double* vs = tb + (size_t)B * ldp;
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    if ((pos + b) % C == crank) {                      // the owner of row pos + b hands it to everyone
                        const int lr = (pos + b) / C;
                        const int jend = SYM ? pos + b + 1 : Ks;
                        for (int j = tid; j < jend; j += NTHR) store_all<C>(cl, &tb[b * ldp + j], PR(lr)[j]);
                        if (tid < LAC) store_all<C>(cl, &vs[b * LAC + tid], Tloc[lr * LAC + tid]);
                    }
                }
                if (SYM) {
                    // lower triangle: P[j][pos + b] for j > pos + b sits in row j -- every CTA sends its rows' entries
                    for (int e = tid; e < RL * B; e += NTHR) {
                        const int lr = e / B, b = e - lr * B, j = lr * C + crank;
                        if (j < Ks && j > pos + b) store_all<C>(cl, &tb[b * ldp + j], PR(lr)[pos + b]);
                    }
                }
                if (fw && lane == 0) {
#pragma unroll
                    for (int i = 0; i < B; ++i) {
                        grsh[i] = w.gr[i];
#pragma unroll
                        for (int k = 0; k < B; ++k) Gsh[i * B + k] = w.G[i][k];
                    }
                }
                cl.sync();
                if (tid < B * LAC) {                                    // G v,  v = T[pos block, :]
                    const int b = tid / LAC, c2 = tid - b * LAC;
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < B; ++k) s += Gsh[b * B + k] * vs[k * LAC + c2];
                    GEsh[tid] = s;
                }
                __syncthreads();
                {
                    const int nrt = ((Ks - crank + C - 1) / C + 7) / 8;
                    tile_update<4>(PR, PLEN, RL, nrt, 0, (Ks + 7) / 8, warp, lane,
                        [&](int lr, int k) -> double {
                            const int i = lr * C + crank;
                            if (k >= B || i >= Ks || (i >= pos && i < pos + B)) return 0.0;
                            double v = 0.0;
#pragma unroll
                            for (int j2 = 0; j2 < B; ++j2) v -= tb[j2 * ldp + i] * Gsh[j2 * B + k];
                            return v;
                        },
                        [&](int k, int j) -> double { return (k < B) ? tb[k * ldp + j] : 0.0; },
                        [&](int rt) -> int { return SYM ? ((rt * 8 + 7) * C + crank) / 8 : (1 << 30); });
                    __syncthreads();
                    for (int e = tid; e < RL * B; e += NTHR) {       // tombstone: zero columns ...
                        const int lr = e / B, b = e - lr * B, i = lr * C + crank;
                        if (i < Ks && (!SYM || i > pos + b)) PR(lr)[pos + b] = 0.0;
                    }
#pragma unroll
                    for (int b = 0; b < B; ++b) {                     // ... and rows
                        if ((pos + b) % C == crank) {
                            double* row = PR((pos + b) / C);
                            const int jend = SYM ? pos + b + 1 : Ks;
                            for (int j = tid; j < jend; j += NTHR) row[j] = 0.0;
                        }
                    }
                }
                for (int e = tid; e < RL * LAC; e += NTHR) {
                    const int lr = e / LAC, c2 = e - lr * LAC, i = lr * C + crank, g2 = c2 / B;
                    if (i >= Ks) continue;
                    if (i >= pos && i < pos + B) {
                        Tloc[e] = 0.0;
                        Cloc[e] = 0.0;
                    } else if ((live >> g2) & 1u) {
                        double x = Tloc[e];
#pragma unroll
                        for (int b = 0; b < B; ++b) x -= tb[b * ldp + i] * GEsh[b * LAC + c2];
                        Tloc[e] = x;
                    }
                }
                for (int i = tid; i < Ks; i += NTHR) {
                    if (i >= pos && i < pos + B) {
                        mu[i] = 0.0;
#pragma unroll
                        for (int k = 0; k < B; ++k) Pd[i * B + k] = 0.0;
                    } else {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < B; ++k) s += tb[k * ldp + i] * grsh[k];
                        mu[i] -= s;
                        const int bs = (i == 0) ? 0 : i - ((i - 1) % B);
                        const int nb = (i == 0) ? 1 : B;
                        double gi[B];
#pragma unroll
                        for (int k2 = 0; k2 < B; ++k2) {
                            double gik = 0.0;
#pragma unroll
                            for (int j2 = 0; j2 < B; ++j2) gik += tb[j2 * ldp + i] * Gsh[j2 * B + k2];
                            gi[k2] = gik;
                        }
#pragma unroll
                        for (int k = 0; k < B; ++k) {
                            if (k < nb) {
                                double x = 0.0;
#pragma unroll
                                for (int k2 = 0; k2 < B; ++k2) x += gi[k2] * tb[k2 * ldp + bs + k];
                                Pd[i * B + k] -= x;
                            }
                        }
                    }
                }
                if (tid < LAC * LAC) {                                               // M -= v^T G v,  r += v^T G mu_m
                    const int c1 = tid / LAC, c2 = tid - c1 * LAC;
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < B; ++b) s += vs[b * LAC + c1] * GEsh[b * LAC + c2];
                    Mt[tid] -= s;
                } else if (tid < LAC * LAC + LAC) {
                    const int c2 = tid - LAC * LAC;
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < B; ++b) s += vs[b * LAC + c2] * grsh[b];
                    rv[c2] += s;
                }
                if (tid == NTHR - 1) {
#pragma unroll
                    for (int b = 0; b < B; ++b) cidx[pos + b] = -1;
                    freel[nfree] = pos;
                    slot[m] = -1;
                    a_s[m] = 0;
                }
                ++nfree;
                par ^= 1;
                __syncthreads();
                DBG_ACC(6);
                if (dbgon) ++dbg[10];
            }
        }
        if (A.debug) clk2 = clock64();
    }
    cl.sync();

    // =================================================================== DRAW: [W_S; b] = L^-T (L^-1 hp + z), Jp_SS = L L^T
    double ml = 0.0;
    int Kd = 0;
    if (!fail) {
        if (tid == 0) {
            int K = 0;
            for (int m = 0; m < N; ++m)
                if (a_s[m])
                    for (int b = 0; b < B; ++b) cidx[K++] = m * B + b;
            cidx[K++] = D - 1;
            scal[0] = K;
        }
        __syncthreads();
        const int K = Kd = scal[0];
        const int nrt = ((K - crank + C - 1) / C + 7) / 8;
        for (int lr = warp; lr * C + crank < K; lr += NWARP) {
            const int i = lr * C + crank, ci = cidx[i];
            double* row = PR(lr);
            for (int j = lane; j <= i; j += 32) row[j] = Jp(ci, cidx[j]);
        }
        for (int j = tid; j < K; j += NTHR) hv[j] = hp(cidx[j]);
        __syncthreads();
        double hld = 0.0;
        const int pr = tid / GW, pc = tid - pr * GW;                // pivot entry of threads < 64
        int mb = 0;
        {   // the first pivot block goes to every CTA; later ones travel with the barrier that ends the previous step
            const int w0 = min(GW, K);
            if (tid < GW * GW && pr < w0 && pc <= pr && pr % C == crank) store_all<C>(cl, &M8[tid], PR(pr / C)[pc]);
            cl.sync();
        }
        for (int k0 = 0; k0 < K; k0 += GW) {
            const int w = min(GW, K - k0), k1 = k0 + w;
            double* M8 = M8base + mb * GW * GW;
            const bool mine = tid < GW * GW && pr < w && pc <= pr && (k0 + pr) % C == crank;
            if (tid < GW * GW) {
                if (pr >= w || pc >= w) M8[tid] = (pr == pc) ? 1.0 : 0.0;
                else if (pc > pr) M8[tid] = 0.0;
            }
            __syncthreads();
            if (warp == 0) {
                const bool ok = chol8(M8, lane, w);
                if (lane == 0) *flag = ok ? 1.0 : 0.0;
                __syncwarp();
                if (lane < GW) rd8[lane] = 1.0 / M8[lane * GW + lane];
            }
            __syncthreads();
            if (*flag == 0.0) { fail = 1; break; }
            double y[GW];
            double dprod = 1.0;
#pragma unroll
            for (int b = 0; b < GW; ++b) {
                double x = (b < w) ? hv[k0 + b] : 0.0;
#pragma unroll
                for (int c = 0; c < b; ++c) x -= M8[b * GW + c] * y[c];
                y[b] = x * rd8[b];
                dprod *= M8[b * GW + b];                            // identity padding beyond w
            }
            hld += log(dprod);
            if (tid < w) rdg[k0 + tid] = rd8[tid];
            if (mine) PR((k0 + pr) / C)[k0 + pc] = M8[tid];
            // panel L[j][k0 + b] = A[j][k0..] L_pp^-T for my rows j >= k1, handed to every CTA
            for (int lr = tid; lr < RL; lr += NTHR) {
                const int j = lr * C + crank;
                if (j >= k1 && j < K) {
                    double* row = PR(lr);
                    double x[GW];
#pragma unroll
                    for (int b = 0; b < GW; ++b) x[b] = (b < w) ? row[k0 + b] : 0.0;
#pragma unroll
                    for (int b = 0; b < GW; ++b) {
                        double s = x[b];
#pragma unroll
                        for (int c = 0; c < b; ++c) s -= x[c] * M8[b * GW + c];
                        x[b] = s * rd8[b];
                        if (b < w) row[k0 + b] = x[b];
                        store_all<C>(cl, &Fp[b * ldp + j], x[b]);
                    }
                }
            }
            cl.sync();
            for (int j = k1 + tid; j < K; j += NTHR) {                  // forward solve of hp, replicated
                double hj = hv[j];
#pragma unroll
                for (int b = 0; b < GW; ++b) hj -= Fp[b * ldp + j] * y[b];
                hv[j] = hj;
            }
            if (tid < w) {
                double yv = 0.0;
#pragma unroll
                for (int b = 0; b < GW; ++b) yv = (b == tid) ? y[b] : yv;
                hv[k0 + tid] = yv;
            }
            // trailing update of my rows >= k1, lower triangle (tile granularity), on the tensor cores
            tile_update<8>(PR, PLEN, RL, nrt, k1 / 8, (K + 7) / 8, warp, lane,
                [&](int lr, int k) -> double {
                    const int i = lr * C + crank;
                    return (i >= k1 && i < K) ? -Fp[k * ldp + i] : 0.0;
                },
                [&](int k, int j) -> double { return Fp[k * ldp + j]; },
                [&](int rt) -> int { return ((rt * 8 + 7) * C + crank) / 8; });
            __syncthreads();
            if (k1 < K) {                                               // next pivot block -> every CTA's other pivot buffer
                const int w1 = min(GW, K - k1);
                if (tid < GW * GW && pr < w1 && pc <= pr && (k1 + pr) % C == crank)
                    store_all<C>(cl, &M8base[(mb ^ 1) * GW * GW + tid], PR((k1 + pr) / C)[k1 + pc]);
            }
            cl.sync();                                                  // Fp is free again, the next pivot has arrived
            mb ^= 1;
        }
        if (!fail) {
            // |L^-1 hp|^2 in a fixed order (every CTA gets the same bits), then xs = L^-1 hp + z
            __syncthreads();
            double q = 0.0;
            for (int j = lane; j < K; j += 32) q += hv[j] * hv[j];
            q = warp_sum(q);
            for (int j = tid; j < K; j += NTHR) { xs[j] = hv[j] + zc[cidx[j]]; sp[j] = 0.0; }
            __syncthreads();
            int bp = 0;
            if (dbgon) dbg[7] = clock64();
            for (int k0 = ((K - 1) / GW) * GW; k0 >= 0; k0 -= GW) {
                const int w = min(GW, K - k0);
                double* m8 = M8 + bp * GW * GW;
                double* spb = spart + bp * C * GW;
                const int pr = tid / GW, pc = tid - pr * GW;
                if (tid < GW * GW && pr < w && pc <= pr && (k0 + pr) % C == crank)
                    store_all<C>(cl, &m8[tid], PR((k0 + pr) / C)[k0 + pc]);
                if (tid >= GW * GW && tid < GW * GW + GW) {
                    const int b = tid - GW * GW;
                    store_all<C>(cl, &spb[crank * GW + b], b < w ? sp[k0 + b] : 0.0);
                }
                cl.sync();
                double o[GW];
#pragma unroll
                for (int b = GW - 1; b >= 0; --b) {
                    double t = 0.0;
                    if (b < w) {
                        t = xs[k0 + b];
#pragma unroll
                        for (int rr = 0; rr < C; ++rr) t -= spb[rr * GW + b];
#pragma unroll
                        for (int c = b + 1; c < GW; ++c)
                            if (c < w) t -= m8[c * GW + b] * o[c];
                        t *= rdg[k0 + b];
                    }
                    o[b] = t;
                }
#pragma unroll
                for (int b = 0; b < GW; ++b) {
                    if (b < w && (k0 + b) % C == crank) {
                        const double* row = PR((k0 + b) / C);
                        for (int j = tid; j < k0; j += NTHR) sp[j] += row[j] * o[b];
                    }
                }
                __syncthreads();
                if (tid < w) {
                    double ov = 0.0;
#pragma unroll
                    for (int b = 0; b < GW; ++b) ov = (b == tid) ? o[b] : ov;
                    xs[k0 + tid] = ov;
                }
                __syncthreads();
                bp ^= 1;
            }
            if (dbgon) dbg[7] = clock64() - dbg[7];
            ml = -hld + 0.5 * q + 0.5 * log(J0b) - 0.5 * h0b * h0b / J0b;
        }
    }
    __syncthreads();
    if (crank == 0) {
        double* Wn = A.W + (size_t)ln * N * B;
        for (int e = tid; e < NB; e += NTHR) Wn[e] = 0.0;
        __syncthreads();
        if (!fail) {
            for (int k = tid; k < Kd; k += NTHR) {
                const int d = cidx[k];
                if (d < NB) Wn[d] = xs[k]; else A.bias[ln] = xs[k];
            }
            for (int m = tid; m < N; m += NTHR) a_g[m] = a_s[m];
        }
        if (tid == 0) {
            if (A.ml) {
                if (!fail) {
                    const double* cprior = A.cprior + (size_t)ln * N;
                    for (int m = 0; m < N; ++m)
                        if (a_s[m]) ml += cprior[m];
                }
                A.ml[ln] = fail ? nan("") : ml;
            }
            A.status[ln] = fail;
            if (A.debug && ln == 0)
                printf("spike_slab_dsm cluster0 (C=%d, LAG=%d): build %lld cyc  scan %lld cyc  draw %lld cyc  (K draw = %d)\n"
                       "   build: panel stores %lld  cluster syncs %lld  pivot %lld  row update %lld\n"
                       "   scan: %lld refills %lld cyc  %lld adds %lld cyc  %lld removals %lld cyc;  draw: backsolve %lld cyc\n",
                       C, LAG, clk1 - clk0, clk2 - clk1, clock64() - clk2, Kd, dbg[0], dbg[1], dbg[2], dbg[3],
                       dbg[8], dbg[4], dbg[9], dbg[5], dbg[10], dbg[6], dbg[7]);
            if (A.debug && ln == 0)
                printf("   refill cumulative: gather %lld  T=PC %lld  partials %lld  sync %lld;  add cumulative: stores+E %lld  sync %lld  P rows %lld\n",
                       dbg[11], dbg[12], dbg[13], dbg[14], dbg[15], dbg[16], dbg[17]);
        }
    }
    cl.sync();                                          // no CTA exits while a peer may still address its shared memory
}

template <int B, int C, bool SYM>
int launch_dsm_bc(const SpikeSlabArgs& A, cudaStream_t stream) {
    const DsmGeom g = dsm_geom(A.N, B, C, SYM);
    if (g.LAG < 1) return PYGLM_ERR_UNSUPPORTED;
    PYGLM_CUDA(cudaFuncSetAttribute(spike_slab_dsm_kernel<B, C, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(A.n_loc * C, 1, 1);
    cfg.blockDim = dim3(NTHR, 1, 1);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PYGLM_CUDA(cudaLaunchKernelEx(&cfg, spike_slab_dsm_kernel<B, C, SYM>, A, g.RL, g.ldp, g.LAG, g.XSZ, g.FPN, g.PSZ, g.TG));
    return PYGLM_OK;
}

// The cluster shapes that are built: 2 and 4 CTAs with the lower triangle of P (half the shared memory per neuron, so a
// neuron fits fewer CTAs and more clusters are resident at once), 8 CTAs with full rows (the first version; what is left
// when D is too large for the others).  pick_cluster: the smallest one whose shared memory holds the state.
inline bool dsm_sym(int c) { return c != 8; }
inline int pick_cluster(int N, int B) {
    for (int c : {2, 4, 8})
        if (dsm_geom(N, B, c, dsm_sym(c)).LAG >= 1) return c;
    return 0;
}

template <int B>
int launch_dsm_b(const SpikeSlabArgs& A, int csize, cudaStream_t stream) {
    if (csize == 0) csize = pick_cluster(A.N, B);
    switch (csize) {
        case 2: return launch_dsm_bc<B, 2, true>(A, stream);
        case 4: return launch_dsm_bc<B, 4, true>(A, stream);
        case 8: return launch_dsm_bc<B, 8, false>(A, stream);
        default: return PYGLM_ERR_UNSUPPORTED;
    }
}

}  // namespace

int choose_cluster(const SpikeSlabArgs& A);

int spike_slab_dsm_launch(const SpikeSlabArgs& A, int csize, cudaStream_t stream) {
    if (csize == 0) csize = choose_cluster(A);          // 0 again: launch_dsm_b falls back to the smallest shape that fits
    switch (A.B) {
        case 1: return launch_dsm_b<1>(A, csize, stream);
        case 2: return launch_dsm_b<2>(A, csize, stream);
        case 3: return launch_dsm_b<3>(A, csize, stream);
        case 4: return launch_dsm_b<4>(A, csize, stream);
        default: return PYGLM_ERR_UNSUPPORTED;
    }
}

namespace {
// Clusters of c CTAs with the kernel's shared-memory footprint that the device holds at once (GPC granularity: 13-14
// clusters of 8 on a B200, not 148 / 8); 0 when the query fails.
template <int B, int C, bool SYM>
int dsm_active_clusters(const DsmGeom& g) {
    if (cudaFuncSetAttribute(spike_slab_dsm_kernel<B, C, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * 64, 1, 1);
    cfg.blockDim = dim3(NTHR, 1, 1);
    cfg.dynamicSmemBytes = g.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, spike_slab_dsm_kernel<B, C, SYM>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

template <int B>
int dsm_active_clusters_b(int N, int c) {
    const DsmGeom g = dsm_geom(N, B, c, dsm_sym(c));
    switch (c) {
        case 2: return dsm_active_clusters<B, 2, true>(g);
        case 4: return dsm_active_clusters<B, 4, true>(g);
        default: return dsm_active_clusters<B, 8, false>(g);
    }
}
}  // namespace

// Which cluster shape, if any, should run n_loc neurons: 0 = none (one CTA per neuron wins).  Estimated time = waves of
// resident clusters x the per-cluster time measured at cfg3's shape (D = 401) on a state captured at the chain's
// EQUILIBRIUM density (0.5, K ~ 280; profiles/r02ag_probe_scan.log -- an earlier table measured on the sparse first
// sweeps underrated the clusters): 1.35 ms for 4 CTAs with the lower triangle (~30 resident; 25 / 50 neurons 1.35 /
// 2.57 ms), ~1.05 ms for 8 CTAs with full rows (13-14 resident), ~2 ms for 2 CTAs; against 2.9-3.1 ms for <= 100
// single-CTA neurons, 3.4 for <= 148 and 3.2 per wave beyond.  With these, 50 local neurons (cfg3 on 4 GPUs) run as two
// waves of 4-CTA clusters: scan 2.86 -> 2.46 ms inside the sweep, 149.1 -> 159.1 sweeps/s (profiles/r02ah_bench_4gpu_*).
// The ratios, not the absolute values, decide; PYGLM_SS_DSM_MAX_WAVES caps the waves a cluster shape may need (default 2).
int choose_cluster(const SpikeSlabArgs& A) {
    if (A.B < 1 || A.B > 4 || A.N * A.B + 1 < 24) return 0;
    static int max_waves = -1;
    if (max_waves < 0) { const char* e = getenv("PYGLM_SS_DSM_MAX_WAVES"); max_waves = e ? atoi(e) : 2; }
    static int cap_key = -1, caps[3] = {0, 0, 0};
    const int shapes[3] = {2, 4, 8};
    const double t_shape[3] = {2.0, 1.35, 1.05};
    const int key = A.N * 8 + A.B;
    if (key != cap_key) {
        for (int k = 0; k < 3; ++k) {
            const int c = shapes[k];
            caps[k] = 0;
            if (dsm_geom(A.N, A.B, c, dsm_sym(c)).LAG < 1) continue;
            switch (A.B) {
                case 1: caps[k] = dsm_active_clusters_b<1>(A.N, c); break;
                case 2: caps[k] = dsm_active_clusters_b<2>(A.N, c); break;
                case 3: caps[k] = dsm_active_clusters_b<3>(A.N, c); break;
                default: caps[k] = dsm_active_clusters_b<4>(A.N, c); break;
            }
        }
        cap_key = key;
    }
    // small problems (cfg2: D = 82) are barrier-bound whatever the cluster size: there the shapes cost the same per
    // neuron and the smallest one that fits wins by needing the fewest SMs
    const bool small = A.N * A.B + 1 <= 200;
    const int w1 = (A.n_loc + 147) / 148;
    double best = (A.n_loc <= 100) ? 3.0 : (w1 == 1 ? 3.4 : 3.2 * w1);
    int pick = 0;
    for (int k = 0; k < 3; ++k) {
        if (caps[k] <= 0) continue;
        const int waves = (A.n_loc + caps[k] - 1) / caps[k];
        if (waves > max_waves) continue;
        const double t = waves * (small ? 1.0 : t_shape[k]);
        if (t < best) { best = t; pick = shapes[k]; }
    }
    return pick;
}

bool spike_slab_dsm_preferred(const SpikeSlabArgs& A) { return choose_cluster(A) != 0; }

}  // namespace pyglm_ss
