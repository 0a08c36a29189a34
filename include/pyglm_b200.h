/*
 * pyglm_b200 -- C ABI of the B200-native Gibbs hot path for PyGLM's sparse Bernoulli network GLM.
 *
 * The reference (slinderman/pyglm) has no FFI of its own: its boundary is the Python class API
 * (pyglm/models.py, pyglm/regression.py) and its only native call is pypolyagamma.pgdrawvpar.
 * Each entry point below states which reference code it stands for (paths relative to the reference
 * repository).  INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pyglm_last_error() returns the message
 *     (thread-local);
 *   - every pointer is a DEVICE pointer owned by the caller unless marked [host]; nothing is allocated,
 *     freed or synchronised inside; work is enqueued on the caller's stream;
 *   - all arithmetic is IEEE float64; matrices are row-major; `ld*` are row pitches in elements;
 *   - padded layouts:  Xp  (T x ldx)   design matrix, columns [0,N*B) = X (n-major, b-minor), column N*B = 1,
 *                                       the rest 0; ldx = N*B+1 rounded up to a multiple of 32;
 *                      Wt  (ldx x ldn) Wt[d, j] = coefficient d of local neuron j (row N*B = bias), zero padded;
 *                      psi, omega (T x ldn); ldn = n_local rounded up to a multiple of 64;
 *                      J   (n_local x ldx x ldx), lower triangle valid; h (n_local x ldx).
 */
#ifndef PYGLM_B200_H
#define PYGLM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pyglm_stream_t; /* == cudaStream_t */

/* library */
const char* pyglm_last_error(void);
int pyglm_abi_version(void);
int pyglm_device_check(void); /* 0 iff the current device is sm_100 */

/* (1) X = Y * basis.  pyglm/utils/basis.py:5-34 (convolve_with_basis); writes the padded layout Xp. */
int pyglm_filter_spikes(const double* S, const double* basis, int T, int N, int L, int B, int clip,
                        double* Xp, int ldx, pyglm_stream_t stream);
/* user-supplied regressors (models.py:76-77, add_data(data, X=...)): dense (T x NB) <-> padded */
int pyglm_pack_design(const double* X, int T, int NB, double* Xp, int ldx, pyglm_stream_t stream);
int pyglm_unpack_design(const double* Xp, int T, int NB, int ldx, double* X, pyglm_stream_t stream);

/* (5) psi = X.vec(a o W) + b for n neurons at once.  pyglm/regression.py:195-201 (activation). */
int pyglm_activation(const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n,
                     double* psi, int ldpsi, pyglm_stream_t stream);
/* sum_{t,j} y psi - log(1+e^psi).  pyglm/regression.py:491-494 summed as in models.py:82-96.
 * workspace: ceil(T/128) * ceil(n/8) doubles.  ll: one double. */
int pyglm_loglik(const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n,
                 const double* Y, int ldy, int y_col0, double* ll, double* workspace, pyglm_stream_t stream);
/* logistic(psi).  pyglm/regression.py:524-526 (mean), models.py:153-163 (means). */
int pyglm_means(const double* Xp, int ldx, const double* Wt, int ldw, int T, int D, int n,
                double* mu, int ldmu, pyglm_stream_t stream);

/* (2) omega ~ PG(1, psi).  pypolyagamma.pgdrawvpar as called at pyglm/regression.py:501-508. */
int pyglm_pg_draw(const double* psi, int ldpsi, long long T, int n_valid, double* omega, int ld_out,
                  unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
                  pyglm_stream_t stream);
/* The same draws (element by element) through the branch-compacted two-pass kernels: pass 1 takes the proposal
 * choice for every element and finishes the exponential-tail draws, pass 2 finishes the inverse-Gaussian draws from a
 * compacted index list so that all lanes of a warp run the same sampler.  workspace: pyglm_pg_draw_ws_bytes(T, n_valid)
 * bytes of device memory, contents irrelevant.  Falls back to pyglm_pg_draw when T * n_valid >= 2^32 - 1. */
size_t pyglm_pg_draw_ws_bytes(long long T, int n_valid);
int pyglm_pg_draw_ws(const double* psi, int ldpsi, long long T, int n_valid, double* omega, int ld_out,
                     unsigned long long seed, unsigned call_id, long long t_off, int n_off, int n_total,
                     void* workspace, size_t workspace_bytes, pyglm_stream_t stream);
int pyglm_philox_uniforms(unsigned long long seed, unsigned call_id, unsigned long long elem0, int n_elem,
                          int count, double* out, pyglm_stream_t stream); /* test hook */

/* (3) J_n = [X,1]^T diag(omega_n) [X,1], lower triangle.  pyglm/regression.py:251-256.
 * With Om := kappa = y - 1/2 and mode-1 tiles the bias row of the result is h = [X,1]^T kappa (:259-260). */
int pyglm_gram_tiles(int D, int mode, int* tiles /*[host]*/, int capacity);
int pyglm_gram_slabs(int ntiles, int n_valid, int T);
int pyglm_weighted_gram(const double* Xp, int ldx, int T, const double* Om, int ldo, int n_valid,
                        const int* tiles, int ntiles, int nslabs, double* J, long long stride_n, int ldj,
                        int i_base, double* workspace, pyglm_stream_t stream);

/* (3') The same weighted Gram (pyglm/regression.py:251-256) on the tcgen05 tensor cores, for non-negative designs:
 * both operands are split into S radix-256 digits (digit 0 unsigned, the rest signed), every digit product of order
 * a+b < S is an exact int8 x int8 -> int32 tcgen05.mma, orders are recombined in int64 and scaled to FP64 once.
 * S = 4 (or 5) agrees with the FP64 kernel to ~1e-10 relative (tests: 1e-9).  The Khatri-Rao operand is formed in
 * integer arithmetic, Z_fix = (xq_i xq_j + rnd(t)) >> 8S from the fixed-point design xq = rint(Xp 2^(8S-e) + dither),
 * so the resident digit planes (build_z) and the tiles the streaming kernel builds in shared memory (mma_stream) are
 * the same numbers and both kernels return the same Jint bit for bit.
 *   pyglm_gram_tc_geometry   out[8] = {M pairs, Mpad, Tpad, Npad, neurons per tile, neuron tiles, time chunks,
 *                                      64-byte K blocks per chunk}                                   [host]
 *   pyglm_column_max         cmax[c] = max_t A[t,c] (A >= 0), *neg_flag = 1 if any entry is negative
 *   pyglm_gram_tc_build_z    digit planes Zs[S][Mpad][Tpad] of Z[t,(i,j)] = Xp[t,i] Xp[t,j], i >= j, pair index
 *                            i(i+1)/2 + j; once per dataset (Z does not depend on the Gibbs state); Zs zeroed by caller
 *   pyglm_gram_tc_build_z_slab  the same for a time slab starting at global bin t_off (time-sharded runs): cmax is the
 *                            maximum over the WHOLE recording (all-reduced by the caller) and the rounding dither is
 *                            keyed by the global bin, so the digits do not depend on how the time axis is cut
 *   pyglm_gram_tc_slice_digits  the digit planes of omega for a GIVEN per-neuron scale omax (all-reduced slab maxima)
 *   pyglm_gram_tc_slice_omega  per sweep: omax[n] = max_t Om[t,n] and digit planes Os[S][Npad][Tpad] (tiled = 0, for
 *                            pyglm_gram_tc_mma) or K-block-major Os[S][Tpad/64][Npad][64] (tiled = 1, for mma_stream)
 *   pyglm_gram_tc_mma        Jint[n][pair] = sum_t sum_{a+b<S} 256^(S-1-a-b) Zs[a][pair][t] Os[b][n][t]  (exact int64)
 *   pyglm_gram_tc_finalize   J[n][i][j] = Jint[n][pair] * 2^(ex_i + ex_j + eo_n - 8S - 8), lower triangle
 *   pyglm_gram_tc_stream_tiles  (i block, j block) of the 8 x 16 pair tiles of the streaming kernel          [host]
 *   pyglm_gram_tc_quantize   xq[Tpad/32][Dp][32] uint32 fixed-point design (S = 4; Dp = D rounded up to 16; zeroed by
 *                            the caller) and rw[Tpad] uint64 rounding addends of a time slab; once per dataset
 *   pyglm_gram_tc_mma_stream the integer GEMM of pyglm_gram_tc_mma with NO resident Z: digit tiles of Z are built in
 *                            shared memory from xq / rw inside the kernel (S = 4); same Jint                     */
int pyglm_gram_tc_geometry(int D, int n_valid, long long T, int S, long long* out /*[host]*/);
int pyglm_column_max(const double* A, int ld, long long T, int ncols, double* cmax, int* neg_flag,
                     pyglm_stream_t stream);
int pyglm_gram_tc_build_z(const double* Xp, int ldx, long long T, int D, const double* cmax, int S,
                          unsigned char* Zs, long long Mpad, long long Tpad, pyglm_stream_t stream);
int pyglm_gram_tc_build_z_slab(const double* Xp, int ldx, long long T, long long t_off, int D, const double* cmax,
                               int S, unsigned char* Zs, long long Mpad, long long Tpad, pyglm_stream_t stream);
int pyglm_gram_tc_slice_digits(const double* Om, int ldo, long long T, int n_valid, int S, const double* omax,
                               unsigned char* Os, int Npad, long long Tpad, int tiled, pyglm_stream_t stream);
int pyglm_gram_tc_slice_omega(const double* Om, int ldo, long long T, int n_valid, int S, double* omax,
                              int* neg_flag, unsigned char* Os, int Npad, long long Tpad, int tiled,
                              pyglm_stream_t stream);
int pyglm_gram_tc_mma(const unsigned char* Zs, const unsigned char* Os, int D, int n_valid, long long T, int S,
                      long long* Jint, long long ldjint, int max_ctas, pyglm_stream_t stream);
/* measurement hook: the MMA schedule above without operand loads / result atomics (int8 tensor-pipe peak) */
int pyglm_gram_tc_mma_probe(const unsigned char* Zs, const unsigned char* Os, int D, int n_valid, long long T, int S,
                            long long* Jint, long long ldjint, pyglm_stream_t stream);
int pyglm_gram_tc_finalize(const long long* Jint, long long ldjint, const double* cmax, const double* omax,
                           int D, int n_valid, int S, double* J, long long stride_n, int ldj, pyglm_stream_t stream);
/* time-sharded runs: the exact int64 reduce-scatter of the slabs' partial sums fused into the finalize pass.  peers is a
 * DEVICE array of `world` base pointers to the ranks' Jint buffers (peer-mapped / symmetric memory, row pitch ldjint);
 * rows [row_off, row_off + n_valid) of every buffer are added and scaled into this rank's J.  Stands for the
 * reduce-scatter SURVEY 8(e) places after regression.py:251-256 on a time-sharded recording. */
int pyglm_gram_tc_finalize_peers(const long long* const* peers /*[device]*/, int world, long long row_off, long long ldjint,
                                 const double* cmax, const double* omax, int D, int n_valid, int S, double* J,
                                 long long stride_n, int ldj, pyglm_stream_t stream);
/* the all-gather of the new (a, W, b) rows (models.py:169-171) as NVLink stores: nbytes (multiple of 16) from
 * src to byte offset dst_off_bytes of each of the `world` peer-mapped buffers in the DEVICE array peers */
int pyglm_peer_push(const void* src, long long nbytes, void* const* peers /*[device]*/, int world, long long dst_off_bytes,
                    pyglm_stream_t stream);
int pyglm_gram_tc_stream_tiles(int D, int* tiles /*[host]*/, int capacity);
int pyglm_gram_tc_quantize(const double* Xp, int ldx, long long T, long long t_off, int D, const double* cmax,
                           unsigned int* xq, unsigned long long* rw, long long Tpad, pyglm_stream_t stream);
int pyglm_gram_tc_mma_stream(const unsigned int* xq, const unsigned long long* rw, const unsigned char* Os, int D,
                             int n_valid, long long T, int S, const int* tiles /*[device]*/, int n_tiles,
                             long long* Jint, long long ldjint, int max_ctas, pyglm_stream_t stream);

/* (4) spike-and-slab update of (a, W, b).  pyglm/regression.py:265-340 (_collapsed_resample_a,
 * _marginal_likelihood, _resample_W) with pybasicbayes' sample_discrete_from_log / sample_gaussian. */
size_t pyglm_spike_slab_workspace_doubles(int N, int B, int n_loc);
int pyglm_scan_randomness(int N, int B, int n_loc, int n_off, unsigned long long seed, unsigned call_id,
                          int* perm, double* us, double* z, int ldz, pyglm_stream_t stream);
int pyglm_spike_slab_update(int N, int B, int n_loc,
                            const double* J, long long stride_n, int ldj, const double* h, int ldh,
                            const double* J0w, const double* h0w, const double* J0b, const double* h0b,
                            const double* cprior, const double* logit_rho,
                            const int* perm, const double* us, const double* z, int ldz,
                            const unsigned char* do_scan, unsigned char* a, double* W, double* bias,
                            double* P_workspace, double* logodds, double* ml, int* status,
                            pyglm_stream_t stream);

/* (6) forward simulation.  pyglm/models.py:98-151 (generate) with pyglm/regression.py:528-541 (rvs): T sequential
 * steps x[t] = Y[t-L:t]^T flipud(basis), psi = Wm x[t] + bias, Y[t] = u < logistic(psi), as one persistent thread-block
 * cluster.  Wm (N x N*B) = model.weights reshaped (models.py:124); Xp (T x ldx) must arrive zeroed with the bias column
 * set (row 0 is the zero-history row); columns [0, N*B) of the other rows are written with the same fma order as
 * pyglm_filter_spikes, so Xp equals the filter of Y bit for bit.  Y (T x N) receives 0/1; U (T x N) or NULL receives
 * the uniforms (Philox stream (seed, call_id, t*N + n)), which lets a test replay the recursion on the host.
 * gauss_sd < 0: Bernoulli spikes as above; gauss_sd >= 0: Gaussian observations Y[t] = psi + gauss_sd * z
 * (SparseGaussianRegression.rvs, pyglm/regression.py:406-417), U then receives the standard normals z. */
int pyglm_generate(const double* Wm, const double* bias, const double* basis, int N, int B, int L, long long T,
                   unsigned long long seed, unsigned call_id, double gauss_sd, double* Xp, int ldx, double* Y, double* U,
                   pyglm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PYGLM_B200_H */
