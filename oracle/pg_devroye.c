/*
 * ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the Polya-gamma PG(1, z) sampler that the reference reaches
 * through `pypolyagamma.pgdrawvpar` (reference call sites: pyglm/regression.py:474-477
 * and :501-508).  pypolyagamma is a third-party PyPI dependency that is NOT vendored
 * under /root/reference and is unpinned in setup.py:13 (latest upstream known: 1.2.x).
 * What is restated here is its published algorithm: Devroye's exact rejection sampler
 * for J*(1, z) as given in Polson, Scott & Windle (JASA 2013, "Bayesian inference for
 * logistic models using Polya-Gamma latent variables", Sec. 4 + supplement) with the
 * truncation point t = 0.64:
 *
 *   z <- |psi| / 2,  fz = pi^2/8 + z^2/2
 *   repeat:
 *     with prob. p/(p+q): X = t + Exp(1)/fz            (truncated exponential tail)
 *     else              : X ~ IG(1/z, 1) truncated to (0, t]
 *     S = a_0(X), Y = U*S; alternate the series partial sums S -= a_1, S += a_2, ...
 *     accept (return X/4) when Y <= S after an odd term, restart when Y > S after an
 *     even term.
 *   a_n(x) = pi(n+1/2) exp(-(n+1/2)^2 pi^2 x / 2)                  for x >  t
 *          = pi(n+1/2) (2/(pi x))^{3/2} exp(-2 (n+1/2)^2 / x)      for 0 < x <= t
 *
 * PARITY UNPINNED: the reference's tests hold no golden vector for PG draws and
 * pypolyagamma itself cannot be installed here (no network, no wheel).  This file is
 * therefore validated distributionally (analytic mean/variance, KS vs. the series CDF)
 * in tests/test_oracle_pg.py, and the CUDA sampler is then validated against THIS file
 * on the same Philox stream.
 *
 * Two RNG modes:
 *   rng_kind 0: counter-based Philox4x32-10, one stream per element keyed by
 *               (seed, call_id, element index) -- bitwise the stream the CUDA kernel uses,
 *               so GPU and CPU draws agree element by element (up to libm ulps).
 *   rng_kind 1: one xoshiro256++ generator per OpenMP thread over a static partition of
 *               the flat array -- the structure of pgdrawvpar (one RNG object per thread);
 *               used only for the timed CPU baseline.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o libpg_oracle.so pg_devroye.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PG_TRUNC 0.64
#define PG_TRUNC_RECIP 1.5625
#define PG_PI 3.141592653589793238462643383279502884

/* ---------------------------------------------------------------- Philox4x32-10 */
typedef struct { uint32_t c[4]; uint32_t k[2]; uint32_t out[4]; int have; } philox_t;

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* stream convention shared with pyglm_b200/csrc/philox.cuh:
 *   key = (seed_lo, seed_hi); counter = (elem_lo, elem_hi, call_id, block#) */
static inline void philox_seed(philox_t* s, uint64_t seed, uint32_t call_id, uint64_t elem) {
    s->k[0] = (uint32_t)seed; s->k[1] = (uint32_t)(seed >> 32);
    s->c[0] = (uint32_t)elem; s->c[1] = (uint32_t)(elem >> 32);
    s->c[2] = call_id; s->c[3] = 0; s->have = 0;
}
static inline uint32_t philox_u32(philox_t* s) {
    if (s->have == 0) { philox4x32_10(s->c, s->k, s->out); s->c[3] += 1; s->have = 4; }
    return s->out[4 - (s->have--)];
}

/* ---------------------------------------------------------------- xoshiro256++ */
typedef struct { uint64_t s[4]; } xo_t;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(xo_t* g) {
    uint64_t* s = g->s;
    uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static inline void xo_seed(xo_t* g, uint64_t seed) { /* splitmix64 expansion */
    for (int i = 0; i < 4; ++i) {
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        g->s[i] = z ^ (z >> 31);
    }
}

/* ---------------------------------------------------------------- unified RNG front */
typedef struct { int kind; philox_t ph; xo_t xo; } rng_t;

/* uniform on [0,1) with 53 random bits; two 32-bit words a (high 27 bits) and b (high 26) */
static inline double rng_unif(rng_t* r) {
    if (r->kind == 0) {
        uint32_t a = philox_u32(&r->ph), b = philox_u32(&r->ph);
        return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
    }
    return (double)(xo_next(&r->xo) >> 11) * (1.0 / 9007199254740992.0);
}
static inline double rng_expon(rng_t* r) { return -log1p(-rng_unif(r)); }
/* square of a standard normal via Box-Muller: (sqrt(-2 log(1-u1)) cos(2 pi u2))^2 */
static inline double rng_norm_sq(rng_t* r) {
    double u1 = rng_unif(r), u2 = rng_unif(r);
    double c = cos(2.0 * PG_PI * u2);
    return -2.0 * log1p(-u1) * c * c;
}

/* ---------------------------------------------------------------- sampler pieces */
static inline double log_pnorm(double x) { return log(0.5 * erfc(-x * 0.70710678118654752440)); }

/* a_n(x), the alternating-series coefficients of the J*(1,0) density */
static inline double pg_a(int n, double x) {
    double K = (n + 0.5) * PG_PI;
    if (x > PG_TRUNC) return K * exp(-0.5 * K * K * x);
    if (x > 0.0) {
        double e = -1.5 * (log(0.5 * PG_PI) + log(x)) + log(K) - 2.0 * (n + 0.5) * (n + 0.5) / x;
        return exp(e);
    }
    return 0.0;
}

/* probability of proposing from the exponential tail, 1/(1+q/p) */
static inline double pg_mass_texpon(double z) {
    const double t = PG_TRUNC;
    double fz = 0.125 * PG_PI * PG_PI + 0.5 * z * z;
    double b = sqrt(1.0 / t) * (t * z - 1.0);
    double a = -sqrt(1.0 / t) * (t * z + 1.0);
    double x0 = log(fz) + fz * t;
    double xb = x0 - z + log_pnorm(b);
    double xa = x0 + z + log_pnorm(a);
    double qdivp = 4.0 / PG_PI * (exp(xb) + exp(xa));
    return 1.0 / (1.0 + qdivp);
}

/* inverse-Gaussian(1/z, 1) truncated to (0, t] */
static inline double pg_rtigauss(double z, rng_t* r) {
    const double t = PG_TRUNC;
    double X = t + 1.0;
    if (PG_TRUNC_RECIP > z) {           /* mu = 1/z > t: propose from the 1/chi^2 tail */
        double alpha = 0.0;
        while (rng_unif(r) > alpha) {
            double E1 = rng_expon(r), E2 = rng_expon(r);
            while (E1 * E1 > 2.0 * E2 / t) { E1 = rng_expon(r); E2 = rng_expon(r); }
            X = 1.0 + E1 * t;
            X = t / (X * X);
            alpha = exp(-0.5 * z * z * X);
        }
    } else {                            /* mu <= t: Michael-Schucany-Haas, reject X > t */
        double mu = 1.0 / z;
        while (X > t) {
            double Y = rng_norm_sq(r);
            double half_mu = 0.5 * mu, mu_Y = mu * Y;
            X = mu + half_mu * mu_Y - half_mu * sqrt(4.0 * mu_Y + mu_Y * mu_Y);
            if (rng_unif(r) > mu / (mu + X)) X = mu * mu / X;
        }
    }
    return X;
}

static double pg1_draw(double psi, rng_t* r) {
    double z = fabs(psi) * 0.5;
    double fz = 0.125 * PG_PI * PG_PI + 0.5 * z * z;
    double mass = pg_mass_texpon(z);
    for (;;) {
        double X;
        if (rng_unif(r) < mass) X = PG_TRUNC + rng_expon(r) / fz;
        else X = pg_rtigauss(z, r);
        double S = pg_a(0, X);
        double Y = rng_unif(r) * S;
        int n = 0;
        for (;;) {
            ++n;
            if (n & 1) { S -= pg_a(n, X); if (Y <= S) return 0.25 * X; }
            else       { S += pg_a(n, X); if (Y > S) break; }
        }
    }
}

/* ---------------------------------------------------------------- entry points */
/* out[i] ~ PG(1, psi[i]) for i in [0, n).  Stands where pgdrawvpar(ppgs, ones, psi, out)
 * stands in pyglm/regression.py:503-506 (b == 1 only, which is all the Bernoulli path uses). */
void pg1_drawv(const double* psi, int64_t n, uint64_t seed, uint32_t call_id,
               int rng_kind, int nthreads, double* out) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    if (rng_kind == 0) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int64_t i = 0; i < n; ++i) {
            rng_t r; r.kind = 0; philox_seed(&r.ph, seed, call_id, (uint64_t)i);
            out[i] = pg1_draw(psi[i], &r);
        }
    } else {
#pragma omp parallel num_threads(nthreads)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
            int tid = 0, nt = 1;
#endif
            rng_t r; r.kind = 1; xo_seed(&r.xo, seed * 0x9E3779B97F4A7C15ull + call_id * 1000003ull + tid);
            int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
            for (int64_t i = lo; i < hi; ++i) out[i] = pg1_draw(psi[i], &r);
        }
    }
}

int pg_omp_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* raw uniforms of one Philox element stream, for stream-parity tests of the CUDA generator */
void philox_stream_unif(uint64_t seed, uint32_t call_id, uint64_t elem, int count, double* out) {
    rng_t r; r.kind = 0; philox_seed(&r.ph, seed, call_id, elem);
    for (int i = 0; i < count; ++i) out[i] = rng_unif(&r);
}
