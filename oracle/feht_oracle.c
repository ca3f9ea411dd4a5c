/*
 * feht_oracle.c -- CPU restatement of feht's hot path.  TEST INFRASTRUCTURE ONLY
 * (see feht_oracle.h for who may use it and for the parity status: PINNED).
 *
 * All floating point is IEEE double evaluated exactly in the written order.
 * Build with -ffp-contract=off (oracle/Makefile does) so that gcc never fuses a*b+c;
 * exp/log are the host libm's, as in the reference binary (static glibc, feht.cabal:34).
 */
#include "feht_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * [3P] math-functions 0.2.1.0 -- Numeric.Polynomial.Chebyshev.chebyshevBroucke
 * foldr' over the coefficient vector (last coefficient first):
 *   step k (b0,b1,_) = (k + x2*b0 - b1, b0, b1);  result (b0 - b2) * 0.5
 * Serves: logGammaCorrection, log1p (below) -> logBeta -> choose -> FET.hs:108,111.
 * ------------------------------------------------------------------------------------------ */
double fo_chebyshev_broucke(double x, const double *cs, int n)
{
    double x2 = x * 2.0;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    for (int i = n - 1; i >= 0; --i) {
        double nb0 = (cs[i] + x2 * b0) - b1;
        b2 = b1;
        b1 = b0;
        b0 = nb0;
    }
    return (b0 - b2) * 0.5;
}

/* [3P] math-functions 0.2.1.0 -- Numeric.SpecFunctions.log1p (its own, not libm's). */
static const double FO_LOG1P_COEFFS[21] = {
    0.10378693562743769800686267719098e+1, -0.13364301504908918098766041553133e+0,
    0.19408249135520563357926199374750e-1, -0.30107551127535777690376537776592e-2,
    0.48694614797154850090456366509137e-3, -0.81054881893175356066809943008622e-4,
    0.13778847799559524782938251496059e-4, -0.23802210894358970251369992914935e-5,
    0.41640416213865183476391859901989e-6, -0.73595828378075994984266837031998e-7,
    0.13117611876241674949152294345011e-7, -0.23546709317742425136696092330175e-8,
    0.42522773276034997775638052962567e-9, -0.77190894134840796826108107493300e-10,
    0.14075746481359069909215356472191e-10, -0.25907105182504968258367235930752e-11,
    0.47890342136145783458424838262418e-12, -0.88986704750629086228437658354573e-13,
    0.16582880735152484063190219487513e-13, -0.31021103529441495264823896592188e-14,
    0.58220336232428664306184718162466e-15};

double fo_log1p(double x)
{
    const double m_epsilon = 2.2204460492503131e-16;
    double ax = fabs(x);
    if (x == 0.0) return 0.0;
    if (x == -1.0) return -INFINITY;
    if (x < -1.0) return NAN;
    if (ax < m_epsilon * 0.5) return x;
    if ((x >= 0.0 && x < 1e-8) || (x >= -1e-9 && x < 0.0)) return x * (1.0 - x * 0.5);
    if (ax < 0.375) return x * (1.0 - x * fo_chebyshev_broucke(x / 0.375, FO_LOG1P_COEFFS, 21));
    return log(1.0 + x);
}

/* [3P] Numeric.SpecFunctions.logGammaCorrection (10 <= x). */
static const double FO_LGC_COEFFS[7] = {
    0.1666389480451863247205729650822e+0,  -0.1384948176067563840732986059135e-4,
    0.9810825646924729426157171547487e-8,  -0.1809129475572494194263306266719e-10,
    0.6221098041892605227126015543416e-13, -0.3399615005417721944303330599666e-15,
    0.2683181998482698748957538846666e-17};

double fo_log_gamma_correction(double x)
{
    const double big = 94906265.62425156;
    if (x < 10.0) return NAN;
    if (x < big) {
        double t = 10.0 / x;
        return fo_chebyshev_broucke(t * t * 2.0 - 1.0, FO_LGC_COEFFS, 7) / x;
    }
    return 1.0 / (x * 12.0);
}

/* [3P] Numeric.SpecFunctions.logGamma -- Algorithm AS 245 (about 10 significant digits).
 * Call sites: incompleteGamma below (arguments 0.5 and 1.5 only on feht's path). */
double fo_log_gamma(double x)
{
    static const double r1[9] = {-2.66685511495, -24.4387534237, -21.9698958928, 11.1667541262,
                                 3.13060547623,  0.607771387771, 11.9400905721,  31.4690115749,
                                 15.2346874070};
    static const double r2[9] = {-78.3359299449, -142.046296688, 137.519416416, 78.6994924154,
                                 4.16438922228,  47.0668766060,  313.399215894, 263.505074721,
                                 43.3400022514};
    static const double r3[9] = {-2.12159572323e5, 2.30661510616e5, 2.74647644705e4,
                                 -4.02621119975e4, -2.29660729780e3, -1.16328495004e5,
                                 -1.46025937511e5, -2.42357409629e4, -5.70691009324e2};
    static const double r4[5] = {0.279195317918525, 0.4917317610505968, 0.0692910599291889,
                                 3.350343815022304, 6.012459259764103};
    const double alr2pi = 0.918938533204673;
    if (x <= 0.0) return INFINITY;
    if (x > 1e308) return INFINITY;
    if (x < 1.5) {
        double a, b, c;
        if (x < 0.5) { a = -log(x); b = x + 1.0; c = x; }
        else         { a = 0.0;     b = x;       c = x - 1.0; }
        double num = (((r1[4] * b + r1[3]) * b + r1[2]) * b + r1[1]) * b + r1[0];
        double den = (((b + r1[8]) * b + r1[7]) * b + r1[6]) * b + r1[5];
        return a + c * num / den;
    }
    if (x < 4.0) {
        double num = (((r2[4] * x + r2[3]) * x + r2[2]) * x + r2[1]) * x + r2[0];
        double den = (((x + r2[8]) * x + r2[7]) * x + r2[6]) * x + r2[5];
        return (x - 2.0) * num / den;
    }
    if (x < 12.0) {
        double num = (((r3[4] * x + r3[3]) * x + r3[2]) * x + r3[1]) * x + r3[0];
        double den = (((x + r3[8]) * x + r3[7]) * x + r3[6]) * x + r3[5];
        return num / den;
    }
    {
        double y = log(x);
        double k = x * (y - 1.0) - 0.5 * y + alr2pi;
        if (x > 3e6) return k;
        double x1 = 1.0 / x;
        double x2 = x1 * x1;
        return k + x1 * ((r4[2] * x2 + r4[1]) * x2 + r4[0]) / ((x2 + r4[4]) * x2 + r4[3]);
    }
}

/* [3P] Numeric.SpecFunctions.logBeta.  On feht's path both arguments are >= 51 (p >= 10 branch);
 * the other branches are restated for completeness. */
double fo_log_beta(double a, double b)
{
    const double m_ln_sqrt_2_pi = 0.9189385332046727418;
    double p = a < b ? a : b;
    double q = a < b ? b : a;
    double pq = p + q;
    double ppq = p / pq;
    if (p < 0.0) return NAN;
    if (p == 0.0) return INFINITY;
    if (p >= 10.0) {
        double c = fo_log_gamma_correction(q) - fo_log_gamma_correction(pq);
        return log(q) * (-0.5) + m_ln_sqrt_2_pi + fo_log_gamma_correction(p) + c +
               (p - 0.5) * log(ppq) + q * fo_log1p(-ppq);
    }
    if (q >= 10.0) {
        double c = fo_log_gamma_correction(q) - fo_log_gamma_correction(pq);
        return fo_log_gamma(p) + c + p - p * log(pq) + (q - 0.5) * fo_log1p(-ppq);
    }
    return fo_log_gamma(p) + fo_log_gamma(q) - fo_log_gamma(pq);
}

/* [3P] Numeric.SpecFunctions.choose:
 *   k > n -> 0;  k' = min k (n-k);  k' < 50 -> chooseExact (fold a*(nk+i)/i);
 *   else approx = exp(logChooseFast n k'), rounded to Int64 when < 2^63.
 * Call site: hypergeometric probability, FET.hs:108,111,112. */
static int g_choose_cache_enabled = 1;
#define FO_CACHE_N 1000
static double *g_choose_cache = NULL; /* [n][k'] dense FO_CACHE_N x (FO_CACHE_N/2+1), NaN = empty */
static pthread_mutex_t g_cache_mu = PTHREAD_MUTEX_INITIALIZER;

void fo_set_choose_cache(int enabled) { g_choose_cache_enabled = enabled; }

static double fo_choose_uncached(int n, int k)
{
    if (k > n) return 0.0;
    int kk = k < (n - k) ? k : (n - k);
    if (kk < 50) {
        double a = 1.0;
        double nk = (double)(n - kk);
        for (int i = 1; i <= kk; ++i) {
            double j = (double)i;
            a = a * (nk + j) / j;
        }
        return a;
    }
    {
        double dn = (double)n, dk = (double)kk;
        double lcf = -log(dn + 1.0) - fo_log_beta(dn - dk + 1.0, dk + 1.0);
        double approx = exp(lcf);
        const double max64 = 9223372036854775807.0; /* fromIntegral (maxBound :: Int64) = 2^63 */
        if (approx < max64) return (double)llrint(approx); /* round = banker's, as Haskell round */
        return approx;
    }
}

double fo_choose(int n, int k)
{
    if (!g_choose_cache_enabled || n < 0 || n >= FO_CACHE_N || k < 0 || k > n)
        return fo_choose_uncached(n, k);
    int kk = k < (n - k) ? k : (n - k);
    const int stride = FO_CACHE_N / 2 + 1;
    if (!g_choose_cache) {
        pthread_mutex_lock(&g_cache_mu);
        if (!g_choose_cache) {
            double *t = (double *)malloc(sizeof(double) * FO_CACHE_N * stride);
            for (long i = 0; i < (long)FO_CACHE_N * stride; ++i) t[i] = NAN;
            __atomic_store_n(&g_choose_cache, t, __ATOMIC_RELEASE);
        }
        pthread_mutex_unlock(&g_cache_mu);
    }
    double *slot = &g_choose_cache[(long)n * stride + kk];
    double v = *slot;
    if (v != v) { /* benign race: every thread writes the same bits */
        v = fo_choose_uncached(n, kk);
        *slot = v;
    }
    return v;
}

/* [3P] Numeric.SpecFunctions.incompleteGamma -- Algorithm AS 239.
 * Call site: complCumulative (chiSquared 1), FET.hs:113-114, with p = 0.5. */
double fo_incomplete_gamma(double p, double x)
{
    const double limit = -88.0, tolerance = 1e-14, overflow = 1e37;
    if (isnan(p) || isnan(x)) return NAN;
    if (x < 0.0 || p <= 0.0) return INFINITY;
    if (x == 0.0) return 0.0;
    if (p >= 2e5) {
        double a = 3.0 * sqrt(p) * (pow(x / p, 1.0 / 3.0) + 1.0 / (9.0 * p) - 1.0);
        return erfc(-a / 1.4142135623730950488) / 2.0;
    }
    if (p >= 500.0) { /* "approx" branch; unreachable on feht's path (p = 0.5) */
        double a = 3.0 * sqrt(p) * (pow(x / p, 1.0 / 3.0) + 1.0 / (9.0 * p) - 1.0);
        return erfc(-a / 1.4142135623730950488) / 2.0;
    }
    if (x >= 1e8) return 1.0;
    if (x <= 1.0 || x < p) {
        /* Pearson's series */
        double a0 = p * log(x) - x - fo_log_gamma(p + 1.0);
        double a = p, c = 1.0, g = 1.0;
        for (;;) {
            double a1 = a + 1.0;
            double c1 = c * x / a1;
            double g1 = g + c1;
            if (c1 <= tolerance) { g = g1; break; }
            a = a1; c = c1; g = g1;
        }
        double gg = a0 + log(g);
        return gg > limit ? exp(gg) : 0.0;
    }
    {
        /* continued fraction */
        double a = 1.0 - p;
        double b = a + x + 1.0;
        double c = 0.0;
        double p1 = 1.0, p2 = x, p3 = x + 1.0, p4 = x * b;
        double g = p3 / p4;
        for (;;) {
            double a1 = a + 1.0, b1 = b + 2.0, c1 = c + 1.0;
            double an = a1 * c1;
            double p5 = b1 * p3 - an * p1;
            double p6 = b1 * p4 - an * p2;
            double rn = p5 / p6;
            double tol = tolerance < tolerance * rn ? tolerance : tolerance * rn;
            if (fabs(g - rn) <= tol) break; /* returns g (the previous convergent) */
            if (fabs(p5) > overflow) {
                p1 = p3 / overflow; p2 = p4 / overflow; p3 = p5 / overflow; p4 = p6 / overflow;
            } else {
                p1 = p3; p2 = p4; p3 = p5; p4 = p6;
            }
            a = a1; b = b1; c = c1; g = rn;
        }
        double gg = p * log(x) - x - fo_log_gamma(p) + log(g);
        return gg > limit ? 1.0 - exp(gg) : 1.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * [3P] statistics 0.13.3.0 -- Statistics.Distribution.Hypergeometric
 * ------------------------------------------------------------------------------------------ */
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* probability (HD m l k) n.  Call sites: FET.hs:108 (probX) and :111 (every n in [0..k]). */
double fo_hypergeometric_probability(int m, int l, int k, int n)
{
    if (n < imax(0, m + k - l) || n > imin(m, k)) return 0.0;
    return fo_choose(m, n) * fo_choose(l - m, k - n) / fo_choose(l, k);
}

/* cumulative (HD m l k) x.  Call site: FET.hs:107 (one-tail). */
double fo_hypergeometric_cumulative(int m, int l, int k, double x)
{
    int minN = imax(0, m + k - l), maxN = imin(m, k);
    if (isnan(x) || isinf(x)) return x > 0 ? 1.0 : (x < 0 ? 0.0 : NAN);
    int n = (int)floor(x);
    if (n < minN) return 0.0;
    if (n >= maxN) return 1.0;
    double s = 0.0;
    for (int i = minN; i <= n; ++i) s += fo_hypergeometric_probability(m, l, k, i);
    return s < 1.0 ? s : 1.0; /* D.sumProbabilities = min 1 . sum */
}

/* complCumulative (chiSquared 1) x = 1 - cumulative; cumulative x | x <= 0 = 0
 *                                   | otherwise = incompleteGamma (ndf/2) (x/2).  FET.hs:113-114. */
double fo_chi_squared1_compl_cumulative(double x)
{
    double cum = (x <= 0.0) ? 0.0 : fo_incomplete_gamma(0.5, x / 2.0);
    return 1.0 - cum;
}

/* ------------------------------------------------------------------------------------------
 * FET.hs
 * ------------------------------------------------------------------------------------------ */
/* FET.hs:127-139 -- (x'*x'*(a+b+c+d)) / ((a+b)*(c+d)*(b+d)*(a+c)), x' = a*d - b*c, in Double. */
double fo_chi_square(int64_t a_, int64_t b_, int64_t c_, int64_t d_)
{
    double a = (double)a_, b = (double)b_, c = (double)c_, d = (double)d_;
    double x = (a * d) - (b * c);
    return (x * x * (a + b + c + d)) / ((a + b) * (c + d) * (b + d) * (a + c));
}

/* FET.hs:91-125. */
double fo_fet(int64_t goa, int64_t gob, int64_t gta, int64_t gtb, int mode)
{
    int64_t m = goa + gta;        /* FET.hs:115 */
    int64_t l = m + gob + gtb;    /* FET.hs:116 */
    int64_t k = goa + gob;        /* FET.hs:117 */
    int64_t gt = gta + gtb;       /* FET.hs:118 */
    if (!(l > 0 && goa >= 0 && gob >= 0 && gta >= 0 && gtb >= 0 && k > 0 && gt > 0))
        return 1.0;               /* FET.hs:99,103,123 */
    if (l >= 1000) {              /* FET.hs:105,109 */
        double chi_p = fo_chi_squared1_compl_cumulative(fo_chi_square(goa, gob, gta, gtb));
        return mode == FO_ONE_TAIL ? chi_p / 2.0 : chi_p;
    }
    if (mode == FO_ONE_TAIL)      /* FET.hs:107 */
        return fo_hypergeometric_cumulative((int)m, (int)l, (int)k, (double)goa);
    {
        /* FET.hs:108,111: sum . filter (<= probX) . fmap (probability d) $ [0..k] */
        double prob_x = fo_hypergeometric_probability((int)m, (int)l, (int)k, (int)goa);
        double s = 0.0;
        for (int n = 0; n <= (int)k; ++n) {
            double t = fo_hypergeometric_probability((int)m, (int)l, (int)k, n);
            if (t <= prob_x) s = s + t;
        }
        return s;
    }
}

/* ------------------------------------------------------------------------------------------
 * Comparison.hs
 * ------------------------------------------------------------------------------------------ */
int64_t fo_count_char_in_vector_by_indices(const uint8_t *v, int64_t len, uint8_t match,
                                           const int32_t *idx, int64_t n_idx)
{
    int64_t rt = 0;
    for (int64_t i = 0; i < n_idx; ++i) {
        int64_t vi = idx[i];
        if (vi < 0 || vi >= len) return -1; /* V.! bounds error */
        if (v[vi] == match) rt += 1;
    }
    return rt;
}

double fo_calculate_ratio(int64_t goa_, int64_t gob_, int64_t gta_, int64_t gtb_)
{
    double goa = (double)goa_, gob = (double)gob_, gta = (double)gta_, gtb = (double)gtb_;
    double go = goa + gob, gt = gta + gtb;
    if (go > 0 && gt > 0 && goa >= 0 && gob >= 0 && gta >= 0 && gtb >= 0)
        return (goa / go) - (gta / gt);
    return 0.0;
}

double fo_bonferroni(double p, int64_t n)
{
    double new_p = p * (double)n;
    return new_p > 1.0 ? 1.0 : new_p; /* NaN > 1 is False -> NaN stays */
}

uint8_t fo_replace_char(uint8_t match_char, uint8_t current_char)
{
    uint8_t n = current_char;
    if (n >= 'a' && n <= 'z') n = (uint8_t)(n - 'a' + 'A'); /* toUpper, ASCII range */
    if (n == match_char) return '1';
    if (n == 'A' || n == 'T' || n == 'C' || n == 'G') return '0';
    return n;
}

static int run_rows(const uint8_t *chars, const int64_t *row_offsets, int64_t r0, int64_t r1,
                    const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2, int correction,
                    int64_t bonferroni_n, int32_t *goa, int32_t *gob, int32_t *gta, int32_t *gtb,
                    double *p, double *ratio)
{
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *v = chars + row_offsets[r];
        int64_t len = row_offsets[r + 1] - row_offsets[r];
        /* Comparison.hs:53-56: four separate passes, as the reference does */
        int64_t a = fo_count_char_in_vector_by_indices(v, len, '1', g1, n1);
        int64_t c = fo_count_char_in_vector_by_indices(v, len, '1', g2, n2);
        int64_t b = fo_count_char_in_vector_by_indices(v, len, '0', g1, n1);
        int64_t d = fo_count_char_in_vector_by_indices(v, len, '0', g2, n2);
        if (a < 0 || b < 0 || c < 0 || d < 0) return -1;
        double pv = fo_fet(a, b, c, d, FO_TWO_TAIL); /* Comparison.hs:50-51 hard-codes TwoTail */
        if (correction == 1) pv = fo_bonferroni(pv, bonferroni_n);
        goa[r] = (int32_t)a; gob[r] = (int32_t)b; gta[r] = (int32_t)c; gtb[r] = (int32_t)d;
        p[r] = pv;
        ratio[r] = fo_calculate_ratio(a, b, c, d);
    }
    return 0;
}

int fo_run_comparison(const uint8_t *chars, const int64_t *row_offsets, int64_t n_rows,
                      const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2, int correction,
                      int64_t bonferroni_n, int32_t *goa, int32_t *gob, int32_t *gta, int32_t *gtb,
                      double *p, double *ratio)
{
    return run_rows(chars, row_offsets, 0, n_rows, g1, n1, g2, n2, correction, bonferroni_n, goa,
                    gob, gta, gtb, p, ratio);
}

typedef struct {
    const uint8_t *chars; const int64_t *row_offsets; int64_t r0, r1;
    const int32_t *g1; int64_t n1; const int32_t *g2; int64_t n2;
    int correction; int64_t bonferroni_n;
    int32_t *goa, *gob, *gta, *gtb; double *p, *ratio; int rc;
} mt_job;

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    j->rc = run_rows(j->chars, j->row_offsets, j->r0, j->r1, j->g1, j->n1, j->g2, j->n2,
                     j->correction, j->bonferroni_n, j->goa, j->gob, j->gta, j->gtb, j->p, j->ratio);
    return NULL;
}

int fo_run_comparison_mt(const uint8_t *chars, const int64_t *row_offsets, int64_t n_rows,
                         const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2,
                         int correction, int64_t bonferroni_n, int32_t *goa, int32_t *gob,
                         int32_t *gta, int32_t *gtb, double *p, double *ratio, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * n_threads);
    for (int t = 0; t < n_threads; ++t) {
        mt_job j = {chars, row_offsets, n_rows * t / n_threads, n_rows * (t + 1) / n_threads,
                    g1, n1, g2, n2, correction, bonferroni_n, goa, gob, gta, gtb, p, ratio, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(th);
    free(jobs);
    return rc;
}

/* Comparison.hs:148, 155, 160-162 */
typedef struct { double key; int64_t tie; int64_t row; } ord_item;

static int ord_cmp(const void *pa, const void *pb)
{
    const ord_item *a = (const ord_item *)pa, *b = (const ord_item *)pb;
    if (a->key > b->key) return -1; /* |ratio| descending */
    if (a->key < b->key) return 1;
    if (a->tie < b->tie) return -1; /* then name rank ascending */
    if (a->tie > b->tie) return 1;
    return 0;
}

int64_t fo_filter_and_order(const double *ratio, const int32_t *name_rank, int64_t n_rows,
                            double ratio_filter, int64_t *out_rows)
{
    ord_item *items = (ord_item *)malloc(sizeof(ord_item) * (n_rows > 0 ? n_rows : 1));
    int64_t n = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        double ar = fabs(ratio[r]);
        if (ratio_filter <= ar) { /* Comparison.hs:148 */
            items[n].key = ar;
            items[n].tie = name_rank ? name_rank[r] : r;
            items[n].row = r;
            ++n;
        }
    }
    qsort(items, (size_t)n, sizeof(ord_item), ord_cmp);
    for (int64_t i = 0; i < n; ++i) out_rows[i] = items[i].row;
    free(items);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs.  Integer-only, so CPU and GPU agree bit for bit.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t fo_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint64_t fo_row_hash(const fo_synth_spec *sp, int64_t row)
{
    return fo_mix64(sp->seed ^ ((uint64_t)row * 0x9E3779B97F4A7C15ULL));
}

uint8_t fo_synth_binary_cell(const fo_synth_spec *sp, int64_t row, int64_t col)
{
    uint64_t hr = fo_row_hash(sp, row);
    uint64_t hc = fo_mix64(hr + (uint64_t)col);
    if ((uint32_t)(hc >> 32) < sp->missing_thr) return '-';
    uint32_t thr;
    if (((hr >> 8) % 1000000ULL) < sp->planted_per_million)
        thr = col < sp->split_col ? sp->thr_hi : sp->thr_lo;
    else
        thr = sp->freq_thr[hr & 255];
    return ((uint32_t)hc < thr) ? '1' : '0';
}

uint8_t fo_synth_snp_cell(const fo_synth_spec *sp, int64_t row, int64_t col)
{
    static const char NUC[4] = {'A', 'C', 'G', 'T'};
    uint64_t hr = fo_row_hash(sp, row);
    uint64_t hc = fo_mix64(hr + (uint64_t)col);
    if ((uint32_t)(hc >> 32) < sp->missing_thr) return '-';
    uint32_t major = (uint32_t)(hr >> 40) & 3u;
    uint32_t minor = (major + 1u + (uint32_t)((hr >> 42) % 3ULL)) & 3u;
    uint32_t thr = sp->freq_thr[hr & 255];
    return (uint8_t)NUC[((uint32_t)hc < thr) ? minor : major];
}

void fo_synth_fill(const fo_synth_spec *sp, int snp_mode, int64_t row0, int64_t n_rows, uint8_t *out)
{
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t c = 0; c < sp->n_samples; ++c)
            out[r * sp->n_samples + c] = snp_mode ? fo_synth_snp_cell(sp, row0 + r, c)
                                                  : fo_synth_binary_cell(sp, row0 + r, c);
}
