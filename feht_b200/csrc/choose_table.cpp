// choose_table.cpp -- host-built table of the reference's binomial coefficients.
//
// feht's two-sided Fisher test (src/FET.hs:108,111,112 under /root/reference) evaluates
//     probability (hypergeometric m l k) n = choose m n * choose (l-m) (k-n) / choose l k
// with `choose` from math-functions 0.2.1.0 (cabal.config:1381), whose values are NOT the
// exactly rounded binomials: k' = min k (n-k) < 50 uses a running product a*(nk+i)/i, larger k'
// uses exp(-log(n+1) - logBeta(n-k'+1, k'+1)).  Parity with feht (1e-12 relative, and identical
// `<= probX` decisions) therefore needs these very doubles, so the host evaluates that
// arithmetic once for every (n, k') with n < 1000 (FET.hs:105: larger tables take the chi-square
// branch) and the device gathers from the 2 MB table.  exp/log are the host libm's, as in the
// reference binary (static glibc, feht.cabal:34).
//
// Compiled with -ffp-contract=off: no fused multiply-add anywhere in this file.
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace feht {
namespace {

// Broucke's Chebyshev summation, coefficients consumed last-to-first
// (Numeric.Polynomial.Chebyshev.chebyshevBroucke).
template <int N>
double broucke(double x, const double (&c)[N]) {
    const double twice = x * 2.0;
    double cur = 0.0, prev = 0.0, prev2 = 0.0;
    for (int i = N; i-- > 0;) {
        const double next = (c[i] + twice * cur) - prev;
        prev2 = prev;
        prev = cur;
        cur = next;
    }
    return (cur - prev2) * 0.5;
}

constexpr double kLgcCoef[7] = {
    0.1666389480451863247205729650822e+0,  -0.1384948176067563840732986059135e-4,
    0.9810825646924729426157171547487e-8,  -0.1809129475572494194263306266719e-10,
    0.6221098041892605227126015543416e-13, -0.3399615005417721944303330599666e-15,
    0.2683181998482698748957538846666e-17};

constexpr double kLog1pCoef[21] = {
    0.10378693562743769800686267719098e+1,  -0.13364301504908918098766041553133e+0,
    0.19408249135520563357926199374750e-1,  -0.30107551127535777690376537776592e-2,
    0.48694614797154850090456366509137e-3,  -0.81054881893175356066809943008622e-4,
    0.13778847799559524782938251496059e-4,  -0.23802210894358970251369992914935e-5,
    0.41640416213865183476391859901989e-6,  -0.73595828378075994984266837031998e-7,
    0.13117611876241674949152294345011e-7,  -0.23546709317742425136696092330175e-8,
    0.42522773276034997775638052962567e-9,  -0.77190894134840796826108107493300e-10,
    0.14075746481359069909215356472191e-10, -0.25907105182504968258367235930752e-11,
    0.47890342136145783458424838262418e-12, -0.88986704750629086228437658354573e-13,
    0.16582880735152484063190219487513e-13, -0.31021103529441495264823896592188e-14,
    0.58220336232428664306184718162466e-15};

// logGammaCorrection for 10 <= x < 9.49e7 (always true here: arguments are 51 .. 1001)
double lgamma_correction(double x) {
    const double t = 10.0 / x;
    return broucke(t * t * 2.0 - 1.0, kLgcCoef) / x;
}

// math-functions' own log1p; only arguments in (-1, 0) occur (x = -p/(p+q), p <= q).
double log1p_mf(double x) {
    const double ax = std::fabs(x);
    if (x == 0.0) return 0.0;
    if (ax < 2.2204460492503131e-16 * 0.5) return x;
    if ((x >= 0.0 && x < 1e-8) || (x >= -1e-9 && x < 0.0)) return x * (1.0 - x * 0.5);
    if (ax < 0.375) return x * (1.0 - x * broucke(x / 0.375, kLog1pCoef));
    return std::log(1.0 + x);
}

// logBeta a b for min(a,b) >= 10
double log_beta_big(double a, double b) {
    const double p = a < b ? a : b, q = a < b ? b : a;
    const double pq = p + q, ppq = p / pq;
    const double c = lgamma_correction(q) - lgamma_correction(pq);
    return std::log(q) * (-0.5) + 0.9189385332046727418 + lgamma_correction(p) + c +
           (p - 0.5) * std::log(ppq) + q * log1p_mf(-ppq);
}

std::vector<double> build() {
    std::vector<double> tab(static_cast<size_t>(kChooseEntries));
    for (int n = 0; n < kChooseN; ++n) {
        double *row = tab.data() + choose_offset(n);
        const int kmax = n / 2;
        for (int kk = 0; kk <= kmax; ++kk) {
            double v;
            if (kk < 50) {
                // chooseExact: the product restarts for every k' because nk = n-k' changes
                v = 1.0;
                const double nk = static_cast<double>(n - kk);
                for (int i = 1; i <= kk; ++i) {
                    const double j = static_cast<double>(i);
                    v = v * (nk + j) / j;
                }
            } else {
                const double dn = n, dk = kk;
                const double approx = std::exp(-std::log(dn + 1.0) - log_beta_big(dn - dk + 1.0, dk + 1.0));
                // `approx < 2^63 -> round` never fires for k' >= 50 (C(100,50) ~ 1e29); kept for fidelity
                v = approx < 9223372036854775807.0 ? std::nearbyint(approx) : approx;
            }
            row[kk] = v;
        }
    }
    return tab;
}

}  // namespace

const std::vector<double> &host_choose_table() {
    static std::vector<double> tab;
    static std::once_flag once;
    std::call_once(once, [] { tab = build(); });
    return tab;
}

}  // namespace feht
