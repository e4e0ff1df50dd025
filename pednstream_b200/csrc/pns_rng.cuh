// Counter-based sampling for the on-device ("philox") draw mode.
//
// The reference draws from numpy's global MT19937 stream in node-visiting order
// (SURVEY.md "RNG ledger": R1/R2 link.py:337,343,356; R3 link.py:382; R4 functions.py:133), which
// cannot be generated in parallel.  The device mode replaces the *generator*, not the
// distributions: every draw is an exact Binomial(n, p) / Normal(0, sigma) variate keyed by
// (seed, replica; t, link, site), so results do not depend on thread scheduling.
//
// Every arithmetic step below is a single IEEE-754 operation (the library is compiled with
// -fmad=false; the fused multiply-adds that are wanted are written explicitly) and uses no libdevice
// transcendental, so `oracle/philox.py`, which restates the same operation sequence in Python floats
// (fused multiply-adds through exact rationals), reproduces the device samples bit for bit.
#pragma once
#include <stdint.h>

namespace pns {

struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1).
__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.v[0] = c[0]; out.v[1] = c[1]; out.v[2] = c[2]; out.v[3] = c[3];
    return out;
}

// 53-bit uniform in [0, 1) from two words.
__host__ __device__ inline double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);   // exact: power of two
}

// ---- deterministic elementary functions (fdlibm-style kernels, basic operations only) ------
__host__ __device__ inline double det_log(double x) {  // x > 0, normal
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    union { double d; uint64_t u; } cv;
    cv.d = x;
    int e = (int)((cv.u >> 52) & 0x7ffu) - 1023;
    cv.u = (cv.u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = cv.d;  // [1, 2)
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double hfsq = (0.5 * f) * f;
    const double dk = (double)e;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

__host__ __device__ inline double det_exp(double x) {  // |x| < 700
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double invln2 = 1.44269504088896338700e+00;
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    const int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));  // nearest integer (C cast truncates)
    const double dk = (double)k;
    const double hi = x - dk * ln2_hi;
    const double lo = dk * ln2_lo;
    const double r = hi - lo;
    const double t = r * r;
    const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    union { double d; uint64_t u; } sc;
    sc.u = (uint64_t)(k + 1023) << 52;  // 2^k, |k| <= 1010
    return y * sc.d;
}

__host__ __device__ inline double det_sin_k(double x) {  // |x| <= pi/4
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x + v * (S1 + z * r);
}

__host__ __device__ inline double det_cos_k(double x) {  // |x| <= pi/4
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    return (1.0 - 0.5 * z) + z * r;
}

// (cos, sin)(2*pi*u), u in [0, 1)
__host__ __device__ inline void det_sincos2pi(double u, double* cos_out, double* sin_out) {
    const double half_pi = 1.5707963267948966;
    const double u4 = u * 4.0;
    const int q = (int)u4;  // 0..3
    const double f = u4 - (double)q;
    double c, s;
    if (f <= 0.5) {
        const double a = f * half_pi;
        c = det_cos_k(a);
        s = det_sin_k(a);
    } else {
        const double a = (1.0 - f) * half_pi;
        c = det_sin_k(a);
        s = det_cos_k(a);
    }
    if (q == 0) { *cos_out = c; *sin_out = s; }
    else if (q == 1) { *cos_out = -s; *sin_out = c; }
    else if (q == 2) { *cos_out = -c; *sin_out = -s; }
    else { *cos_out = s; *sin_out = -c; }
}

// cos(2*pi*u), u in [0, 1)
__host__ __device__ inline double det_cos2pi(double u) {
    const double half_pi = 1.5707963267948966;
    const double u4 = u * 4.0;
    const int q = (int)u4;  // 0..3
    const double f = u4 - (double)q;
    double c, s;
    if (f <= 0.5) {
        const double a = f * half_pi;
        c = det_cos_k(a);
        s = det_sin_k(a);
    } else {
        const double a = (1.0 - f) * half_pi;
        c = det_sin_k(a);
        s = det_cos_k(a);
    }
    if (q == 0) return c;
    if (q == 1) return -s;
    if (q == 2) return -c;
    return s;
}

// x**0.8 for the release probability (link.py:317), x in [0, 1] fp32
__host__ __device__ inline float det_pow08(float x) {
    if (!(x > 0.0f)) return 0.0f;
    if (x >= 1.0f) return 1.0f;
    return (float)det_exp((double)0.8f * det_log((double)x));   // numpy demotes the exponent to float32
}

// ---- samplers ---------------------------------------------------------------------------------
struct DrawKey {
    uint32_t t, link, replica, k0, k1;
};

// ---- sampler tables (filled on the device by init_sampler_tables, once per device) ------------
// kInvK[k] = 1/k correctly rounded, k = 1..PNS_INV_TABLE: the pmf recurrences multiply by these instead of
// dividing (a division is an 80-130 cycle dependent chain per step; a table load is off the critical path).
// kPmfMode09[n] = pmf of Binomial(n, 1 - 0.9) at its mode, n <= PNS_MODE_TABLE_N: the blockers draw
// (link.py:382) always has p = 0.9, so its mode-centred search starts from a table load.
#ifndef PNS_INV_TABLE
#define PNS_INV_TABLE 2048
#endif
#define PNS_MODE_TABLE_N 4096
#define PNS_MODE_MEAN_TABULATED 12.0   // mean from which the search starts at the mode (mode pmf tabulated)
#define PNS_MODE_MEAN_GENERIC 48.0     // ... for any other probability (mode pmf evaluated per draw)
// kPow09[n] = 0.9^n as binomial_from_zero forms it (by squaring), n <= PNS_MODE_TABLE_N: pmf(0) of a blockers draw.
#ifdef __CUDACC__
__device__ double kInvK[PNS_INV_TABLE + 1];
__device__ double kPmfMode09[PNS_MODE_TABLE_N + 1];
__device__ double kPow09[PNS_MODE_TABLE_N + 1];
#define PNS_TABLE(a, i) __ldg((a) + (i))
#else
static double kInvK[PNS_INV_TABLE + 1];
static double kPmfMode09[PNS_MODE_TABLE_N + 1];
static double kPow09[PNS_MODE_TABLE_N + 1];
#define PNS_TABLE(a, i) ((a)[i])
#endif
__host__ __device__ inline double pns_inv_k(int k) {   // k >= 1
#ifdef __CUDA_ARCH__
    return k <= PNS_INV_TABLE ? __ldg(kInvK + k) : 1.0 / (double)k;
#else
    return 1.0 / (double)k;
#endif
}
__host__ __device__ inline double pns_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
__host__ __device__ inline double pns_sqrt(double x) {
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}

// Binomial(n, pp) by CDF inversion from 0, 0 < pp <= 0.5, u in [0, 1); used for small means n*pp, where q^n is
// far from underflow.  pmf(k) = pmf(k-1) * f_k with f_k = ratio*(n-k+1)/k = A/k - ratio, A = ratio*(n+1): one
// table load and one fused multiply-add, formed off the critical path (the loop-carried chain is one multiply,
// one subtract and a compare).  Four steps per trip, because an in-order warp cannot overlap loop trips by
// itself (restated in oracle/philox.py).
__host__ __device__ inline double pow_by_squaring(double q, int n) {
    double pk = 1.0, b = q;
    for (int e = n; e; e >>= 1) {
        if (e & 1) pk = pk * b;
        b = b * b;
    }
    return pk;
}
__host__ __device__ inline int binomial_from_zero(int n, double pp, double u, double p0_tabulated = -1.0) {
    const double q = 1.0 - pp;
    const double ratio = pp / q;
    const double A = ratio * (double)(n + 1);
    double pk = p0_tabulated >= 0.0 ? p0_tabulated : pow_by_squaring(q, n);
    int k = 0;
    while (u > pk && k < n) {
        const double f1 = pns_fma(A, pns_inv_k(k + 1), -ratio);
        const double f2 = pns_fma(A, pns_inv_k(k + 2), -ratio);
        const double f3 = pns_fma(A, pns_inv_k(k + 3), -ratio);
        const double f4 = pns_fma(A, pns_inv_k(k + 4), -ratio);
        u = u - pk; k += 1; pk = pk * f1;
        if (!(u > pk && k < n)) break;
        u = u - pk; k += 1; pk = pk * f2;
        if (!(u > pk && k < n)) break;
        u = u - pk; k += 1; pk = pk * f3;
        if (!(u > pk && k < n)) break;
        u = u - pk; k += 1; pk = pk * f4;
    }
    return k;
}

// log(x!) - log(sqrt(2 pi x) (x/e)^x) for an integer-valued x >= 0 (Loader, "Fast and accurate computation of
// binomial probabilities", 2000): table below 16, the asymptotic series above
#ifdef __CUDACC__
__device__
#endif
const double kStirlingErr[16] = {
    0x0.0p+0, 0x1.4c071bcda0a5bp-4, 0x1.52a9b923ea649p-5, 0x1.c579a268d80b3p-6, 0x1.54a2662fd78a9p-6,
    0x1.10b4e513fcbedp-6, 0x1.c6b167bebdf36p-7, 0x1.85d4d612e4a86p-7, 0x1.552805e7b3076p-7,
    0x1.2f4871b12ab64p-7, 0x1.10f9d4c0743a7p-7, 0x1.f0593088014f8p-8, 0x1.c7018733aa9c6p-8,
    0x1.a40514700f36cp-8, 0x1.86076c002d4a7p-8, 0x1.6c08f6f194a10p-8};
#ifdef __CUDACC__
__device__
#endif
inline double pns_stirlerr(double x) {
    if (x < 16.0) return kStirlingErr[(int)x];
    const double r = 1.0 / x;
    const double rr = r * r;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) * rr) * rr) * rr) * rr) * r;
}

// deviance term x*log(x/np) + np - x by its series in v = (x-np)/(x+np) (|x - np| << x + np at the mode)
#ifdef __CUDACC__
__device__
#endif
inline double pns_bd0(double x, double np) {
    const double d = x - np;
    double v = d / (x + np);
    double s = d * v;
    if (fabs(s) < 2.2250738585072014e-308) return s;
    double ej = (2.0 * x) * v;
    v = v * v;
    for (int j = 1; j < 64; ++j) {
        ej = ej * v;
        const double s1 = s + ej * (1.0 / (double)(2 * j + 1));
        if (s1 == s) return s1;
        s = s1;
    }
    return s;
}

// pmf of Binomial(n, pp) at m (0 < m < n), saddle-point form (Loader 2000); relative error ~2e-15
#ifdef __CUDACC__
__device__
#endif
inline double binomial_pmf_mode(int n, int m, double pp, double q) {
    const double nd = (double)n, md = (double)m, kd = (double)(n - m);
    const double lc = (((pns_stirlerr(nd) - pns_stirlerr(md)) - pns_stirlerr(kd)) - pns_bd0(md, nd * pp)) -
                      pns_bd0(kd, nd * q);
    return det_exp(lc) * pns_sqrt(nd / ((6.283185307179586 * md) * kd));
}

// Binomial(n, pp) for large means: search outward from the mode m = floor((n+1) pp) -- m, then the pairs
// {m+1, m-1}, {m+2, m-2}, ... -- subtracting the pmf values from u.  Expected work ~1.6 standard deviations instead of the mean (a jammed link's
// blockers draw, Binomial(1200, 0.9): ~17 candidates instead of a 120-step walk).  Up and down recurrences:
// pmf(k+1)/pmf(k) = A/(k+1) - ratio, pmf(k-1)/pmf(k) = B/(n-k+1) - 1/ratio, A = ratio (n+1), B = (n+1)/ratio.
#ifdef __CUDACC__
__device__
#endif
inline int binomial_from_mode(int n, double pp, double u, double pm_tabulated) {
    const double q = 1.0 - pp;
    const double ratio = pp / q;
    const double iratio = q / pp;
    const int m = (int)((double)(n + 1) * pp);
    const double pm = pm_tabulated > 0.0 ? pm_tabulated : binomial_pmf_mode(n, m, pp, q);
    if (!(u > pm)) return m;
    u = u - pm;
    const double A = ratio * (double)(n + 1);
    const double B = iratio * (double)(n + 1);
    double pu = pm, pd = pm;
    // candidates m+i and m-i are tested as a pair: one subtraction and one compare for two pmf values.  Both
    // exist for i <= m (the mode is at most n/2), which covers all but the far tail of u.
    const int paired = m < n - m ? m : n - m;
    int i = 1;
    for (; i <= paired; ++i) {
        pu = pu * pns_fma(A, pns_inv_k(m + i), -ratio);
        pd = pd * pns_fma(B, pns_inv_k(n - m + i), -iratio);
        const double both = pu + pd;
        if (!(u > both)) return !(u > pu) ? m + i : m - i;
        u = u - both;
    }
    // far tail: one side is exhausted
    int ku = m + paired, kd = m - paired;
    for (;;) {
        if (ku < n) {
            ku += 1;
            pu = pu * pns_fma(A, pns_inv_k(ku), -ratio);
            if (!(u > pu)) return ku;
            u = u - pu;
        } else {
            pu = 0.0;
        }
        if (kd > 0) {
            pd = pd * pns_fma(B, pns_inv_k(n - kd + 1), -iratio);
            kd -= 1;
            if (!(u > pd)) return kd;
            u = u - pd;
        } else {
            pd = 0.0;
        }
        if (!(pu > 0.0) && !(pd > 0.0)) return m;   // u fell into the rounding leftover of the total mass
    }
}

// Exact Binomial(n, p), 0 < p < 1, n > 0: CDF inversion of one uniform -- from zero for small means n*min(p,1-p),
// outward from the mode for large ones; 0.9^n and the mode pmf of the blockers draw (p = 0.9) come from tables.
// The two searches are separate out-of-line functions with scalar arguments (Philox block included): draws happen
// on a minority of the links and must not bloat the callers' register footprint, and the rare, register-hungry
// mode search must not tax the call site of the common one (inlined, or behind one shared entry point, the
// 64-register link kernels spill).  (A rejection sampler for large means -- BTRS -- was tried in round 1: spills on
// the main path, and a warp runs its slow path almost every time because some lane needs it.)
// The release (R1) and blockers (R3) draws of a link and step take their uniforms from one Philox block (site 1):
// words 0,1 and words 2,3; a draw with a block of its own (the activity draw R2, test hooks) names its site.
__host__ __device__ inline double site_uniform(uint32_t t, uint32_t link, uint32_t replica, uint32_t k0, uint32_t k1,
                                               uint32_t site, uint32_t word) {
    const Philox4 w = philox4x32_10(t, link, site, replica, k0, k1);
    return word ? u53(w.v[2], w.v[3]) : u53(w.v[0], w.v[1]);
}
#ifdef __CUDACC__
__device__ __noinline__
#else
inline
#endif
int binomial_zero_core(uint32_t t, uint32_t link, uint32_t replica, uint32_t k0, uint32_t k1, uint32_t site,
                       uint32_t word, int n, double pp, bool tabulated) {
    return binomial_from_zero(n, pp, site_uniform(t, link, replica, k0, k1, site, word),
                              tabulated ? PNS_TABLE(kPow09, n) : -1.0);
}
#ifdef __CUDACC__
__device__ __noinline__
#else
inline
#endif
int binomial_mode_core(uint32_t t, uint32_t link, uint32_t replica, uint32_t k0, uint32_t k1, uint32_t site,
                       uint32_t word, int n, double pp, bool tabulated) {
    return binomial_from_mode(n, pp, site_uniform(t, link, replica, k0, k1, site, word),
                              tabulated ? PNS_TABLE(kPmfMode09, n) : 0.0);
}

#ifdef __CUDACC__
#define PNS_DEVFN __device__ __forceinline__
#else
#define PNS_DEVFN inline
#endif
PNS_DEVFN int binomial_draw(const DrawKey& key, uint32_t site, uint32_t word, int n, double p) {
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    const double pp = flip ? 1.0 - p : p;
    const bool tabulated = pp == (1.0 - 0.9) && n <= PNS_MODE_TABLE_N;
    const int k = (double)n * pp >= (tabulated ? PNS_MODE_MEAN_TABULATED : PNS_MODE_MEAN_GENERIC)
                      ? binomial_mode_core(key.t, key.link, key.replica, key.k0, key.k1, site, word, n, pp, tabulated)
                      : binomial_zero_core(key.t, key.link, key.replica, key.k0, key.k1, site, word, n, pp, tabulated);
    return flip ? n - k : k;
}
PNS_DEVFN int binomial_philox(const DrawKey& key, uint32_t site, int n, double p) { return binomial_draw(key, site, 0u, n, p); }
PNS_DEVFN int binomial_release(const DrawKey& key, int n, double p) { return binomial_draw(key, 1u, 0u, n, p); }   // R1
PNS_DEVFN int binomial_blockers(const DrawKey& key, int n) { return binomial_draw(key, 1u, 1u, n, 0.9); }          // R3

// one entry of the sampler tables (see kInvK / kPmfMode09)
#ifdef __CUDACC__
__device__
#endif
inline void init_sampler_table_entry(int i) {
    if (i >= 1 && i <= PNS_INV_TABLE) kInvK[i] = 1.0 / (double)i;
    if (i == 0) kInvK[0] = 0.0;
    if (i <= PNS_MODE_TABLE_N) {
        const double pp = 1.0 - 0.9, q = 1.0 - pp;
        const int m = (int)((double)(i + 1) * pp);
        kPmfMode09[i] = (m > 0 && m < i) ? binomial_pmf_mode(i, m, pp, q) : 0.0;
        kPow09[i] = pow_by_squaring(q, i);
    }
}

// Poisson(lam), 0 <= lam <= 700, by CDF inversion from 0 (demand draws of the batched environment:
// lam is a pedestrian arrival rate per step, a few tens): pmf(0) = exp(-lam), pmf(k) = pmf(k-1)*lam/k.
__host__ __device__ inline int poisson_inversion(double lam, double u) {
    if (!(lam > 0.0)) return 0;
    double pk = det_exp(-lam);
    int k = 0;
    while (u > pk && k < 4096) {
        u = u - pk;
        k += 1;
        pk = (pk * lam) / (double)k;
    }
    return k;
}

// Standard normal by Box-Muller on one Philox block.
__host__ __device__ inline double normal_philox(const DrawKey& key, uint32_t site) {
    const Philox4 w = philox4x32_10(key.t, key.link, site, key.replica, key.k0, key.k1);
    const double u1 = 1.0 - u53(w.v[0], w.v[1]);  // (0, 1]
    const double u2 = u53(w.v[2], w.v[3]);
    double rad = -2.0 * det_log(u1);
#ifdef __CUDA_ARCH__
    rad = __dsqrt_rn(rad);
#else
    rad = __builtin_sqrt(rad);
#endif
    return rad * det_cos2pi(u2);
}

// ---- float32 Box-Muller (speed noise, functions.py:132-133) ---------------------------------------
// The noise sample feeds a float32 history value (speed[t]), so it is generated in float32: one
// Philox block -> two 24-bit uniforms -> (radius, angle) -> two independent standard normals.
// Polynomials are the Cephes single-precision kernels, evaluated with explicit fused
// multiply-adds (one rounding each) so that oracle/philox.py can restate them exactly.
__host__ __device__ inline float pns_fmaf(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
__host__ __device__ inline float pns_sqrtf(float x) {
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}

__host__ __device__ inline float det_logf(float x) {   // x in (0, 1], normal
    union { float f; uint32_t u; } cv;
    cv.f = x;
    int e = (int)((cv.u >> 23) & 0xffu) - 127;
    cv.u = (cv.u & 0x007fffffu) | 0x3f800000u;
    float m = cv.f;                                      // [1, 2)
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    const float f = m - 1.0f;
    const float z = f * f;
    float y = 7.0376836292e-2f;
    y = pns_fmaf(y, f, -1.1514610310e-1f);
    y = pns_fmaf(y, f, 1.1676998740e-1f);
    y = pns_fmaf(y, f, -1.2420140846e-1f);
    y = pns_fmaf(y, f, 1.4249322787e-1f);
    y = pns_fmaf(y, f, -1.6668057665e-1f);
    y = pns_fmaf(y, f, 2.0000714765e-1f);
    y = pns_fmaf(y, f, -2.4999993993e-1f);
    y = pns_fmaf(y, f, 3.3333331174e-1f);
    y = (y * f) * z;
    const float fe = (float)e;
    y = pns_fmaf(-2.12194440e-4f, fe, y);
    y = pns_fmaf(-0.5f, z, y);
    return pns_fmaf(0.693359375f, fe, f + y);
}

// (cos, sin)(2*pi*u), u in [0, 1), float32
__host__ __device__ inline void det_sincos2pif(float u, float* cos_out, float* sin_out) {
    const float half_pi = 1.57079632679f;
    const float u4 = u * 4.0f;
    const int q = (int)u4;                               // 0..3
    const float f = u4 - (float)q;
    const bool lo = f <= 0.5f;
    const float a = (lo ? f : 1.0f - f) * half_pi;       // [0, pi/4]
    const float z = a * a;
    float ps = pns_fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = pns_fmaf(ps, z, -1.6666654611e-1f);
    const float sn = pns_fmaf(ps * z, a, a);
    float pc = pns_fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = pns_fmaf(pc, z, 4.166664568298827e-2f);
    const float cs = pns_fmaf(pc * z, z, pns_fmaf(-0.5f, z, 1.0f));
    const float c = lo ? cs : sn, s = lo ? sn : cs;
    if (q == 0) { *cos_out = c; *sin_out = s; }
    else if (q == 1) { *cos_out = -s; *sin_out = c; }
    else if (q == 2) { *cos_out = -c; *sin_out = -s; }
    else { *cos_out = s; *sin_out = -c; }
}

// Box-Muller on two 32-bit words: (cosine branch, sine branch)
__host__ __device__ inline void box_muller_f32(uint32_t wa, uint32_t wb, float* g_cos, float* g_sin) {
    const float u1 = (float)((wa >> 8) + 1u) * 5.9604644775390625e-8f;   // (0, 1], 24 bits
    const float u2 = (float)(wb >> 8) * 5.9604644775390625e-8f;          // [0, 1)
    const float rad = pns_sqrtf(-2.0f * det_logf(u1));
    float cs, sn;
    det_sincos2pif(u2, &cs, &sn);
    *g_cos = rad * cs;
    *g_sin = rad * sn;
}

// Speed noise is drawn per *quad* of links (two adjacent corridors, links 4q .. 4q+3): one Philox
// block keyed by the quad's first link gives four independent standard normals -- words 0,1 serve
// links 4q (cosine branch) and 4q+1 (sine branch), words 2,3 serve links 4q+2 and 4q+3.  Nothing of
// the block is wasted, and a warp of the single-replica link kernel can draw for four warps.
__host__ __device__ inline void normal_quad_philox(const DrawKey& key, uint32_t site, float g[4]) {
    const Philox4 w = philox4x32_10(key.t, key.link, site, key.replica, key.k0, key.k1);
    box_muller_f32(w.v[0], w.v[1], &g[0], &g[1]);
    box_muller_f32(w.v[2], w.v[3], &g[2], &g[3]);
}

// The two normals of one corridor (key.link = its even link): the half of the quad's block it owns.
__host__ __device__ inline void normal_pair_philox(const DrawKey& key, uint32_t site, double* g0, double* g1) {
    const Philox4 w = philox4x32_10(key.t, key.link & ~3u, site, key.replica, key.k0, key.k1);
    const bool second = (key.link & 2u) != 0;
    float a, b;
    box_muller_f32(second ? w.v[2] : w.v[0], second ? w.v[3] : w.v[1], &a, &b);
    *g0 = (double)a;
    *g1 = (double)b;
}

}  // namespace pns
