// Counter-based sampling for the on-device ("philox") draw mode.
//
// The reference draws from numpy's global MT19937 stream in node-visiting order
// (SURVEY.md "RNG ledger": R1/R2 link.py:337,343,356; R3 link.py:382; R4 functions.py:133), which
// cannot be generated in parallel.  The device mode replaces the *generator*, not the
// distributions: every draw is an exact Binomial(n, p) / Normal(0, sigma) variate keyed by
// (seed, replica; t, link, site), so results do not depend on thread scheduling.
//
// Every arithmetic step below is a single IEEE-754 operation (the library is compiled with
// -fmad=false) and uses no libdevice transcendental, so `oracle/philox.py`, which restates the
// same operation sequence in Python floats, reproduces the device samples bit for bit.
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

// 1/k, k = 1..512, correctly rounded (what the host's 1.0 / k gives): the pmf recurrence of the
// inversion multiplies by these instead of dividing.  On the device they come from a 4 KB table:
// a division or __drcp_rn is an 80-130 cycle dependent chain per step, and the 120-step walk of
// a jammed link's blockers draw then keeps its CTA alive for 9 us (measured: the whole step got
// 9 us slower once the first links jammed); a table load is off the critical path.
#ifdef __CUDACC__
__device__ const double kInvK[513] = {
    0x0p+0, 0x1.0000000000000p+0, 0x1.0000000000000p-1, 0x1.5555555555555p-2, 0x1.0000000000000p-2, 0x1.999999999999ap-3,
    0x1.5555555555555p-3, 0x1.2492492492492p-3, 0x1.0000000000000p-3, 0x1.c71c71c71c71cp-4, 0x1.999999999999ap-4, 0x1.745d1745d1746p-4,
    0x1.5555555555555p-4, 0x1.3b13b13b13b14p-4, 0x1.2492492492492p-4, 0x1.1111111111111p-4, 0x1.0000000000000p-4, 0x1.e1e1e1e1e1e1ep-5,
    0x1.c71c71c71c71cp-5, 0x1.af286bca1af28p-5, 0x1.999999999999ap-5, 0x1.8618618618618p-5, 0x1.745d1745d1746p-5, 0x1.642c8590b2164p-5,
    0x1.5555555555555p-5, 0x1.47ae147ae147bp-5, 0x1.3b13b13b13b14p-5, 0x1.2f684bda12f68p-5, 0x1.2492492492492p-5, 0x1.1a7b9611a7b96p-5,
    0x1.1111111111111p-5, 0x1.0842108421084p-5, 0x1.0000000000000p-5, 0x1.f07c1f07c1f08p-6, 0x1.e1e1e1e1e1e1ep-6, 0x1.d41d41d41d41dp-6,
    0x1.c71c71c71c71cp-6, 0x1.bacf914c1bad0p-6, 0x1.af286bca1af28p-6, 0x1.a41a41a41a41ap-6, 0x1.999999999999ap-6, 0x1.8f9c18f9c18fap-6,
    0x1.8618618618618p-6, 0x1.7d05f417d05f4p-6, 0x1.745d1745d1746p-6, 0x1.6c16c16c16c17p-6, 0x1.642c8590b2164p-6, 0x1.5c9882b931057p-6,
    0x1.5555555555555p-6, 0x1.4e5e0a72f0539p-6, 0x1.47ae147ae147bp-6, 0x1.4141414141414p-6, 0x1.3b13b13b13b14p-6, 0x1.3521cfb2b78c1p-6,
    0x1.2f684bda12f68p-6, 0x1.29e4129e4129ep-6, 0x1.2492492492492p-6, 0x1.1f7047dc11f70p-6, 0x1.1a7b9611a7b96p-6, 0x1.15b1e5f75270dp-6,
    0x1.1111111111111p-6, 0x1.0c9714fbcda3bp-6, 0x1.0842108421084p-6, 0x1.0410410410410p-6, 0x1.0000000000000p-6, 0x1.f81f81f81f820p-7,
    0x1.f07c1f07c1f08p-7, 0x1.e9131abf0b767p-7, 0x1.e1e1e1e1e1e1ep-7, 0x1.dae6076b981dbp-7, 0x1.d41d41d41d41dp-7, 0x1.cd85689039b0bp-7,
    0x1.c71c71c71c71cp-7, 0x1.c0e070381c0e0p-7, 0x1.bacf914c1bad0p-7, 0x1.b4e81b4e81b4fp-7, 0x1.af286bca1af28p-7, 0x1.a98ef606a63bep-7,
    0x1.a41a41a41a41ap-7, 0x1.9ec8e951033d9p-7, 0x1.999999999999ap-7, 0x1.948b0fcd6e9e0p-7, 0x1.8f9c18f9c18fap-7, 0x1.8acb90f6bf3aap-7,
    0x1.8618618618618p-7, 0x1.8181818181818p-7, 0x1.7d05f417d05f4p-7, 0x1.78a4c8178a4c8p-7, 0x1.745d1745d1746p-7, 0x1.702e05c0b8170p-7,
    0x1.6c16c16c16c17p-7, 0x1.6816816816817p-7, 0x1.642c8590b2164p-7, 0x1.6058160581606p-7, 0x1.5c9882b931057p-7, 0x1.58ed2308158edp-7,
    0x1.5555555555555p-7, 0x1.51d07eae2f815p-7, 0x1.4e5e0a72f0539p-7, 0x1.4afd6a052bf5bp-7, 0x1.47ae147ae147bp-7, 0x1.446f86562d9fbp-7,
    0x1.4141414141414p-7, 0x1.3e22cbce4a902p-7, 0x1.3b13b13b13b14p-7, 0x1.3813813813814p-7, 0x1.3521cfb2b78c1p-7, 0x1.323e34a2b10bfp-7,
    0x1.2f684bda12f68p-7, 0x1.2c9fb4d812ca0p-7, 0x1.29e4129e4129ep-7, 0x1.27350b8812735p-7, 0x1.2492492492492p-7, 0x1.21fb78121fb78p-7,
    0x1.1f7047dc11f70p-7, 0x1.1cf06ada2811dp-7, 0x1.1a7b9611a7b96p-7, 0x1.1811811811812p-7, 0x1.15b1e5f75270dp-7, 0x1.135c81135c811p-7,
    0x1.1111111111111p-7, 0x1.0ecf56be69c90p-7, 0x1.0c9714fbcda3bp-7, 0x1.0a6810a6810a7p-7, 0x1.0842108421084p-7, 0x1.0624dd2f1a9fcp-7,
    0x1.0410410410410p-7, 0x1.0204081020408p-7, 0x1.0000000000000p-7, 0x1.fc07f01fc07f0p-8, 0x1.f81f81f81f820p-8, 0x1.f44659e4a4271p-8,
    0x1.f07c1f07c1f08p-8, 0x1.ecc07b301ecc0p-8, 0x1.e9131abf0b767p-8, 0x1.e573ac901e574p-8, 0x1.e1e1e1e1e1e1ep-8, 0x1.de5d6e3f8868ap-8,
    0x1.dae6076b981dbp-8, 0x1.d77b654b82c34p-8, 0x1.d41d41d41d41dp-8, 0x1.d0cb58f6ec074p-8, 0x1.cd85689039b0bp-8, 0x1.ca4b3055ee191p-8,
    0x1.c71c71c71c71cp-8, 0x1.c3f8f01c3f8f0p-8, 0x1.c0e070381c0e0p-8, 0x1.bdd2b899406f7p-8, 0x1.bacf914c1bad0p-8, 0x1.b7d6c3dda338bp-8,
    0x1.b4e81b4e81b4fp-8, 0x1.b2036406c80d9p-8, 0x1.af286bca1af28p-8, 0x1.ac5701ac5701bp-8, 0x1.a98ef606a63bep-8, 0x1.a6d01a6d01a6dp-8,
    0x1.a41a41a41a41ap-8, 0x1.a16d3f97a4b02p-8, 0x1.9ec8e951033d9p-8, 0x1.9c2d14ee4a102p-8, 0x1.999999999999ap-8, 0x1.970e4f80cb872p-8,
    0x1.948b0fcd6e9e0p-8, 0x1.920fb49d0e229p-8, 0x1.8f9c18f9c18fap-8, 0x1.8d3018d3018d3p-8, 0x1.8acb90f6bf3aap-8, 0x1.886e5f0abb04ap-8,
    0x1.8618618618618p-8, 0x1.83c977ab2beddp-8, 0x1.8181818181818p-8, 0x1.7f405fd017f40p-8, 0x1.7d05f417d05f4p-8, 0x1.7ad2208e0ecc3p-8,
    0x1.78a4c8178a4c8p-8, 0x1.767dce434a9b1p-8, 0x1.745d1745d1746p-8, 0x1.724287f46debcp-8, 0x1.702e05c0b8170p-8, 0x1.6e1f76b4337c7p-8,
    0x1.6c16c16c16c17p-8, 0x1.6a13cd1537290p-8, 0x1.6816816816817p-8, 0x1.661ec6a5122f9p-8, 0x1.642c8590b2164p-8, 0x1.623fa77016240p-8,
    0x1.6058160581606p-8, 0x1.5e75bb8d015e7p-8, 0x1.5c9882b931057p-8, 0x1.5ac056b015ac0p-8, 0x1.58ed2308158edp-8, 0x1.571ed3c506b3ap-8,
    0x1.5555555555555p-8, 0x1.5390948f40febp-8, 0x1.51d07eae2f815p-8, 0x1.5015015015015p-8, 0x1.4e5e0a72f0539p-8, 0x1.4cab88725af6ep-8,
    0x1.4afd6a052bf5bp-8, 0x1.49539e3b2d067p-8, 0x1.47ae147ae147bp-8, 0x1.460cbc7f5cf9ap-8, 0x1.446f86562d9fbp-8, 0x1.42d6625d51f87p-8,
    0x1.4141414141414p-8, 0x1.3fb013fb013fbp-8, 0x1.3e22cbce4a902p-8, 0x1.3c995a47babe7p-8, 0x1.3b13b13b13b14p-8, 0x1.3991c2c187f63p-8,
    0x1.3813813813814p-8, 0x1.3698df3de0748p-8, 0x1.3521cfb2b78c1p-8, 0x1.33ae45b57bcb2p-8, 0x1.323e34a2b10bfp-8, 0x1.30d190130d190p-8,
    0x1.2f684bda12f68p-8, 0x1.2e025c04b8097p-8, 0x1.2c9fb4d812ca0p-8, 0x1.2b404ad012b40p-8, 0x1.29e4129e4129ep-8, 0x1.288b01288b013p-8,
    0x1.27350b8812735p-8, 0x1.25e22708092f1p-8, 0x1.2492492492492p-8, 0x1.23456789abcdfp-8, 0x1.21fb78121fb78p-8, 0x1.20b470c67c0d9p-8,
    0x1.1f7047dc11f70p-8, 0x1.1e2ef3b3fb874p-8, 0x1.1cf06ada2811dp-8, 0x1.1bb4a4046ed29p-8, 0x1.1a7b9611a7b96p-8, 0x1.19453808ca29cp-8,
    0x1.1811811811812p-8, 0x1.16e0689427379p-8, 0x1.15b1e5f75270dp-8, 0x1.1485f0e0acd3bp-8, 0x1.135c81135c811p-8, 0x1.12358e75d3033p-8,
    0x1.1111111111111p-8, 0x1.0fef010fef011p-8, 0x1.0ecf56be69c90p-8, 0x1.0db20a88f4696p-8, 0x1.0c9714fbcda3bp-8, 0x1.0b7e6ec259dc8p-8,
    0x1.0a6810a6810a7p-8, 0x1.0953f39010954p-8, 0x1.0842108421084p-8, 0x1.073260a47f7c6p-8, 0x1.0624dd2f1a9fcp-8, 0x1.05197f7d73404p-8,
    0x1.0410410410410p-8, 0x1.03091b51f5e1ap-8, 0x1.0204081020408p-8, 0x1.0101010101010p-8, 0x1.0000000000000p-8, 0x1.fe01fe01fe020p-9,
    0x1.fc07f01fc07f0p-9, 0x1.fa11caa01fa12p-9, 0x1.f81f81f81f820p-9, 0x1.f6310aca0dbb5p-9, 0x1.f44659e4a4271p-9, 0x1.f25f644230ab5p-9,
    0x1.f07c1f07c1f08p-9, 0x1.ee9c7f8458e02p-9, 0x1.ecc07b301ecc0p-9, 0x1.eae807aba01ebp-9, 0x1.e9131abf0b767p-9, 0x1.e741aa59750e4p-9,
    0x1.e573ac901e574p-9, 0x1.e3a9179dc1a73p-9, 0x1.e1e1e1e1e1e1ep-9, 0x1.e01e01e01e01ep-9, 0x1.de5d6e3f8868ap-9, 0x1.dca01dca01dcap-9,
    0x1.dae6076b981dbp-9, 0x1.d92f2231e7f8ap-9, 0x1.d77b654b82c34p-9, 0x1.d5cac807572b2p-9, 0x1.d41d41d41d41dp-9, 0x1.d272ca3fc5b1ap-9,
    0x1.d0cb58f6ec074p-9, 0x1.cf26e5c44bfc6p-9, 0x1.cd85689039b0bp-9, 0x1.cbe6d9601cbe7p-9, 0x1.ca4b3055ee191p-9, 0x1.c8b265afb8a42p-9,
    0x1.c71c71c71c71cp-9, 0x1.c5894d10d4986p-9, 0x1.c3f8f01c3f8f0p-9, 0x1.c26b5392ea01cp-9, 0x1.c0e070381c0e0p-9, 0x1.bf583ee868d8bp-9,
    0x1.bdd2b899406f7p-9, 0x1.bc4fd65883e7bp-9, 0x1.bacf914c1bad0p-9, 0x1.b951e2b18ff23p-9, 0x1.b7d6c3dda338bp-9, 0x1.b65e2e3beee05p-9,
    0x1.b4e81b4e81b4fp-9, 0x1.b37484ad806cep-9, 0x1.b2036406c80d9p-9, 0x1.b094b31d922a4p-9, 0x1.af286bca1af28p-9, 0x1.adbe87f94905ep-9,
    0x1.ac5701ac5701bp-9, 0x1.aaf1d2f87ebfdp-9, 0x1.a98ef606a63bep-9, 0x1.a82e65130e159p-9, 0x1.a6d01a6d01a6dp-9, 0x1.a574107688a4ap-9,
    0x1.a41a41a41a41ap-9, 0x1.a2c2a87c51ca0p-9, 0x1.a16d3f97a4b02p-9, 0x1.a01a01a01a01ap-9, 0x1.9ec8e951033d9p-9, 0x1.9d79f176b682dp-9,
    0x1.9c2d14ee4a102p-9, 0x1.9ae24ea5510dap-9, 0x1.999999999999ap-9, 0x1.9852f0d8ec0ffp-9, 0x1.970e4f80cb872p-9, 0x1.95cbb0be377aep-9,
    0x1.948b0fcd6e9e0p-9, 0x1.934c67f9b2ce6p-9, 0x1.920fb49d0e229p-9, 0x1.90d4f120190d5p-9, 0x1.8f9c18f9c18fap-9, 0x1.8e6527af1373fp-9,
    0x1.8d3018d3018d3p-9, 0x1.8bfce8062ff3ap-9, 0x1.8acb90f6bf3aap-9, 0x1.899c0f601899cp-9, 0x1.886e5f0abb04ap-9, 0x1.87427bcc092b9p-9,
    0x1.8618618618618p-9, 0x1.84f00c2780614p-9, 0x1.83c977ab2beddp-9, 0x1.82a4a0182a4a0p-9, 0x1.8181818181818p-9, 0x1.8060180601806p-9,
    0x1.7f405fd017f40p-9, 0x1.7e225515a4f1dp-9, 0x1.7d05f417d05f4p-9, 0x1.7beb3922e017cp-9, 0x1.7ad2208e0ecc3p-9, 0x1.79baa6bb6398bp-9,
    0x1.78a4c8178a4c8p-9, 0x1.77908119ac60dp-9, 0x1.767dce434a9b1p-9, 0x1.756cac201756dp-9, 0x1.745d1745d1746p-9, 0x1.734f0c541fe8dp-9,
    0x1.724287f46debcp-9, 0x1.713786d9c7c09p-9, 0x1.702e05c0b8170p-9, 0x1.6f26016f26017p-9, 0x1.6e1f76b4337c7p-9, 0x1.6d1a62681c861p-9,
    0x1.6c16c16c16c17p-9, 0x1.6b1490aa31a3dp-9, 0x1.6a13cd1537290p-9, 0x1.691473a88d0c0p-9, 0x1.6816816816817p-9, 0x1.6719f3601671ap-9,
    0x1.661ec6a5122f9p-9, 0x1.6524f853b4aa3p-9, 0x1.642c8590b2164p-9, 0x1.63356b88ac0dep-9, 0x1.623fa77016240p-9, 0x1.614b36831ae94p-9,
    0x1.6058160581606p-9, 0x1.5f66434292dfcp-9, 0x1.5e75bb8d015e7p-9, 0x1.5d867c3ece2a5p-9, 0x1.5c9882b931057p-9, 0x1.5babcc647fa91p-9,
    0x1.5ac056b015ac0p-9, 0x1.59d61f123ccaap-9, 0x1.58ed2308158edp-9, 0x1.5805601580560p-9, 0x1.571ed3c506b3ap-9, 0x1.56397ba7c52e2p-9,
    0x1.5555555555555p-9, 0x1.54725e6bb82fep-9, 0x1.5390948f40febp-9, 0x1.52aff56a8054bp-9, 0x1.51d07eae2f815p-9, 0x1.50f22e111c4c5p-9,
    0x1.5015015015015p-9, 0x1.4f38f62dd4c9bp-9, 0x1.4e5e0a72f0539p-9, 0x1.4d843bedc2c4cp-9, 0x1.4cab88725af6ep-9, 0x1.4bd3edda68fe1p-9,
    0x1.4afd6a052bf5bp-9, 0x1.4a27fad76014ap-9, 0x1.49539e3b2d067p-9, 0x1.4880522014880p-9, 0x1.47ae147ae147bp-9, 0x1.46dce34596066p-9,
    0x1.460cbc7f5cf9ap-9, 0x1.453d9e2c776cap-9, 0x1.446f86562d9fbp-9, 0x1.43a2730abee4dp-9, 0x1.42d6625d51f87p-9, 0x1.420b5265e5951p-9,
    0x1.4141414141414p-9, 0x1.40782d10e6566p-9, 0x1.3fb013fb013fbp-9, 0x1.3ee8f42a5af07p-9, 0x1.3e22cbce4a902p-9, 0x1.3d5d991aa75c6p-9,
    0x1.3c995a47babe7p-9, 0x1.3bd60d9232955p-9, 0x1.3b13b13b13b14p-9, 0x1.3a524387ac822p-9, 0x1.3991c2c187f63p-9, 0x1.38d22d366088ep-9,
    0x1.3813813813814p-9, 0x1.3755bd1c945eep-9, 0x1.3698df3de0748p-9, 0x1.35dce5f9f2af8p-9, 0x1.3521cfb2b78c1p-9, 0x1.34679ace01346p-9,
    0x1.33ae45b57bcb2p-9, 0x1.32f5ced6a1dfap-9, 0x1.323e34a2b10bfp-9, 0x1.3187758e9ebb6p-9, 0x1.30d190130d190p-9, 0x1.301c82ac40260p-9,
    0x1.2f684bda12f68p-9, 0x1.2eb4ea1fed14bp-9, 0x1.2e025c04b8097p-9, 0x1.2d50a012d50a0p-9, 0x1.2c9fb4d812ca0p-9, 0x1.2bef98e5a3711p-9,
    0x1.2b404ad012b40p-9, 0x1.2a91c92f3c105p-9, 0x1.29e4129e4129ep-9, 0x1.293725bb804a5p-9, 0x1.288b01288b013p-9, 0x1.27dfa38a1ce4dp-9,
    0x1.27350b8812735p-9, 0x1.268b37cd60127p-9, 0x1.25e22708092f1p-9, 0x1.2539d7e9177b2p-9, 0x1.2492492492492p-9, 0x1.23eb79717605bp-9,
    0x1.23456789abcdfp-9, 0x1.22a0122a0122ap-9, 0x1.21fb78121fb78p-9, 0x1.21579804855e6p-9, 0x1.20b470c67c0d9p-9, 0x1.2012012012012p-9,
    0x1.1f7047dc11f70p-9, 0x1.1ecf43c7fb84cp-9, 0x1.1e2ef3b3fb874p-9, 0x1.1d8f5672e4abdp-9, 0x1.1cf06ada2811dp-9, 0x1.1c522fc1ce059p-9,
    0x1.1bb4a4046ed29p-9, 0x1.1b17c67f2bae3p-9, 0x1.1a7b9611a7b96p-9, 0x1.19e0119e0119ep-9, 0x1.19453808ca29cp-9, 0x1.18ab083902bdbp-9,
    0x1.1811811811812p-9, 0x1.1778a191bd684p-9, 0x1.16e0689427379p-9, 0x1.1648d50fc3201p-9, 0x1.15b1e5f75270dp-9, 0x1.151b9a3fdd5c9p-9,
    0x1.1485f0e0acd3bp-9, 0x1.13f0e8d344724p-9, 0x1.135c81135c811p-9, 0x1.12c8b89edc0acp-9, 0x1.12358e75d3033p-9, 0x1.11a3019a74826p-9,
    0x1.1111111111111p-9, 0x1.107fbbe011080p-9, 0x1.0fef010fef011p-9, 0x1.0f5edfab325a2p-9, 0x1.0ecf56be69c90p-9, 0x1.0e40655826011p-9,
    0x1.0db20a88f4696p-9, 0x1.0d24456359e3ap-9, 0x1.0c9714fbcda3bp-9, 0x1.0c0a7868b4171p-9, 0x1.0b7e6ec259dc8p-9, 0x1.0af2f722eecb5p-9,
    0x1.0a6810a6810a7p-9, 0x1.09ddba6af8360p-9, 0x1.0953f39010954p-9, 0x1.08cabb37565e2p-9, 0x1.0842108421084p-9, 0x1.07b9f29b8eae2p-9,
    0x1.073260a47f7c6p-9, 0x1.06ab59c7912fbp-9, 0x1.0624dd2f1a9fcp-9, 0x1.059eea0727586p-9, 0x1.05197f7d73404p-9, 0x1.04949cc1664c5p-9,
    0x1.0410410410410p-9, 0x1.038c6b78247fcp-9, 0x1.03091b51f5e1ap-9, 0x1.02864fc7729e9p-9, 0x1.0204081020408p-9, 0x1.0182436517a37p-9,
    0x1.0101010101010p-9, 0x1.0080402010080p-9, 0x1.0000000000000p-9,
};
#endif
__host__ __device__ inline double pns_inv_k(int k) {   // 1 <= k <= 512
#ifdef __CUDA_ARCH__
    return __ldg(kInvK + k);
#else
    return 1.0 / (double)k;
#endif
}

// Binomial(m, pp) by CDF inversion, m <= 512, 0 < pp <= 0.5, u in [0, 1).  The pmf recurrence is
// pmf(k) = pmf(k-1) * f_k with f_k = (ratio * (m-k+1)) * (1/k), the factor formed off the critical
// path: the loop-carried chain is one multiply, one subtract and a compare.  Four steps per trip,
// because an in-order warp cannot overlap loop trips by itself (restated in oracle/philox.py).
__host__ __device__ inline int binomial_inversion(int m, double pp, double u) {
    const double q = 1.0 - pp;
    const double ratio = pp / q;
    double pk = 1.0, b = q;
    for (int e = m; e; e >>= 1) {
        if (e & 1) pk = pk * b;
        b = b * b;
    }
    int k = 0;
    double mk = (double)m;                                   // m - k, exact
    while (u > pk && k < m) {
        const double f1 = (ratio * mk) * pns_inv_k(k + 1);
        const double f2 = (ratio * (mk - 1.0)) * pns_inv_k(k + 2 <= 512 ? k + 2 : 512);
        const double f3 = (ratio * (mk - 2.0)) * pns_inv_k(k + 3 <= 512 ? k + 3 : 512);
        const double f4 = (ratio * (mk - 3.0)) * pns_inv_k(k + 4 <= 512 ? k + 4 : 512);
        u = u - pk; k += 1; pk = pk * f1;
        if (!(u > pk && k < m)) break;
        u = u - pk; k += 1; pk = pk * f2;
        if (!(u > pk && k < m)) break;
        u = u - pk; k += 1; pk = pk * f3;
        if (!(u > pk && k < m)) break;
        u = u - pk; k += 1; pk = pk * f4;
        mk = mk - 4.0;
    }
    return k;
}

// Exact Binomial(n, p), 0 < p < 1, n > 0; chunks of 512 trials keep q^m representable (binomial
// additivity).  Out of line and with scalar arguments: it is called on a minority of the links and
// must not bloat the callers' register footprint.  (A rejection sampler for large means -- BTRS,
// 1.2 rounds instead of n*min(p,1-p) inversion steps -- was tried: its register needs made ptxas
// spill 170 bytes per thread on the link kernels' main path, and behind a separately compiled ABI
// boundary the call overhead cost as much as it saved; see DESIGN.md.)
#ifdef __CUDACC__
__device__ __noinline__
#else
inline
#endif
int binomial_philox_core(uint32_t t, uint32_t link, uint32_t replica, uint32_t k0, uint32_t k1, uint32_t site,
                         int n, double p) {
    const bool flip = p > 0.5;
    const double pp = flip ? 1.0 - p : p;
    int total = 0, left = n;
    uint32_t chunk = 0;
    while (left > 0) {
        const Philox4 w = philox4x32_10(t, link, site | ((chunk >> 1) << 8), replica, k0, k1);
        for (int h = 0; h < 2 && left > 0; ++h) {
            const int m = left < 512 ? left : 512;
            total += binomial_inversion(m, pp, u53(w.v[2 * h], w.v[2 * h + 1]));
            left -= m;
            chunk += 1;
        }
    }
    return flip ? n - total : total;
}

#ifdef __CUDACC__
__device__ __forceinline__
#else
inline
#endif
int binomial_philox(const DrawKey& key, uint32_t site, int n, double p) {
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    return binomial_philox_core(key.t, key.link, key.replica, key.k0, key.k1, site, n, p);
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
