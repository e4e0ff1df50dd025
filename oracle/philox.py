"""TEST INFRASTRUCTURE ONLY -- Python restatement of the device-side counter-based samplers.

Mirrors pednstream_b200/csrc/pns_rng.cuh operation for operation in IEEE doubles (Python floats
never contract multiply-adds), so device samples can be checked bit for bit:
Philox4x32-10 (Salmon et al., SC'11), 53-bit uniforms, fdlibm-style log/exp/sin/cos kernels,
chunked CDF-inversion binomial, Box-Muller normal.  There is no reference-side counterpart: the
reference draws from numpy's sequential global stream (SURVEY.md "RNG ledger"); this mode replaces
the generator, not the distributions.
"""
from __future__ import annotations

import math
import struct

import numpy as np

M32 = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [c0 & M32, c1 & M32, c2 & M32, c3 & M32]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def u53(hi, lo):
    return (float(hi >> 5) * 67108864.0 + float(lo >> 6)) / 9007199254740992.0


def det_log(x):
    Lg = (6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
          2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
          1.479819860511658591e-01)
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    bits = struct.unpack("<Q", struct.pack("<d", x))[0]
    e = ((bits >> 52) & 0x7FF) - 1023
    m = struct.unpack("<d", struct.pack("<Q", (bits & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000))[0]
    if m > 1.4142135623730951:
        m = m * 0.5
        e += 1
    f = m - 1.0
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]))
    t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])))
    R = t2 + t1
    hfsq = (0.5 * f) * f
    dk = float(e)
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f)


def det_exp(x):
    P = (1.66666666666666019037e-01, -2.77777777770155933842e-03, 6.61375632143793436117e-05,
         -1.65339022054652515390e-06, 4.13813679705723846039e-08)
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    invln2 = 1.44269504088896338700e+00
    if x < -700.0:
        return 0.0
    if x > 700.0:
        x = 700.0
    k = int(invln2 * x + (-0.5 if x < 0.0 else 0.5))          # truncation toward zero, as the C cast
    dk = float(k)
    hi = x - dk * ln2_hi
    lo = dk * ln2_lo
    r = hi - lo
    t = r * r
    c = r - t * (P[0] + t * (P[1] + t * (P[2] + t * (P[3] + t * P[4]))))
    y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi)
    return y * struct.unpack("<d", struct.pack("<Q", (k + 1023) << 52))[0]


def det_sin_k(x):
    S = (-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
         2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10)
    z = x * x
    v = z * x
    r = S[1] + z * (S[2] + z * (S[3] + z * (S[4] + z * S[5])))
    return x + v * (S[0] + z * r)


def det_cos_k(x):
    Cc = (4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
          -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11)
    z = x * x
    r = z * (Cc[0] + z * (Cc[1] + z * (Cc[2] + z * (Cc[3] + z * (Cc[4] + z * Cc[5])))))
    return (1.0 - 0.5 * z) + z * r


def det_sincos2pi(u):
    """(cos, sin)(2*pi*u)."""
    half_pi = 1.5707963267948966
    u4 = u * 4.0
    q = int(u4)
    f = u4 - float(q)
    if f <= 0.5:
        a = f * half_pi
        c, s = det_cos_k(a), det_sin_k(a)
    else:
        a = (1.0 - f) * half_pi
        c, s = det_sin_k(a), det_cos_k(a)
    return ((c, s), (-s, c), (-c, -s), (s, -c))[q]


def det_cos2pi(u):
    half_pi = 1.5707963267948966
    u4 = u * 4.0
    q = int(u4)
    f = u4 - float(q)
    if f <= 0.5:
        a = f * half_pi
        c, s = det_cos_k(a), det_sin_k(a)
    else:
        a = (1.0 - f) * half_pi
        c, s = det_sin_k(a), det_cos_k(a)
    return (c, -s, -c, s)[q]


def det_pow08(x32):
    """float32 -> float32, x**0.8 on [0, 1]."""
    x = float(x32)
    if not x > 0.0:
        return np.float32(0.0)
    if x >= 1.0:
        return np.float32(1.0)
    return np.float32(det_exp(float(np.float32(0.8)) * det_log(x)))   # numpy demotes the exponent to float32


def binomial_inversion(m, pp, u):
    q = 1.0 - pp
    ratio = pp / q
    pk, b, e = 1.0, q, m
    while e:
        if e & 1:
            pk = pk * b
        b = b * b
        e >>= 1
    k = 0
    while u > pk and k < m:
        u = u - pk
        k += 1
        pk = pk * ((ratio * float(m - k + 1)) * (1.0 / float(k)))
    return k


def binomial_philox(seed, t, link, replica, site, n, p):
    n = int(n)
    p = float(p)
    if n <= 0 or not p > 0.0:
        return 0
    if p >= 1.0:
        return n
    flip = p > 0.5
    pp = 1.0 - p if flip else p
    k0, k1 = seed & M32, (seed >> 32) & M32
    total, left, chunk = 0, n, 0
    while left > 0:
        w = philox4x32_10(t, link, site | ((chunk >> 1) << 8), replica, k0, k1)
        for h in range(2):
            if left <= 0:
                break
            m = min(left, 512)
            total += binomial_inversion(m, pp, u53(w[2 * h], w[2 * h + 1]))
            left -= m
            chunk += 1
    return n - total if flip else total


def poisson_philox(seed, t, row, replica, lam):
    """Demand draw of the batched environment (pns_kernels.cu k_demand_draw): Poisson(lam) by inversion from 0,
    uniform from words 0,1 of the block keyed (t, row, site 8, replica)."""
    w = philox4x32_10(t, row, 8, replica, seed & M32, (seed >> 32) & M32)
    u = u53(w[0], w[1])
    if not lam > 0.0:
        return 0
    pk, k = det_exp(-lam), 0
    while u > pk and k < 4096:
        u = u - pk
        k += 1
        pk = (pk * lam) / float(k)
    return k


def sudden_burst_philox(seed, row, replica, steps):
    """(start, period, height) of a sudden-demand burst (k_demand_draw, site 9)."""
    b = philox4x32_10(0, row, 9, replica, seed & M32, (seed >> 32) & M32)
    period = 10 + b[0] % 10
    span = max(1, steps - period)
    return b[1] % span, period, 20 + b[2] % 30


def normal_philox(seed, t, link, replica, site):
    w = philox4x32_10(t, link, site, replica, seed & M32, (seed >> 32) & M32)
    u1 = 1.0 - u53(w[0], w[1])
    u2 = u53(w[2], w[3])
    return math.sqrt(-2.0 * det_log(u1)) * det_cos2pi(u2)


# ---- float32 Box-Muller with exact fused multiply-adds (mirrors pns_rng.cuh) -----------------------
from fractions import Fraction

F32 = np.float32


def _round_f32(fr: Fraction) -> np.float32:
    """Correctly rounded (nearest, ties to even) float32 of an exact rational."""
    if fr == 0:
        return F32(0.0)
    guess = F32(float(fr))                       # may be off by one ulp through double rounding
    cands = {guess, np.nextafter(guess, F32(np.inf)), np.nextafter(guess, F32(-np.inf))}
    best = None
    for c in cands:
        if not np.isfinite(c):
            continue
        err = abs(Fraction(float(c)) - fr)
        key = (err, int(np.float32(c).view(np.uint32)) & 1)      # ties -> even mantissa
        if best is None or key < best[0]:
            best = (key, c)
    return F32(best[1])


def fmaf(a, b, c) -> np.float32:
    return _round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def det_logf(x):
    x = F32(x)
    bits = int(x.view(np.uint32))
    e = ((bits >> 23) & 0xFF) - 127
    m = np.uint32((bits & 0x007FFFFF) | 0x3F800000).view(np.float32)
    if m > F32(1.41421356):
        m = m * F32(0.5)
        e += 1
    f = m - F32(1.0)
    z = f * f
    y = F32(7.0376836292e-2)
    for coef in (-1.1514610310e-1, 1.1676998740e-1, -1.2420140846e-1, 1.4249322787e-1, -1.6668057665e-1,
                 2.0000714765e-1, -2.4999993993e-1, 3.3333331174e-1):
        y = fmaf(y, f, F32(coef))
    y = (y * f) * z
    fe = F32(e)
    y = fmaf(F32(-2.12194440e-4), fe, y)
    y = fmaf(F32(-0.5), z, y)
    return fmaf(F32(0.693359375), fe, f + y)


def det_sincos2pif(u):
    u = F32(u)
    u4 = u * F32(4.0)
    q = int(u4)
    f = u4 - F32(q)
    lo = f <= F32(0.5)
    a = (f if lo else F32(1.0) - f) * F32(1.57079632679)
    z = a * a
    ps = fmaf(F32(-1.9515295891e-4), z, F32(8.3321608736e-3))
    ps = fmaf(ps, z, F32(-1.6666654611e-1))
    sn = fmaf(ps * z, a, a)
    pc = fmaf(F32(2.443315711809948e-5), z, F32(-1.388731625493765e-3))
    pc = fmaf(pc, z, F32(4.166664568298827e-2))
    cs = fmaf(pc * z, z, fmaf(F32(-0.5), z, F32(1.0)))
    c, s = (cs, sn) if lo else (sn, cs)
    return ((c, s), (-s, c), (-c, -s), (s, -c))[q]


def box_muller_f32(wa, wb):
    """Box-Muller on two 32-bit words in float32: (cosine branch, sine branch)."""
    u1 = F32((wa >> 8) + 1) * F32(5.9604644775390625e-8)
    u2 = F32(wb >> 8) * F32(5.9604644775390625e-8)
    rad = np.sqrt(F32(-2.0) * det_logf(u1))
    cs, sn = det_sincos2pif(u2)
    return float(rad * cs), float(rad * sn)


def normal_quad_philox(seed, t, quad_link, replica, site):
    """Four standard normals from the block keyed by `quad_link`: words 0,1 -> first corridor of the
    quad (cosine, sine branch), words 2,3 -> second corridor (pns_rng.cuh normal_quad_philox)."""
    w = philox4x32_10(t, quad_link, site, replica, seed & M32, (seed >> 32) & M32)
    return box_muller_f32(w[0], w[1]) + box_muller_f32(w[2], w[3])


def normal_pair_philox(seed, t, link, replica, site):
    """The two normals of the corridor whose even link is `link`: its half of the quad's block."""
    g = normal_quad_philox(seed, t, link & ~3, replica, site)
    return (g[2], g[3]) if link & 2 else (g[0], g[1])


class PhiloxDraws:
    """Draw provider for oracle.ltm_oracle.LtmOracle matching the kernels' PHILOX mode."""

    def __init__(self, seed=0, replica=0):
        self.seed, self.replica = int(seed), int(replica)

    _SITE = {"R1": 1, "R2": 2, "R3": 3}

    def release_prob(self, rel32):
        return np.float32(0.7) + np.float32(0.15) * det_pow08(rel32)

    def exp(self, values):
        """Logit exponentials of the route-choice kernel in PHILOX mode (det_exp, not numpy's exp)."""
        return np.array([det_exp(float(v)) for v in values], dtype=np.float64)

    def binomial(self, site, link, t, n, p):
        # the kernels key the draw by the step being computed (= time index + 1)
        return binomial_philox(self.seed, t + 1, link.col, self.replica, self._SITE[site], int(n), float(p))

    def normal(self, link, t, sigma):
        # the two directions of a corridor share one block keyed by the even link of the pair
        pair = normal_pair_philox(self.seed, t, link.col & ~1, self.replica, 4)
        return sigma * pair[link.col & 1]
