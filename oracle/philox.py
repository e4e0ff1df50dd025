"""TEST INFRASTRUCTURE ONLY -- Python restatement of the device-side counter-based samplers.

Mirrors pednstream_b200/csrc/pns_rng.cuh operation for operation in IEEE doubles (Python floats
never contract multiply-adds), so device samples can be checked bit for bit:
Philox4x32-10 (Salmon et al., SC'11), 53-bit uniforms, fdlibm-style log/exp/sin/cos kernels,
CDF-inversion binomial (from zero for small means, outward from the mode for large ones), Box-Muller normal.  There is no reference-side counterpart: the
reference draws from numpy's sequential global stream (SURVEY.md "RNG ledger"); this mode replaces
the generator, not the distributions.
"""
from __future__ import annotations

import math
import struct
from fractions import Fraction

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


def _fma_exact(a, b, c):
    """Correctly rounded a*b + c (what __fma_rn / fma() give): exact through rationals."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _load_libm_fma():
    """C99 fma() of the host's libm (correctly rounded by definition) when it can be loaded and agrees with the
    exact evaluation on a few probes; it is ~50x faster than going through rationals."""
    try:
        import ctypes
        import ctypes.util
        lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        f = lib.fma
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double] * 3
        probes = [(0.1, 0.3, -0.03), (1.0 + 2.0 ** -52, 1.0 - 2.0 ** -53, -1.0), (3.7e10, 1.0 / 1153.0, -0.1111111111111111)]
        if all(f(*p) == _fma_exact(*p) for p in probes):
            return f
    except Exception:
        pass
    return _fma_exact


fma = _load_libm_fma()


INV_TABLE = 2048          # pns_rng.cuh kInvK: correctly rounded reciprocals of 1 .. 2048


def inv_k(k):
    return 1.0 / float(k)          # the table holds exactly these values; beyond it the kernel divides


def binomial_from_zero(n, pp, u):
    """CDF inversion from 0 (pns_rng.cuh binomial_from_zero): pmf(0) = q^n by squaring, then
    pmf(k) = pmf(k-1) * (A/k - ratio) with A = ratio*(n+1), the factor formed by one fused multiply-add."""
    q = 1.0 - pp
    ratio = pp / q
    A = ratio * float(n + 1)
    pk, b, e = 1.0, q, n
    while e:
        if e & 1:
            pk = pk * b
        b = b * b
        e >>= 1
    k = 0
    while u > pk and k < n:
        u = u - pk
        k += 1
        pk = pk * fma(A, inv_k(k), -ratio)
    return k


SFE = tuple(float.fromhex(h) for h in (
    "0x0.0p+0", "0x1.4c071bcda0a5bp-4", "0x1.52a9b923ea649p-5", "0x1.c579a268d80b3p-6", "0x1.54a2662fd78a9p-6",
    "0x1.10b4e513fcbedp-6", "0x1.c6b167bebdf36p-7", "0x1.85d4d612e4a86p-7", "0x1.552805e7b3076p-7",
    "0x1.2f4871b12ab64p-7", "0x1.10f9d4c0743a7p-7", "0x1.f0593088014f8p-8", "0x1.c7018733aa9c6p-8",
    "0x1.a40514700f36cp-8", "0x1.86076c002d4a7p-8", "0x1.6c08f6f194a10p-8"))


def stirlerr(x):
    """log(x!) - log(sqrt(2 pi x) (x/e)^x) for an integer-valued x >= 0 (Loader 2000): table below 16, the
    asymptotic series above."""
    if x < 16.0:
        return SFE[int(x)]
    r = 1.0 / x
    rr = r * r
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) * rr) * rr) * rr) * rr) * r


def bd0(x, np_):
    """Deviance term x*log(x/np) + np - x by its series in v = (x-np)/(x+np) (|x - np| << x + np at the mode)."""
    d = x - np_
    v = d / (x + np_)
    s = d * v
    if abs(s) < 2.2250738585072014e-308:
        return s
    ej = (2.0 * x) * v
    v = v * v
    for j in range(1, 64):
        ej = ej * v
        s1 = s + ej * (1.0 / float(2 * j + 1))
        if s1 == s:
            return s1
        s = s1
    return s


def binomial_pmf_mode(n, m, pp, q):
    """pmf of Binomial(n, pp) at m (its mode; 0 < m < n), saddle-point form (Loader 2000)."""
    nd, md, kd = float(n), float(m), float(n - m)
    lc = (((stirlerr(nd) - stirlerr(md)) - stirlerr(kd)) - bd0(md, nd * pp)) - bd0(kd, nd * q)
    return det_exp(lc) * math.sqrt(nd / ((6.283185307179586 * md) * kd))


PP09 = 1.0 - 0.9               # the flipped probability of the blockers draw (link.py:382)
MODE_MEAN_TABULATED = 12.0     # mean from which the search starts at the mode: blockers draw (mode pmf tabulated)
MODE_MEAN_GENERIC = 48.0       # ... any other probability (mode pmf evaluated per draw)
MODE_TABLE_N = 4096


def binomial_from_mode(n, pp, u):
    """Search outward from the mode m = floor((n+1) pp): m, then the pairs {m+1, m-1}, {m+2, m-2}, ... subtracting the pmf values from u
    (pns_rng.cuh binomial_from_mode).  Expected work ~ 1.6 standard deviations instead of the mean."""
    q = 1.0 - pp
    ratio = pp / q
    iratio = q / pp
    m = int(float(n + 1) * pp)
    pm = binomial_pmf_mode(n, m, pp, q)
    if not u > pm:
        return m
    u = u - pm
    A = ratio * float(n + 1)
    B = iratio * float(n + 1)
    pu = pd = pm
    paired = min(m, n - m)
    for i in range(1, paired + 1):           # candidates m+i and m-i tested as a pair
        pu = pu * fma(A, inv_k(m + i), -ratio)
        pd = pd * fma(B, inv_k(n - m + i), -iratio)
        both = pu + pd
        if not u > both:
            return m + i if not u > pu else m - i
        u = u - both
    ku, kd = m + paired, m - paired
    while True:                              # far tail: one side is exhausted
        if ku < n:
            ku += 1
            pu = pu * fma(A, inv_k(ku), -ratio)
            if not u > pu:
                return ku
            u = u - pu
        else:
            pu = 0.0
        if kd > 0:
            pd = pd * fma(B, inv_k(n - kd + 1), -iratio)
            kd -= 1
            if not u > pd:
                return kd
            u = u - pd
        else:
            pd = 0.0
        if not pu > 0.0 and not pd > 0.0:
            return m               # u fell into the rounding leftover of the total mass


def binomial_u(n, p, u):
    """Exact Binomial(n, p) from one uniform (pns_rng.cuh binomial_u / binomial_core; binomial09_u is the same
    algorithm with the constants of p = 0.9 folded and 0.9^n, the mode pmf tabulated)."""
    n = int(n)
    p = float(p)
    if n <= 0 or not p > 0.0:
        return 0
    if p >= 1.0:
        return n
    flip = p > 0.5
    pp = 1.0 - p if flip else p
    mean = float(n) * pp
    tabulated = pp == PP09 and n <= MODE_TABLE_N
    if mean >= (MODE_MEAN_TABULATED if tabulated else MODE_MEAN_GENERIC):
        k = binomial_from_mode(n, pp, u)
    else:
        k = binomial_from_zero(n, pp, u)
    return n - k if flip else k


def binomial_philox(seed, t, link, replica, site, n, p):
    """A draw with a Philox block of its own (site in the counter): the activity draw R2, test hooks."""
    w = philox4x32_10(t, link, site, replica, seed & M32, (seed >> 32) & M32)
    return binomial_u(n, p, u53(w[0], w[1]))


def link_draws(seed, t, link, replica):
    """(u1, u3): the uniforms of a link's release (R1) and blockers (R3) draws of step t -- one block, site 1."""
    w = philox4x32_10(t, link, 1, replica, seed & M32, (seed >> 32) & M32)
    return u53(w[0], w[1]), u53(w[2], w[3])


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


def device_scenario(seed, replica, n_corridors, n_change, base_params, n_od, origin_rows):
    """Restatement of pns_kernels.cu k_scenario_draw for one replica (global replica index `replica`):
    base_params[c] = (k_critical, k_jam, free_flow_speed) of corridor c; origin_rows = indices of the demand rows
    that are origins.  Returns (corridor overrides {c: (kc, kj, vf)} in draw order, OD weights [n_od],
    demand parameters {row: (pattern code, base, peak)})."""
    k0, k1 = seed & M32, (seed >> 32) & M32
    order = list(range(n_corridors))
    over = {}
    for k in range(n_change):
        w = philox4x32_10(k, 0, 10, replica, k0, k1)
        j = min(k + int(u53(w[0], w[1]) * float(n_corridors - k)), n_corridors - 1)
        order[k], order[j] = order[j], order[k]
        c = order[k]
        a = philox4x32_10(k, 0, 11, replica, k0, k1)
        b = philox4x32_10(k, 0, 12, replica, k0, k1)
        kc, kj, vf = (float(x) for x in base_params[c])
        if a[0] & 1:
            f = 0.6 + 0.6 * u53(a[2], a[3])
            kc_new = max(0.5, kc * f)
            kj = max(kc_new * 2.0, kj * f)
            kc = kc_new
        if a[1] & 1:
            vf = vf * (0.6 + 0.3 * u53(b[0], b[1]))
        over[c] = (kc, kj, vf)
    weights = []
    for od in range(n_od):
        w = philox4x32_10(od, 0, 13, replica, k0, k1)
        weights.append(1.0 + 9.0 * u53(w[0], w[1]))
    demand = {}
    for row in origin_rows:
        w = philox4x32_10(row, 0, 14, replica, k0, k1)
        v = philox4x32_10(row, 0, 15, replica, k0, k1)
        base = 2.0 + 8.0 * u53(w[1], w[2])
        peak = max(10.0 + 20.0 * u53(v[0], v[1]), base + 5.0)
        demand[row] = (w[0] % 3, base, peak)
    return over, weights, demand


# ---- float32 Box-Muller with exact fused multiply-adds (mirrors pns_rng.cuh) -----------------------

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
        # the kernels key the draw by the step being computed (= time index + 1); the release and blockers draws of
        # a link share one block, the activity draw has its own
        if site == "R2":
            return binomial_philox(self.seed, t + 1, link.col, self.replica, 2, int(n), float(p))
        u1, u3 = link_draws(self.seed, t + 1, link.col, self.replica)
        return binomial_u(int(n), float(p), u1 if site == "R1" else u3)

    def normal(self, link, t, sigma):
        # the two directions of a corridor share one block keyed by the even link of the pair
        pair = normal_pair_philox(self.seed, t, link.col & ~1, self.replica, 4)
        return sigma * pair[link.col & 1]
