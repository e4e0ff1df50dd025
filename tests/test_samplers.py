"""Counter-based samplers of the on-device draw mode: the Python restatement (oracle/philox.py) is
checked against the exact distributions, and the kernel source (host-emulation build) against the
restatement.  The GPU build is checked against the same restatement in test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import philox as ph


def _selftest(lib, kind, site, trials, p, seed, t):
    n = len(trials)
    tr = np.ascontiguousarray(trials, dtype=np.int32)
    pp = np.ascontiguousarray(p, dtype=np.float64)
    oi = np.zeros(n, dtype=np.int32)
    od = np.zeros(4 * n, dtype=np.float64)
    rc = lib.pns_rng_selftest(kind, n, tr.ctypes.data_as(C.c_void_p), pp.ctypes.data_as(C.c_void_p),
                              C.c_uint64(seed), t, site, oi.ctypes.data_as(C.c_void_p),
                              od.ctypes.data_as(C.c_void_p), None)
    assert rc == 0
    return oi, od


def test_emulated_samplers_match_python_restatement(emu_lib):
    n = 600
    rng = np.random.default_rng(11)
    trials = rng.integers(0, 1400, n).astype(np.int32)
    trials[:40] = rng.integers(0, 12, 40)
    p = rng.uniform(0.01, 0.99, n)
    p[100:160] = 0.9                                  # R3 blockers (link.py:382)
    trials[100:130] = 1200                            # a jammed link: three chunks of inversion
    seed, t = 0xFEEDFACE12345678, 41
    for site in (1, 3):
        oi, _ = _selftest(emu_lib, 0, site, trials, p, seed, t)
        want = [ph.binomial_philox(seed, t, i, 0, site, int(trials[i]), float(p[i])) for i in range(n)]
        assert oi.tolist() == want
    # a link's release (any p) and blockers (p = 0.9, tabulated fast path) draws share one Philox block
    for site in (1, 3):
        oi, _ = _selftest(emu_lib, 3, site, trials, p, seed, t)
        want = [ph.binomial_u(int(trials[i]), 0.9 if site == 3 else float(p[i]), ph.link_draws(seed, t, i, 0)[site // 2])
                for i in range(n)]
        assert oi.tolist() == want
    _, od = _selftest(emu_lib, 1, 4, trials, p, seed, t)
    want = np.array([ph.normal_quad_philox(seed, t, i, 0, 4) for i in range(n)]).reshape(-1)
    assert np.array_equal(od, want)
    # a corridor's normals are its half of the quad's block: links 4q, 4q+1 | 4q+2, 4q+3
    for link in (0, 2, 4, 6, 10):
        quad = ph.normal_quad_philox(seed, t, link & ~3, 0, 4)
        assert ph.normal_pair_philox(seed, t, link, 0, 4) == (quad[2:] if link & 2 else quad[:2])


@pytest.mark.parametrize("n,p", [(1200, 0.9), (130, 0.9), (88, 0.775), (40, 0.25), (300, 0.5), (2000, 0.02), (9, 0.3),
                                 (5000, 0.6)])
def test_binomial_sampler_follows_the_binomial_law(n, p):
    """CDF inversion (from zero / outward from the mode, with the p > 0.5 flip) against scipy's pmf: moments within 5 standard
    errors and a chi-square test over pooled cells."""
    from scipy import stats
    N = 6000
    x = np.array([ph.binomial_philox(1234, 7, i, 0, 3, n, p) for i in range(N)])
    assert x.min() >= 0 and x.max() <= n
    mean, var = n * p, n * p * (1 - p)
    assert abs(x.mean() - mean) < 5 * np.sqrt(var / N)
    assert abs(x.var() - var) < 5 * var * np.sqrt(2.0 / N) + 0.05 * var
    lo, hi = int(stats.binom.ppf(0.001, n, p)), int(stats.binom.ppf(0.999, n, p))
    edges = np.arange(lo, hi + 2)
    obs = np.histogram(np.clip(x, lo, hi), bins=edges)[0].astype(float)
    exp = stats.binom.pmf(np.arange(lo, hi + 1), n, p)
    exp[0] += stats.binom.cdf(lo - 1, n, p)
    exp[-1] += stats.binom.sf(hi, n, p)
    exp *= N
    # pool cells with small expectation
    o2, e2, ao, ae = [], [], 0.0, 0.0
    for o, e in zip(obs, exp):
        ao += o; ae += e
        if ae >= 8:
            o2.append(ao); e2.append(ae); ao = ae = 0.0
    if ae > 0:
        o2[-1] += ao; e2[-1] += ae
    chi2 = float(((np.array(o2) - np.array(e2)) ** 2 / np.array(e2)).sum())
    assert chi2 < stats.chi2.ppf(1 - 1e-4, len(o2) - 1), (chi2, len(o2))


def test_speed_noise_sampler_is_standard_normal():
    from scipy import stats
    g = np.array([ph.normal_quad_philox(99, 3, 4 * i, 0, 4) for i in range(3000)])
    assert g.shape == (3000, 4)
    flat = g.reshape(-1)
    assert abs(flat.mean()) < 5 / np.sqrt(flat.size) and abs(flat.var() - 1) < 0.06
    assert stats.kstest(flat, "norm").pvalue > 1e-4
    c = np.corrcoef(g.T)
    assert np.abs(c - np.eye(4)).max() < 0.08


def _draw_demand(lib, S, rows, R, replica_base, seed, base, peak, pattern):
    t = np.arange(S)
    spread = 2 * (S / 20) ** 2
    b1 = np.ascontiguousarray(np.exp(-(t - S / 4) ** 2 / spread))
    b2 = np.ascontiguousarray(np.exp(-(t - 3 * S / 4) ** 2 / spread))
    base = np.ascontiguousarray(base, dtype=np.float64).reshape(-1)
    peak = np.ascontiguousarray(peak, dtype=np.float64).reshape(-1)
    pattern = np.ascontiguousarray(pattern, dtype=np.int32).reshape(-1)
    out = np.zeros((S + 1, rows * R))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.pns_env_draw_demand(S, rows, R, replica_base, C.c_uint64(seed), ptr(b1), ptr(b2), ptr(base), ptr(peak),
                                 ptr(pattern), ptr(out), None)
    assert rc == 0
    return out.reshape(S + 1, rows, R), b1, b2


def test_device_demand_draws(emu_lib):
    """pns_env_draw_demand (kernel source, host build): equal to the Python restatement, Poisson around the
    two-peak rate, constant and sudden-demand variants, and independent of the sharding of replicas."""
    from scipy import stats
    S, rows, R, seed = 120, 3, 16, 0xABCDEF0123
    base = np.tile(np.array([[5.0], [8.0], [3.0]]), (1, R))
    peak = np.tile(np.array([[20.0], [12.0], [30.0]]), (1, R))
    pattern = np.tile(np.array([[0], [1], [2]], dtype=np.int32), (1, R))
    pattern[0, 5] = -1
    d, b1, b2 = _draw_demand(emu_lib, S, rows, R, 0, seed, base, peak, pattern)
    # restatement, value by value
    for r in (0, 7, 15):
        for t in (0, 30, 31, 90, 119):
            lam = (base[0, r] + peak[0, r] * b1[t]) + peak[0, r] * b2[t]
            assert d[t, 0, r] == ph.poisson_philox(seed, t, 0, r, lam)
        start, period, height = ph.sudden_burst_philox(seed, 2, r, S)
        for t in range(S):
            lam = (base[2, r] + peak[2, r] * b1[t]) + peak[2, r] * b2[t]
            want = ph.poisson_philox(seed, t, 2, r, lam) + (height if start <= t < start + period else 0)
            assert d[t, 2, r] == want
    assert (d[:, 1, :] == 8.0).all()                       # constant: all S+1 entries (od_manager.py:106-109)
    assert (d[:, 0, 5] == 0).all() and (d[S, 0, :] == 0).all() and (d[S, 2, :] == 0).all()
    # sharding: replicas 8..15 drawn as a second shard are the same numbers
    d2, _, _ = _draw_demand(emu_lib, S, rows, 8, 8, seed, base[:, 8:], peak[:, 8:], pattern[:, 8:])
    assert np.array_equal(d2, d[:, :, 8:])
    # Poisson law at a fixed rate: many replicas, constant lam (peak 0)
    Rn = 4000
    flat, _, _ = _draw_demand(emu_lib, 4, 1, Rn, 0, 99, np.full((1, Rn), 37.5), np.zeros((1, Rn)),
                              np.zeros((1, Rn), dtype=np.int32))
    x = flat[:4].reshape(-1)
    assert abs(x.mean() - 37.5) < 5 * np.sqrt(37.5 / x.size) and abs(x.var() - 37.5) < 0.1 * 37.5
    lo, hi = int(stats.poisson.ppf(0.001, 37.5)), int(stats.poisson.ppf(0.999, 37.5))
    obs = np.histogram(np.clip(x, lo, hi), bins=np.arange(lo, hi + 2))[0].astype(float)
    exp = stats.poisson.pmf(np.arange(lo, hi + 1), 37.5)
    exp[0] += stats.poisson.cdf(lo - 1, 37.5); exp[-1] += stats.poisson.sf(hi, 37.5)
    exp *= x.size
    keep = exp >= 8
    chi2 = float(((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum())
    assert chi2 < stats.chi2.ppf(1 - 1e-4, int(keep.sum()) - 1), chi2
