"""GPU, tier T2 (statistical): Philox4x32-10 + Box-Muller AWGN (src/noise.rs).  The RNG stream
necessarily differs from the reference's ChaCha20 + ziggurat, so moments, whiteness and
independence are checked with the 4-sigma bounds of SURVEY App. B at N = 2^24, the Philox block
function against the Random123 known answers, and the stream against the oracle's restatement of
the same counter layout."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import load

pytestmark = pytest.mark.gpu
G = load()
N = 1 << 24


@pytest.mark.parametrize("case", G["philox_kat"], ids=lambda c: str(c["ctr"][0]))
def test_philox_kat(ae, case):
    assert ae.noise.philox4x32_10(case["ctr"], case["key"]) == case["want"]


def moments(x, sigma):
    n = x.size
    x = x.astype(np.float64)
    m = x.mean()
    v = x.var()
    z = (x - m) / np.sqrt(v)
    assert abs(m) < 4 * sigma / np.sqrt(n)
    assert abs(v / sigma**2 - 1) < 4 * np.sqrt(2.0 / n)
    assert abs(np.mean(z**3)) < 4 * np.sqrt(6.0 / n)
    assert abs(np.mean(z**4) - 3) < 4 * np.sqrt(24.0 / n)


def test_fill_statistics(ae):
    from scipy import stats as sps

    power = 0.01
    g = ae.noise.new(power, 815)
    v = ae.DeviceVec.with_capacity(N)
    g.fill(v)                      # while len < capacity { push(next()) } (:62-66)
    assert len(v) == N and g.tell() == N
    g.fill(v)                      # full vector: no-op (App. A.5)
    assert g.tell() == N
    z = v.to_numpy()
    sigma = np.sqrt(power)         # scale = sqrt(power) (:35)
    moments(z.real, sigma)
    moments(z.imag, sigma)
    # whiteness: autocorrelation lags 1..64 and re/im cross-correlation < 4/sqrt(N)
    for comp in (z.real.astype(np.float64), z.imag.astype(np.float64)):
        c = comp - comp.mean()
        den = np.dot(c, c)
        for lag in list(range(1, 65)):
            assert abs(np.dot(c[:-lag], c[lag:]) / den) < 4.5 / np.sqrt(N)
    re, im = z.real.astype(np.float64), z.imag.astype(np.float64)
    assert abs(np.corrcoef(re, im)[0, 1]) < 4 / np.sqrt(N)
    # Kolmogorov-Smirnov against N(0, sigma^2) on a 2^20 subsample
    assert sps.kstest(re[: 1 << 20] / sigma, "norm").pvalue > 1e-3
    assert sps.kstest(im[: 1 << 20] / sigma, "norm").pvalue > 1e-3


def test_apply_scales_twice_like_the_reference(ae):
    """Awgn::apply multiplies by scale a second time (:41-42 and :58): component
    variance = power^2 in compat=reference, power in corrected."""
    n, power = 1 << 22, 0.25
    for compat, var in ((ae.COMPAT_REFERENCE, power**2), (ae.COMPAT_CORRECTED, power)):
        g = ae.noise.new(power, 7)
        s = ae.DeviceVec.zeros(n)
        g.apply(s, compat)
        z = s.to_numpy()
        assert abs(z.real.astype(np.float64).var() / var - 1) < 5 * np.sqrt(2.0 / n)
        assert abs(z.imag.astype(np.float64).var() / var - 1) < 5 * np.sqrt(2.0 / n)


def test_stream_is_a_pure_function_of_seed_stream_index(ae):
    n = 100_003
    g = ae.noise.new(1.0, 42)
    a = ae.DeviceVec.with_capacity(n)
    g.fill(a)
    whole = a.to_numpy()
    # same samples when generated in odd-sized pieces (continuation across calls, odd offsets)
    g2 = ae.noise.new(1.0, 42)
    parts = []
    for k in (1, 2, 3, 1000, 4097, n - 5103):
        p = ae.DeviceVec.with_capacity(k)
        g2.fill(p)
        parts.append(p.to_numpy())
    assert np.array_equal(np.concatenate(parts).view(np.uint32), whole.view(np.uint32))
    # seek reproduces; different seeds / streams are different
    g2.seek(5)
    p = ae.DeviceVec.with_capacity(10)
    g2.fill(p)
    assert np.array_equal(p.to_numpy().view(np.uint32), whole[5:15].view(np.uint32))
    g3 = ae.noise.new(1.0, 43)
    q = ae.DeviceVec.with_capacity(1000)
    g3.fill(q)
    assert abs(np.corrcoef(q.to_numpy().real, whole[:1000].real)[0, 1]) < 0.15
    g4 = ae.noise.new(1.0, 42)
    g4.set_stream_id(1)
    r = ae.DeviceVec.with_capacity(1000)
    g4.fill(r)
    assert abs(np.corrcoef(r.to_numpy().real, whole[:1000].real)[0, 1]) < 0.15


def test_matches_oracle_restatement(ae):
    """Same Philox counters, same Box-Muller; only libm (logf/sincospif vs glibc) differs."""
    n = 1 << 16
    g = ae.noise.new(0.5, 815)
    g.seek(3)
    v = ae.DeviceVec.with_capacity(n)
    g.fill(v)
    want = o.awgn_fill(n, 0.5, 815, offset=3)
    got = v.to_numpy()
    assert np.max(np.abs(got - want)) < 2e-6 * 6
    sig = (np.arange(n) % 7).astype(np.complex64)
    g = ae.noise.new(0.5, 815)
    d = ae.DeviceVec.from_numpy(sig)
    g.apply(d)
    assert np.max(np.abs(d.to_numpy() - o.awgn_apply(sig, 0.5, 815))) < 1e-5


def test_generator_defaults(ae):
    g = ae.noise.generator()       # power 1, seed 815 (:9-11)
    v = ae.DeviceVec.with_capacity(1 << 20)
    g.fill(v)
    z = v.to_numpy()
    assert abs(z.real.var() - 1) < 0.01 and abs(z.imag.var() - 1) < 0.01
    g.set_power(4.0)               # :47-50
    w = ae.DeviceVec.with_capacity(1 << 20)
    g.fill(w)
    assert abs(w.to_numpy().real.var() / 4 - 1) < 0.01


def test_iter_and_next_host(ae):
    g = ae.noise.new(2.0, 9)
    a = g.next_host(1000)
    it = g.iter(block=64)
    b = np.array([next(it) for _ in range(100)], dtype=np.complex64)
    g2 = ae.noise.new(2.0, 9)
    v = ae.DeviceVec.with_capacity(1100)
    g2.fill(v)
    whole = v.to_numpy()
    assert np.array_equal(np.concatenate([a, b]).view(np.uint32), whole.view(np.uint32))
