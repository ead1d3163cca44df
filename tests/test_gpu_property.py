"""Randomised (hypothesis, derandomised so every run sees the same cases) parity through the C ABI:
shapes nobody hand-picked — ragged lengths, odd factors, streaming splits — against the oracle."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from tests import oracle as o
from tests.golden_util import evm_db, same_bits

pytestmark = pytest.mark.gpu
CFG = dict(max_examples=30, deadline=None, derandomize=True)


def cx(rng, n, scale=1.0):
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


@settings(**CFG)
@given(n=st.integers(1, 3000), howmany=st.integers(1, 7), bwd=st.booleans(), kind=st.integers(0, 3), seed=st.integers(0, 2**31))
def test_fft_any_length_any_batch(ae, n, howmany, bwd, kind, seed):
    rng = np.random.default_rng(seed)
    x = cx(rng, n * howmany)
    sc = [ae.Scale.None_, ae.Scale.SN, ae.Scale.N, ae.Scale.X(0.25)][kind]
    want = o.cfft(x, n, bwd=bwd, scale_kind=kind, scale_x=0.25, compat=ae.COMPAT_REFERENCE)
    d = ae.DeviceVec.from_numpy(x)
    f = ae.Cfft.with_len(n)
    (f.ibwd if bwd else f.ifwd)(d, sc, howmany=howmany)
    got = d.to_numpy()
    if n == 1:
        assert same_bits(got, want)
    else:
        assert evm_db(got, want) <= -80.0


@settings(**CFG)
@given(ntaps=st.integers(1, 150), n=st.integers(1, 6000), cut=st.floats(0.0, 1.0), mode=st.sampled_from([1, 2]), seed=st.integers(0, 2**31))
def test_fir_streaming_split_equals_one_shot(ae, ntaps, n, cut, mode, seed):
    """y = h * x with the history carried across calls: filtering x in two pieces gives what the oracle
    gives for the whole signal (direct form T0-close, overlap-save within -80 dB), any split point."""
    from aether_primitives_b200 import fir as F

    rng = np.random.default_rng(seed)
    x, h = cx(rng, n), cx(rng, ntaps, 1.0 / ntaps)
    want = o.fir_f64(x, h)
    try:
        filt = F.Fir(h, mode)
    except ae.AeError:
        return                                   # overlap-save refuses tap counts it has no block length for
    k = int(round(cut * n))
    out = np.empty(n, np.complex64)
    for lo, hi in ((0, k), (k, n)):
        if hi > lo:
            di = ae.DeviceVec.from_numpy(x[lo:hi])
            do = ae.DeviceVec.zeros(hi - lo)
            filt.filter(di, do)
            out[lo:hi] = do.to_numpy()
    ref_pow = float(np.sum(np.abs(want) ** 2))
    err_pow = float(np.sum(np.abs(out - want) ** 2))
    assert err_pow <= 1e-8 * ref_pow + 1e-10 * n     # -80 dB, with an absolute floor for tiny signals


@settings(**CFG)
@given(n_dst=st.integers(1, 4000), dec=st.integers(1, 9), seed=st.integers(0, 2**31))
def test_downsample_bit_exact(ae, n_dst, dec, seed):
    rng = np.random.default_rng(seed)
    x = cx(rng, n_dst * dec)
    dst = ae.DeviceVec.zeros(n_dst)
    ae.sampling.downsample(ae.DeviceVec.from_numpy(x), dst)
    assert same_bits(dst.to_numpy(), o.downsample(x, n_dst))


@settings(**CFG)
@given(n=st.integers(1, 3000), k=st.integers(0, 9), compat=st.sampled_from([0, 1]), seed=st.integers(0, 2**31))
def test_interpolate_bit_exact(ae, n, k, compat, seed):
    rng = np.random.default_rng(seed)
    x = cx(rng, n, 100.0)
    dst = ae.DeviceVec.with_capacity(1)
    ae.sampling.interpolate(ae.DeviceVec.from_numpy(x), dst, k, compat)
    assert same_bits(dst.to_numpy(), o.interpolate(x, k, compat))


@settings(**CFG)
@given(nsym=st.integers(1, 5000), qpsk=st.booleans(), compat=st.sampled_from([0, 1]), noise=st.floats(0.0, 2.0), seed=st.integers(0, 2**31))
def test_modulate_demod_bit_exact(ae, nsym, qpsk, compat, noise, seed):
    rng = np.random.default_rng(seed)
    m = ae.modulation.qpsk() if qpsk else ae.modulation.bpsk()
    table = o.QPSK if qpsk else o.BPSK
    bits = rng.integers(0, 2, nsym * (2 if qpsk else 1), dtype=np.uint8)
    sym = m.modulate(ae.DeviceBits.from_numpy(bits))
    want_sym = o.modulate(table, bits)
    assert same_bits(sym.to_numpy(), want_sym)
    noisy = (want_sym + cx(rng, nsym, noise)).astype(np.complex64)
    out = ae.DeviceBits.with_capacity(1)
    m.demod_naive(ae.DeviceVec.from_numpy(noisy), out, compat)
    assert np.array_equal(out.to_numpy(), o.demod(table, noisy, compat))


@settings(**CFG)
@given(n=st.integers(1, 20000), seed=st.integers(0, 2**31), with_specials=st.booleans())
def test_vec_stats_against_oracle(ae, n, seed, with_specials):
    rng = np.random.default_rng(seed)
    x = cx(rng, n)
    if with_specials and n > 4:
        x[rng.integers(0, n)] = complex(np.inf, 0.0)
        x[rng.integers(0, n)] = 0
        x[rng.integers(0, n)] = x[rng.integers(0, n)]          # a duplicate: tie on norm_sqr
    got = ae.DeviceVec.from_numpy(x).vec_stats()
    want = o.vec_stats(x)
    assert (got.min[1] if got.min else n) == want["min_idx"]
    assert (got.max[1] if got.max else n) == want["max_idx"]
