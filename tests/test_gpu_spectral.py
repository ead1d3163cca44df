"""GPU parity (tier T1) for the SURVEY §8(f) rows: spectrogram core (src/util/plot.rs:46-68) and the
frequency-domain correlator (benches/benches.rs:382-423)."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import evm_db

pytestmark = pytest.mark.gpu


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


@pytest.mark.parametrize("fft_len,n", [(1024, 1024 * 7), (1024, 5000), (256, 256), (2048, 2048 * 3 + 17), (100, 950), (64, 1), (8192, 20000)])
def test_spectrogram_vs_oracle(ae, fft_len, n):
    x = rnd(n, fft_len + n)
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        fft = ae.Cfft.with_len(fft_len, compat)
        d = ae.DeviceVec.from_numpy(x)
        lin = ae.spectral.spectrogram(d, fft, use_db=False).to_numpy()
        want = o.spectrogram(x, fft_len, use_db=False, compat=compat)
        assert lin.size == want.size == ((n + fft_len - 1) // fft_len) * fft_len
        assert 10 * np.log10(np.sum((lin - want) ** 2) / np.sum(want ** 2)) <= -80.0
        db = ae.spectral.spectrogram(d, fft, use_db=True).to_numpy()
        wdb = o.spectrogram(x, fft_len, use_db=True, compat=compat)
        big = want > 1e-3 * want.max()          # dB of a near-zero bin amplifies f32 rounding without bound
        assert np.max(np.abs(db[big] - wdb[big])) < 1e-3
        assert np.array_equal(d.to_numpy().view(np.uint32), x.view(np.uint32)), "input must be preserved"


def test_spectrogram_tone_lands_in_the_shifted_bin(ae):
    n, k0 = 1024, 37
    t = np.arange(n)
    x = np.exp(2j * np.pi * k0 * t / n).astype(np.complex64)
    fft = ae.Cfft.with_len(n, ae.COMPAT_CORRECTED)   # exp(-): a +k0 tone lands in bin k0
    lv = ae.spectral.spectrogram(ae.DeviceVec.from_numpy(x), fft, use_db=True).to_numpy()
    assert int(np.argmax(lv)) == k0 + n // 2           # vec_mirror puts DC at n/2
    assert abs(lv.max() - 10 * np.log10(np.sqrt(n))) < 1e-3   # Scale::SN: |X[k0]| = sqrt(n)


@pytest.mark.parametrize("n", [512, 1024, 2048, 100])
def test_correlator_bench_shape(ae, n):
    """Exactly the reference benchmark's setup (benches/benches.rs:391-416)."""
    sig4 = np.array([-1 + 1j, 0, 1 - 1j, 1 - 1j], dtype=np.complex64)
    frames = 5
    inp = np.tile(sig4, n * frames // 4 + 1)[: n * frames].astype(np.complex64)
    sig = np.zeros(n, np.complex64)
    sig[:4] = np.conj(sig4)
    for scale, ok in ((ae.Scale.None_, o.SCALE_NONE), (ae.Scale.SN, o.SCALE_SN)):
        fft = ae.Cfft.with_len(n)
        d = ae.DeviceVec.from_numpy(inp)
        before = ae.launch_count()
        ae.spectral.correlate(d, ae.DeviceVec.from_numpy(sig), fft, scale, howmany=frames)
        if n != 100:
            assert ae.launch_count() == before + 1   # one fused kernel
        want = o.correlate(inp, n, sig, scale_kind=ok)
        assert evm_db(d.to_numpy(), want) <= -80.0
        # equals the call-by-call device composition
        e = ae.DeviceVec.from_numpy(inp)
        for fr in range(frames):
            v = e.view(fr * n, (fr + 1) * n)
            v.vec_rfft(fft, scale).vec_mul(ae.DeviceVec.from_numpy(sig)).vec_rifft(fft, scale)
        assert evm_db(d.to_numpy(), e.to_numpy()) <= -100.0


def test_correlator_random_and_errors(ae):
    n, frames = 1024, 64
    x, s = rnd(n * frames, 1), rnd(n, 2)
    fft = ae.Cfft.with_len(n)
    d = ae.DeviceVec.from_numpy(x)
    ae.spectral.correlate(d, ae.DeviceVec.from_numpy(s), fft, ae.Scale.N, howmany=frames)
    assert evm_db(d.to_numpy(), o.correlate(x, n, s, scale_kind=o.SCALE_N)) <= -80.0
    with pytest.raises(ae.AeError):
        ae.spectral.correlate(d, ae.DeviceVec.zeros(n - 1), fft, ae.Scale.N, howmany=frames)
    with pytest.raises(ae.AeError):
        ae.spectral.correlate(ae.DeviceVec.zeros(n + 1), ae.DeviceVec.zeros(n), fft)
