"""GPU parity, tier T1: FIR direct form and overlap-save vs the f64 direct-form truth
(src/fir.rs:1-22 holds no filter: parity unpinned, oracle = textbook definition)."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import evm_db

pytestmark = pytest.mark.gpu


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def taps(t, seed=3):
    """windowed-sinc * exp(j phi), unit DC gain"""
    rng = np.random.default_rng(seed)
    k = np.arange(t) - (t - 1) / 2
    h = np.sinc(k / 4.3) * np.hamming(t) * np.exp(1j * rng.uniform(0, 2 * np.pi))
    return (h / np.abs(h.sum())).astype(np.complex64)


@pytest.mark.parametrize("mode", ["direct", "os"])
@pytest.mark.parametrize("t", [1, 2, 7, 8, 9, 16, 64, 65, 200, 1024])
@pytest.mark.parametrize("n", [1, 5, 63, 2048, 5000, 20011])
def test_stream_vs_truth(ae, mode, t, n):
    from aether_primitives_b200 import fir as F

    x, h = rnd(n, n + t), taps(t) if t > 2 else rnd(t, 9)
    filt = F.Fir(h, F.DIRECT if mode == "direct" else F.OVERLAP_SAVE)
    assert filt.ntaps() == t
    din, dout = ae.DeviceVec.from_numpy(x), ae.DeviceVec.zeros(n)
    filt.filter(din, dout)
    truth = o.fir_f64(x, h)
    got = dout.to_numpy()
    if mode == "os" and n < t:
        # a block shorter than the filter only sees the first taps (tiny for a windowed sinc) while
        # overlap-save rounds relative to the whole FFT block: bound the ABSOLUTE error instead
        assert np.max(np.abs(got - truth)) <= 2e-6 * np.sum(np.abs(h)) * np.max(np.abs(x))
        return
    assert evm_db(got, truth) <= -80.0   # T1 gate
    if mode == "direct":
        # same algorithm as the f32 oracle, different summation order: same accuracy class
        e_ora = evm_db(o.fir(x, h), truth)
        assert evm_db(got, truth) <= max(e_ora + 3.02, -130.0)
    elif n >= 2048:
        # overlap-save rounds relative to the energy of the whole FFT block, not of the output sample
        assert evm_db(got, truth) <= -120.0


@pytest.mark.parametrize("mode", ["direct", "os"])
def test_carried_history(ae, mode):
    from aether_primitives_b200 import fir as F

    t, n = 64, 10000
    x, h = rnd(n, 1), taps(t)
    truth = o.fir_f64(x, h)
    filt = F.Fir(h, F.DIRECT if mode == "direct" else F.OVERLAP_SAVE)
    got = []
    cuts = [0, 10, 30, 31, 1000, 1001, 4096, 9999, n]  # includes blocks shorter than the history
    for a, b in zip(cuts[:-1], cuts[1:]):
        din, dout = ae.DeviceVec.from_numpy(x[a:b]), ae.DeviceVec.zeros(b - a)
        filt.filter(din, dout)
        got.append(dout.to_numpy())
    assert evm_db(np.concatenate(got), truth) <= -120.0
    filt.reset()
    din, dout = ae.DeviceVec.from_numpy(x[:100]), ae.DeviceVec.zeros(100)
    filt.filter(din, dout)
    assert evm_db(dout.to_numpy(), truth[:100]) <= -120.0


@pytest.mark.parametrize("mode", ["direct", "os", "auto"])
@pytest.mark.parametrize("frame_len,t", [(1024, 64), (256, 16), (1000, 33), (2048, 64)])
def test_framed_zero_state(ae, mode, frame_len, t):
    from aether_primitives_b200 import fir as F

    frames = 7
    x, h = rnd(frame_len * frames, 5), taps(t)
    filt = F.Fir(h, {"direct": F.DIRECT, "os": F.OVERLAP_SAVE, "auto": F.AUTO}[mode])
    din, dout = ae.DeviceVec.from_numpy(x), ae.DeviceVec.zeros(x.size)
    filt.filter(din, dout, frame_len=frame_len)
    truth = o.fir_f64(x, h, frame_len=frame_len)
    assert evm_db(dout.to_numpy(), truth) <= -120.0


def test_in_place_and_errors(ae):
    from aether_primitives_b200 import fir as F

    x, h = rnd(5000, 2), taps(32)
    filt = F.Fir(h, F.DIRECT)
    d = ae.DeviceVec.from_numpy(x)
    filt.filter(d, d)
    assert evm_db(d.to_numpy(), o.fir_f64(x, h)) <= -120.0
    with pytest.raises(ae.AeError):
        filt.filter(ae.DeviceVec.zeros(10), ae.DeviceVec.zeros(11))
    with pytest.raises(ae.AeError):
        F.Fir(np.zeros(0, np.complex64))


def test_large_linearity_and_impulse(ae):
    """2^22 samples, 64 and 1024 taps (BASELINE config 3 shapes, reduced): impulse response
    reproduces the taps; direct and overlap-save agree."""
    from aether_primitives_b200 import fir as F

    n = 1 << 22
    for t in (64, 1024):
        h = taps(t)
        x = rnd(n, t)
        d_in = ae.DeviceVec.from_numpy(x)
        y1, y2 = ae.DeviceVec.zeros(n), ae.DeviceVec.zeros(n)
        F.Fir(h, F.DIRECT).filter(d_in, y1)
        F.Fir(h, F.OVERLAP_SAVE).filter(d_in, y2)
        a, b = y1.to_numpy(), y2.to_numpy()
        assert evm_db(a, b) <= -110.0
        seg = slice(n // 2, n // 2 + 4096)
        truth = o.fir_f64(x[seg.start - t + 1: seg.stop], h)[t - 1:]
        assert evm_db(a[seg], truth) <= -120.0 and evm_db(b[seg], truth) <= -120.0
        imp = np.zeros(4096, np.complex64)
        imp[0] = 1
        yi = ae.DeviceVec.zeros(4096)
        F.Fir(h, F.AUTO).filter(ae.DeviceVec.from_numpy(imp), yi)
        assert evm_db(yi.to_numpy()[:t], h) <= -120.0
