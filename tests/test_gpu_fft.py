"""GPU parity, tier T1 (power EVM <= -80 dB vs the oracle; error vs f64 truth <= 2x the oracle's):
Cfft / Fft trait (src/fft.rs:48-235), vec_fft/vec_ifft/vec_rfft/vec_rifft (src/vecops.rs:301-325)."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import cx, evm_db, load, same_bits, truth

pytestmark = pytest.mark.gpu
G = load()
EVM_LIMIT_DB = -80.0


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


POW2 = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
ODD = [1, 2, 3, 4, 5, 8, 12, 100, 360, 1000, 1009, 3 * 1024]


@pytest.mark.parametrize("n", POW2 + ODD)
def test_all_entry_points_vs_oracle(ae, n):
    frames = 3 if n <= 4096 else 2
    x = rnd(n * frames, n)
    scales = [(ae.Scale.None_, o.SCALE_NONE, 1.0), (ae.Scale.SN, o.SCALE_SN, 1.0), (ae.Scale.N, o.SCALE_N, 1.0), (ae.Scale.X(0.37), o.SCALE_X, 0.37)]
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        f = ae.Cfft.with_len(n, compat)
        assert f.len() == n
        for bwd in (False, True):
            for sc, ok, ox in scales:
                want = o.cfft(x, n, bwd=bwd, scale_kind=ok, scale_x=ox, compat=compat)
                din = ae.DeviceVec.from_numpy(x)
                dout = ae.DeviceVec.zeros(x.size)
                (f.bwd if bwd else f.fwd)(din, dout, sc, howmany=frames)
                got = dout.to_numpy()
                assert evm_db(got, want) <= EVM_LIMIT_DB
                assert same_bits(din.to_numpy(), x), "fwd/bwd must not modify the input"
                # in place
                (f.ibwd if bwd else f.ifwd)(din, sc, howmany=frames)
                assert same_bits(din.to_numpy(), got)
                # temp variant: result in the plan's scratch, input preserved
                din2 = ae.DeviceVec.from_numpy(x)
                view = (f.tbwd if bwd else f.tfwd)(din2, sc, howmany=frames)
                assert same_bits(view.to_numpy(), got)
                assert same_bits(din2.to_numpy(), x)


@pytest.mark.parametrize("n", [8, 16, 64, 100, 128, 256, 360, 512, 1009, 1024, 2048, 4096, 8192])
def test_accuracy_vs_f64_truth(ae, n):
    t = truth()
    x = t["x_%d" % n]
    f = ae.Cfft.with_len(n, ae.COMPAT_CORRECTED)
    for bwd, key, sign in ((False, "neg", -1), (True, "pos", +1)):
        d = ae.DeviceVec.from_numpy(x)
        (f.ibwd if bwd else f.ifwd)(d, ae.Scale.None_)
        got = d.to_numpy()
        ref = t["%s_%d" % (key, n)]
        e_gpu = evm_db(got, ref)
        e_ora = evm_db(o.fft_raw(x, sign), ref)
        assert e_gpu <= EVM_LIMIT_DB
        assert e_gpu <= e_ora + 3.02, "GPU error power must be within 2x of the oracle's (%.1f vs %.1f dB)" % (e_gpu, e_ora)


def test_golden_roundtrip_100_and_vec_fft(ae):
    c = G["fft_roundtrip_100"]
    v = cx(c["v"])
    d = ae.DeviceVec.from_numpy(v)
    d.vec_fft(ae.Scale.SN).vec_ifft(ae.Scale.SN)               # src/vecops.rs:445-452
    # the reference's assertion, literally: assert_evm!(c, v) at -80 "dB" = |err| <= 1e-8 |ref| = equality to the bit
    # (the DC-exact odd butterflies of fft.cu make every non-DC bin of a constant input exactly zero)
    assert o.assert_evm(d.to_numpy(), v, -80.0)[0] == o.OK, "macro-worst %.1f dB" % o.evm_macro_worst_db(d.to_numpy(), v)
    f = ae.Cfft.with_len(100)
    d = ae.DeviceVec.from_numpy(v)
    d.vec_rfft(f, ae.Scale.SN).vec_rifft(f, ae.Scale.SN)       # src/vecops.rs:454-463
    assert o.assert_evm(d.to_numpy(), v, -80.0)[0] == o.OK, "macro-worst %.1f dB" % o.evm_macro_worst_db(d.to_numpy(), v)
    # other constant frames and lengths with odd factors: the same exactness
    for n, val in ((100, 3 - 2j), (75, 1 + 1j), (1000, 0.5 + 4j), (360, -1 + 1j), (1001, 2 + 2j)):
        w = np.full(n, val, np.complex64)
        d = ae.DeviceVec.from_numpy(w)
        d.vec_fft(ae.Scale.None_)
        spec = d.to_numpy()
        assert np.all(spec[1:] == 0), "non-DC bins of a constant frame must be exactly zero (n=%d)" % n
        assert spec[0] == np.complex64(n * val)


def test_golden_doctest_128(ae):
    c = G["fft_doctest_128"]
    d = ae.DeviceVec.from_numpy(cx(c["v"]))
    d.vec_fft(ae.Scale.None_)
    assert o.assert_evm(d.to_numpy(), cx(c["spectrum"]), -80.0)[0] == o.OK   # the doctest's literal assertion
    f = ae.Cfft.with_len(128)
    f.ibwd(d, ae.Scale.N)
    assert o.assert_evm(d.to_numpy(), cx(c["v"]), -80.0)[0] == o.OK
    d.vec_rfft(f, ae.Scale.SN).vec_scale(2.0).vec_rifft(f, ae.Scale.SN)
    assert o.assert_evm(d.to_numpy(), np.full(128, 2 + 0j, np.complex64), c["twos_db"])[0] == o.OK


def test_sign_quirk_f3(ae):
    n = G["fft_sign_quirk"]["n"]
    x = np.zeros(n, np.complex64)
    x[1] = 1
    k = np.arange(n)
    for nn in (n, 64, 1024):
        x = np.zeros(nn, np.complex64)
        x[1] = 1
        k = np.arange(nn)
        d = ae.DeviceVec.from_numpy(x)
        ae.Cfft.with_len(nn, ae.COMPAT_REFERENCE).ifwd(d, ae.Scale.None_)
        assert np.allclose(d.to_numpy(), np.exp(+2j * np.pi * k / nn), atol=1e-6)
        d = ae.DeviceVec.from_numpy(x)
        ae.Cfft.with_len(nn, ae.COMPAT_REFERENCE).ibwd(d, ae.Scale.None_)
        assert np.allclose(d.to_numpy(), np.exp(-2j * np.pi * k / nn), atol=1e-6)


def test_length_assert(ae):
    f = ae.Cfft.with_len(64)
    with pytest.raises(ae.AeError) as e:
        f.ifwd(ae.DeviceVec.zeros(65), ae.Scale.None_)
    assert e.value.status == ae._lib.AE_ELEN and "Input and FFT must be the same length" in e.value.message
    with pytest.raises(ae.AeError):
        f.fwd(ae.DeviceVec.zeros(64), ae.DeviceVec.zeros(63), ae.Scale.None_)
    with pytest.raises(ae.AeError):
        ae.Cfft.with_len(0)


def test_scale_fold_equals_separate_pass(ae):
    """Scale folded into the last FFT pass == transform(None) then Scale::scale, bit for bit."""
    n, frames = 1024, 5
    x = rnd(n * frames, 3)
    f = ae.Cfft.with_len(n)
    for sc in (ae.Scale.SN, ae.Scale.N, ae.Scale.X(1.234)):
        a = ae.DeviceVec.from_numpy(x)
        f.ifwd(a, sc, howmany=frames)
        b = ae.DeviceVec.from_numpy(x)
        f.ifwd(b, ae.Scale.None_, howmany=frames)
        b.vec_scale(sc.factor(n))
        assert same_bits(a.to_numpy(), b.to_numpy())


def test_large_batch_properties(ae):
    """BASELINE config 2 shape (reduced frames): fwd(SN) then bwd(SN) returns the input; Parseval;
    linearity — size-independent properties at a size the oracle does not need to see."""
    n, frames = 1024, 1 << 14
    x = rnd(n * frames, 11)
    f = ae.Cfft.with_len(n)
    d = ae.DeviceVec.from_numpy(x)
    f.ifwd(d, ae.Scale.SN, howmany=frames)
    X = d.to_numpy()
    assert abs(np.sum(np.abs(X.astype(np.complex128)) ** 2) / np.sum(np.abs(x.astype(np.complex128)) ** 2) - 1) < 1e-5
    f.ibwd(d, ae.Scale.SN, howmany=frames)
    assert evm_db(d.to_numpy(), x) <= -120
    # spot-check 8 random frames against the oracle
    rng = np.random.default_rng(0)
    for fr in rng.integers(0, frames, 8):
        want = o.cfft(x[fr * n:(fr + 1) * n], n, scale_kind=o.SCALE_SN)
        assert evm_db(X[fr * n:(fr + 1) * n], want) <= EVM_LIMIT_DB


@pytest.mark.parametrize("n", [1 << 15, 1 << 16, 1 << 17, 1 << 20])
def test_large_power_of_two_four_step(ae, n):
    """2^15..2^24: four-step path (two column-FFT passes); same contract as every other length."""
    frames = 3 if n <= (1 << 17) else 1
    x = rnd(n * frames, n % 1000)
    for compat, bwd, sc, ok in ((ae.COMPAT_REFERENCE, False, ae.Scale.SN, o.SCALE_SN), (ae.COMPAT_CORRECTED, True, ae.Scale.None_, o.SCALE_NONE)):
        f = ae.Cfft.with_len(n, compat)
        want = o.cfft(x, n, bwd=bwd, scale_kind=ok, compat=compat)
        din, dout = ae.DeviceVec.from_numpy(x), ae.DeviceVec.zeros(x.size)
        (f.bwd if bwd else f.fwd)(din, dout, sc, howmany=frames)
        got = dout.to_numpy()
        assert evm_db(got, want) <= EVM_LIMIT_DB
        truth = np.concatenate([(np.fft.ifft(x[i * n:(i + 1) * n].astype(np.complex128)) * n) if ((compat == ae.COMPAT_REFERENCE) != bwd)
                                else np.fft.fft(x[i * n:(i + 1) * n].astype(np.complex128)) for i in range(frames)])
        truth = truth * (1.0 / np.sqrt(np.float32(n)) if sc.kind == 1 else 1.0)
        assert evm_db(got, truth) <= evm_db(want, truth) + 3.02
        assert same_bits(din.to_numpy(), x)
        (f.ibwd if bwd else f.ifwd)(din, sc, howmany=frames)      # in place
        assert same_bits(din.to_numpy(), got)


@pytest.mark.parametrize("n", [6, 7, 11, 13, 14, 17, 18, 20, 24, 30, 34, 49, 60, 77, 85, 91, 96, 143, 169, 210, 289, 600, 1001, 1200, 1536,
                               3000, 4199, 6000, 6144])
def test_mixed_radix_lengths_with_ragged_frame_groups(ae, n):
    """Any-length path: register butterflies for 2/3/4/5/7/8/11/13 and the prime-factor composites 6/10/12 (5*2, 3*4, 3*2 are
    merged: 18 = 6*3, 20 = 10*2, 24 = 12*2, 60 = 10*6, 600 = 10*10*6, 1200 = 10*10*12, 3000 = 3*10*10*10), per-output fallback
    for other primes (17, 19), several frames per CTA — frame counts that do not fill the last CTA's slots, in place and
    out of place, both directions."""
    frames = 45 if n <= 1100 else 5
    x = rnd(n * frames, n)
    f = ae.Cfft.with_len(n)
    for bwd in (False, True):
        want = o.cfft(x, n, bwd=bwd, scale_kind=o.SCALE_SN, scale_x=1.0, compat=ae.COMPAT_REFERENCE)
        din = ae.DeviceVec.from_numpy(x)
        dout = ae.DeviceVec.zeros(x.size)
        (f.bwd if bwd else f.fwd)(din, dout, ae.Scale.SN, howmany=frames)
        got = dout.to_numpy()
        assert evm_db(got, want) <= EVM_LIMIT_DB
        for fr in range(frames):     # per frame, so one bad frame cannot hide in the average
            assert evm_db(got[fr * n:(fr + 1) * n], want[fr * n:(fr + 1) * n]) <= EVM_LIMIT_DB
        (f.ibwd if bwd else f.ifwd)(din, ae.Scale.SN, howmany=frames)
        assert same_bits(din.to_numpy(), got)
    # round trip: bwd(fwd(x)) with Scale::SN both ways returns x (src/fft.rs:93-117 doctest pattern)
    d = ae.DeviceVec.from_numpy(x)
    f.ifwd(d, ae.Scale.SN, howmany=frames)
    f.ibwd(d, ae.Scale.SN, howmany=frames)
    assert evm_db(d.to_numpy(), x) <= EVM_LIMIT_DB


@pytest.mark.parametrize("n", [67, 134, 1009, 4099, 6145, 7919, 8198, 10000, 12288, 12289, 20000, 3 * 16384 + 1])
def test_bluestein_lengths(ae, n):
    """Non-power-of-two lengths above 6144 and lengths with a prime factor > 61 run as a chirp-z
    convolution over the power-of-two kernels: same contract as every other length (all scales,
    both directions, in place and out of place, EVM <= -80 dB vs the oracle per frame)."""
    frames = 3
    x = rnd(n * frames, n)
    f = ae.Cfft.with_len(n)
    for bwd in (False, True):
        for sc, ok, ox in ((ae.Scale.None_, o.SCALE_NONE, 1.0), (ae.Scale.SN, o.SCALE_SN, 1.0)):
            want = o.cfft(x, n, bwd=bwd, scale_kind=ok, scale_x=ox, compat=ae.COMPAT_REFERENCE)
            din = ae.DeviceVec.from_numpy(x)
            dout = ae.DeviceVec.zeros(x.size)
            (f.bwd if bwd else f.fwd)(din, dout, sc, howmany=frames)
            got = dout.to_numpy()
            for fr in range(frames):
                assert evm_db(got[fr * n:(fr + 1) * n], want[fr * n:(fr + 1) * n]) <= EVM_LIMIT_DB
            assert same_bits(din.to_numpy(), x)
            (f.ibwd if bwd else f.ifwd)(din, sc, howmany=frames)
            assert same_bits(din.to_numpy(), got)
    d = ae.DeviceVec.from_numpy(x)
    f.ifwd(d, ae.Scale.SN, howmany=frames)
    f.ibwd(d, ae.Scale.SN, howmany=frames)
    assert evm_db(d.to_numpy(), x) <= EVM_LIMIT_DB
