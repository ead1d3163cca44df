"""GPU parity for the fused chains: modem loop-back (examples/modem.rs:15-32), the headline
FFT -> FIR -> QPSK demod chain, and the OFDM-like chain (BASELINE config 5)."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import evm_db

pytestmark = pytest.mark.gpu


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def taps(t, seed=3):
    rng = np.random.default_rng(seed)
    k = np.arange(t) - (t - 1) / 2
    h = np.sinc(k / 4.3) * np.hamming(t) * np.exp(1j * rng.uniform(0, 2 * np.pi))
    return (h / np.abs(h.sum())).astype(np.complex64)


# ---------------------------------------------------------------- modem
@pytest.mark.parametrize("kind", ["bpsk", "qpsk"])
@pytest.mark.parametrize("nbits", [2, 8, 100, 1000, 2_000_000])
def test_modem_fused_equals_unfused_and_oracle(ae, kind, nbits):
    from aether_primitives_b200.stats import DeviceStats

    rng = np.random.default_rng(815)
    bits = rng.integers(0, 2, nbits, dtype=np.uint8)
    m, table = (ae.modulation.bpsk(), o.BPSK) if kind == "bpsk" else (ae.modulation.qpsk(), o.QPSK)
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        for power in (0.01, 1.5):
            # call-by-call on the device, exactly like the example
            sym = m.modulate(ae.DeviceBits.from_numpy(bits))
            g = ae.noise.new(power, 815)
            g.apply(sym, compat)
            out = ae.DeviceBits.with_capacity(nbits)
            m.demod_naive(sym, out, compat)
            unfused = out.to_numpy()
            # one fused kernel
            g = ae.noise.new(power, 815)
            st = DeviceStats()
            fused = ae.DeviceBits.with_capacity(1)
            before = ae.launch_count()
            ae.chain.modem_fused(m, g, ae.DeviceBits.from_numpy(bits), fused, st, compat)
            assert ae.launch_count() == before + 1
            f = fused.to_numpy()
            assert f.tolist() == unfused.tolist()          # bit-exact: same float expressions, same noise
            errs = int(np.sum((bits != 0) != (f != 0)))
            r = st.read()
            assert r["bit_errors"] == errs and r["n_bits"] == nbits
            if power == 0.01:
                assert errs == 0                           # the example's assert_eq!(b, bits)
            # the oracle uses glibc logf/sin/cos: decisions may differ only for symbols on a boundary
            ob, osym, oerr = o.modem(table, bits, power, 815, compat=compat)
            mism = np.nonzero((ob != 0) != (f != 0))[0]
            bps = m.bits_per_symbol()
            for i in mism:
                s = osym[i // bps]
                assert min(abs(s.real), abs(s.imag)) < 1e-4, "decision differs away from a boundary"
            assert len(mism) <= max(2, nbits // 100000)


def test_modem_continues_noise_stream(ae):
    bits = np.zeros(2000, dtype=np.uint8)
    m = ae.modulation.qpsk()
    g = ae.noise.new(1.0, 1)
    a, b = ae.DeviceBits.with_capacity(1), ae.DeviceBits.with_capacity(1)
    ae.chain.modem_fused(m, g, ae.DeviceBits.from_numpy(bits[:1000 - 2]), a, None)
    assert g.tell() == 499
    ae.chain.modem_fused(m, g, ae.DeviceBits.from_numpy(bits[1000 - 2:]), b, None)   # odd offset
    g2 = ae.noise.new(1.0, 1)
    c = ae.DeviceBits.with_capacity(1)
    ae.chain.modem_fused(m, g2, ae.DeviceBits.from_numpy(bits), c, None)
    assert np.concatenate([a.to_numpy(), b.to_numpy()]).tolist() == c.to_numpy().tolist()


# ---------------------------------------------------------------- FFT -> FIR -> demod
def check_chain(ae, n, frames, t, scale, compat, seed):
    from aether_primitives_b200.chain import FftFirDemod

    x, h = rnd(n * frames, seed), taps(t) if t > 1 else np.array([0.8 - 0.3j], np.complex64)
    okind = {0: o.SCALE_NONE, 1: o.SCALE_SN, 2: o.SCALE_N, 3: o.SCALE_X}[scale.kind]
    want_bits, want_sym = o.chain_fft_fir_demod(x, n, h, scale_kind=okind, scale_x=scale.x, compat=compat)
    ch = FftFirDemod(n, h, scale, compat)
    d_in = ae.DeviceVec.from_numpy(x)
    fused = ae.DeviceBits.with_capacity(1)
    before = ae.launch_count()
    ch.run(d_in, fused)
    launches = ae.launch_count() - before
    got = fused.to_numpy()
    assert got.size == 2 * x.size
    # composition of the stand-alone kernels: symbols must meet the T1 EVM gate
    unf = ae.DeviceBits.with_capacity(1)
    sym = ae.DeviceVec.zeros(x.size)
    ch.run_unfused(d_in, unf, sym)
    assert evm_db(sym.to_numpy(), want_sym) <= -80.0
    # decisions: mismatches only where the oracle's symbol sits within eps of a decision boundary (H4)
    amp = np.sqrt(np.mean(np.abs(want_sym) ** 2))
    for bits in (got, unf.to_numpy()):
        mism = np.nonzero((bits != 0) != (want_bits != 0))[0]
        for i in mism:
            s = want_sym[i // 2]
            comp = s.real if i % 2 == 0 else s.imag
            assert abs(comp) < 2e-5 * amp + 1e-30, "bit %d differs away from a boundary (%r)" % (i, s)
        assert len(mism) <= max(3, x.size // 20000)
    vals = set(np.unique(got).tolist())
    assert vals <= ({0, 1, 2} if compat == ae.COMPAT_REFERENCE else {0, 1})
    return launches


@pytest.mark.parametrize("n,t", [(1024, 64), (256, 16), (512, 33), (2048, 64), (4096, 100), (1024, 1), (1024, 2), (256, 256)])
def test_chain_vs_oracle(ae, n, t):
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        launches = check_chain(ae, n, 9, t, ae.Scale.SN, compat, n + t)
        assert launches == 1, "the chain must be ONE kernel"


def test_chain_scales_and_fallback(ae):
    for sc in (ae.Scale.None_, ae.Scale.N, ae.Scale.X(0.05)):
        check_chain(ae, 1024, 5, 64, sc, ae.COMPAT_REFERENCE, 77)
    # non-power-of-two frame: falls back to the composition of stand-alone kernels
    check_chain(ae, 1000, 4, 24, ae.Scale.SN, ae.COMPAT_REFERENCE, 5)
    from aether_primitives_b200.chain import FftFirDemod

    ch = FftFirDemod(1024, taps(64))
    with pytest.raises(ae.AeError):
        ch.run(ae.DeviceVec.zeros(1000), ae.DeviceBits.with_capacity(1))


def test_chain_host_pipeline_matches_device_path(ae):
    from aether_primitives_b200.chain import FftFirDemod

    n, frames = 1024, 20000  # > one 64 MiB chunk (8192 frames) so the 3-stage pipeline wraps
    x = rnd(n * frames, 9)
    ch = FftFirDemod(n, taps(64))
    dev = ae.DeviceBits.with_capacity(1)
    ch.run(ae.DeviceVec.from_numpy(x), dev)
    host_bits = np.empty(2 * x.size, dtype=np.uint8)
    ch.run_host(x.ctypes.data, x.size, host_bits.ctypes.data)
    assert np.array_equal(host_bits, dev.to_numpy())


def test_chain_full_size_properties(ae):
    """2^16 frames x 1024 (reduced from BASELINE's 2^20 only to keep host RAM modest): determinism,
    frame independence (any frame alone gives the same bits) and oracle spot checks."""
    from aether_primitives_b200.chain import FftFirDemod

    n, frames = 1024, 1 << 16
    x = rnd(n * frames, 21)
    h = taps(64)
    ch = FftFirDemod(n, h)
    d = ae.DeviceVec.from_numpy(x)
    a, b = ae.DeviceBits.with_capacity(1), ae.DeviceBits.with_capacity(1)
    ch.run(d, a)
    ch.run(d, b)
    A = a.to_numpy()
    assert np.array_equal(A, b.to_numpy())
    rng = np.random.default_rng(1)
    for fr in rng.integers(0, frames, 6):
        one = ae.DeviceBits.with_capacity(1)
        ch.run(ae.DeviceVec.from_numpy(x[fr * n:(fr + 1) * n]), one)
        assert np.array_equal(one.to_numpy(), A[2 * fr * n: 2 * (fr + 1) * n])
        wb, ws = o.chain_fft_fir_demod(x[fr * n:(fr + 1) * n], n, h)
        mism = np.nonzero((wb != 0) != (one.to_numpy() != 0))[0]
        assert all(abs(ws[i // 2].real if i % 2 == 0 else ws[i // 2].imag) < 2e-5 for i in mism)


# ---------------------------------------------------------------- OFDM-like chain
@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_ofdm_chain_vs_oracle(ae, n):
    from aether_primitives_b200.stats import DeviceStats, evm_db as sdb

    frames, first = 6, 3
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        for power in (0.0, 0.05, 0.6):
            tx, rx = ae.DeviceBits.with_capacity(1), ae.DeviceBits.with_capacity(1)
            st = DeviceStats()
            before = ae.launch_count()
            ae.chain.ofdm_chain(n, frames, first, power, 5, st, tx, rx, compat)
            assert ae.launch_count() == before + 1
            otx, orx, ostats, osym = o.ofdm_chain(n, frames, first, power, 5, compat)
            assert tx.to_numpy().tolist() == otx.tolist()            # M-sequence + layout: bit-exact
            R = rx.to_numpy()
            mism = np.nonzero((R != 0) != (orx != 0))[0]
            for i in mism:
                s = osym[i // 2]
                assert abs(s.real if i % 2 == 0 else s.imag) < 1e-4
            r = st.read()
            assert r["n_bits"] == 2 * n * frames
            assert abs(r["bit_errors"] - ostats[0]) <= len(mism)
            assert r["bit_errors"] == int(np.sum((R != 0) != (otx != 0)))
            assert abs(r["ref_pow"] - ostats[3]) < 1e-6 * ostats[3]
            if power > 0:
                assert abs(sdb(r["err_pow"], r["ref_pow"]) - sdb(ostats[2], ostats[3])) < 0.01
            else:
                assert r["bit_errors"] == 0 and sdb(r["err_pow"], r["ref_pow"]) < -120
    # frames are independent of how the job is split (what the multi-GPU sharding relies on)
    whole = ae.DeviceBits.with_capacity(1)
    ae.chain.ofdm_chain(n, 8, 0, 0.3, 5, None, None, whole)
    lo, hi = ae.DeviceBits.with_capacity(1), ae.DeviceBits.with_capacity(1)
    ae.chain.ofdm_chain(n, 5, 0, 0.3, 5, None, None, lo)
    ae.chain.ofdm_chain(n, 3, 5, 0.3, 5, None, None, hi)
    assert np.concatenate([lo.to_numpy(), hi.to_numpy()]).tolist() == whole.to_numpy().tolist()


def test_streaming_pipeline_matches_device_path_and_reports_stages(ae):
    """ae_pipe_* (pipeline.rs / pool.rs analogue): blocks streamed through depth-3 slots give the device
    path's bits, in order; the per-stage report counts every block and keeps utilisation within 0..100 %."""
    import torch
    from aether_primitives_b200.chain import ChainPipeline, FftFirDemod

    n, block_frames, blocks = 1024, 256, 11
    rng = np.random.default_rng(77)
    taps = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64) / 8
    x = (rng.standard_normal(blocks * block_frames * n) + 1j * rng.standard_normal(blocks * block_frames * n)).astype(np.complex64)
    ch = FftFirDemod(n, taps)
    want = ae.DeviceBits.with_capacity(2 * x.size)
    ch.run(ae.DeviceVec.from_numpy(x), want)
    want = want.to_numpy()

    h_in = torch.from_numpy(x).pin_memory()
    h_out = torch.zeros(2 * x.size, dtype=torch.uint8).pin_memory()
    pipe = ChainPipeline(ch, block_frames, depth=3)
    with pytest.raises(ae.AeError):
        pipe.recv()                                           # nothing in flight
    bs, bo = block_frames * n * 8, block_frames * n * 2
    received = []
    for b in range(blocks):                                    # more blocks than slots: send() recycles the oldest slot
        pipe.send(h_in.data_ptr() + b * bs, h_out.data_ptr() + b * bo)
    assert pipe.in_flight() == blocks
    for b in range(blocks):
        received.append(pipe.recv())
    assert received == [h_out.data_ptr() + b * bo for b in range(blocks)]
    assert pipe.in_flight() == 0
    assert np.array_equal(h_out.numpy(), want)
    rep = pipe.report(reset=True)
    assert [r["name"] for r in rep] == ["h2d", "fft-fir-demod", "d2h"]
    for r in rep:
        assert r["processed"] == blocks and 0.0 < r["utilisation_pct"] <= 100.0 + 1e-6 and r["per_second"] > 0
        assert r["active_ms"] <= r["elapsed_ms"] + 1e-6
    assert all(r["processed"] == 0 for r in pipe.report())
    # a second window after the reset works and interleaved send/recv keeps order
    for b in range(4):
        pipe.send(h_in.data_ptr() + b * bs, h_out.data_ptr() + b * bo)
        assert pipe.recv() == h_out.data_ptr() + b * bo
    assert pipe.report()[1]["processed"] == 4
    with pytest.raises(ae.AeError):
        ChainPipeline(ch, 0)
