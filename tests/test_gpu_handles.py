"""GPU: Vec-like handle semantics behind the C ABI (len/capacity/append/grow/views/wrapped foreign
memory), i.e. what makes the device types behave like the reference's `Vec<cf32>` / `&mut [cf32]`."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import same_bits

pytestmark = pytest.mark.gpu


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def test_append_and_growth_keep_contents(ae):
    x = rnd(1000, 1)
    dst = ae.DeviceVec.with_capacity(4)                 # far too small: must grow like Vec::push
    assert len(dst) == 0 and dst.capacity() == 4
    ae.sampling.interpolate(ae.DeviceVec.from_numpy(x[:10]), dst, 2)
    first = dst.to_numpy()
    ae.sampling.interpolate(ae.DeviceVec.from_numpy(x), dst, 1)     # reallocates; earlier samples survive
    got = dst.to_numpy()
    assert len(dst) == first.size + 1999 and dst.capacity() >= len(dst)
    assert same_bits(got[: first.size], first)
    assert same_bits(got[first.size:], o.interpolate(x, 1))
    dst.clear()
    assert len(dst) == 0 and dst.capacity() >= 2000     # clear keeps the allocation


def test_pending_tape_survives_growth_and_set_len(ae):
    x = rnd(100, 2)
    v = ae.DeviceVec.with_capacity(100)
    g = ae.noise.new(1.0, 3)
    g.fill(v)                                            # len -> capacity
    base = v.to_numpy()
    v.vec_scale(2.0)                                     # pending
    ae.sampling.interpolate(ae.DeviceVec.from_numpy(x), v, 0, ae.COMPAT_CORRECTED)   # append forces a flush + realloc
    got = v.to_numpy()
    assert same_bits(got[:100], o.vec_scale(base, 2.0)) and same_bits(got[100:], x)


def test_wrapped_foreign_memory(ae):
    torch = pytest.importorskip("torch")
    t = torch.view_as_complex(torch.randn(4097, 2, device="cuda"))
    ref = t.cpu().numpy().copy()
    v = ae.DeviceVec.from_torch(t)
    v.vec_conj().vec_scale(3.0).flush()
    ae.sync()
    torch.cuda.synchronize()
    assert same_bits(t.cpu().numpy(), o.vec_scale(o.vec_conj(ref), 3.0))      # the library wrote into torch's buffer
    with pytest.raises(ae.AeError):
        ae.sampling.interpolate(ae.DeviceVec.from_numpy(ref), v, 1)           # borrowed memory cannot grow
    # an 8-byte aligned (not 16) slice of foreign memory takes the scalar paths
    u = ae.DeviceVec.wrap(t.data_ptr() + 8, 4096, owner=t)
    u.vec_mirror().flush()
    ae.sync()
    want = o.vec_scale(o.vec_conj(ref), 3.0)
    want[1:] = o.vec_mirror(want[1:])
    assert same_bits(t.cpu().numpy(), want)


def test_bits_handles(ae):
    b = ae.DeviceBits.with_capacity(2)
    m = ae.modulation.bpsk()
    s = rnd(1001, 4)
    m.demod_naive(ae.DeviceVec.from_numpy(s), b)         # append + grow
    m.demod_naive(ae.DeviceVec.from_numpy(s[:7]), b)
    got = b.to_numpy()
    assert got.size == 1008
    assert got[:1001].tolist() == o.demod(o.BPSK, s).tolist() and got[1001:].tolist() == o.demod(o.BPSK, s[:7]).tolist()
    b.clear()
    assert len(b) == 0


def test_pinned_host_alloc_and_stream_switch(ae):
    lib = ae._lib.lib()
    p = C.c_void_p()
    ae._lib.call("ae_host_alloc", 1 << 20, C.byref(p))
    assert p.value
    ae._lib.call("ae_host_free", p)
    torch = pytest.importorskip("torch")
    x = rnd(5000, 5)
    v = ae.DeviceVec.from_numpy(x)
    v.vec_scale(2.0)                                     # recorded on the library stream
    s = torch.cuda.Stream()
    ae.set_stream(s.cuda_stream)                         # later work is ordered after earlier work
    v.vec_conj()
    got = v.to_numpy()
    ae.set_stream(None)
    assert same_bits(got, o.vec_conj(o.vec_scale(x, 2.0)))
    assert int(lib.ae_launch_count()) > 0


def test_empty_inputs_behave_like_empty_slices(ae):
    """Empty Vec / slice: element-wise ops are no-ops, collectors return empty, and the calls the
    reference panics on for empty input still fail."""
    e = ae.DeviceVec.zeros(0)
    f = ae.DeviceVec.zeros(0)
    e.vec_scale(2.0).vec_mul(f).vec_conj().vec_mirror().vec_add(f).vec_zero().vec_clone(f).flush()
    assert len(e) == 0 and e.to_numpy().size == 0
    m = ae.modulation.qpsk()
    sym = m.modulate(ae.DeviceBits.zeros(0))                  # chunks() of an empty slice -> empty Vec
    assert len(sym) == 0
    out = ae.DeviceBits.with_capacity(1)
    m.demod_naive(sym, out)
    assert len(out) == 0
    fft = ae.Cfft.with_len(64)
    fft.ifwd(e, ae.Scale.SN, howmany=0)                       # zero frames
    with pytest.raises(ae.AeError):
        fft.ifwd(e, ae.Scale.SN)                              # one frame expected: length assert
    with pytest.raises(ae.AeError):
        ae.sampling.interpolate(e, ae.DeviceVec.with_capacity(1), 3)      # src.last().unwrap() on empty
    with pytest.raises(ae.AeError):
        ae.sampling.downsample(e, ae.DeviceVec.zeros(0))                  # division by zero
    g = ae.noise.new(1.0, 1)
    g.apply(e)
    assert g.tell() == 0
    from aether_primitives_b200.chain import FftFirDemod

    ch = FftFirDemod(1024, np.ones(4, np.complex64))
    bits = ae.DeviceBits.with_capacity(1)
    ch.run(e, bits)                                           # zero frames
    assert len(bits) == 0
    assert len(ae.sequence.expand(5, 0)) == 0


def test_csv_round_trip(ae, tmp_path):
    """util::file csv reader/writer for cf32 (src/util/file.rs:109-124, test :176-214): `re,im` per line, no header."""
    x = (np.arange(200) + 1j * np.arange(200)).astype(np.complex64)          # the reference test's sequence
    x[7] = complex(np.float32(0.1), np.float32(-1e-7))
    p = str(tmp_path / "seq.csv")
    ae.DeviceVec.from_numpy(x).to_csv(p)
    assert open(p).readline().strip() == "0.0,0.0"
    back = ae.DeviceVec.from_csv(p).to_numpy()
    assert back.tobytes() == x.tobytes()


def test_cuda_graph_records_then_replays(ae):
    """ae_graph_*: launches made inside the bracket are recorded, not run; every ae_graph_launch replays them all"""
    x = (np.arange(1000) + 1j).astype(np.complex64)
    a = ae.DeviceVec.from_numpy(x)
    with ae.Graph() as g:
        for _ in range(3):
            a.vec_scale(2.0).flush()
    assert np.array_equal(a.to_numpy(), x)              # nothing ran while recording
    g.launch()
    assert np.array_equal(a.to_numpy(), x * 8)
    g.launch()
    g.launch()
    assert np.array_equal(a.to_numpy(), x * 512)
    # a fused chain step recorded 4 times: the replay gives the bits of a plain call
    from aether_primitives_b200.chain import FftFirDemod

    rng = np.random.default_rng(4)
    xs = (rng.standard_normal(8 * 1024) + 1j * rng.standard_normal(8 * 1024)).astype(np.complex64)
    h = (np.hamming(16) / 16).astype(np.complex64)
    ch = FftFirDemod(1024, h)
    d = ae.DeviceVec.from_numpy(xs)
    want = ae.DeviceBits.with_capacity(1)
    ch.run(d, want)
    got = ae.DeviceBits.zeros(2 * xs.size)              # sized before recording: nothing may reallocate inside
    with ae.Graph() as g2:
        for _ in range(4):
            ch.run(d, got)
    assert not got.to_numpy().any()
    g2.launch()
    assert np.array_equal(got.to_numpy(), want.to_numpy())
    g.close()
    g2.close()
    # VecOps recorded BEFORE the bracket run once, now; VecOps recorded INSIDE and never flushed still belong to the graph
    b = ae.DeviceVec.from_numpy(x)
    b.vec_scale(3.0)                                    # pending when the bracket opens
    with ae.Graph() as g3:
        b.vec_scale(2.0)                                # pending when the bracket closes
    assert np.array_equal(b.to_numpy(), x * 3)
    g3.launch()
    g3.launch()
    assert np.array_equal(b.to_numpy(), x * 12)
    g3.close()
