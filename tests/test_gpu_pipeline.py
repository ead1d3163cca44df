"""General stage pipeline (ae_pipeline_*, src/pipeline.rs:26-137) with pooled pinned/device buffers (src/pool.rs): three
stages on three CUDA streams — H2D copy, transform + VecOps + demod, D2H copy — against the same work done call by call."""
import numpy as np
import pytest

from tests import oracle as o

pytestmark = pytest.mark.gpu


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


class Block:
    """what travels through the pipeline: pinned staging buffers + device buffers, made once by the pool"""

    def __init__(self, ae, n):
        self.n = n
        self.h_in = ae.PinnedBuf(n, "complex64")
        self.h_bits = ae.PinnedBuf(2 * n, "uint8")
        self.d = ae.DeviceVec.zeros(n)
        self.bits = ae.DeviceBits.with_capacity(2 * n)      # demod_naive appends, like the reference's output.extend
        self.tag = None


def test_three_stage_pipeline_matches_serial_calls(ae):
    n, frames, blocks, depth = 1024, 64, 7, 3
    ns = n * frames
    fft = ae.Cfft.with_len(n)                       # used by the middle stage only
    w = ae.DeviceVec.from_numpy(rnd(ns, 99))
    qpsk = ae.modulation.qpsk()
    made = []

    def maker():
        b = Block(ae, ns)
        made.append(b)
        return b

    def resetter(b):                                # what `|o| o.clear()` is in the reference's pool examples
        b.tag = None
        b.bits.clear()

    pool = ae.pool.make(0, maker, resetter)

    def h2d(e):
        e.val.d.upload_async(e.val.h_in.ptr, ns)
        return e

    def work(e):
        b = e.val
        fft.ifwd(b.d, ae.Scale.SN, howmany=frames)
        b.d.vec_mul(w).vec_conj()                   # recorded; the pipeline flushes it on this stage's stream
        qpsk.demod_naive(b.d, b.bits)
        return e

    def d2h(e):
        e.val.bits.download_async(e.val.h_bits.ptr, 2 * ns)
        return e

    tx, rx = ae.pipeline.Pipeline.new("h2d", h2d, depth=depth).add_stage("fft-mul-demod", work).add_stage("d2h", d2h).finish()
    xs = [rnd(ns, 10 + i) for i in range(blocks)]
    got = []

    def take_result():
        e = rx.recv()
        got.append((e.val.tag, e.val.h_bits.array.copy()))
        e.release()

    for i, x in enumerate(xs):
        if rx.in_flight() == depth:
            take_result()
        e = pool.take_or_make()
        e.val.tag = i
        e.val.h_in.array[:] = x
        tx.send(e)
    while rx.in_flight():
        take_result()
    assert [t for t, _ in got] == list(range(blocks))              # in order
    assert pool.cap() <= depth and pool.len() == pool.cap()         # buffers were reused, all returned
    # the same work, call by call
    for i, x in enumerate(xs):
        d = ae.DeviceVec.from_numpy(x)
        fft.ifwd(d, ae.Scale.SN, howmany=frames)
        d.vec_mul(w).vec_conj()
        bits = ae.DeviceBits.with_capacity(1)
        qpsk.demod_naive(d, bits)
        assert np.array_equal(got[i][1], bits.to_numpy()), "block %d" % i
    rep = rx.report()
    assert [r["name"] for r in rep] == ["h2d", "fft-mul-demod", "d2h"]
    assert all(r["processed"] == blocks and r["active_ms"] > 0 and 0 < r["utilisation_pct"] <= 100.0 for r in rep)
    assert rx.report_lines()[0].startswith("Stage: h2d")
    rx.report(reset=True)
    assert rx.report()[0]["processed"] == 0


def test_stage_error_and_empty_recv(ae):
    def boom(x):
        raise ValueError("stage failed")

    tx, rx = ae.pipeline.Pipeline.new("a", lambda x: x, depth=2).add_stage("b", boom).finish()
    with pytest.raises(ValueError):
        tx.send(1)
    assert rx.in_flight() == 0
    with pytest.raises(ae.AeError):
        rx.recv()
    # values flow from stage to stage (FnMut(I) -> O)
    tx, rx = ae.pipeline.Pipeline.new("inc", lambda x: x + 1, depth=2).add_stage("dbl", lambda x: 2 * x).finish()
    for i in range(5):
        tx.send(i)
        if rx.in_flight() == 2:
            assert rx.recv() == 2 * (i - 1 + 1)
    assert rx.recv() == 10


@pytest.mark.parametrize("depth", [1, 2, 5])
def test_many_stages_and_depths(ae, depth):
    """five VecOps stages on five streams; every depth, more items than slots: each item sees every stage exactly once, in order"""
    n, items = 4096, 11
    xs = [rnd(n, 300 + i) for i in range(items)]
    p = ae.pipeline.Pipeline.new("s0", lambda v: v.vec_scale(2.0), depth=depth)
    for k in range(1, 5):
        p = p.add_stage("s%d" % k, (lambda kk: (lambda v: v.vec_scale(float(kk + 2))))(k))
    tx, rx = p.finish()
    out = []
    for x in xs:
        if rx.in_flight() == depth:
            out.append(rx.recv().to_numpy())
        tx.send(ae.DeviceVec.from_numpy(x))
    while rx.in_flight():
        out.append(rx.recv().to_numpy())
    assert len(out) == items
    for x, y in zip(xs, out):
        want = x.copy()
        for f in (2.0, 3.0, 4.0, 5.0, 6.0):
            want = (want * np.float32(f)).astype(np.complex64)
        assert np.array_equal(y, want)
    rep = rx.report()
    assert len(rep) == 5 and all(r["processed"] == items for r in rep)
