"""GPU parity, tier T0 (bit-exact): sampling, modulation, hard demod, sequence, statistics."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import cx, load, same_bits

pytestmark = pytest.mark.gpu
G = load()


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


# ---------------------------------------------------------------- sampling (src/sampling.rs)
@pytest.mark.parametrize("case", G["interpolate"], ids=lambda c: c["cite"])
def test_interpolate_golden(ae, case):
    src = ae.DeviceVec.from_numpy(cx(case["src"]))
    dst = ae.DeviceVec.with_capacity(1)
    ae.sampling.interpolate(src, dst, case["k"])
    assert len(dst) == len(case["src"]) + (len(case["src"]) - 1) * case["k"]
    assert same_bits(dst.to_numpy(), cx(case["want"]))


def test_interpolate_quirk_f4_and_append(ae):
    q = G["interpolate_quirk"]
    src = ae.DeviceVec.from_numpy(cx(q["src"]))
    dst = ae.DeviceVec.with_capacity(2)
    ae.sampling.interpolate(src, dst, q["k"], ae.COMPAT_REFERENCE)
    ae.sampling.interpolate(src, dst, q["k"], ae.COMPAT_CORRECTED)  # appends, like Vec::push (:16-23)
    assert same_bits(dst.to_numpy(), np.concatenate([cx(q["reference"]), cx(q["corrected"])]))
    with pytest.raises(ae.AeError):
        ae.sampling.interpolate(ae.DeviceVec.zeros(0), dst, 1)       # src.last().unwrap() (:23)


@pytest.mark.parametrize("n", [1, 2, 3, 400, 1024, 100001])
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 7, 10])
def test_interpolate_vs_oracle(ae, n, k):
    x = rnd(n, n * 31 + k)
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        dst = ae.DeviceVec.with_capacity(1)
        ae.sampling.interpolate(ae.DeviceVec.from_numpy(x), dst, k, compat)
        assert same_bits(dst.to_numpy(), o.interpolate(x, k, compat))


@pytest.mark.parametrize("case", G["downsample"], ids=lambda c: c["cite"])
def test_downsample_golden(ae, case):
    src = np.array(case["src"], dtype=np.float32)
    for fn in (ae.sampling.downsample, ae.sampling.downsample_sb):
        d_src = ae.DeviceVec.from_numpy(src + 0j)
        d_dst = ae.DeviceVec.zeros(case["n_dst"])
        if case["want"] is None:
            with pytest.raises(ae.AeError) as e:
                fn(d_src, d_dst)
            assert e.value.status == ae._lib.AE_ELEN and "Only even decimations are supported" in e.value.message
        else:
            fn(d_src, d_dst)
            assert d_dst.to_numpy().real.tolist() == case["want"]
    if case["want"] is not None:  # generic T: Copy — bytes
        b = ae.DeviceBits.from_numpy(np.array(case["src"], dtype=np.uint8))
        out = ae.DeviceBits.zeros(case["n_dst"])
        ae.sampling.downsample(b, out)
        assert out.to_numpy().tolist() == case["want"]


@pytest.mark.parametrize("n_dst,dec", [(1, 1), (7, 3), (1024, 30), (512, 16), (4097, 4), (100000, 2), (65536, 4), (3, 1000)])
def test_downsample_vs_oracle(ae, n_dst, dec):
    x = rnd(n_dst * dec, n_dst)
    d = ae.DeviceVec.zeros(n_dst)
    ae.sampling.downsample(ae.DeviceVec.from_numpy(x), d)
    assert same_bits(d.to_numpy(), o.downsample(x, n_dst))


def test_downsample_edge_cases(ae):
    with pytest.raises(ae.AeError):
        ae.sampling.downsample(ae.DeviceVec.zeros(4), ae.DeviceVec.zeros(0))   # divide by zero (:38)
    x = rnd(3, 1)
    d = ae.DeviceVec.zeros(5)
    ae.sampling.downsample(ae.DeviceVec.from_numpy(x), d, strict=False)       # dec = 0 -> src[0] everywhere
    assert same_bits(d.to_numpy(), o.downsample(x, 5, strict=False))


# ---------------------------------------------------------------- modulation (src/modulation.rs)
@pytest.mark.parametrize("case", G["modulate"], ids=lambda c: c["cite"])
def test_modulate_golden(ae, case):
    m = ae.modulation.bpsk() if case["table"] == "bpsk" else ae.modulation.qpsk()
    out = m.modulate(ae.DeviceBits.from_numpy(case["bits"]))
    assert same_bits(out.to_numpy(), cx(case["want"]))


def test_naive_demod_golden(ae):
    bits = np.array(G["naive_demod"]["bits"], dtype=np.uint8)
    m = ae.modulation.qpsk()
    sym = m.modulate(ae.DeviceBits.from_numpy(bits))
    out = ae.DeviceBits.with_capacity(100)
    m.demod_naive(sym, out)
    assert out.to_numpy().tolist() == bits.tolist()


@pytest.mark.parametrize("case", G["demod_quirks"], ids=lambda c: c["cite"])
def test_demod_quirks(ae, case):
    m = ae.modulation.qpsk()
    for compat, key in ((ae.COMPAT_REFERENCE, "reference"), (ae.COMPAT_CORRECTED, "corrected")):
        out = ae.DeviceBits.with_capacity(2)
        m.demod_naive(ae.DeviceVec.from_numpy(cx(case["sym"])), out, compat)
        assert out.to_numpy().tolist() == case[key]


@pytest.mark.parametrize("nbits", [2, 4, 6, 10, 1000, 8000, 100002])
@pytest.mark.parametrize("kind", ["bpsk", "qpsk"])
def test_modulate_demod_vs_oracle(ae, nbits, kind):
    rng = np.random.default_rng(nbits)
    bits = rng.integers(0, 2, nbits, dtype=np.uint8)
    m, table = (ae.modulation.bpsk(), o.BPSK) if kind == "bpsk" else (ae.modulation.qpsk(), o.QPSK)
    assert m.bits_per_symbol() == (1 if kind == "bpsk" else 2)
    sym = m.modulate(ae.DeviceBits.from_numpy(bits))
    assert same_bits(sym.to_numpy(), o.modulate(table, bits))
    # modulate_into truncates silently (:123-131)
    short = ae.DeviceVec.zeros(max(1, len(sym) // 2))
    m.modulate_into(ae.DeviceBits.from_numpy(bits), short)
    assert same_bits(short.to_numpy(), o.modulate(table, bits, out_cap=len(short)))
    # hard decisions on noisy + adversarial symbols, appended to existing output
    s = sym.to_numpy() + 0.7 * (rng.standard_normal(len(sym)) + 1j * rng.standard_normal(len(sym))).astype(np.complex64)
    adv = np.array([0, -0.0, 1e-9, -1e-9, 3e-8, -3e-8, 1e30, -1e30, np.nan, np.inf, -np.inf, 1.0, -1.0, 1e-45], dtype=np.float32)
    k = min(len(s), 200)
    s.real[:k] = adv[rng.integers(0, adv.size, k)]
    s.imag[:k] = adv[rng.integers(0, adv.size, k)]
    for compat in (ae.COMPAT_REFERENCE, ae.COMPAT_CORRECTED):
        out = ae.DeviceBits.from_numpy(np.array([9, 9, 9], dtype=np.uint8))  # unaligned append position
        m.demod_naive(ae.DeviceVec.from_numpy(s), out, compat)
        got = out.to_numpy()
        assert got[:3].tolist() == [9, 9, 9]
        assert got[3:].tolist() == o.demod(table, s, compat).tolist()


def test_modulate_panics(ae):
    q = ae.modulation.qpsk()
    with pytest.raises(ae.AeError) as e:
        q.modulate(ae.DeviceBits.from_numpy([0, 1, 1]))      # ragged tail -> bits[1] out of bounds (:24)
    assert e.value.status == ae._lib.AE_EIDX
    b = ae.modulation.bpsk()
    b.modulate(ae.DeviceBits.from_numpy([0, 2, 1, 0]))       # index 2 of a 2-entry table (:14)
    with pytest.raises(ae.AeError) as e:
        ae.sync()
    assert e.value.status == ae._lib.AE_EIDX
    ae.sync()  # flag is cleared after being reported
    with pytest.raises(ae.AeError):
        ae.modulation.Modulation(np.ones(3, np.complex64))   # only [cf32;2] and [cf32;4] implement the trait


def test_custom_tables(ae):
    rng = np.random.default_rng(4)
    table = rnd(4, 99)
    m = ae.modulation.Modulation(table)
    bits = rng.integers(0, 2, 4096, dtype=np.uint8)
    assert same_bits(m.modulate(ae.DeviceBits.from_numpy(bits)).to_numpy(), o.modulate(table, bits))
    s = rnd(3000, 5)
    out = ae.DeviceBits.with_capacity(1)
    m.demod_naive(ae.DeviceVec.from_numpy(s), out)
    assert out.to_numpy().tolist() == o.demod(table, s).tolist()


# ---------------------------------------------------------------- sequence (src/sequence.rs)
def test_sequence_golden(ae):
    e = G["expand"]
    assert ae.sequence.expand(e["seed"], e["len"]).to_numpy().tolist() == e["want"]
    g = G["generate"]
    assert ae.sequence.generate(g["init"], g["back"], g["len"]).to_numpy().tolist() == g["want"]
    l = G["generate_lte"]
    init = ae.sequence.expand(l["seed"], 31).to_numpy()
    seq = ae.sequence.generate(init, l["back"], l["len"]).to_numpy()
    assert len(seq) == l["len"]
    assert seq.tolist() == o.mseq_generate(init, l["back"], l["len"]).tolist()
    with pytest.raises(ae.AeError):
        ae.sequence.expand(1, 65)


@pytest.mark.parametrize("back", [[1, 2], [28, 31], [3, 31], [1, 3, 4, 64], [5, 5, 7], [1], [63], [2, 4, 6],
                                  [16], [16, 17], [16, 64], [20, 33, 64], [17, 17, 40], [15, 31], [64]])  # >= 16 everywhere: 16 elements per step
@pytest.mark.parametrize("length", [10, 100, 32769, 300007])
def test_mseq_vs_oracle(ae, back, length):
    rng = np.random.default_rng(sum(back) + length)
    deg = max(back)
    for n_init in (deg, deg + 5):
        init = rng.integers(0, 2, n_init, dtype=np.uint8)
        if init.sum() == 0:
            init[0] = 1
        got = ae.sequence.generate(init, back, length).to_numpy()
        assert got.tolist() == o.mseq_generate(init, back, length).tolist()


def test_mseq_edge_cases(ae):
    # init longer than len: returned unchanged (src/sequence.rs:48)
    assert ae.sequence.generate([1, 0, 1, 1], [1, 2], 2).to_numpy().tolist() == [1, 0, 1, 1]
    # init shorter than the deepest tap: the closure indexes out of bounds
    with pytest.raises(ae.AeError):
        ae.sequence.generate([1], [1, 2], 10)
    # non-binary init bytes are kept verbatim; only their parity feeds the recurrence
    got = ae.sequence.generate([3, 2], [1, 2], 8).to_numpy().tolist()
    assert got == o.mseq_generate([3, 2], [1, 2], 8).tolist()


# ---------------------------------------------------------------- statistics
def test_bit_errors_and_evm(ae):
    from aether_primitives_b200.stats import DeviceStats, evm_db

    rng = np.random.default_rng(6)
    n = 1_000_003
    a = rng.integers(0, 2, n, dtype=np.uint8)
    b = a.copy()
    flips = rng.choice(n, 12345, replace=False)
    b[flips] ^= 1
    a2 = a.copy()
    a2[a2 == 1] = 2  # the reference's QPSK demod emits {0,2}: non-zero means 1
    st = DeviceStats()
    st.count_bit_errors(ae.DeviceBits.from_numpy(a2), ae.DeviceBits.from_numpy(b))
    r = st.read()
    assert r["bit_errors"] == 12345 and r["n_bits"] == n
    x, y = rnd(100000, 1), rnd(100000, 2)
    st.zero()
    st.evm_accumulate(ae.DeviceVec.from_numpy(x), ae.DeviceVec.from_numpy(y))
    r = st.read()
    e = np.sum(np.abs(x.astype(np.complex128) - y) ** 2)
    p = np.sum(np.abs(y.astype(np.complex128)) ** 2)
    assert abs(r["err_pow"] / e - 1) < 1e-5 and abs(r["ref_pow"] / p - 1) < 1e-5
    assert abs(evm_db(r["err_pow"], r["ref_pow"]) - o.evm_power_db(x, y)) < 1e-3
