"""CPU: the oracle (oracle/aether_oracle.cpp) against every golden vector the reference's own
tests and doctests hold for the hot path, plus the quirk regression cases.
This is what pins the oracle before the CUDA path is compared with it."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import cx, evm_db, load, same_bits, truth

G = load()


@pytest.mark.parametrize("case", G["evm"], ids=lambda c: c["cite"])
def test_assert_evm_macro(case):
    rc, _ = o.assert_evm(cx(case["act"]), cx(case["ref"]), case["db"])
    assert (rc == o.OK) == case["pass"]


def test_assert_evm_arg_checks():
    assert o.assert_evm(cx([[1, 0]]), cx([[1, 0], [1, 0]]))[0] == o.ELEN       # src/lib.rs:34
    assert o.assert_evm(cx([[1, 0]]), cx([[1, 0]]), 3.0)[0] == o.EARG          # src/lib.rs:35


@pytest.mark.parametrize("case", G["vecops"], ids=lambda c: c["cite"])
def test_vecops_golden(case):
    v = cx(case["v"])
    op = case["op"]
    if op == "scale":
        got = o.vec_scale(v, case["s"])
    elif op in ("mul", "div", "add", "sub", "clone"):
        got = getattr(o, "vec_" + op)(v, cx(case["o"]))
    else:
        got = getattr(o, "vec_" + op)(v)
    assert same_bits(got, cx(case["want"]))
    assert o.assert_evm(got, cx(case["want"]), -80.0)[0] == o.OK  # the reference's own assertion


def test_vecops_length_mismatch_panics():
    for name in ("mul", "div", "add", "sub", "clone"):
        with pytest.raises(o.OracleError) as e:
            getattr(o, "vec_" + name)(np.ones(4, np.complex64), np.ones(5, np.complex64))
        assert e.value.status == o.ELEN


def test_vecops_doctest_chain():
    c = G["vecops_chain"]
    v, twos, ones = cx(c["v"]), cx(c["twos"]), cx(c["ones"])
    v = o.vec_div(v, twos)
    v = o.vec_mul(v, twos)
    v = o.vec_zero(v)
    v = o.vec_add(v, ones)
    v = o.vec_sub(v, twos)
    v = o.vec_clone(v, ones)
    v = (v.real - 1j).astype(np.complex64)  # vec_mutate(|c| c.im = -1.0)
    v = o.vec_conj(v)
    v = o.vec_mirror(v)
    assert o.assert_evm(v, cx(c["want"]), c["db"])[0] == o.OK


def test_scale_golden():
    s = G["scale"]
    v = cx(s["v"])
    assert same_bits(o.scale(o.SCALE_NONE, v), cx(s["none"]))
    assert same_bits(o.scale(o.SCALE_SN, v), cx(s["sn"]))
    assert same_bits(o.scale(o.SCALE_N, v), cx(s["n"]))
    assert same_bits(o.scale(o.SCALE_X, v, 2.0), cx(s["x2"]))
    # SN is sqrt().recip() then multiply, computed in f32 (src/fft.rs:26)
    assert o.scale_factor(o.SCALE_SN, 1000) == float(np.float32(1.0) / np.sqrt(np.float32(1000.0)))


def test_fft_roundtrip_100():
    c = G["fft_roundtrip_100"]
    v = cx(c["v"])
    for compat in (o.REFERENCE, o.CORRECTED):
        f = o.cfft(v, c["n"], bwd=False, scale_kind=o.SCALE_SN, compat=compat)
        b = o.cfft(f, c["n"], bwd=True, scale_kind=o.SCALE_SN, compat=compat)
        assert evm_db(b, v) < -120.0
        # the reference's own assertion (src/vecops.rs:445-463), literally: assert_evm!(c, v) at the default -80,
        # i.e. |err| <= 1e-8 |ref| = equality to the bit.  It holds because the odd butterflies (rustfft's form,
        # restated in the oracle) give exact zeros for a constant input.
        assert o.assert_evm(b, v, -80.0)[0] == o.OK
        assert np.array_equal(b.view(np.uint32), v.view(np.uint32))


def test_fft_doctest_128():
    c = G["fft_doctest_128"]
    v = cx(c["v"])
    spec = o.cfft(v, 128, bwd=False)
    assert o.assert_evm(spec, cx(c["spectrum"]), -80.0)[0] == o.OK   # exact: DC = 128, all other bins 0
    back = o.cfft(spec, 128, bwd=True, scale_kind=o.SCALE_N)
    assert o.assert_evm(back, v, -80.0)[0] == o.OK
    d = o.cfft(back, 128, bwd=False, scale_kind=o.SCALE_SN)
    d = o.vec_scale(d, 2.0)
    d = o.cfft(d, 128, bwd=True, scale_kind=o.SCALE_SN)
    assert o.assert_evm(d, np.full(128, 2 + 0j, np.complex64), c["twos_db"])[0] == o.OK


def test_fft_sign_quirk_f3():
    n = G["fft_sign_quirk"]["n"]
    x = np.zeros(n, np.complex64)
    x[1] = 1
    k = np.arange(n)
    fwd = o.cfft(x, n, bwd=False, compat=o.REFERENCE)
    bwd = o.cfft(x, n, bwd=True, compat=o.REFERENCE)
    assert np.allclose(fwd, np.exp(+2j * np.pi * k / n), atol=1e-6)
    assert np.allclose(bwd, np.exp(-2j * np.pi * k / n), atol=1e-6)
    assert np.allclose(o.cfft(x, n, bwd=False, compat=o.CORRECTED), np.exp(-2j * np.pi * k / n), atol=1e-6)


@pytest.mark.parametrize("n", [8, 16, 64, 100, 128, 256, 360, 512, 1009, 1024, 2048, 4096, 8192])
def test_fft_against_f64_truth(n):
    t = truth()
    x = t["x_%d" % n]
    for sign, key in ((-1, "neg"), (+1, "pos")):
        got = o.fft_raw(x, sign)
        assert evm_db(got, t["%s_%d" % (key, n)]) < -120.0
        assert evm_db(o.fft_raw_f64(x, sign), t["%s_%d" % (key, n)]) < -250.0


@pytest.mark.parametrize("case", G["interpolate"], ids=lambda c: c["cite"])
def test_interpolate_golden(case):
    got = o.interpolate(cx(case["src"]), case["k"])
    assert len(got) == len(case["src"]) + (len(case["src"]) - 1) * case["k"]
    assert same_bits(got, cx(case["want"]))  # reference uses assert_eq! (exact)


def test_interpolate_quirk_f4():
    q = G["interpolate_quirk"]
    assert same_bits(o.interpolate(cx(q["src"]), q["k"], o.REFERENCE), cx(q["reference"]))
    assert same_bits(o.interpolate(cx(q["src"]), q["k"], o.CORRECTED), cx(q["corrected"]))
    with pytest.raises(o.OracleError):
        o.interpolate(np.zeros(0, np.complex64), 1)  # src.last().unwrap() on empty (:23)


@pytest.mark.parametrize("case", G["downsample"], ids=lambda c: c["cite"])
def test_downsample_golden(case):
    src = np.array(case["src"], dtype=np.int32)
    if case["want"] is None:
        with pytest.raises(o.OracleError) as e:
            o.downsample(src, case["n_dst"], strict=True)
        assert e.value.status == o.ELEN
    else:
        assert o.downsample(src, case["n_dst"]).tolist() == case["want"]


def test_downsample_edge_cases():
    with pytest.raises(o.OracleError):
        o.downsample(np.arange(4, dtype=np.int32), 0)   # division by zero (:38)
    # release-build semantics: dst longer than src -> dec = 0 -> every output is src[0] (App. A.6)
    assert o.downsample(np.arange(3, dtype=np.int32) + 7, 5, strict=False).tolist() == [7] * 5


@pytest.mark.parametrize("case", G["modulate"], ids=lambda c: c["cite"])
def test_modulate_golden(case):
    table = o.BPSK if case["table"] == "bpsk" else o.QPSK
    assert same_bits(o.modulate(table, case["bits"]), cx(case["want"]))


def test_modulate_panics():
    with pytest.raises(o.OracleError) as e:
        o.modulate(o.QPSK, [0, 1, 1])          # ragged tail: bits[1] out of bounds (:24)
    assert e.value.status == o.EIDX
    with pytest.raises(o.OracleError):
        o.modulate(o.BPSK, [0, 2])             # table index 2 of a 2-entry table (:14)
    # modulate_into zips with the output and silently truncates (:123-131)
    assert len(o.modulate(o.QPSK, [0, 0, 1, 0, 0, 1], out_cap=2)) == 2


def test_naive_demod_golden():
    bits = np.array(G["naive_demod"]["bits"], dtype=np.uint8)
    sym = o.modulate(o.QPSK, bits)
    assert o.demod(o.QPSK, sym).tolist() == bits.tolist()


@pytest.mark.parametrize("case", G["demod_quirks"], ids=lambda c: c["cite"])
def test_demod_quirks(case):
    sym = cx(case["sym"])
    assert o.demod(o.QPSK, sym, o.REFERENCE).tolist() == case["reference"]
    assert o.demod(o.QPSK, sym, o.CORRECTED).tolist() == case["corrected"]


def test_demod_roundtrip_all_symbols():
    rng = np.random.default_rng(815)
    bits = rng.integers(0, 2, 4000, dtype=np.uint8)
    q = o.demod(o.QPSK, o.modulate(o.QPSK, bits), o.CORRECTED)
    assert q.tolist() == bits.tolist()
    b = o.demod(o.BPSK, o.modulate(o.BPSK, bits))
    assert b.tolist() == bits.tolist()


def test_sequence_golden():
    e = G["expand"]
    assert o.expand(e["seed"], e["len"]).tolist() == e["want"]
    g = G["generate"]
    assert o.mseq_generate(g["init"], g["back"], g["len"]).tolist() == g["want"]
    l = G["generate_lte"]
    seq = o.mseq_generate(o.expand(l["seed"], 31), l["back"], l["len"])
    assert len(seq) == l["len"]
    # independent restatement of the closure |n, seq| (seq[n-28] + seq[n-31]) % 2
    ref = list(o.expand(l["seed"], 31))
    while len(ref) < l["len"]:
        n = len(ref)
        ref.append((ref[n - 28] + ref[n - 31]) % 2)
    assert seq.tolist() == ref
    with pytest.raises(o.OracleError):
        o.expand(1, 65)  # seed >> 64 overflows


@pytest.mark.parametrize("case", G["philox_kat"], ids=lambda c: str(c["ctr"][0]))
def test_philox_kat(case):
    assert o.philox(case["ctr"], case["key"]) == case["want"]


def test_awgn_oracle_statistics():
    n = 1 << 18
    z = o.awgn_fill(n, 0.25, 815)
    for comp in (z.real, z.imag):
        assert abs(comp.mean()) < 4 * 0.5 / np.sqrt(n)
        assert abs(comp.var() / 0.25 - 1) < 4 * np.sqrt(2.0 / n)
    # apply(): reference scales twice -> sigma = power
    s = o.awgn_apply(np.zeros(n, np.complex64), 0.25, 815, compat=o.REFERENCE)
    assert abs(s.real.var() / 0.25**2 - 1) < 4 * np.sqrt(2.0 / n)
    s = o.awgn_apply(np.zeros(n, np.complex64), 0.25, 815, compat=o.CORRECTED)
    assert abs(s.real.var() / 0.25 - 1) < 4 * np.sqrt(2.0 / n)


def test_fir_oracle_matches_numpy():
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(500) + 1j * rng.standard_normal(500)).astype(np.complex64)
    h = (rng.standard_normal(17) + 1j * rng.standard_normal(17)).astype(np.complex64)
    ref = np.convolve(x.astype(np.complex128), h.astype(np.complex128))[:500]
    assert evm_db(o.fir(x, h), ref) < -120
    assert evm_db(o.fir_f64(x, h), ref) < -280
    # framed: zero state at every frame start
    yf = o.fir(x, h, frame_len=100)
    for f in range(5):
        seg = np.convolve(x[100 * f:100 * f + 100].astype(np.complex128), h.astype(np.complex128))[:100]
        assert evm_db(yf[100 * f:100 * f + 100], seg) < -120
    # carried history == filtering the concatenation
    y2 = o.fir(x[250:], h, state=x[250 - 16:250])
    assert evm_db(y2, ref[250:]) < -120


def test_chain_oracle_consistency():
    rng = np.random.default_rng(3)
    n, frames = 256, 3
    x = (rng.standard_normal(n * frames) + 1j * rng.standard_normal(n * frames)).astype(np.complex64)
    h = (rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex64)
    bits, sym = o.chain_fft_fir_demod(x, n, h)
    bits2, _ = o.chain_fft_fir_demod(x, n, h, nthreads=3)
    assert bits.tolist() == bits2.tolist()
    X = o.cfft(x, n, scale_kind=o.SCALE_SN)
    assert same_bits(sym, o.fir(X, h, frame_len=n))
    assert bits.tolist() == o.demod(o.QPSK, sym).tolist()
