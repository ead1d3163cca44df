"""GPU parity, tier T0 (bit-exact): fused VecOps tape vs the oracle (src/vecops.rs:94-182)."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import cx, load, same_bits

pytestmark = pytest.mark.gpu
G = load()


def rnd(n, seed, special=True):
    rng = np.random.default_rng(seed)
    v = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    if special and n >= 64:
        sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3e38, -3e38, 1e-38, 1.0, -1.0], dtype=np.float32)
        idx = rng.integers(0, n, 24)
        v.real[idx] = sp[rng.integers(0, sp.size, 24)]
        idx = rng.integers(0, n, 24)
        v.imag[idx] = sp[rng.integers(0, sp.size, 24)]
    return v


@pytest.mark.parametrize("case", G["vecops"], ids=lambda c: c["cite"])
def test_vecops_golden(ae, case):
    v = ae.DeviceVec.from_numpy(cx(case["v"]))
    op = case["op"]
    if op == "scale":
        v.vec_scale(case["s"])
    elif op in ("mul", "div", "add", "sub", "clone"):
        getattr(v, "vec_" + op)(ae.DeviceVec.from_numpy(cx(case["o"])))
    else:
        getattr(v, "vec_" + op)()
    assert same_bits(v.to_numpy(), cx(case["want"]))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 1023, 4096, 100003])
@pytest.mark.parametrize("op", ["scale", "mul", "div", "add", "sub", "conj", "mirror", "clone", "zero"])
def test_single_op_bit_exact(ae, n, op):
    a, b = rnd(n, 10 + n), rnd(n, 20 + n)
    v = ae.DeviceVec.from_numpy(a)
    if op == "scale":
        v.vec_scale(1.7)
        want = o.vec_scale(a, 1.7)
    elif op in ("mul", "div", "add", "sub", "clone"):
        getattr(v, "vec_" + op)(ae.DeviceVec.from_numpy(b))
        want = getattr(o, "vec_" + op)(a, b)
    else:
        getattr(v, "vec_" + op)()
        want = getattr(o, "vec_" + op)(a)
    assert v.pending_ops() == 1
    assert same_bits(v.to_numpy(), want)


def test_length_mismatch_is_the_reference_panic(ae):
    v = ae.DeviceVec.zeros(4)
    w = ae.DeviceVec.zeros(5)
    for name in ("mul", "div", "add", "sub", "clone"):
        with pytest.raises(ae.AeError) as e:
            getattr(v, "vec_" + name)(w)
        assert e.value.status == ae._lib.AE_ELEN
        assert "Vectors must have same length" in e.value.message  # src/vecops.rs:100-104


def test_doctest_chain_fused_into_one_kernel(ae):
    c = G["vecops_chain"]
    v = ae.DeviceVec.from_numpy(cx(c["v"]))
    twos = ae.DeviceVec.from_numpy(cx(c["twos"]))
    ones = ae.DeviceVec.from_numpy(cx(c["ones"]))
    before = ae.launch_count()
    v.vec_div(twos).vec_mul(twos).vec_zero().vec_add(ones).vec_sub(twos).vec_clone(ones)
    assert ae.launch_count() == before            # nothing ran yet: ops are on the tape
    v.vec_mutate(lambda z: complex(z.real, -1.0))  # host slow path: flushes (1 launch), round trip
    v.vec_conj().vec_mirror()
    got = v.to_numpy()
    assert ae.launch_count() == before + 2
    assert o.assert_evm(got, cx(c["want"]), c["db"])[0] == o.OK


@pytest.mark.parametrize("n", [2, 5, 64, 1001, 65536 + 3])
def test_random_chains_bit_exact(ae, n):
    rng = np.random.default_rng(n)
    ops = ["scale", "mul", "div", "add", "sub", "conj", "mirror", "clone", "zero"]
    for trial in range(6):
        a = rnd(n, 100 + trial)
        operands = [rnd(n, 200 + trial * 10 + i, special=False) for i in range(4)]
        dv = ae.DeviceVec.from_numpy(a)
        dops = [ae.DeviceVec.from_numpy(x) for x in operands]
        want = a.copy()
        before = ae.launch_count()
        k = int(rng.integers(2, 15))
        for _ in range(k):
            op = ops[int(rng.integers(0, len(ops) - (2 if trial < 4 else 0)))]  # mostly no clone/zero
            if op == "scale":
                s = float(np.float32(rng.uniform(-2, 2)))
                dv.vec_scale(s)
                want = o.vec_scale(want, s)
            elif op in ("mul", "div", "add", "sub", "clone"):
                i = int(rng.integers(0, 4))
                getattr(dv, "vec_" + op)(dops[i])
                want = getattr(o, "vec_" + op)(want, operands[i])
            else:
                getattr(dv, "vec_" + op)()
                want = getattr(o, "vec_" + op)(want)
        got = dv.to_numpy()
        assert ae.launch_count() == before + 1, "a chain must run as ONE fused kernel"
        assert same_bits(got, want)


def test_tape_overflow_and_operand_mutation_order(ae):
    n = 1000
    a, b = rnd(n, 1, False), rnd(n, 2, False)
    v, w = ae.DeviceVec.from_numpy(a), ae.DeviceVec.from_numpy(b)
    want = a.copy()
    for i in range(40):  # > 16 ops: the tape flushes itself
        v.vec_add(w)
        want = o.vec_add(want, b)
    # operand changes AFTER being recorded: the recorded op must see the old value
    v.vec_mul(w)
    want = o.vec_mul(want, b)
    w.vec_scale(3.0)
    v.vec_sub(w)
    want = o.vec_sub(want, o.vec_scale(b, 3.0))
    assert same_bits(v.to_numpy(), want)
    assert same_bits(w.to_numpy(), o.vec_scale(b, 3.0))


def test_views_behave_like_slices(ae):
    a = rnd(64, 5, False)
    v = ae.DeviceVec.from_numpy(a)
    lo = v.view(0, 31)   # odd length, 8-byte aligned only from element 1 on
    hi = v.view(31, 64)
    lo.vec_mirror()
    hi.vec_conj().vec_scale(2.0)
    want = a.copy()
    want[:31] = o.vec_mirror(a[:31])
    want[31:] = o.vec_scale(o.vec_conj(a[31:]), 2.0)
    assert same_bits(v.to_numpy(), want)
    with pytest.raises(ae.AeError):
        v.view(10, 65)


def test_scale_kinds(ae):
    s = G["scale"]
    for kind, key in ((ae.Scale.None_, "none"), (ae.Scale.SN, "sn"), (ae.Scale.N, "n"), (ae.Scale.X(2.0), "x2")):
        v = ae.DeviceVec.from_numpy(cx(s["v"]))
        kind.scale(v)
        assert same_bits(v.to_numpy(), cx(s[key]))
    for n in (3, 100, 1000, 1024, 12345):
        assert ae.Scale.SN.factor(n) == o.scale_factor(o.SCALE_SN, n)
        assert ae.Scale.N.factor(n) == o.scale_factor(o.SCALE_N, n)


def test_large_fused_chain_property(ae):
    """2^24 samples: mul . conj . mirror (BASELINE config 4 chain) against the oracle."""
    n = 1 << 24
    a, b = rnd(n, 7, False), rnd(n, 8, False)
    v, w = ae.DeviceVec.from_numpy(a), ae.DeviceVec.from_numpy(b)
    v.vec_mul(w).vec_conj().vec_mirror()
    want = o.vec_mirror(o.vec_conj(o.vec_mul(a, b)))
    assert same_bits(v.to_numpy(), want)
    # mirror twice = identity; conj twice = identity
    v.vec_mirror().vec_conj().vec_conj().vec_mirror()
    assert same_bits(v.to_numpy(), want)


def test_raw_sample_file_round_trip(ae, tmp_path):
    """util::file format (src/util/file.rs): raw native-endian cf32 back to back, no header."""
    x = rnd(3 * (8 << 20) // 2 + 12345, 77, special=False)     # > one 64 MiB staging chunk
    p = tmp_path / "samples.cf32"
    x.tofile(p)
    v = ae.DeviceVec.from_file(str(p))
    assert len(v) == x.size and same_bits(v.to_numpy(), x)
    v.vec_conj()
    q = tmp_path / "out.cf32"
    v.to_file(str(q))
    assert same_bits(np.fromfile(q, dtype=np.complex64), o.vec_conj(x))
    bad = tmp_path / "bad.cf32"
    bad.write_bytes(b"\\x00" * 13)
    with pytest.raises(ae.AeError) as e:
        ae.DeviceVec.from_file(str(bad))
    assert "integer number of the requested struct" in e.value.message   # src/util/file.rs:19-22
