"""CPU: the C-ABI library loads, exports every symbol include/aether_b200.h declares, fails loudly
without a GPU (no CPU fallback), and the product package never touches oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "aether_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ae_[a-z0-9_]+)\s*\((?!\s*\*)", src)))      # not `ae_status (*ae_stage_fn)(...)`


def test_library_exports_every_declared_symbol():
    from aether_primitives_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) > 80
    for s in syms:
        assert hasattr(lib, s), "libaether_b200.so does not export %s" % s
    # and the Python binding knows each of them
    assert set(syms) == set(_lib.EXPORTS)


def test_no_cpu_fallback_without_gpu():
    import aether_primitives_b200 as ae

    if ae.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ae.AeError) as e:
        ae.init(0)
    assert e.value.status == ae._lib.AE_ECUDA
    with pytest.raises(ae.AeError) as e:
        ae.DeviceVec.zeros(16)
    assert e.value.status == ae._lib.AE_ECUDA
    with pytest.raises(ae.AeError):
        ae.Cfft.with_len(1024)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "aether_primitives_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in txt and "aether_oracle" not in txt and "tests.oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
    mk = open(os.path.join(ROOT, "Makefile")).read()
    lib_rule = mk[mk.index("$(LIB): $(OBJS)"):].split("\n\n")[0]
    assert "oracle" not in lib_rule


def test_philox_host_entry_point_needs_no_gpu():
    from aether_primitives_b200 import noise

    assert noise.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_scale_factor_host_entry_point():
    import aether_primitives_b200 as ae
    from tests import oracle as o

    for n in (1, 4, 100, 1024, 2048, 99999):
        assert ae.Scale.SN.factor(n) == o.scale_factor(o.SCALE_SN, n)
        assert ae.Scale.N.factor(n) == o.scale_factor(o.SCALE_N, n)
    assert ae.Scale.X(2.5).factor(7) == 2.5


def test_rust_sys_binding_is_generated_from_the_current_header():
    """rust/aether-b200-sys/src/lib.rs is emitted by tools/gen_rust_sys.py; it must match the header
    (no Rust toolchain here, so this is the only guard against drift) and declare every export."""
    import re
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    from aether_primitives_b200 import _lib

    text = open(os.path.join(root, "rust", "aether-b200-sys", "src", "lib.rs")).read()
    assert set(re.findall(r"pub fn (ae_\w+)\(", text)) == set(_lib.EXPORTS)


def test_rust_shim_calls_every_entry_point_and_has_no_stubs():
    """rust/aether-b200/src/lib.rs cannot be compiled here (no toolchain), so it is checked mechanically: every ae_*
    function the header declares is called through the -sys crate, nothing is left unimplemented, the reference's
    traits are implemented (VecOps, Fft, Modulation) and every trait method is present."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "rust", "aether-b200", "src", "lib.rs")).read()
    code = re.sub(r"//[^\n]*", "", text)
    called = set(re.findall(r"sys::(ae_[a-z0-9_]+)\s*\(", code))
    missing = sorted(set(header_symbols()) - called)
    assert not missing, "the Rust shim never calls: %s" % ", ".join(missing)
    assert not (called - set(header_symbols())), "the Rust shim calls symbols the header does not declare"
    for stub in ("unimplemented!", "todo!", "unreachable!"):
        assert stub not in code, stub
    assert "unsafe impl Send" not in code and "unsafe impl Sync" not in code      # the device context is not locked
    for impl in ("impl VecOps for DeviceVec", "impl Fft for CudaFft", "impl Modulation for $name"):
        assert impl in code, impl
    # every method of the reference traits (src/vecops.rs:39-89, src/fft.rs:48-77, src/modulation.rs:94-149)
    for m in ("vec_scale", "vec_mul", "vec_div", "vec_conj", "vec_mirror", "vec_clone", "vec_zero", "vec_mutate", "vec_add",
              "vec_sub", "vec_fft", "vec_ifft", "vec_rfft", "vec_rifft", "fwd", "bwd", "ifwd", "ibwd", "tfwd", "tbwd", "len",
              "symbol", "modulate", "modulate_into", "demod_naive"):
        assert re.search(r"fn %s\s*[(<]" % m, code), m
    for f in ("pub fn generator()", "pub fn new(power: f32, seed: u64)", "pub fn set_power", "pub fn apply(", "pub fn fill(",
              "pub fn iter(", "pub fn interpolate(", "pub fn downsample(", "pub fn downsample_sb(", "pub fn expand(", "pub fn generate_taps("):
        assert f in code, f
    # balanced delimiters: the cheapest syntax check available without rustc
    for a, b in ("()", "[]", "{}"):
        assert code.count(a) == code.count(b), "unbalanced %s%s" % (a, b)


def test_cpp_mirror_reaches_every_entry_point():
    """include/aether_b200.hpp is the compiled host side of this build (no Rust toolchain): every ae_* function of the C
    header is called from it, and the reference's trait methods exist under their own names (src/vecops.rs:39-89,
    src/fft.rs:48-77, src/modulation.rs:94-149, src/noise.rs:29-70, src/sampling.rs, src/sequence.rs)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "aether_b200.hpp")).read()
    code = re.sub(r"//[^\n]*", "", text)
    called = set(re.findall(r"\b(ae_[a-z0-9_]+)\s*\(", code))
    missing = sorted(set(header_symbols()) - called)
    assert not missing, "the C++ mirror never calls: %s" % ", ".join(missing)
    for m in ("vec_scale", "vec_mul", "vec_div", "vec_conj", "vec_mirror", "vec_clone", "vec_zero", "vec_mutate", "vec_add",
              "vec_sub", "vec_fft", "vec_ifft", "vec_rfft", "vec_rifft", "fwd", "bwd", "ifwd", "ibwd", "tfwd", "tbwd", "len",
              "modulate", "modulate_into", "demod_naive", "set_power", "apply", "fill", "next", "generator", "interpolate",
              "downsample", "downsample_sb", "expand", "generate", "bpsk", "qpsk"):
        assert re.search(r"\b%s\s*\(" % m, code), m


def test_util_db_matches_the_reference_tests():
    """src/util/mod.rs:14-22 (doctest) and :53-66 (db_to_ratio, ratio_to_db)."""
    from aether_primitives_b200.util import DB

    db = DB.from_ratio(100)
    assert db.ratio() == 100.0 and db.db() == 20.0
    assert DB(30.0).ratio() == 1000.0 and DB(0.0).ratio() == 1.0
    assert abs(DB.from_ratio(100.0).db() - 20.0) < 1e-6 and abs(DB.from_ratio(0.1).db() + 10.0) < 1e-6
    assert DB.from_ratio(0.0).db() == float("-inf")


def test_committed_bench_line_carries_the_contract_keys():
    """The latest bench line under profiles/ (written by `python bench.py` on a B200) has every key the
    bench contract names; guards against a key being dropped by a later edit of bench.py."""
    import glob
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = sorted(glob.glob(os.path.join(root, "profiles", "r[0-9]_bench_v[0-9]*.json")), key=lambda p: (int(re.findall(r"r(\d+)_bench", p)[0]), int(re.findall(r"_v(\d+)", p)[0])))
    d = json.load(open(lines[-1]))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert "workload" in d["config"] and d["gpu_launches"] > 0 and d["n_gpus"] == 1
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9


def test_util_assert_evm_agrees_with_the_oracle_restatement():
    """assert_evm! (src/lib.rs:26-49; its tests :87-118): same pass/fail and same first failing element as the oracle."""
    import numpy as np

    from aether_primitives_b200.util import assert_evm
    from tests import oracle as o

    ones = np.full(100, 1 + 1j, np.complex64)
    assert_evm(ones, ones)                                              # evm_passes
    assert_evm(ones, ones * np.float32(1 + 1e-9), -80.0)
    with pytest.raises(AssertionError, match="EVM limit exceeded"):     # evm_fails
        assert_evm(ones, ones * np.float32(1.1), -80.0)
    with pytest.raises(AssertionError, match="same length"):
        assert_evm(ones, ones[:-1])
    with pytest.raises(AssertionError, match="must be negative"):
        assert_evm(ones, ones, 3.0)
    rng = np.random.default_rng(3)
    for db in (-80.0, -40.0, -20.0, -6.0):
        r = (rng.standard_normal(500) + 1j * rng.standard_normal(500)).astype(np.complex64)
        a = (r * np.float32(1 + 10 ** (db / 10)) + np.float32(1e-7)).astype(np.complex64)   # straddles the limit
        rc, bad = o.assert_evm(a, r, db)
        try:
            assert_evm(a, r, db)
            assert rc == o.OK
        except AssertionError as ex:
            assert rc == o.EVM_EXCEEDED and ("for element %d." % bad) in str(ex)
