"""CPU check of K14b's DEVICE code (aether_primitives_b200/csrc/chain_x2.cuh): the kernel body is compiled
for the host by tests/cpp/chain_x2_emu.cpp (std::thread per CUDA thread, barrier per __syncwarp, a model of
mma.sync m16n8k8 TF32 and of the TMA copy) and its decision bytes are compared with the oracle's chain
(Cfft::fwd(scale) -> zero-state FIR -> QPSK demod_naive).  This pins index maps, twiddle rows, the Toeplitz
GEMM of the wrap-around correction and the fragment layouts without a GPU; the GPU parity tests
(tests/test_gpu_chain.py) then check the real kernel."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "build", "chain_x2_emu")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu():
    if not os.path.isdir(CUDA_INC):
        pytest.skip("CUDA headers (vector_types.h) not found")
    src = os.path.join(ROOT, "tests", "cpp", "chain_x2_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "aether_primitives_b200", "csrc", f)
                    for f in ("chain_x2.cuh", "chain_x2_host.h", "fft_device.cuh", "common.cuh")]
    if not os.path.exists(EMU) or any(os.path.getmtime(d) > os.path.getmtime(EMU) for d in deps):
        os.makedirs(os.path.dirname(EMU), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wno-attributes", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"),
                               src, "-o", EMU, "-pthread"])
    return EMU


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def taps(t, seed=3):
    if t == 1:
        return np.array([0.8 - 0.3j], np.complex64)
    rng = np.random.default_rng(seed)
    k = np.arange(t) - (t - 1) / 2
    h = np.sinc(k / 4.3) * np.hamming(t) * np.exp(1j * rng.uniform(0, 2 * np.pi))
    return (h / np.abs(h.sum())).astype(np.complex64)


def run_emu(emu, tmp_path, x, h, n, compat, scale, staged, warps, blocks):
    frames = x.size // n
    inverse = 1 if compat == o.REFERENCE else 0
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<8i", n, h.size, frames, compat, inverse, staged, warps, blocks))
        f.write(struct.pack("<f", scale))
        f.write(np.ascontiguousarray(x, np.complex64).tobytes())
        f.write(np.ascontiguousarray(h, np.complex64).tobytes())
    subprocess.check_call([emu, str(fin), str(fout)], timeout=600)
    return np.fromfile(fout, dtype=np.uint8)


def check(got, want_bits, want_sym, compat):
    assert got.size == want_bits.size
    amp = np.sqrt(np.mean(np.abs(want_sym) ** 2))
    mism = np.nonzero((got != 0) != (want_bits != 0))[0]
    for i in mism:
        s = want_sym[i // 2]
        comp = s.real if i % 2 == 0 else s.imag
        assert abs(comp) < 2e-5 * amp + 1e-30, "bit %d differs away from a boundary (%r)" % (i, s)
    assert len(mism) <= max(3, got.size // 40000)
    same = (got != 0) == (want_bits != 0)
    assert np.array_equal(got[same], want_bits[same])          # byte VALUES too ({0,2} quirk of compat=reference)
    assert set(np.unique(got).tolist()) <= ({0, 1, 2} if compat == o.REFERENCE else {0, 1})


@pytest.mark.parametrize("t", [64, 1, 2, 9, 33, 63])
@pytest.mark.parametrize("compat", [o.REFERENCE, o.CORRECTED])
def test_emulated_kernel_vs_oracle(emu, tmp_path, t, compat):
    n, frames = 1024, 7
    x, h = rnd(n * frames, 100 + t), taps(t)
    want_bits, want_sym = o.chain_fft_fir_demod(x, n, h, scale_kind=o.SCALE_SN, compat=compat)
    scale = np.float32(1.0) / np.sqrt(np.float32(n))
    got = run_emu(emu, tmp_path, x, h, n, compat, float(scale), staged=1, warps=2, blocks=2)
    check(got, want_bits, want_sym, compat)


@pytest.mark.parametrize("compat", [o.REFERENCE, o.CORRECTED])
def test_emulated_lean_build(emu, tmp_path, compat):
    """the 14+ warp build: plain loads, first-stage twiddles from the shared table"""
    n, frames = 1024, 5
    x, h = rnd(n * frames, 31), taps(47)
    want_bits, want_sym = o.chain_fft_fir_demod(x, n, h, scale_kind=o.SCALE_N, compat=compat)
    got = run_emu(emu, tmp_path, x, h, n, compat, 1.0 / n, staged=2, warps=2, blocks=1)
    check(got, want_bits, want_sym, compat)


def test_emulated_kernel_plain_loads_and_special_values(emu, tmp_path):
    """non-staged variant; frames with zeros, tiny values (decisions on an axis take the exact path), a NaN and an inf frame"""
    n, frames = 1024, 6
    x, h = rnd(n * frames, 5), taps(64)
    x[0:n] = 0                                    # every symbol exactly on both axes -> reference picks index 0
    x[n:2 * n] *= np.float32(1e-30)
    x[2 * n + 17] = np.complex64(complex(np.nan, 1.0))
    x[3 * n + 900] = np.complex64(complex(np.inf, 0.0))
    want_bits, want_sym = o.chain_fft_fir_demod(x, n, h, scale_kind=o.SCALE_NONE, compat=o.REFERENCE)
    got = run_emu(emu, tmp_path, x, h, n, o.REFERENCE, 1.0, staged=0, warps=3, blocks=1)
    ok = np.ones(frames, bool)
    ok[2] = ok[3] = False                         # NaN / inf frames: every bin is NaN; only the byte range is checked
    sel = np.repeat(ok, 2 * n)
    assert np.array_equal(got[:2 * n], want_bits[:2 * n])          # the all-zero frame is exact
    sym_sel = np.repeat(ok, n)
    check(got[sel], want_bits[sel], want_sym[sym_sel], o.REFERENCE)
    assert set(np.unique(got).tolist()) <= {0, 1, 2}
    # NaN symbols: "later index wins" -> idx 3 -> bytes (1, 2) (src/modulation.rs:44-49).  A NaN sample makes
    # every bin of its frame NaN in both components whatever the FFT algorithm; an inf sample does not (inf - inf
    # appears in algorithm-dependent places), so the inf frame is only range-checked above.
    fr = slice(2 * n * 2, 2 * n * 3)
    assert np.all(np.isnan(want_sym[2 * n:3 * n].real) & np.isnan(want_sym[2 * n:3 * n].imag))
    assert np.array_equal(got[fr], want_bits[fr]) and set(np.unique(got[fr]).tolist()) == {1, 2}
