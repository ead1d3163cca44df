"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never by aether_primitives_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "liboracle.so")

OK, ELEN, EARG, EIDX, EVM_EXCEEDED = 0, 1, 2, 6, 100
REFERENCE, CORRECTED = 0, 1
SCALE_NONE, SCALE_SN, SCALE_N, SCALE_X = 0, 1, 2, 3

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", ROOT, "oracle"])
        l = C.CDLL(SO)
        l.ora_evm_power_db.restype = C.c_double
        l.ora_evm_macro_worst_db.restype = C.c_double
        l.ora_scale_factor.restype = C.c_float
        l.ora_scale_factor.argtypes = [C.c_int, C.c_size_t, C.c_float]
        l.ora_scale.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_size_t]
        l.ora_vec_scale.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        l.ora_assert_evm.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
        l.ora_evm_power_db.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        l.ora_evm_macro_worst_db.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        for name in ("mul", "div", "add", "sub", "clone"):
            getattr(l, "ora_vec_" + name).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        for name in ("conj", "mirror", "zero"):
            getattr(l, "ora_vec_" + name).argtypes = [C.c_void_p, C.c_size_t]
        l.ora_fft_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        l.ora_fft_raw_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        l.ora_cfft_exec.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_int]
        l.ora_fir.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        l.ora_fir_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        l.ora_interpolate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        l.ora_downsample.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
        l.ora_modulate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        l.ora_demod.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        l.ora_expand.argtypes = [C.c_uint64, C.c_size_t, C.c_void_p]
        l.ora_mseq_generate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        l.ora_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.ora_awgn_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64]
        l.ora_awgn_apply.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        l.ora_chain_fft_fir_demod.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                              C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        l.ora_modem.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.ora_ofdm_chain.argtypes = [C.c_size_t, C.c_size_t, C.c_uint64, C.c_float, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
        l.ora_spectrogram.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        l.ora_vec_stats.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        l.ora_f32_stats.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        l.ora_correlate.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_int]
        _lib = l
    return _lib


class OracleError(RuntimeError):
    def __init__(self, status):
        super().__init__("oracle status %d" % status)
        self.status = status


def _ck(rc):
    if rc != OK:
        raise OracleError(rc)


def c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64).copy()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


QPSK = np.array([1 + 1j, -1 + 1j, 1 - 1j, -1 - 1j], dtype=np.complex64)
BPSK = np.array([1 + 1j, -1 - 1j], dtype=np.complex64)


# ---- assert_evm! ---------------------------------------------------------------------------
def assert_evm(actual, ref, db=-80.0):
    a, r = c64(actual), c64(ref)
    bad = C.c_size_t(0)
    return lib().ora_assert_evm(_p(a), a.size, _p(r), r.size, float(db), C.byref(bad)), bad.value


def evm_power_db(actual, ref):
    a, r = c64(actual), c64(ref)
    assert a.size == r.size
    return lib().ora_evm_power_db(_p(a), _p(r), a.size)


def evm_macro_worst_db(actual, ref):
    a, r = c64(actual), c64(ref)
    return lib().ora_evm_macro_worst_db(_p(a), _p(r), a.size)


# ---- VecOps --------------------------------------------------------------------------------
def vec_scale(v, s):
    v = c64(v); lib().ora_vec_scale(_p(v), v.size, float(s)); return v


def _bin(name, v, o):
    v, o = c64(v), c64(o)
    _ck(getattr(lib(), "ora_vec_" + name)(_p(v), v.size, _p(o), o.size))
    return v


def vec_mul(v, o): return _bin("mul", v, o)
def vec_div(v, o): return _bin("div", v, o)
def vec_add(v, o): return _bin("add", v, o)
def vec_sub(v, o): return _bin("sub", v, o)
def vec_clone(v, o): return _bin("clone", v, o)


def _un(name, v):
    v = c64(v); getattr(lib(), "ora_vec_" + name)(_p(v), v.size); return v


def vec_conj(v): return _un("conj", v)
def vec_mirror(v): return _un("mirror", v)
def vec_zero(v): return _un("zero", v)


def scale_factor(kind, n, x=1.0):
    return float(lib().ora_scale_factor(kind, n, float(x)))


def scale(kind, v, x=1.0):
    v = c64(v); lib().ora_scale(kind, float(x), _p(v), v.size); return v


# ---- FFT -----------------------------------------------------------------------------------
def fft_raw(x, sign):
    x = c64(x); out = np.empty_like(x)
    _ck(lib().ora_fft_raw(_p(x), _p(out), x.size, sign)); return out


def fft_raw_f64(x, sign):
    x = c64(x); out = np.empty(x.size, dtype=np.complex128)
    _ck(lib().ora_fft_raw_f64(_p(x), _p(out), x.size, sign)); return out


def cfft(x, n, bwd=False, scale_kind=SCALE_NONE, scale_x=1.0, compat=REFERENCE):
    x = c64(x); out = np.empty_like(x)
    howmany = x.size // n if n else 0
    _ck(lib().ora_cfft_exec(_p(x), x.size, _p(out), n, howmany, int(bwd), scale_kind, float(scale_x), compat))
    return out


# ---- FIR -----------------------------------------------------------------------------------
def fir(x, taps, state=None, frame_len=0):
    x, h = c64(x), c64(taps); y = np.empty_like(x)
    st = c64(state) if state is not None else None
    _ck(lib().ora_fir(_p(x), x.size, _p(h), h.size, _p(y), _p(st) if st is not None else None, frame_len)); return y


def fir_f64(x, taps, state=None, frame_len=0):
    x, h = c64(x), c64(taps); y = np.empty(x.size, dtype=np.complex128)
    st = c64(state) if state is not None else None
    _ck(lib().ora_fir_f64(_p(x), x.size, _p(h), h.size, _p(y), _p(st) if st is not None else None, frame_len)); return y


# ---- sampling ------------------------------------------------------------------------------
def interpolate(src, n_between, compat=REFERENCE):
    src = c64(src)
    if src.size == 0:
        raise OracleError(EARG)
    dst = np.empty((src.size - 1) * (n_between + 1) + 1, dtype=np.complex64)
    _ck(lib().ora_interpolate(_p(src), src.size, _p(dst), n_between, compat)); return dst


def downsample(src, n_dst, strict=True):
    src = np.ascontiguousarray(src); dst = np.empty(n_dst, dtype=src.dtype)
    _ck(lib().ora_downsample(_p(src), src.size, _p(dst), n_dst, src.dtype.itemsize, int(strict))); return dst


# ---- modulation ----------------------------------------------------------------------------
def modulate(table, bits, out_cap=None):
    table = c64(table); bits = np.ascontiguousarray(bits, dtype=np.uint8)
    bps = 1 if table.size == 2 else 2
    cap = (bits.size + bps - 1) // bps if out_cap is None else out_cap
    out = np.zeros(max(cap, 1), dtype=np.complex64); n_out = C.c_size_t(0)
    _ck(lib().ora_modulate(_p(table), table.size, _p(bits), bits.size, _p(out), cap, C.byref(n_out)))
    return out[: n_out.value]


def demod(table, sym, compat=REFERENCE):
    table, sym = c64(table), c64(sym)
    bps = 1 if table.size == 2 else 2
    bits = np.empty(sym.size * bps, dtype=np.uint8)
    _ck(lib().ora_demod(_p(table), table.size, _p(sym), sym.size, _p(bits), compat)); return bits


# ---- sequence ------------------------------------------------------------------------------
def expand(seed, length):
    out = np.empty(length, dtype=np.uint8)
    _ck(lib().ora_expand(C.c_uint64(seed), length, _p(out))); return out


def mseq_generate(init, back, length):
    init = np.ascontiguousarray(init, dtype=np.uint8); back = np.ascontiguousarray(back, dtype=np.uint32)
    out = np.empty(max(length, init.size), dtype=np.uint8)
    out[: init.size] = init
    _ck(lib().ora_mseq_generate(_p(init), init.size, _p(back), back.size, length, _p(out))); return out


# ---- noise ---------------------------------------------------------------------------------
def philox(ctr, key):
    c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
    lib().ora_philox4x32_10(_p(c), _p(k), _p(o)); return [int(v) for v in o]


def awgn_fill(n, power, seed, stream=0, offset=0):
    out = np.empty(n, dtype=np.complex64)
    lib().ora_awgn_fill(_p(out), n, float(power), C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(offset)); return out


def awgn_apply(sig, power, seed, stream=0, offset=0, compat=REFERENCE):
    sig = c64(sig)
    lib().ora_awgn_apply(_p(sig), sig.size, float(power), C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(offset), compat); return sig


# ---- chains --------------------------------------------------------------------------------
def chain_fft_fir_demod(x, n, taps, scale_kind=SCALE_SN, scale_x=1.0, compat=REFERENCE, nthreads=1, want_symbols=True):
    x, h = c64(x), c64(taps); frames = x.size // n
    bits = np.empty(2 * x.size, dtype=np.uint8)
    sym = np.empty_like(x) if want_symbols else None
    _ck(lib().ora_chain_fft_fir_demod(_p(x), n, frames, _p(h), h.size, _p(QPSK), scale_kind, float(scale_x), compat, _p(bits),
                                      _p(sym) if sym is not None else None, nthreads))
    return bits, sym


def modem(table, bits, power, seed, stream=0, offset=0, compat=REFERENCE):
    table = c64(table); bits = np.ascontiguousarray(bits, dtype=np.uint8)
    bps = 1 if table.size == 2 else 2
    out = np.empty(bits.size, dtype=np.uint8); sym = np.empty(bits.size // bps, dtype=np.complex64); errs = C.c_uint64(0)
    _ck(lib().ora_modem(_p(table), table.size, _p(bits), bits.size, float(power), C.c_uint64(seed), C.c_uint64(stream),
                        C.c_uint64(offset), compat, _p(out), _p(sym), C.byref(errs)))
    return out, sym, errs.value


def ofdm_chain(n, frames, first_frame, noise_power, seed, compat=REFERENCE):
    tx = np.empty(2 * n * frames, dtype=np.uint8); rx = np.empty_like(tx)
    stats = np.zeros(4, dtype=np.float64)  # bit_errors, n_bits, err_pow, ref_pow
    sym = np.empty(n * frames, dtype=np.complex64)
    _ck(lib().ora_ofdm_chain(n, frames, C.c_uint64(first_frame), float(noise_power), C.c_uint64(seed), compat, _p(tx), _p(rx),
                             _p(stats), _p(sym)))
    return tx, rx, stats, sym


# ---- SURVEY 8(f) rows ----------------------------------------------------------------------
def spectrogram(sym, fft_len, use_db=True, compat=REFERENCE):
    sym = c64(sym)
    chunks = (sym.size + fft_len - 1) // fft_len
    out = np.empty(chunks * fft_len, dtype=np.float64)
    _ck(lib().ora_spectrogram(_p(sym), sym.size, fft_len, int(use_db), compat, _p(out)))
    return out


def correlate(data, n, sig, scale_kind=SCALE_NONE, scale_x=1.0, compat=REFERENCE):
    data, sig = c64(data), c64(sig)
    _ck(lib().ora_correlate(_p(data), n, data.size // n, _p(sig), sig.size, scale_kind, float(scale_x), compat))
    return data


def vec_stats(v):
    """-> dict(n, min_idx, max_idx, min_val, max_val, sum_re, sum_im, sum_pow); complex64 or float32 input."""
    v = np.ascontiguousarray(v)
    cplx = np.iscomplexobj(v)
    v = c64(v) if cplx else np.ascontiguousarray(v, dtype=np.float32)
    out = np.zeros(3, np.uint64)
    vals = np.zeros(2, np.float32)
    sums = np.zeros(3, np.float64)
    fn = lib().ora_vec_stats if cplx else lib().ora_f32_stats
    _ck(fn(_p(v), v.size, _p(out), _p(vals), _p(sums)))
    return dict(n=int(out[0]), min_idx=int(out[1]), max_idx=int(out[2]), min_val=float(vals[0]), max_val=float(vals[1]),
                sum_re=float(sums[0]), sum_im=float(sums[1]), sum_pow=float(sums[2]))
