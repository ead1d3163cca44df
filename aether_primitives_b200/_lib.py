"""ctypes binding of libaether_b200.so (the C ABI in include/aether_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing this module raises
on import of any symbol, and every compute call fails with AE_ECUDA when no B200 is present.
Nothing here imports or calls oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libaether_b200.so")

AE_OK, AE_ELEN, AE_EARG, AE_ECUDA, AE_ENCCL, AE_EOOM, AE_EIDX = range(7)
COMPAT_REFERENCE, COMPAT_CORRECTED = 0, 1
SCALE_NONE, SCALE_SN, SCALE_N, SCALE_X = 0, 1, 2, 3
FFT_FWD, FFT_BWD = 0, 1
FIR_AUTO, FIR_DIRECT, FIR_OVERLAP_SAVE = 0, 1, 2


class AeError(RuntimeError):
    """A non-zero ae_status.  `message` carries the text the reference would panic with."""

    def __init__(self, status: int, message: str):
        super().__init__(f"ae_status {status}: {message}")
        self.status = status
        self.message = message


class Stats(C.Structure):
    _fields_ = [("bit_errors", C.c_uint64), ("n_bits", C.c_uint64), ("err_pow", C.c_double), ("ref_pow", C.c_double)]


class VecStatsRaw(C.Structure):
    _fields_ = [("n", C.c_uint64), ("min_idx", C.c_uint64), ("max_idx", C.c_uint64), ("min_val", C.c_float), ("max_val", C.c_float),
                ("sum_re", C.c_double), ("sum_im", C.c_double), ("sum_pow", C.c_double)]


class PipeStage(C.Structure):
    _fields_ = [("name", C.c_char * 16), ("processed", C.c_uint64), ("active_ms", C.c_double), ("elapsed_ms", C.c_double),
                ("per_second", C.c_double), ("utilisation_pct", C.c_double)]


_P = C.c_void_p
_SZ = C.c_size_t
_I = C.c_int
_F = C.c_float
_U64 = C.c_uint64

# name -> (restype, argtypes); restype None means ae_status
_SIGS = {
    "ae_version": (C.c_char_p, []),
    "ae_device_count": (None, [C.POINTER(_I)]),
    "ae_init": (None, [_I]),
    "ae_set_stream": (None, [_P]),
    "ae_get_stream": (_P, []),
    "ae_sync": (None, []),
    "ae_last_error_string": (C.c_char_p, []),
    "ae_sm_count": (None, [C.POINTER(_I)]),
    "ae_launch_count": (_U64, []),
    "ae_host_alloc": (None, [_SZ, C.POINTER(_P)]),
    "ae_host_free": (None, [_P]),
    "ae_vec_alloc": (None, [_SZ, _SZ, C.POINTER(_P)]),
    "ae_vec_wrap": (None, [_P, _SZ, C.POINTER(_P)]),
    "ae_vec_view": (None, [_P, _SZ, _SZ, C.POINTER(_P)]),
    "ae_vec_free": (None, [_P]),
    "ae_vec_len": (_SZ, [_P]),
    "ae_vec_capacity": (_SZ, [_P]),
    "ae_vec_set_len": (None, [_P, _SZ]),
    "ae_vec_reserve": (None, [_P, _SZ]),
    "ae_vec_device_ptr": (None, [_P, C.POINTER(_P)]),
    "ae_vec_upload": (None, [_P, _P, _SZ]),
    "ae_vec_download": (None, [_P, _P, _SZ]),
    "ae_vec_scale": (None, [_P, _F]),
    "ae_vec_mul": (None, [_P, _P]),
    "ae_vec_div": (None, [_P, _P]),
    "ae_vec_conj": (None, [_P]),
    "ae_vec_add": (None, [_P, _P]),
    "ae_vec_sub": (None, [_P, _P]),
    "ae_vec_mirror": (None, [_P]),
    "ae_vec_clone": (None, [_P, _P]),
    "ae_vec_zero": (None, [_P]),
    "ae_vec_mutate": (None, [_P, _P, _P]),
    "ae_vec_flush": (None, [_P]),
    "ae_vec_pending_ops": (_SZ, [_P]),
    "ae_scale_factor": (None, [_I, _SZ, _F, C.POINTER(_F)]),
    "ae_vec_scale_kind": (None, [_P, _I, _F]),
    "ae_vec_fft": (None, [_P, _I, _F, _I]),
    "ae_vec_ifft": (None, [_P, _I, _F, _I]),
    "ae_bits_alloc": (None, [_SZ, _SZ, C.POINTER(_P)]),
    "ae_bits_wrap": (None, [_P, _SZ, C.POINTER(_P)]),
    "ae_bits_free": (None, [_P]),
    "ae_bits_len": (_SZ, [_P]),
    "ae_bits_capacity": (_SZ, [_P]),
    "ae_bits_set_len": (None, [_P, _SZ]),
    "ae_bits_device_ptr": (None, [_P, C.POINTER(_P)]),
    "ae_bits_upload": (None, [_P, _P, _SZ]),
    "ae_bits_download": (None, [_P, _P, _SZ]),
    "ae_fft_create": (None, [_SZ, C.POINTER(_P)]),
    "ae_fft_destroy": (None, [_P]),
    "ae_fft_len": (_SZ, [_P]),
    "ae_fft_set_compat": (None, [_P, _I]),
    "ae_fft_exec": (None, [_P, _I, _P, _P, _I, _F, _SZ]),
    "ae_fft_exec_tmp": (None, [_P, _I, _P, _I, _F, _SZ, C.POINTER(_P)]),
    "ae_fir_create": (None, [_P, _SZ, _I, C.POINTER(_P)]),
    "ae_fir_destroy": (None, [_P]),
    "ae_fir_ntaps": (_SZ, [_P]),
    "ae_fir_block_hop": (_SZ, [_P]),
    "ae_fir_reset": (None, [_P]),
    "ae_fir_exec": (None, [_P, _P, _P, _SZ]),
    "ae_interpolate": (None, [_P, _P, _SZ, _I]),
    "ae_downsample": (None, [_P, _P, _I]),
    "ae_downsample_sb": (None, [_P, _P, _I]),
    "ae_downsample_bits": (None, [_P, _P, _I]),
    "ae_mod_create": (None, [_P, _SZ, C.POINTER(_P)]),
    "ae_mod_bpsk": (None, [C.POINTER(_P)]),
    "ae_mod_qpsk": (None, [C.POINTER(_P)]),
    "ae_mod_destroy": (None, [_P]),
    "ae_mod_bits_per_symbol": (_SZ, [_P]),
    "ae_mod_modulate": (None, [_P, _P, _P]),
    "ae_mod_modulate_into": (None, [_P, _P, _P]),
    "ae_mod_demod": (None, [_P, _P, _P, _I]),
    "ae_awgn_create": (None, [_F, _U64, C.POINTER(_P)]),
    "ae_awgn_generator": (None, [C.POINTER(_P)]),
    "ae_awgn_destroy": (None, [_P]),
    "ae_awgn_set_power": (None, [_P, _F]),
    "ae_awgn_set_stream_id": (None, [_P, _U64]),
    "ae_awgn_seek": (None, [_P, _U64]),
    "ae_awgn_tell": (_U64, [_P]),
    "ae_awgn_fill": (None, [_P, _P]),
    "ae_awgn_apply": (None, [_P, _P, _I]),
    "ae_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "ae_mseq_expand": (None, [_U64, _SZ, _P]),
    "ae_mseq_generate": (None, [_P, _SZ, _P, _SZ, _SZ, _P]),
    "ae_stats_alloc": (None, [C.POINTER(_P)]),
    "ae_stats_free": (None, [_P]),
    "ae_stats_zero": (None, [_P]),
    "ae_stats_read": (None, [_P, C.POINTER(Stats)]),
    "ae_count_bit_errors": (None, [_P, _P, _P]),
    "ae_evm_accumulate": (None, [_P, _P, _P]),
    "ae_comm_unique_id": (None, [_P]),
    "ae_comm_init_rank": (None, [_P, _I, _I, C.POINTER(_P)]),
    "ae_comm_init_all": (None, [_I, C.POINTER(_P)]),
    "ae_comm_destroy": (None, [_P]),
    "ae_comm_info": (None, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "ae_stats_allreduce": (None, [_P, _P]),
    "ae_stats_allreduce_all": (None, [C.POINTER(_P), C.POINTER(_P), _I]),
    "ae_modem_fused": (None, [_P, _P, _P, _P, _P, _I]),
    "ae_chain_create": (None, [_SZ, _P, _SZ, _I, _F, _I, C.POINTER(_P)]),
    "ae_chain_destroy": (None, [_P]),
    "ae_chain_exec": (None, [_P, _P, _P]),
    "ae_chain_exec_host": (None, [_P, _P, _SZ, _P]),
    "ae_pipe_create": (None, [_P, _SZ, _I, C.POINTER(_P)]),
    "ae_pipe_destroy": (None, [_P]),
    "ae_pipe_send": (None, [_P, _P, _P]),
    "ae_pipe_recv": (None, [_P, C.POINTER(_P)]),
    "ae_pipe_in_flight": (_SZ, [_P]),
    "ae_pipe_report": (None, [_P, C.POINTER(PipeStage), _I]),
    "ae_pipeline_create": (None, [_I, C.POINTER(_P)]),
    "ae_pipeline_add_stage": (None, [_P, C.c_char_p, _P, _P]),
    "ae_pipeline_send": (None, [_P, _P]),
    "ae_pipeline_recv": (None, [_P, C.POINTER(_P)]),
    "ae_pipeline_in_flight": (_SZ, [_P]),
    "ae_pipeline_stages": (_SZ, [_P]),
    "ae_pipeline_report": (None, [_P, C.POINTER(PipeStage), _SZ, _I]),
    "ae_pipeline_destroy": (None, [_P]),
    "ae_vec_upload_async": (None, [_P, _P, _SZ]),
    "ae_vec_download_async": (None, [_P, _P, _SZ]),
    "ae_bits_upload_async": (None, [_P, _P, _SZ]),
    "ae_bits_download_async": (None, [_P, _P, _SZ]),
    "ae_graph_begin": (None, []),
    "ae_graph_end": (None, [C.POINTER(_P)]),
    "ae_graph_launch": (None, [_P]),
    "ae_graph_destroy": (None, [_P]),
    "ae_chain_exec_unfused": (None, [_P, _P, _P, _P]),
    "ae_ofdm_chain": (None, [_SZ, _SZ, _U64, _F, _U64, _I, _P, _P, _P]),
    "ae_f32_alloc": (None, [_SZ, C.POINTER(_P)]),
    "ae_f32_free": (None, [_P]),
    "ae_f32_len": (_SZ, [_P]),
    "ae_f32_device_ptr": (None, [_P, C.POINTER(_P)]),
    "ae_f32_download": (None, [_P, _P, _SZ]),
    "ae_f32_stats": (None, [_P, C.POINTER(VecStatsRaw)]),
    "ae_vec_stats": (None, [_P, C.POINTER(VecStatsRaw)]),
    "ae_spectrogram": (None, [_P, _P, _P, _I]),
    "ae_correlate": (None, [_P, _P, _P, _I, _F, _SZ]),
    "ae_awgn_next_host": (None, [_P, _P, _SZ]),
    "ae_vec_read_raw": (None, [C.c_char_p, C.POINTER(_P)]),
    "ae_vec_write_raw": (None, [_P, C.c_char_p]),
}

EXPORTS = tuple(_SIGS.keys())
_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
                "aether_primitives_b200 has no CPU fallback."
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = _I if res is None else res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != AE_OK:
        raise AeError(status, lib().ae_last_error_string().decode("utf-8", "replace"))


def call(name: str, *args):
    """Call an ae_status-returning entry point and raise AeError on failure."""
    check(getattr(lib(), name)(*args))
