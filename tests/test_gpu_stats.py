"""VecStats (README.md:90-92 TODO): min/max with first index, f64 sums — against the oracle."""
import numpy as np
import pytest

from tests import oracle as ora

pytestmark = pytest.mark.gpu


def _check(st, want, cplx):
    assert st.n == want["n"]
    if want["min_idx"] < want["n"]:
        assert st.min[1] == want["min_idx"]
        assert np.float32(st.min[0]).tobytes() == np.float32(want["min_val"]).tobytes()
    else:
        assert st.min is None
    if want["max_idx"] < want["n"]:
        assert st.max[1] == want["max_idx"]
        assert np.float32(st.max[0]).tobytes() == np.float32(want["max_val"]).tobytes()
    else:
        assert st.max is None
    s = st.sum if cplx else complex(st.sum, 0.0)
    scale = max(1.0, np.sqrt(want["sum_pow"] * want["n"]))
    assert abs(s.real - want["sum_re"]) <= 1e-10 * scale and abs(s.imag - want["sum_im"]) <= 1e-10 * scale
    assert abs(st.power * st.n - want["sum_pow"]) <= 1e-10 * max(1.0, want["sum_pow"])


@pytest.mark.parametrize("n", [1, 2, 31, 255, 256, 257, 4097, 100_003, (1 << 22) + 5])
def test_vec_stats_cf32_matches_oracle(ae, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    st = ae.DeviceVec.from_numpy(x).vec_stats()
    _check(st, ora.vec_stats(x), True)


def test_vec_stats_first_index_wins_ties_and_nan_never_wins(ae):
    x = np.zeros(5000, np.complex64)
    x[:] = 1 + 1j
    x[[7, 4000]] = 3 - 4j                      # equal maxima: first index
    x[[19, 4500]] = 0.5j                       # equal minima
    x[3] = complex(np.nan, 0.0)
    x[4999] = complex(0.0, np.nan)
    st = ae.DeviceVec.from_numpy(x).vec_stats()
    assert st.max == (25.0, 7) and st.min == (0.25, 19)
    assert np.isnan(st.mean.real) and np.isnan(st.power)
    allnan = np.full(300, complex(np.nan, np.nan), np.complex64)
    st = ae.DeviceVec.from_numpy(allnan).vec_stats()
    assert st.min is None and st.max is None
    inf = np.full(70, complex(np.inf, 0.0), np.complex64)
    st = ae.DeviceVec.from_numpy(inf).vec_stats()
    assert st.min == (np.inf, 0) and st.max == (np.inf, 0)


def test_vec_stats_sees_pending_vecops_and_views(ae):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(10_000) + 1j * rng.standard_normal(10_000)).astype(np.complex64)
    v = ae.DeviceVec.from_numpy(x)
    v.vec_scale(2.0).vec_conj()                 # still on the tape when vec_stats is called
    want = ora.vec_stats(np.conj(x * np.float32(2.0)).astype(np.complex64))
    _check(v.vec_stats(), want, True)
    _check(v.view(1001, 9000).vec_stats(), ora.vec_stats(v.to_numpy()[1001:9000]), True)   # 8-byte aligned view


def test_vec_stats_empty_is_an_error(ae):
    with pytest.raises(ae.AeError):
        ae.DeviceVec.zeros(0).vec_stats()


def test_f32_stats_of_spectrogram_levels(ae):
    rng = np.random.default_rng(6)
    x = (rng.standard_normal(64 * 1024) + 1j * rng.standard_normal(64 * 1024)).astype(np.complex64)
    fft = ae.Cfft.with_len(1024)
    lv = ae.spectral.spectrogram(ae.DeviceVec.from_numpy(x), fft, use_db=True)
    _check(lv.vec_stats(), ora.vec_stats(lv.to_numpy()), False)
