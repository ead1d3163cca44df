import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


def truth():
    return np.load(os.path.join(HERE, "golden", "fft_truth.npz"))


def cx(pairs):
    a = np.array(pairs, dtype=np.float32).reshape(-1, 2)
    return (a[:, 0] + 1j * a[:, 1]).astype(np.complex64)


def same_bits(a, b):
    """byte-for-byte equality of two complex64/float arrays, any NaN == any NaN (NaN payloads are
    not preserved identically by x86 and the GPU), -0.0 != +0.0."""
    a = np.ascontiguousarray(a).view(np.float32).ravel()
    b = np.ascontiguousarray(b).view(np.float32).ravel()
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def evm_db(act, ref):
    act = np.asarray(act, dtype=np.complex128)
    ref = np.asarray(ref, dtype=np.complex128)
    e = np.sum(np.abs(act - ref) ** 2)
    r = np.sum(np.abs(ref) ** 2)
    if e == 0:
        return -np.inf
    return 10 * np.log10(e / r)
