"""GPU: fuzz of the lazy op-tape's dependency tracking.  Several device vectors (and views of them)
are used as each other's operands in random order; every result must equal the eager evaluation
with the oracle, i.e. recorded ops must see operand values as of the moment they were recorded."""
import numpy as np
import pytest

from tests import oracle as o
from tests.golden_util import same_bits

pytestmark = pytest.mark.gpu

BIN = ["mul", "add", "sub", "clone", "div"]
UN = ["conj", "mirror", "zero", "scale"]


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0.5, 1.5, n) * np.exp(2j * np.pi * rng.uniform(0, 1, n))).astype(np.complex64)


@pytest.mark.parametrize("seed", range(12))
def test_interleaved_vectors_match_eager_semantics(ae, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([6, 64, 257, 4096]))
    nv = 4
    host = [rnd(n, 100 * seed + i) for i in range(nv)]
    dev = [ae.DeviceVec.from_numpy(h) for h in host]
    for step in range(60):
        a = int(rng.integers(0, nv))
        kind = rng.random()
        if kind < 0.5:
            b = int(rng.integers(0, nv))
            if b == a:
                continue
            op = BIN[int(rng.integers(0, len(BIN)))]
            getattr(dev[a], "vec_" + op)(dev[b])
            host[a] = getattr(o, "vec_" + op)(host[a], host[b])
        elif kind < 0.85:
            op = UN[int(rng.integers(0, len(UN)))]
            if op == "scale":
                s = float(np.float32(rng.uniform(0.5, 1.5)))
                dev[a].vec_scale(s)
                host[a] = o.vec_scale(host[a], s)
            else:
                getattr(dev[a], "vec_" + op)()
                host[a] = getattr(o, "vec_" + op)(host[a])
        elif kind < 0.95 and n >= 6:
            # operate on a view (slice) of a vector, with another vector's slice as operand
            lo = int(rng.integers(0, n // 2))
            hi = int(rng.integers(lo + 1, n + 1))
            b = (a + 1) % nv
            va, vb = dev[a].view(lo, hi), dev[b].view(lo, hi)
            va.vec_add(vb).vec_conj()
            tmp = host[a].copy()
            tmp[lo:hi] = o.vec_conj(o.vec_add(host[a][lo:hi], host[b][lo:hi]))
            host[a] = tmp
            del va, vb      # dropping a view with a pending tape must not lose the ops
        else:
            # observe one vector mid-way (forces its flush only)
            assert same_bits(dev[a].to_numpy(), host[a])
    for i in range(nv):
        assert same_bits(dev[i].to_numpy(), host[i]), "vector %d differs" % i


def test_freeing_an_operand_before_flush_is_safe(ae):
    x, y = rnd(1000, 1), rnd(1000, 2)
    v = ae.DeviceVec.from_numpy(x)
    w = ae.DeviceVec.from_numpy(y)
    v.vec_mul(w)
    del w                      # the tape keeps the operand's allocation alive
    import gc

    gc.collect()
    assert same_bits(v.to_numpy(), o.vec_mul(x, y))
