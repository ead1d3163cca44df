"""GPU: multi-GPU sharding of the path (BASELINE configs 3 and 5).

* the streaming FIR cut into R shards with a (T-1)-sample input halo gives, concatenated, the unsharded output
  BIT FOR BIT (direct form and overlap-save, 64 and 1024 taps) — checked on one GPU by running the R shards one
  after the other, and on R real ranks by tests/dist_gpu_worker.py under torchrun (skipped with fewer than 2 GPUs);
* the OFDM chain's BER/EVM counters reduced over ranks with the library's NCCL communicator (ae_stats_allreduce)
  equal the single-rank counters over the same global frames."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def taps(t, seed=3):
    rng = np.random.default_rng(seed)
    k = np.arange(t) - (t - 1) / 2
    h = np.sinc(k / 4.3) * np.hamming(t) * np.exp(1j * rng.uniform(0, 2 * np.pi))
    return (h / np.abs(h.sum())).astype(np.complex64)


@pytest.mark.parametrize("mode", ["direct", "os"])
@pytest.mark.parametrize("t", [64, 1024])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_fir_equals_unsharded_bit_for_bit(ae, mode, t, world):
    from aether_primitives_b200 import fir as F
    from aether_primitives_b200.sharding import ShardedFir, fir_halo

    n = 300_007 if t == 64 else 120_011
    x, h = rnd(n, 17 * t + world), taps(t)
    m = F.DIRECT if mode == "direct" else F.OVERLAP_SAVE
    whole = ae.DeviceVec.zeros(n)
    F.Fir(h, m).filter(ae.DeviceVec.from_numpy(x), whole)
    want = whole.to_numpy()
    parts = []
    for r in range(world):
        sh = ShardedFir(h, n, r, world, m)
        lo_in, hi = sh.input_range()
        assert sh.halo >= (fir_halo(t) if r else 0) and lo_in % sh.fir.block_hop() == 0
        if r == 0:
            assert lo_in == 0 and sh.halo == 0
        y = sh.filter(ae.DeviceVec.from_numpy(x[lo_in:hi]))
        assert len(y) == sh.hi - sh.lo
        parts.append(y.to_numpy())
    got = np.concatenate(parts)
    assert got.size == n
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "sharded FIR differs from the unsharded stream"


@pytest.mark.parametrize("seed", range(24))
def test_sharded_fir_random_shapes(ae, seed):
    """random stream lengths (shorter than a block, not a multiple of the hop, shards smaller than the halo), tap counts,
    shard counts and both modes: the concatenation of the shards is the unsharded stream, bit for bit"""
    from aether_primitives_b200 import fir as F
    from aether_primitives_b200.sharding import ShardedFir

    rng = np.random.default_rng(1000 + seed)
    t = int(rng.choice([1, 2, 17, 64, 65, 200, 256, 1024]))
    world = int(rng.integers(1, 10))
    n = int(rng.choice([rng.integers(1, 900), rng.integers(900, 5000), rng.integers(5000, 200_000)]))
    m = F.DIRECT if seed % 2 else F.OVERLAP_SAVE
    x, h = rnd(n, seed), (taps(t) if t > 2 else rnd(t, 9))
    whole = ae.DeviceVec.zeros(n)
    F.Fir(h, m).filter(ae.DeviceVec.from_numpy(x), whole)
    want = whole.to_numpy()
    got = []
    for r in range(world):
        sh = ShardedFir(h, n, r, world, m)
        lo_in, hi = sh.input_range()
        assert 0 <= lo_in <= sh.lo <= sh.hi <= hi <= n
        if sh.hi == sh.lo:
            continue                                   # more ranks than samples: this one owns nothing
        got.append(sh.filter(ae.DeviceVec.from_numpy(x[lo_in:hi])).to_numpy())
    got = np.concatenate(got)
    assert got.size == n
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "t=%d world=%d n=%d mode=%d" % (t, world, n, m)


def test_sharded_fir_rejects_wrong_input_length(ae):
    from aether_primitives_b200.sharding import ShardedFir

    sh = ShardedFir(taps(64), 10000, 1, 2)
    with pytest.raises(ValueError):
        sh.filter(ae.DeviceVec.zeros(5000))


def test_two_ranks_nccl_counters_and_fir_shards():
    """torchrun, 2 ranks on 2 GPUs: NCCL-reduced OFDM counters == single-rank counters, FIR shards == unsharded"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    out = os.path.join(ROOT, "gpurun_out", "dist_gpu_worker.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if os.path.exists(out):
        os.remove(out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dist_gpu_worker.py"), out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    res = json.load(open(out))
    assert res["world"] == 2
    assert res["ofdm"]["reduced"]["bit_errors"] == res["ofdm"]["single"]["bit_errors"] > 0
    assert res["ofdm"]["reduced"]["n_bits"] == res["ofdm"]["single"]["n_bits"] == 2 * 2048 * res["ofdm"]["frames"]
    for k in ("err_pow", "ref_pow"):
        assert abs(res["ofdm"]["reduced"][k] - res["ofdm"]["single"][k]) <= 1e-11 * abs(res["ofdm"]["single"][k])
    assert res["fir"]["bit_identical"] is True
    assert res["chain"]["bit_identical"] is True


def test_single_process_communicator_on_one_gpu(ae):
    """ae_comm_init_all(1) + ae_stats_allreduce / _all: the library binds libnccl at run time; with one rank the
    reduction is the identity (the >= 2 rank case is test_two_ranks_nccl_counters_and_fir_shards)"""
    import ctypes as C

    from aether_primitives_b200._lib import call
    from aether_primitives_b200.stats import Comm, DeviceStats

    h = (C.c_void_p * 1)()
    call("ae_comm_init_all", 1, h)
    comm = Comm(C.c_void_p(h[0]))
    assert comm.info() == {"nranks": 1, "rank": 0, "device": 0}
    st = DeviceStats()
    ae.chain.ofdm_chain(1024, 16, 0, 0.5, 5, st, None, None, ae.COMPAT_CORRECTED)
    before = st.read()
    st.allreduce(comm)
    assert st.read() == before and before["n_bits"] == 2 * 1024 * 16
    sp = (C.c_void_p * 1)(st._h.value if hasattr(st._h, "value") else st._h)
    cp = (C.c_void_p * 1)(h[0])
    call("ae_stats_allreduce_all", sp, cp, 1)
    assert st.read() == before
    # one-process-per-GPU form with a single rank
    c2 = Comm.init_rank(Comm.unique_id(), 1, 0)
    st.allreduce(c2)
    assert st.read() == before
    c2.close()
    comm.close()


def test_chain_with_odd_bit_buffer_address(ae):
    """a wrapped bit buffer at an odd address cannot take 16/32-bit stores: the chain falls back to the stand-alone kernels"""
    import torch

    from aether_primitives_b200.chain import FftFirDemod

    x, h = rnd(4 * 1024, 3), taps(64)
    ch = FftFirDemod(1024, h)
    want = ae.DeviceBits.with_capacity(1)
    ch.run(ae.DeviceVec.from_numpy(x), want)
    raw = torch.zeros(2 * x.size + 1, dtype=torch.uint8, device="cuda")
    odd = ae.DeviceBits.wrap(raw.data_ptr() + 1, 2 * x.size, owner=raw)
    ch.run(ae.DeviceVec.from_numpy(x), odd)
    got = raw[1:].cpu().numpy()
    w = want.to_numpy()
    mism = np.nonzero((got != 0) != (w != 0))[0]
    assert len(mism) <= 3 and set(np.unique(got).tolist()) <= {0, 1, 2}
