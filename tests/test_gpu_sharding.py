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
