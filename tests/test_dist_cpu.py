"""CPU, world_size 2 over gloo: frame sharding and the only collective of the path (the <= 32-byte
BER/EVM counter all-reduce).  Each rank computes its shard with the ORACLE (the test's stand-in
for the device kernel on a GPU-less box) and the reduced counters must equal the single-rank run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aether_primitives_b200.sharding import fir_halo, frame_range


def test_frame_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            ranges = [frame_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        frame_range(10, 2, 2)
    assert fir_halo(64) == 63 and fir_halo(1) == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aether_primitives_b200.stats import allreduce
    from tests import oracle as o

    a, b = frame_range(frames, rank, world)
    _tx, _rx, st, _sym = o.ofdm_chain(n, b - a, a, 0.6, 5)
    local = {"bit_errors": int(st[0]), "n_bits": int(st[1]), "err_pow": float(st[2]), "ref_pow": float(st[3])}
    red = allreduce(local)
    if rank == 0:
        out.put(red)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_counter_allreduce_matches_single_rank():
    from tests import oracle as o

    n, frames = 512, 10
    _tx, _rx, st, _sym = o.ofdm_chain(n, frames, 0, 0.6, 5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    red = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert red["bit_errors"] == int(st[0]) and red["bit_errors"] > 0
    assert red["n_bits"] == int(st[1]) == 2 * n * frames
    assert abs(red["err_pow"] - st[2]) <= 1e-12 * st[2]
    assert abs(red["ref_pow"] - st[3]) <= 1e-12 * st[3]


def test_allreduce_without_process_group_is_identity():
    from aether_primitives_b200.stats import allreduce

    v = {"bit_errors": 3, "n_bits": 10, "err_pow": 0.5, "ref_pow": 2.0}
    assert allreduce(v) == v
