"""Worker of tests/test_gpu_sharding.py::test_two_ranks_nccl_counters_and_fir_shards — run under torchrun, one rank
per GPU.  Every rank computes its shard through the C ABI; the BER/EVM counters are reduced with the LIBRARY's NCCL
communicator (ae_comm_init_rank + ae_stats_allreduce; torch.distributed only carries the 128-byte id); rank 0 then
repeats the whole job alone and writes the comparison to argv[1]."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import aether_primitives_b200 as ae
    from aether_primitives_b200 import fir as F
    from aether_primitives_b200.chain import FftFirDemod
    from aether_primitives_b200.sharding import ShardedFir, frame_range
    from aether_primitives_b200.stats import Comm, DeviceStats

    ae.init(local)
    comm = Comm.from_torch_distributed()
    assert comm.info() == {"nranks": world, "rank": rank, "device": local}

    # ---- config 5: OFDM chain over 2^16 global frames, counters reduced through the C ABI
    frames = 1 << 16
    f0, f1 = frame_range(frames, rank, world)
    st = DeviceStats()
    ae.chain.ofdm_chain(2048, f1 - f0, f0, 0.5, 5, st, None, None, ae.COMPAT_CORRECTED)
    st.allreduce(comm)
    reduced = st.read()

    # ---- config 3: streaming FIR, 64 taps overlap-save, 2^22 global samples, each rank its shard + halo
    n = 1 << 22
    rng = np.random.default_rng(99)                     # same global stream on every rank
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    k = np.arange(64) - 31.5
    h = (np.sinc(k / 4.3) * np.hamming(64)).astype(np.complex64)
    sh = ShardedFir(h, n, rank, world, F.OVERLAP_SAVE)
    lo_in, hi = sh.input_range()
    y = sh.filter(ae.DeviceVec.from_numpy(x[lo_in:hi])).to_numpy()  # hi = hi_in: shard + read-ahead
    parts = [None] * world
    dist.all_gather_object(parts, y)

    # ---- headline chain: frames sharded, bits gathered
    cf = 64
    xc = x[: cf * 1024]
    c0, c1 = frame_range(cf, rank, world)
    ch = FftFirDemod(1024, h, ae.Scale.SN)
    bits = ae.DeviceBits.with_capacity(1)
    ch.run(ae.DeviceVec.from_numpy(xc[c0 * 1024:c1 * 1024]), bits)
    cparts = [None] * world
    dist.all_gather_object(cparts, bits.to_numpy())

    if rank == 0:
        st1 = DeviceStats()
        ae.chain.ofdm_chain(2048, frames, 0, 0.5, 5, st1, None, None, ae.COMPAT_CORRECTED)
        single = st1.read()
        whole = ae.DeviceVec.zeros(n)
        F.Fir(h, F.OVERLAP_SAVE).filter(ae.DeviceVec.from_numpy(x), whole)
        fir_same = bool(np.array_equal(np.concatenate(parts).view(np.uint32), whole.to_numpy().view(np.uint32)))
        b1 = ae.DeviceBits.with_capacity(1)
        ch.run(ae.DeviceVec.from_numpy(xc), b1)
        chain_same = bool(np.array_equal(np.concatenate(cparts), b1.to_numpy()))
        json.dump({"world": world, "ofdm": {"frames": frames, "reduced": reduced, "single": single},
                   "fir": {"samples": n, "bit_identical": fir_same}, "chain": {"frames": cf, "bit_identical": chain_same}},
                  open(sys.argv[1], "w"))
    ae.sync()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
