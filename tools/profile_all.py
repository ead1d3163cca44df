#!/usr/bin/env python
"""Runs every stand-alone kernel of the path once or twice at 2^26 samples so that one
`ncu --set full` capture covers them all (summarised into profiles/ with tools/ncu_summary.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aether_primitives_b200 as ae
from aether_primitives_b200 import fir as F
from aether_primitives_b200.stats import DeviceStats
from bench import make_taps

n = 1 << 26
ae.init(0)
ae.use_torch_stream()
x = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
y = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
dx, dy = ae.DeviceVec.from_torch(x), ae.DeviceVec.from_torch(y)
st = DeviceStats()
qpsk = ae.modulation.qpsk()
g = ae.noise.new(0.01, 815)
fft = ae.Cfft.with_len(1024)
for rep in range(1):
    dx.vec_mul(dy).vec_conj().vec_mirror().flush()
    dx.vec_scale(0.5).vec_add(dy).flush()
    ds = ae.DeviceVec.zeros(n // 4)
    ae.sampling.downsample(dx, ds)
    it = ae.DeviceVec.with_capacity(n)
    ae.sampling.interpolate(ds, it, 3)
    del it, ds
    for t, mode in ((64, F.DIRECT), (64, F.OVERLAP_SAVE), (1024, F.OVERLAP_SAVE)):
        F.Fir(make_taps(t), mode).filter(dx, dy)
    bits = ae.DeviceBits.zeros(2 * n)
    sym = ae.DeviceVec.zeros(n)
    qpsk.modulate_into(bits, sym)
    g.apply(sym)
    fv = ae.DeviceVec.with_capacity(n)
    g.fill(fv)
    out = ae.DeviceBits.with_capacity(2 * n)
    qpsk.demod_naive(sym, out)
    ae.chain.modem_fused(qpsk, g, bits, out, st)
    st.count_bit_errors(bits, out)
    st.evm_accumulate(dx, dy)
    del bits, sym, fv, out
    fft.ifwd(dx, ae.Scale.SN, howmany=n // 1024)
    lv = ae.spectral.spectrogram(dx, fft, True)
    sig = ae.DeviceVec.from_torch(torch.view_as_complex(torch.randn(1024, 2, device="cuda")))
    ae.spectral.correlate(dx, sig, fft, ae.Scale.SN, howmany=n // 1024)
    lv.vec_stats()
    del lv
    dx.vec_stats()
    for nn in (100, 1000, 3000):                       # any-length mixed-radix kernel
        fr = n // nn
        ae.Cfft.with_len(nn).ifwd(dx.view(0, fr * nn), ae.Scale.SN, howmany=fr)
    ae.Cfft.with_len(1 << 16).ifwd(dx, ae.Scale.SN, howmany=n >> 16)      # four-step
    ae.Cfft.with_len(1 << 20).ifwd(dx, ae.Scale.SN, howmany=n >> 20)      # four-step, 1024 x 1024
    from aether_primitives_b200.chain import FftFirDemod
    cb = ae.DeviceBits.with_capacity(2 * n)
    FftFirDemod(1024, make_taps(64), ae.Scale.SN).run(dx, cb)              # headline chain
    del cb
    ae.chain.ofdm_chain(2048, 1 << 14, 0, 0.05, 5, st)
    seq = ae.sequence.generate([1] + [0] * 30, [28, 31], 1 << 24)
    del seq
    ae.sync()
print("ok", ae.launch_count())
