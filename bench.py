#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 cf32 path.

Workload (BASELINE.json metric): batched 1024-point cf32 Cfft::fwd(Scale::SN) -> 64-tap complex
FIR (zero state per frame) -> QPSK hard demod, 2^20 frames per GPU, synthetic N(0,1) input.
One "step" = one pass of the fused chain kernel over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--no-extras]

For N > 1 launch with torchrun (one rank per GPU); frames are sharded, no data-path collective
(weak scaling: every rank owns 2^20 frames).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FFT_LEN = 1024
NTAPS = 64
ALG_BYTES_PER_SAMPLE = 10.0   # 8 B cf32 in + 2 B bits out (SURVEY §8d headline row)
FLOP_PER_SAMPLE = 2 * 50 + 6 + 16  # two 1024-pt FFTs + window multiply + triangular fix-up (nominal 5 N log2 N count)
METRIC = "cf32 Gsamples/s, FFT->FIR->QPSK-demod chain (1024-pt fwd FFT, 64-tap FIR, hard demod)"


def make_taps(t: int = NTAPS) -> np.ndarray:
    """windowed-sinc * exp(j phi), unit DC gain, seed 3"""
    rng = np.random.default_rng(3)
    k = np.arange(t) - (t - 1) / 2
    h = np.sinc(k / 4.0) * np.hamming(t) * np.exp(1j * rng.uniform(0, 2 * np.pi))
    return (h / np.abs(h.sum())).astype(np.complex64)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled every 20 ms DURING the timed region (NVML from a thread;
    falls back to `nvidia-smi -lms` when pynvml is missing)."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, index: int):
        self.index = index
        self.samples = []   # (time, sm_mhz, reasons_bitmask)
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        self.proc = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            rs = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((time.time(), mhz, rs))
                    except Exception:
                        pass
                    time.sleep(0.02)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def rd():
                bits = [0x8, 0x4, 0x40, 0x20]
                for line in self.proc.stdout:
                    f = [x.strip() for x in line.split(",")]
                    try:
                        self.max_mhz = float(f[1])
                        rs = sum(b for b, v in zip(bits, f[2:6]) if v.lower().startswith("active"))
                        self.samples.append((time.time(), float(f[0]), rs))
                    except Exception:
                        pass

            self.thread = threading.Thread(target=rd, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        return time.time()

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self, t0: float, t1: float) -> dict:
        rows = [(m, r) for (t, m, r) in self.samples if t0 <= t <= t1] or [(m, r) for (_, m, r) in self.samples[-3:]]
        sm = [m for m, _ in rows]
        mask = 0
        for _, r in rows:
            mask |= r
        reasons = sorted(n for n, b in self.REASONS.items() if mask & b)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's algorithm, all host threads, bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_chain_rate(target_seconds: float, threads: int, steps: int = 1, warmup: int = 0):
    """Times tests.oracle.chain_fft_fir_demod (C++ restatement of the reference, frame-parallel
    std::thread) on a sample sized for ~target_seconds per step.  Returns (Gsamples/s, ms/step, frames)."""
    from tests import oracle as o

    taps = make_taps()
    rng = np.random.default_rng(1)
    cal = 64 * threads
    x = (rng.standard_normal(cal * FFT_LEN) + 1j * rng.standard_normal(cal * FFT_LEN)).astype(np.complex64)
    o.chain_fft_fir_demod(x, FFT_LEN, taps, nthreads=threads, want_symbols=False)
    t0 = time.perf_counter()
    o.chain_fft_fir_demod(x, FFT_LEN, taps, nthreads=threads, want_symbols=False)
    dt = time.perf_counter() - t0
    frames = int(max(cal, min(1 << 20, cal * target_seconds / max(dt, 1e-6))))
    frames = (frames // threads) * threads
    x = (rng.standard_normal(frames * FFT_LEN) + 1j * rng.standard_normal(frames * FFT_LEN)).astype(np.complex64)
    for _ in range(warmup):
        o.chain_fft_fir_demod(x, FFT_LEN, taps, nthreads=threads, want_symbols=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.chain_fft_fir_demod(x, FFT_LEN, taps, nthreads=threads, want_symbols=False)
    dt = (time.perf_counter() - t0) / steps
    return frames * FFT_LEN / dt / 1e9, dt * 1e3, frames


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: ~2 s of CPU work per step so that steps+warmup finish within minutes
    per_step = max(0.5, min(3.0, 120.0 / max(1, args.steps + args.warmup)))
    gs, ms, frames = cpu_chain_rate(per_step, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gs, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fft1024_fwd_SN -> fir64 -> qpsk_demod", "fft_len": FFT_LEN, "ntaps": NTAPS,
                   "frames_per_step": frames, "note": "reference crate is Rust and cannot be built here; this is the C++ restatement (oracle port), not rustfft"},
        "cpu_baseline": {"value": gs, "unit": "Gsamples/s", "cores": threads, "kind": "port",
                         "sample": "%d frames x %d samples per step, %d host threads" % (frames, FFT_LEN, threads)},
        "e2e": {"value": gs, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def timed(torch, fn, steps: int, warmup: int, barrier=None):
    """W warm-ups, then exactly `steps` calls bracketed by barrier + synchronize; CUDA events on the
    launching stream.  Returns seconds for all steps."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1) / 1e3


def extras(torch, ae, d_in, frames, hbm_peak, steps=5, warmup=3, rank=0, world=1):
    """Stand-alone kernels of the other BASELINE configs, device-timed on THIS rank's GPU, each with its
    algorithmic bytes and fraction of the measured HBM peak.  Reported under "extra"; for N > 1 every rank
    runs them on its own shard (weak scaling) and run_ours() combines the ranks (max time, summed units).
    Config 3 is the exception: one 2^28-sample stream cut into `world` shards with a (T-1)-sample halo."""
    from aether_primitives_b200 import fir as F
    from aether_primitives_b200.sharding import ShardedFir
    from aether_primitives_b200.stats import DeviceStats

    out = {}
    n = frames * FFT_LEN

    def rec(name, alg_bytes, secs, units, unit_name):
        gbs = alg_bytes / secs / 1e9
        out[name] = {"ms": secs * 1e3, "G%s/s" % unit_name: units / secs / 1e9, "alg_GB/s": gbs, "frac_hbm": gbs / hbm_peak}

    fft = ae.Cfft.with_len(FFT_LEN)
    # config 2: forward + backward 1024-pt FFT with scaling, in place
    t = timed(torch, lambda: (fft.ifwd(d_in, ae.Scale.SN, howmany=frames), fft.ibwd(d_in, ae.Scale.SN, howmany=frames)), steps, warmup) / steps
    rec("fft1024_fwd+bwd_SN", 32.0 * n, t, n, "samples")
    out["fft1024_fwd+bwd_SN"]["fft_TFLOP/s"] = 2 * 5 * FFT_LEN * 10 * frames / t / 1e12
    t = timed(torch, lambda: fft.ifwd(d_in, ae.Scale.SN, howmany=frames), steps, warmup) / steps
    rec("fft1024_fwd_SN", 16.0 * n, t, n, "samples")
    # config 4: fused mul.conj.mirror, downsample/4, interpolate x4 on 2^28 samples
    m = min(n, 1 << 28)
    a = d_in.view(0, m)
    # unit-modulus operand: eight repetitions of mul/conj/mirror leave the data N(0,1)-like for the
    # kernels timed afterwards (a zero operand would hand the spectrogram all-zero frames)
    b = ae.DeviceVec.from_torch(torch.full((m,), 0.6 + 0.8j, dtype=torch.complex64, device="cuda"))
    t = timed(torch, lambda: a.vec_mul(b).vec_conj().vec_mirror().flush(), steps, warmup) / steps
    rec("vecops_mul_conj_mirror_fused", 24.0 * m, t, m, "samples")
    ds = ae.DeviceVec.zeros(m // 4)
    t = timed(torch, lambda: ae.sampling.downsample(a, ds), steps, warmup) / steps
    rec("downsample_by4", 16.0 * (m // 4), t, m // 4, "outputs")
    # the kept samples are 32 bytes apart = one DRAM sector each, so every sector of the input is fetched
    # (ncu: dram__bytes_read = the whole input, profiles/r2_all_kernels_ncu_summary.txt): the floor is 40 B/output
    out["downsample_by4"]["dram_sector_GB/s"] = 40.0 * (m // 4) / t / 1e9
    out["downsample_by4"]["frac_hbm_dram_sector"] = out["downsample_by4"]["dram_sector_GB/s"] / hbm_peak
    src = d_in.view(0, m // 4)
    dst = ae.DeviceVec.with_capacity(m)

    def interp():
        dst.clear()
        ae.sampling.interpolate(src, dst, 3)
    t = timed(torch, interp, steps, warmup) / steps
    rec("interpolate_x4", 40.0 * (m // 4), t, m // 4, "inputs")
    del dst, ds
    # config 3: 64-tap FIR direct and overlap-save, 1024-tap overlap-save, 2^28 samples (b is the output)
    taps64 = make_taps(64)
    for name, filt in (("fir64_direct", F.Fir(taps64, F.DIRECT)), ("fir64_overlap_save", F.Fir(taps64, F.OVERLAP_SAVE)),
                       ("fir1024_overlap_save", F.Fir(make_taps(1024), F.OVERLAP_SAVE))):
        t = timed(torch, lambda: filt.filter(a, b), steps, warmup) / steps
        rec(name, 16.0 * m, t, m, "samples")
    out["fir64_direct"]["fp32_TFLOP/s"] = 8.0 * 64 * m / (out["fir64_direct"]["ms"] * 1e-3) / 1e12
    # config 3 also names the 1024-tap direct form: 8192 flop per sample, FP32-bound by construction; a 2^24-sample slice
    md = min(m, 1 << 24)
    fd = F.Fir(make_taps(1024), F.DIRECT)
    ad, bd = a.view(0, md), b.view(0, md)
    t = timed(torch, lambda: fd.filter(ad, bd), 2, 1) / 2
    rec("fir1024_direct", 16.0 * md, t, md, "samples")
    out["fir1024_direct"]["fp32_TFLOP/s"] = 8.0 * 1024 * md / t / 1e12
    out["fir1024_direct"]["bytes_note"] = "FP32-bound by construction (74 TFLOP/s peak = 9 Gsamples/s); 2^24-sample slice"
    del fd, ad, bd
    # config 3 across GPUs: ONE stream of 2^28 samples, rank r filters [r n/R - halo, (r+1) n/R) and keeps its n/R outputs
    # (strong scaling; "units" are this rank's outputs, so the combined line is the rate of the whole stream)
    total = 1 << 28
    for name, tp in (("fir64_stream_sharded", 64), ("fir1024_stream_sharded", 1024)):
        sh = ShardedFir(make_taps(tp), total, rank, world, F.OVERLAP_SAVE)
        lo_in, hi = sh.input_range()
        xs = d_in.view(0, hi - lo_in)          # synthetic stream: any N(0,1) samples of the right length
        t = timed(torch, lambda: sh.filter(xs), steps, warmup) / steps
        rec(name, 16.0 * (sh.hi - sh.lo), t, sh.hi - sh.lo, "samples")
        out[name]["halo_samples"] = sh.halo + (hi - sh.hi)
        out[name]["strong_scaling_total_samples"] = total
        del sh
    # config 1: modem loop-back at 1M symbols (launch-bound) and 2^28 symbols
    qpsk = ae.modulation.qpsk()
    g = ae.noise.new(0.01, 815)
    st = DeviceStats()
    for nsym in (1_000_000, 1 << 28):
        bits_in = ae.DeviceBits.zeros(2 * nsym)
        bits_out = ae.DeviceBits.zeros(2 * nsym)
        t = timed(torch, lambda: ae.chain.modem_fused(qpsk, g, bits_in, bits_out, st, ae.COMPAT_REFERENCE), steps, warmup) / steps
        rec("modem_fused_%dsym" % nsym, 4.0 * nsym, t, nsym, "symbols")
        if nsym == 1_000_000:
            # the as-specified size is launch-bound (a ~4 us kernel): 16 steps recorded into one CUDA graph (ae_graph_*)
            reps = 16
            cap = torch.cuda.Stream()          # capture needs a real stream: the legacy default stream cannot be captured
            torch.cuda.synchronize()
            with torch.cuda.stream(cap):
                ae.use_torch_stream()
                ae.chain.modem_fused(qpsk, g, bits_in, bits_out, st, ae.COMPAT_REFERENCE)
                with ae.Graph() as gr:
                    for _ in range(reps):
                        ae.chain.modem_fused(qpsk, g, bits_in, bits_out, st, ae.COMPAT_REFERENCE)
                t = timed(torch, gr.launch, steps, warmup) / steps / reps
                gr.close()
            torch.cuda.synchronize()
            ae.use_torch_stream()
            rec("modem_fused_%dsym_graph16" % nsym, 4.0 * nsym, t, nsym, "symbols")
            out["modem_fused_%dsym_graph16" % nsym]["us_per_step"] = t * 1e6
        del bits_in, bits_out
    # config 1 stand-alone pieces: modulate, Awgn::apply, Awgn::fill, demod on 2^26 symbols
    nsym = 1 << 26
    bits_in = ae.DeviceBits.zeros(2 * nsym)
    sym = ae.DeviceVec.zeros(nsym)
    t = timed(torch, lambda: qpsk.modulate_into(bits_in, sym), steps, warmup) / steps
    rec("modulate_qpsk", 10.0 * nsym, t, nsym, "symbols")
    t = timed(torch, lambda: g.apply(sym), steps, warmup) / steps
    rec("awgn_apply", 16.0 * nsym, t, nsym, "samples")
    fillv = ae.DeviceVec.with_capacity(nsym)

    def fill():
        fillv.clear()
        g.fill(fillv)
    t = timed(torch, fill, steps, warmup) / steps
    rec("awgn_fill", 8.0 * nsym, t, nsym, "samples")
    bits_out = ae.DeviceBits.with_capacity(2 * nsym)

    def demod():
        bits_out.clear()
        qpsk.demod_naive(sym, bits_out)
    t = timed(torch, demod, steps, warmup) / steps
    rec("demod_qpsk", 10.0 * nsym, t, nsym, "symbols")
    del bits_in, sym, fillv, bits_out
    # 8(f): spectrogram core (8 B in, 4 B out) and correlator (16 B) on the chain's input
    lv = ae.spectral.DeviceF32(n)
    t = timed(torch, lambda: ae.spectral.spectrogram(d_in, fft, True, lv), steps, warmup) / steps
    rec("spectrogram1024_dB", 12.0 * n, t, n, "samples")
    del lv
    sigv = ae.DeviceVec.zeros(FFT_LEN)
    t = timed(torch, lambda: ae.spectral.correlate(d_in, sigv, fft, ae.Scale.SN, howmany=frames), steps, warmup) / steps
    rec("correlator1024", 16.0 * n, t, n, "samples")
    # any-length path (mixed-radix shared-memory kernel): N = 100 is the reference's own test length
    for nn in (100, 1000):
        fr = (1 << 27) // nn
        fo = ae.Cfft.with_len(nn)
        vo = d_in.view(0, fr * nn)
        t = timed(torch, lambda: fo.ifwd(vo, ae.Scale.SN, howmany=fr), steps, warmup) / steps
        rec("fft%d_fwd_SN" % nn, 16.0 * fr * nn, t, fr * nn, "samples")
    t = timed(torch, lambda: d_in.vec_stats(), steps, warmup) / steps
    rec("vecstats_cf32", 8.0 * n, t, n, "samples")
    # config 5: OFDM-like chain, 2048-pt, 2^18 frames, counters only
    fr = 1 << 18
    t = timed(torch, lambda: ae.chain.ofdm_chain(2048, fr, 0, 0.05, 5, st), steps, warmup) / steps
    out["ofdm2048_chain"] = {"ms": t * 1e3, "Gsymbols/s": fr * 2048 / t / 1e9, "bytes_note": "counters only; compute-bound by construction"}
    return out


def bind_to_gpu_numa_node(torch, local_rank: int) -> None:
    """Pin this rank's host threads to the CPUs next to its GPU, so the pinned e2e buffers are
    allocated on the GPU's NUMA node (first touch) and H2D/D2H do not cross the socket link."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        if ids:
            os.sched_setaffinity(0, ids)
    except Exception:
        pass


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch

    import aether_primitives_b200 as ae
    from aether_primitives_b200.chain import FftFirDemod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ae.init(local_rank)
    ae.use_torch_stream()
    hbm_peak, peak_src = measured_peaks()

    frames = args.frames
    n = frames * FFT_LEN
    taps = make_taps()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1 + rank)
    x = torch.view_as_complex(torch.randn(n, 2, device="cuda", dtype=torch.float32, generator=gen))
    bits = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    d_in = ae.DeviceVec.from_torch(x)
    d_bits = ae.DeviceBits.wrap(bits.data_ptr(), bits.numel(), owner=bits)
    chain = FftFirDemod(FFT_LEN, taps, ae.Scale.SN, ae.COMPAT_REFERENCE)

    def barrier():
        if dist:
            dist.barrier()

    def step():
        chain.run(d_in, d_bits)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    l0 = ae.launch_count()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l_warm = ae.launch_count() - l0
    t_mark0 = sampler.mark()
    l1 = ae.launch_count()
    secs = timed(torch, step, args.steps, 0, barrier)
    launches = ae.launch_count() - l1
    t_mark1 = sampler.mark()
    tt = torch.tensor([secs], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    secs_max = float(tt.item())
    value = world * n * args.steps / secs_max / 1e9
    ms_per_step = secs_max / args.steps * 1e3
    kernel_s = secs / max(1, launches)  # one kernel per step: launch duration == step duration on this rank
    achieved = ALG_BYTES_PER_SAMPLE * n / kernel_s / 1e9
    # DRAM traffic of the same kernel from the committed ncu --set full capture, scaled to this launch
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "chain_traffic.json")))
        traffic = tj["dram_bytes_per_sample"] * n
    except Exception:
        pass

    # ---- integrity: bits checksum, cross-rank sum (outside the timed region) ----
    ones = torch.count_nonzero(bits).to(torch.int64)
    if dist:
        dist.all_reduce(ones)

    # ---- e2e through the C ABI with HOST buffers (pinned): H2D + kernel + D2H every step ----
    e2e = None
    if not args.no_e2e:
        e2e_frames = frames
        avail = 0
        try:
            for l in open("/proc/meminfo"):
                if l.startswith("MemAvailable"):
                    avail = int(l.split()[1]) * 1024
        except Exception:
            pass
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        need = 10 * e2e_frames * FFT_LEN * local_world
        while avail and need * 2 > avail and e2e_frames > 8192:
            e2e_frames //= 2
            need //= 2
        ne = e2e_frames * FFT_LEN
        affinity0 = os.sched_getaffinity(0)
        bind_to_gpu_numa_node(torch, local_rank)   # pinned buffers on the GPU's NUMA node
        h_in = torch.empty(ne, dtype=torch.complex64).pin_memory()
        h_out = torch.empty(2 * ne, dtype=torch.uint8).pin_memory()
        h_in.copy_(x[:ne])
        torch.cuda.synchronize()

        def e2e_step():
            chain.run_host(h_in.data_ptr(), ne, h_out.data_ptr())   # returns after the D2H of the result

        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        l2 = ae.launch_count()
        for _ in range(e_steps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        e_secs = max(e0.elapsed_time(e1) / 1e3, wall)
        barrier()
        te = torch.tensor([e_secs], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ne * e_steps / float(te.item()) / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": 8 * ne,
               "d2h_bytes_per_step": 2 * ne, "steps": e_steps, "frames_per_step_per_gpu": e2e_frames,
               "launches_per_step": (ae.launch_count() - l2) // e_steps,
               "note": "ae_chain_exec_host: pinned host buffers, 64 MiB chunks, 3-stream H2D/kernel/D2H pipeline; PCIe-bound"}
        # the host path must give the same bits as the device path
        same = bool(torch.equal(h_out.cuda(), bits[: 2 * ne]))
        e2e["matches_device_path"] = same
        # yard-stick: a bare pinned-host -> device copy of the same input buffer (the PCIe ceiling of e2e)
        d_tmp = torch.empty(ne, dtype=torch.complex64, device="cuda")
        d_tmp.copy_(h_in, non_blocking=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            d_tmp.copy_(h_in, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d = 3 * 8 * ne / (c0.elapsed_time(c1) / 1e3) / 1e9
        e2e["h2d_copy_GB/s"] = h2d
        e2e["frac_of_h2d_copy"] = (8 * ne * e_steps / e_secs / 1e9) / h2d
        # yard-stick of the WHOLE job: every rank at once (same barrier), both directions at once (the step moves 8 B in
        # and 2 B out per sample), bare cudaMemcpyAsync on two streams: what the host fabric gives N concurrent ranks
        d_out_tmp = torch.empty(2 * ne, dtype=torch.uint8, device="cuda")
        h_out2 = torch.empty(2 * ne, dtype=torch.uint8).pin_memory()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        torch.cuda.synchronize()
        tw0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(s_in):
                d_tmp.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out2.copy_(d_out_tmp, non_blocking=True)
        torch.cuda.synchronize()
        ty = torch.tensor([time.perf_counter() - tw0], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(ty, op=dist.ReduceOp.MAX)
        yard = world * ne * 3 / float(ty.item()) / 1e9          # Gsamples/s the bare copies of all ranks sustain together
        e2e["concurrent_copy_yardstick_Gsamples/s"] = yard
        e2e["frac_of_concurrent_yardstick"] = e2e["value"] / yard
        del d_tmp, d_out_tmp, h_out2
        # the same job through the GENERAL stage pipeline (ae_pipeline_*: pipeline.rs / pool.rs of the reference on CUDA
        # streams): three user-defined stages over pooled device blocks, fed with 64 MiB slices of the same pinned buffers.
        # No collective inside the try: a rank that fails must not leave the others waiting.
        gp, gp_secs = None, float("inf")
        try:
            chunk_frames = min(8192, e2e_frames)
            cs = chunk_frames * FFT_LEN
            n_chunks = ne // cs

            class _Blk:
                def __init__(self):
                    self.d = ae.DeviceVec.zeros(cs)
                    self.bits = ae.DeviceBits.zeros(2 * cs)
                    self.off = 0

            blk_pool = ae.pool.make(0, _Blk, lambda b: None)
            hin_p, hout_p = h_in.data_ptr(), h_out.data_ptr()

            def st_h2d(e):
                e.val.d.upload_async(hin_p + 8 * e.val.off, cs)
                return e

            def st_chain(e):
                chain.run(e.val.d, e.val.bits)
                return e

            def st_d2h(e):
                e.val.bits.download_async(hout_p + 2 * e.val.off, 2 * cs)
                return e

            h_out.zero_()
            tx, rx = ae.pipeline.Pipeline.new("h2d", st_h2d, depth=3).add_stage("fft-fir-demod", st_chain).add_stage("d2h", st_d2h).finish()

            def gp_step():
                for ci in range(n_chunks):
                    if rx.in_flight() == 3:
                        rx.recv().release()
                    e = blk_pool.take_or_make()
                    e.val.off = ci * cs
                    tx.send(e)
                while rx.in_flight():
                    rx.recv().release()

            gp_step()
            rx.report(reset=True)
            tg0 = time.perf_counter()
            for _ in range(3):
                gp_step()
            gp_secs = time.perf_counter() - tg0
            rep = rx.report()
            gp = {"chunk_frames": chunk_frames, "depth": 3, "samples_per_rank": n_chunks * cs * 3,
                  "matches_device_path": bool(torch.equal(h_out[: 2 * n_chunks * cs].cuda(), bits[: 2 * n_chunks * cs])),
                  "stage_utilisation_pct": {r["name"]: round(r["utilisation_pct"], 1) for r in rep},
                  "note": "ae_pipeline_*: user stages h2d / chain / d2h on three streams, pooled device blocks, host closures in Python"}
            del tx, rx, blk_pool
        except Exception as ex:
            gp = {"error": repr(ex)}
        tg = torch.tensor([gp_secs], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        if "error" not in gp and float(tg.item()) != float("inf"):
            gp["Gsamples/s"] = world * gp.pop("samples_per_rank") / float(tg.item()) / 1e9
        e2e["general_pipeline"] = gp
        del h_in, h_out
        try:
            os.sched_setaffinity(0, affinity0)       # the CPU baseline below uses every host core
        except Exception:
            pass
    sampler.stop()
    clocks = sampler.summary(t_mark0, t_mark1)

    # ---- BASELINE config 5 on every rank's frame shard: OFDM-like chain, BER/EVM counters summed over
    #      ranks with the path's only collective (<= 32 bytes per rank, NCCL through torch.distributed)
    config5 = None
    try:
        from aether_primitives_b200.sharding import frame_range
        from aether_primitives_b200.stats import Comm, DeviceStats, evm_db

        total_frames = (1 << 18) * world                      # 2^18 frames per GPU
        f0, f1 = frame_range(total_frames, rank, world)
        st = DeviceStats()
        ofdm = lambda: ae.chain.ofdm_chain(2048, f1 - f0, f0, 0.5, 5, st, None, None, ae.COMPAT_CORRECTED)
        for _ in range(3):
            ofdm()
        st.zero()
        t5 = timed(torch, ofdm, 5, 0, barrier) / 5
        tt5 = torch.tensor([t5], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tt5, op=dist.ReduceOp.MAX)
            # the path's only collective, through the C ABI: ae_comm_init_rank + ae_stats_allreduce (NCCL inside the
            # library; torch.distributed only carried the 128-byte id to the ranks)
            comm = Comm.from_torch_distributed()
            st.allreduce(comm)
            red = st.read()
            comm.close()
        else:
            red = st.read()
        config5 = {"workload": "mseq -> QPSK -> 2048-pt bwd FFT(SN) -> AWGN -> fwd FFT(SN) -> demod -> BER/EVM", "frames_total": total_frames,
                   "Gsymbols/s": total_frames * 2048 / float(tt5.item()) / 1e9, "ms": float(tt5.item()) * 1e3,
                   "bit_errors": red["bit_errors"], "n_bits": red["n_bits"], "ber": red["bit_errors"] / max(1, red["n_bits"]),
                   "evm_db": evm_db(red["err_pow"], red["ref_pow"]), "collective": "ae_stats_allreduce (ncclAllReduce inside libaether_b200.so), %d ranks" % world if dist else "single rank"}
    except Exception as ex:
        config5 = {"error": repr(ex)}

    extra = None
    cpu = None
    if not args.no_extras:
        try:
            extra = extras(torch, ae, d_in, frames, hbm_peak, rank=rank, world=world)
        except Exception as ex:  # extras never invalidate the headline line
            extra = {"error": repr(ex)}
        if dist:
            # every named shape at N GPUs: slowest rank's time, units summed over ranks; frac_hbm stays per GPU
            gathered = [None] * world
            dist.all_gather_object(gathered, extra)
            if rank == 0 and all(isinstance(g, dict) and "error" not in g for g in gathered):
                comb = {}
                for name, e0 in gathered[0].items():
                    ms = max(g[name]["ms"] for g in gathered)
                    row = {"ms": ms, "n_gpus": world}
                    for k in e0:
                        if k.startswith("G") and k.endswith("/s"):
                            units = sum(g[name][k] * g[name]["ms"] for g in gathered)     # G-units x ms, per rank
                            row[k] = units / ms
                            row[k + "_per_gpu"] = row[k] / world
                    if "frac_hbm" in e0:
                        row["frac_hbm_per_gpu"] = min(g[name]["frac_hbm"] * g[name]["ms"] / ms for g in gathered)
                    for k in ("halo_samples", "strong_scaling_total_samples", "bytes_note"):
                        if k in e0:
                            row[k] = e0[k]
                    comb[name] = row
                extra = comb
    if rank == 0 and world == 1:
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            gs, ms, cf = cpu_chain_rate(10.0, threads)
            gs1, ms1, cf1 = cpu_chain_rate(3.0, 1)
            cpu = {"value": gs, "unit": "Gsamples/s", "cores": threads, "kind": "port",
                   "sample": "%d frames x %d samples, %d host threads, %.1f s; C++ restatement of aether_primitives (not rustfft)" % (cf, FFT_LEN, threads, ms / 1e3),
                   "single_thread_value": gs1}
            try:  # second, independent CPU yard-stick for the FFT alone: scipy's pocketfft, complex64
                import scipy.fft as sfft

                xf = (np.random.default_rng(2).standard_normal((1 << 14, FFT_LEN)) + 0j).astype(np.complex64)
                sfft.fft(xf, axis=1, workers=threads)
                t0 = time.perf_counter()
                for _ in range(5):
                    sfft.fft(xf, axis=1, workers=threads)
                cpu["scipy_pocketfft_fft1024_Gsamples/s"] = 5 * xf.size / (time.perf_counter() - t0) / 1e9
            except Exception:
                pass
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "fft1024_fwd_SN -> fir64 -> qpsk_demod (BASELINE headline chain)", "fft_len": FFT_LEN, "ntaps": NTAPS,
                       "frames_per_gpu": frames, "samples_per_gpu": n, "compat": "reference", "parallelism": "frames sharded, dp%d" % world,
                       "l2_policy": "inputs larger than L2 (%.1f GiB in, %.1f GiB out per step)" % (8 * n / 2**30, 2 * n / 2**30)},
            "roofline": {"bound": "hbm", "kernel": "chain_x2_kernel<1024> (K14b)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": "ncu dram__bytes_read+write per sample from profiles/chain_traffic.json x samples per launch",
                         "peak_source": peak_src,
                         "alg_bytes_per_sample": ALG_BYTES_PER_SAMPLE, "kernel_ms": kernel_s * 1e3,
                         "limiter": "not HBM: instruction dispatch.  Packed FP32 (FFMA2/FADD2), 16-lane ALU and LSU instructions hold the "
                                    "sub-partition's dispatch port for two cycles each; the sum of the SASS stall fields of one frame "
                                    "(~3300 cycles per warp-frame) is the frame time within 10 % (DESIGN.md 5.2, profiles/r2_*)",
                         "fp32_TFLOP/s_nominal": FLOP_PER_SAMPLE * n / kernel_s / 1e12},
            "e2e": e2e, "gpu_launches": launches, "warmup_launches": l_warm, "clocks": clocks,
            "checksum_ones": int(ones.item()),
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        if config5:
            line["config5_ofdm"] = config5
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under
    torchrun), so everything else written to fd 1 during the run is sent to stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per GPU (BASELINE: 2^20)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
