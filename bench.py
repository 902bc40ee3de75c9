#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 DSP hot path (contract: see the task's "Measurement" section).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference] [--no-extra]

One "step" = one pass of the hot path over one batch of synthetic tuner I/Q (1 s of a 10 MS/s stream).
N > 1 is launched by torchrun, one rank per GPU; tuner streams are independent so every rank processes its own
stream (weak scaling, no data-path collective); torch.distributed is used for the barrier and the
max-over-ranks time only.

Keys of the JSON line (rank 0): value = whole-job input complex MS/s with inputs resident in HBM; e2e = the same
through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region); roofline = dominant kernel
against the measured HBM peak; cpu_baseline = the oracle (CPU restatement of the reference's Java path) timed
on this box's host cores on a bounded sample; chain_c4fm = the same measurements for BASELINE configs[2]
(channelizer -> FIR -> AGC -> DQPSK timing recovery on all 400 channels) as a secondary result.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "channelized+demodulated complex MS/s and real-time channel count at 1/2/4/8 GPU"
UNIT = "MS/s"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # FFMA lanes x 2 flop x max SM clock

WORKLOADS = {
    # BASELINE.json configs[1]: polyphase channelizer, 10 MS/s -> 400 x 25 kHz channels
    "channelizer": dict(fs=1e7, taps_per_channel=9, seconds=1.0, demod=None,
                        desc="configs[1]: polyphase channelizer 10 MS/s -> 400 x 25 kHz channels (M=400, T=9), "
                             "all bins kept, gain M, [channel][time] output"),
    # BASELINE.json configs[2]: channelizer + per-channel FIR/AGC + C4FM DQPSK timing recovery on 400 channels
    "c4fm": dict(fs=1e7, taps_per_channel=9, seconds=1.0, demod="c4fm",
                 desc="configs[2]: channelizer 10 MS/s -> 400 channels + 72-tap FIR + block AGC + DQPSK "
                      "decision-directed timing recovery (P25 Phase 1 C4FM) on all 400 channels, dibits out"),
    # BASELINE.json configs[4], one GPU's share: one 20 MS/s tuner -> 800 channels (M = 800)
    "channelizer_20m": dict(fs=2e7, taps_per_channel=9, seconds=1.0, demod=None,
                            desc="configs[4] per-GPU share: polyphase channelizer 20 MS/s -> 800 x 25 kHz channels "
                                 "(M=800, T=9), all bins kept, gain M, [channel][time] output"),
    "c4fm_20m": dict(fs=2e7, taps_per_channel=9, seconds=1.0, demod="c4fm",
                     desc="configs[4] per-GPU share: 20 MS/s tuner -> 800 channels + 72-tap FIR + AGC + C4FM DQPSK "
                          "timing recovery on all 800 channels, dibits out"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel):
    """per-launch DRAM bytes of `kernel` from the committed ncu --set full summary (profiles/traffic.json)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for name, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU baseline
def c4fm_fir_taps():
    import scipy.signal as ss
    # P25P1DecoderC4FM.java:136-148: Remez low-pass 5100/6500 Hz at the 50 kHz channel rate (est. 72 taps).  The
    # Remez designer itself is a SURVEY section 8(f) "next" item; taps are an input of ComplexFIRFilter2.
    return ss.remez(72, [0, 5100, 6500, 25000], [1, 0], fs=50000).astype(np.float32)


def synth_input_numpy(n_complex, m, seed):
    """bounded CPU-side sample of the synthetic workload: tones on bin centres + AWGN, interleaved float32"""
    rng = np.random.default_rng(seed)
    t = np.arange(n_complex)
    z = 1e-3 * (rng.standard_normal(n_complex) + 1j * rng.standard_normal(n_complex))
    for k in rng.choice(m, 8, replace=False):
        f = (k if k < m // 2 else k - m) / m
        z += 0.05 * np.exp(2j * np.pi * (f * t + rng.uniform()))
    out = np.empty(2 * n_complex, np.float32)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


class CpuChain:
    """The oracle (kind "port": CPU restatement of the reference Java path; no JVM here) arranged like the
    reference runs it: one channelizer per tuner stream, one output processor + decoder chain per channel.  One
    independent tuner stream per host thread (the oracle releases the GIL inside its ctypes calls)."""

    def __init__(self, workload, threads):
        import oracle
        self.cfg = WORKLOADS[workload]
        self.workload = workload
        fs = self.cfg["fs"]
        self.m = m = int(fs / 25000) // 2 * 2
        taps = oracle.sinc_m2_channelizer(fs / m, m, self.cfg["taps_per_channel"])
        # bounded sample: 1 M input samples, or 2 whole 1024-sample assembler buffers per channel for the chain
        self.n_complex = 2 * 1024 * (m // 2) if self.cfg["demod"] else 1000000
        self.x = synth_input_numpy(self.n_complex, m, 1)
        fir = c4fm_fir_taps() if self.cfg["demod"] == "c4fm" else None
        self.threads = threads
        self.states = []
        for _ in range(threads):
            st = {"chan": oracle.Channelizer(taps, m)}
            if self.cfg["demod"] == "c4fm":
                st["procs"] = [oracle.OneChannelOutputProcessor(50000.0, k, float(m)) for k in range(m)]
                st["chains"] = [oracle.P25Chain(oracle.C4FM, 50000.0, fir) for _ in range(m)]
            self.states.append(st)

    def one_pass(self, st):
        res = st["chan"].receive(self.x)
        if self.cfg["demod"] == "c4fm":
            for k in range(self.m):
                y = st["procs"][k].process(res)
                st["chains"][k].receive(y[: y.size // 2048 * 2048])
        else:
            # the reference's per-channel extraction (ReusableChannelResultsBuffer.getChannel + gain) is part of
            # configs[1]'s [channel][time] output
            import oracle
            if "procs" not in st:
                st["procs"] = [oracle.OneChannelOutputProcessor(50000.0, k, float(self.m)) for k in range(self.m)]
            for k in range(self.m):
                st["procs"][k].process(res)

    def run(self, reps, threads=None):
        """all threads x reps passes; returns (seconds, complex input samples processed)"""
        threads = threads or self.threads
        ths = [threading.Thread(target=lambda s=self.states[i]: [self.one_pass(s) for _ in range(reps)])
               for i in range(threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0, threads * reps * self.n_complex

    def describe(self, reps):
        return ("%d threads x %d passes over %d complex input samples each (%s); oracle = C restatement of the "
                "reference Java path, JVM unavailable" % (self.threads, reps, self.n_complex, self.workload))


def cpu_baseline(workload, threads, target_seconds=10.0):
    chain = CpuChain(workload, threads)
    chain.run(1, 1)                                   # warm-up
    single_s, n = chain.run(1, 1)
    reps = max(1, int(target_seconds / max(single_s * 1.3, 1e-3)))
    dt, total = chain.run(reps)
    return {"value": total / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "single_thread_value": n / single_s / 1e6, "sample": chain.describe(reps)}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """Times the reference's CPU implementation of the path (the oracle port; the Java itself cannot run here:
    no JVM) on all host threads.  One step = every thread runs `reps` passes over the bounded sample."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg = WORKLOADS[args.workload]
    chain = CpuChain(args.workload, cores)
    chain.run(1, 1)
    single_s, _ = chain.run(1, 1)
    reps = max(1, int(2.5 / max(single_s * 1.3, 1e-3)))     # ~2.5-3 s per step: thread start / join stays < 1 %
    for _ in range(args.warmup):
        chain.run(reps)
    dt, total = 0.0, 0
    for _ in range(args.steps):
        d, n = chain.run(reps)
        dt += d
        total += n
    value = total / dt / 1e6
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": chain.describe(reps)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"]}, "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))



# ---------------------------------------------------------------------------------------------- GPU-side synthetic input
def synth_tones_wideband(torch, dev, m, n_complex, seed):
    """one unit-phase-random tone per bin at bin centre + U(-5, 5) kHz, amplitude 1/M each,
    plus AWGN sigma 1e-4 per axis.  Built in the frequency domain (tones on FFT bins of the n_complex-point grid)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    spec = torch.zeros(n_complex, dtype=torch.complex64, device=dev)
    k = torch.arange(m, device=dev)
    centre = torch.where(k < m // 2, k, k - m).double() / m                       # cycles per sample
    off = (torch.rand(m, device=dev, generator=g, dtype=torch.float64) - 0.5) * (10000.0 / (25000.0 * m))
    bins = torch.round((centre + off) * n_complex).long() % n_complex
    ph = 2 * np.pi * torch.rand(m, device=dev, generator=g, dtype=torch.float64)
    spec[bins] = torch.polar(torch.full((m,), float(n_complex) / m, device=dev, dtype=torch.float64), ph).to(torch.complex64)
    z = torch.fft.ifft(spec)
    del spec
    x = torch.view_as_real(z).reshape(-1).contiguous()
    x += 1e-4 * torch.randn(x.shape, device=dev, generator=g, dtype=torch.float32)
    return x


def synth_c4fm_wideband(torch, dev, m, n_ch, seed, amplitude=0.02, noise=2e-3):
    """every bin carries a continuous C4FM stream (4800 sym/s, deviation levels +/-1, +/-3 ->
    phase steps of level * pi/4 per symbol, raised-cosine frequency pulses), random dibits, carrier offset U(-200, 200)
    Hz and timing phase per channel; channels are multiplexed onto the bin centres in the frequency domain."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    fs_ch, sym_rate = 50000.0, 4800.0
    sps = fs_ch / sym_rate
    n_sym = int(n_ch / sps) + 8
    dibits = torch.randint(0, 4, (m, n_sym), device=dev, generator=g)
    levels = torch.tensor([1.0, 3.0, -1.0, -3.0], device=dev, dtype=torch.float64)[dibits]
    off = (torch.rand(m, 1, device=dev, generator=g, dtype=torch.float64) - 0.5) * 400.0
    tphase = torch.rand(m, 1, device=dev, generator=g, dtype=torch.float64)
    phase0 = 2 * np.pi * torch.rand(m, 1, device=dev, generator=g, dtype=torch.float64)
    n = torch.arange(n_ch, device=dev, dtype=torch.float64).unsqueeze(0)
    n_w = n_ch * m // 2
    spec = torch.zeros(n_w, dtype=torch.complex64, device=dev)
    quarter = n_ch // 4
    j = torch.cat([torch.arange(0, quarter, device=dev), torch.arange(-quarter, 0, device=dev)])
    for lo in range(0, m, 50):                       # 50 channels at a time bounds the float64 temporaries
        hi = min(m, lo + 50)
        t = n / sps - tphase[lo:hi]
        k = torch.floor(t).long()
        valid = (k >= 0) & (k < n_sym)
        u = t - k - 0.5
        pulse = 1.0 + torch.cos(2 * np.pi * u)
        lv = torch.gather(levels[lo:hi], 1, k.clamp(0, n_sym - 1))
        freq = torch.where(valid, lv * pulse, torch.zeros_like(pulse))
        phase = torch.cumsum(freq, dim=1) * (np.pi / 4.0) / sps + 2 * np.pi * off[lo:hi] * n / fs_ch + phase0[lo:hi]
        xf = torch.fft.fft(torch.polar(torch.full_like(phase, amplitude), phase).to(torch.complex64), dim=1)
        for c in range(lo, hi):
            centre = (c if c < m // 2 else c - m) * (n_ch // 2)
            spec[(centre + j) % n_w] += xf[c - lo, j % n_ch]
        del t, k, valid, u, pulse, lv, freq, phase, xf
    z = torch.fft.ifft(spec) * (n_w / n_ch)
    del spec
    x = torch.view_as_real(z).reshape(-1).contiguous()
    x += noise * torch.randn(x.shape, device=dev, generator=g, dtype=torch.float32)
    return x, dibits.to(torch.uint8).cpu().numpy()


def dibit_match(decoded, truth, skip=300):
    """fraction of decoded dibits equal to the transmitted ones after acquisition, best over small lags"""
    best = 0.0
    for lag in range(0, 24):
        n = min(decoded.size - lag, truth.size) - 20
        if n > skip + 100:
            best = max(best, float(np.mean(decoded[lag + skip:lag + n] == truth[skip:n])))
    return best

def synth_dqpsk_channels(torch, dev, n_channels, n, seed, symbol_rate=6000.0, fs=50000.0, beta=0.35, noise=0.03):
    """channel-domain pi/4-DQPSK streams (raised-cosine shaped impulse train evaluated at the
    non-integer 8.33 samples per symbol), carrier offset U(-200, 200) Hz and timing phase per channel, + AWGN."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sps = fs / symbol_rate
    n_sym = int(n / sps) + 16
    out = torch.empty((n_channels, 2 * n), dtype=torch.float32, device=dev)
    truth = torch.empty((n_channels, n_sym), dtype=torch.uint8)
    levels = torch.tensor([1.0, 3.0, -1.0, -3.0], device=dev, dtype=torch.float64)
    idx = torch.arange(n, device=dev, dtype=torch.float64).unsqueeze(0)
    for lo in range(0, n_channels, 256):
        hi = min(n_channels, lo + 256)
        c = hi - lo
        dib = torch.randint(0, 4, (c, n_sym), device=dev, generator=g)
        truth[lo:hi] = dib.to(torch.uint8).cpu()
        sym = torch.polar(torch.ones((c, n_sym), device=dev, dtype=torch.float64),
                          torch.cumsum(levels[dib] * (np.pi / 4.0), dim=1))
        off = (torch.rand(c, 1, device=dev, generator=g, dtype=torch.float64) - 0.5) * 400.0
        tph = torch.rand(c, 1, device=dev, generator=g, dtype=torch.float64)
        t = idx / sps - tph
        k0 = torch.floor(t).long()
        z = torch.zeros((c, n), dtype=torch.complex128, device=dev)
        for dk in range(-6, 8):
            k = k0 + dk
            valid = (k >= 0) & (k < n_sym)
            u = t - k
            den = 1.0 - (2 * beta * u) ** 2
            pulse = torch.where(den.abs() < 1e-9, torch.full_like(u, (np.pi / 4) * np.sinc(1 / (2 * beta))),
                                torch.sinc(u) * torch.cos(np.pi * beta * u) / den)
            z += torch.where(valid, torch.gather(sym, 1, k.clamp(0, n_sym - 1)) * pulse, torch.zeros_like(z))
        z = z * torch.polar(torch.ones_like(idx), 2 * np.pi * off * idx / fs)
        zr = torch.view_as_real(z.to(torch.complex64)).reshape(c, 2 * n)
        out[lo:hi] = zr + noise * torch.randn(zr.shape, device=dev, generator=g, dtype=torch.float32)
        del sym, t, k0, z, zr
    return out, truth.numpy()


def run_config4(args, rank, world, local_rank):
    """BASELINE configs[3]: P25 Phase 2 HDQPSK on >= 1000 channel-domain streams (no channelizer): 154-tap FIR -> AGC ->
    Gardner timing recovery (demodulator layout chosen by the bank size).  Channels are sharded over the ranks (level 3 of SURVEY 8e)."""
    import scipy.signal as ss
    import torch
    import torch.distributed as dist
    from sdrtrunk_b200 import native
    from sdrtrunk_b200.dsp import Bank
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    native.init(local_rank)
    L = native.lib()
    dev = torch.device("cuda", local_rank)
    channels, n = int(os.environ.get("SDRGPU_BENCH_CHANNELS", "4096")), 24 * 1024   # per GPU; 0.49 s of signal per step
    fir = ss.remez(154, [0, 6500, 7200, 25000], [1, 0], fs=50000).astype(np.float32)
    x, truth = synth_dqpsk_channels(torch, dev, channels, n, seed=4 + 1000 * rank)
    bank = Bank.preset(native.PRESET_P25_HDQPSK, channels, 50000.0, fir, max_samples_per_call=n, device=local_rank)
    stream = torch.cuda.Stream(device=dev)
    bank.setStream(stream.cuda_stream)
    stride = n // 7 + 64
    sym = torch.zeros((channels, stride), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(channels, dtype=torch.int32, device=dev)

    def step():
        native.check(L.sdrgpu_bank_process(bank._h, C.c_void_p(x.data_ptr()), 2 * n, n, native.DEVICE,
                                           C.c_void_p(sym.data_ptr()), stride, None, 0, C.c_void_p(cnt.data_ptr()),
                                           native.DEVICE))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    step()
    barrier()
    first = sym.cpu().numpy()
    counts = cnt.cpu().numpy()
    sanity = {str(c): round(dibit_match(first[c, :counts[c]], truth[c], skip=300), 4) for c in (0, channels // 4, channels - 1)}
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.sdrgpu_launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = L.sdrgpu_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    bank.enableTiming(True)
    step()
    k_filter, k_demod = bank.lastKernelMs()
    # the same step with the complete Phase 2 framing (P25P2SuperFrameDetector) running in the demodulator kernel; the
    # synthetic channels carry no sync patterns, so every symbol goes through the sync detector: the expensive state
    bank.setSyncDetector(native.SYNC_P25_PHASE2_FRAMED)
    step()
    _, k_demod_framed = bank.lastKernelMs()
    bank.setSyncDetector(native.SYNC_NONE)
    if rank == 0:
        ms_step = ms / args.steps
        total = channels * n * world
        peak, peak_src = load_peaks()
        alg = 8.0 * channels * n + channels * n * 6000.0 / 50000.0
        line = {"metric": METRIC, "value": total / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[3]: P25 Phase 2 HDQPSK, %d channel-domain streams per GPU at 50 kHz "
                                       "(channel samples, not tuner samples): 154-tap FIR + AGC + Gardner DQPSK "
                                       "timing recovery" % channels,
                           "channels_per_gpu": channels, "samples_per_channel_per_step": n,
                           "sharding": "channel rows per GPU, no collective"},
                "realtime_channels": channels * world * (n / 50000.0) / (ms_step * 1e-3),
                "gpu_launches": launches, "decode_sanity": sanity,
                "kernels_ms": {"fir_agc": k_filter, "psk": k_demod, "psk_with_phase2_framing_searching": k_demod_framed},
                "roofline": {"bound": "hbm", "achieved": alg / (k_demod * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / (k_demod * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "psk_kernel<gardner>",
                             "kernel_ms": k_demod, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                             "note": "latency bound per-symbol feedback loop; lanes per channel chosen by the bank size"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- GPU arm
class GpuWorkload:
    """One workload's device state: synthetic input in HBM + pinned host copy, channelizer (+ bank + pipeline)."""

    def __init__(self, name, rank, local_rank):
        import torch
        from sdrtrunk_b200 import native
        from sdrtrunk_b200.dsp import Bank, ComplexPolyphaseChannelizerM2, Pipeline
        self.torch, self.native, self.L = torch, native, native.lib()
        self.name = name
        self.cfg = cfg = WORKLOADS[name]
        fs = cfg["fs"]
        self.m = m = ComplexPolyphaseChannelizerM2.getChannelCount(fs)
        n_complex = int(fs * cfg["seconds"])
        if cfg["demod"]:
            n_complex = n_complex // (1024 * (m // 2)) * (1024 * (m // 2))    # whole assembler buffers
        self.n_complex, self.n_floats = n_complex, 2 * n_complex
        self.n_blocks = n_complex // (m // 2)
        self.dev = dev = torch.device("cuda", local_rank)

        # synthetic tuner I/Q generated on the device, and its pinned host copy
        self.truth = None
        if cfg["demod"] == "c4fm":
            self.x_dev, self.truth = synth_c4fm_wideband(torch, dev, m, self.n_blocks, seed=3 + 1000 * rank,
                                                         amplitude=8.0 / m)
        else:
            self.x_dev = synth_tones_wideband(torch, dev, m, n_complex, seed=2 + 1000 * rank)
        self.x_host = torch.empty(self.n_floats, dtype=torch.float32, pin_memory=True)
        self.x_host.copy_(self.x_dev)

        self.stream = torch.cuda.Stream(device=dev)
        self.chan = ComplexPolyphaseChannelizerM2(fs, cfg["taps_per_channel"], device=local_rank,
                                                  maxInputFloats=self.n_floats)
        self.chan.setStream(self.stream.cuda_stream)
        self.pipeline = None
        if cfg["demod"] == "c4fm":
            self.bank = Bank.preset(native.PRESET_P25_C4FM, m, 2 * fs / m, c4fm_fir_taps(),
                                    max_samples_per_call=self.n_blocks, device=local_rank)
            self.bank.setStream(self.stream.cuda_stream)
            self.pipeline = Pipeline(self.chan, self.bank)
            self.chunks = int(os.environ.get("SDRGPU_BENCH_CHUNKS", "8"))
            self.pipeline.setChunks(self.chunks)
            self.sym_stride = self.n_blocks // 8 + 64          # > 4800/50000 symbols per sample
            self.sym_dev = torch.zeros((m, self.sym_stride), dtype=torch.uint8, device=dev)
            self.cnt_dev = torch.zeros(m, dtype=torch.int32, device=dev)
            self.sym_host = torch.zeros((m, self.sym_stride), dtype=torch.uint8, pin_memory=True)
            self.cnt_host = torch.zeros(m, dtype=torch.int32, pin_memory=True)
            self.h2d, self.d2h = 4 * self.n_floats, self.sym_host.numel() + 4 * m
        else:
            self.out_dev = torch.empty((m, 2 * self.n_blocks), dtype=torch.float32, device=dev)
            self.out_host = torch.empty((m, 2 * self.n_blocks), dtype=torch.float32, pin_memory=True)
            self.h2d, self.d2h = 4 * self.n_floats, self.out_host.numel() * 4

    def step_device(self):
        n, L = self.native, self.L
        if self.pipeline is None:
            n.check(L.sdrgpu_chan_process(self.chan._h, C.c_void_p(getattr(self, '_dev_in', self.x_dev).data_ptr()), self.n_floats, n.DEVICE,
                                          C.c_void_p(self.out_dev.data_ptr()), 2 * self.n_blocks, n.DEVICE,
                                          n.LAYOUT_CHANNELS, None))
        else:
            n.check(L.sdrgpu_pipeline_process(self.pipeline._h, C.c_void_p(getattr(self, '_dev_in', self.x_dev).data_ptr()), self.n_floats,
                                              n.DEVICE, C.c_void_p(self.sym_dev.data_ptr()), self.sym_stride, None, 0,
                                              C.c_void_p(self.cnt_dev.data_ptr()), n.DEVICE))

    def step_host(self):
        n, L = self.native, self.L
        if self.pipeline is None:
            n.check(L.sdrgpu_chan_process(self.chan._h, C.c_void_p(getattr(self, '_host_in', self.x_host).data_ptr()), self.n_floats, n.HOST,
                                          C.c_void_p(self.out_host.data_ptr()), 2 * self.n_blocks, n.HOST,
                                          n.LAYOUT_CHANNELS, None))
        else:
            n.check(L.sdrgpu_pipeline_process(self.pipeline._h, C.c_void_p(getattr(self, '_host_in', self.x_host).data_ptr()),
                                              self.n_floats, n.HOST, C.c_void_p(self.sym_host.data_ptr()), self.sym_stride, None, 0,
                                              C.c_void_p(self.cnt_host.data_ptr()), n.HOST))

    def enable_u8_input(self, on):
        """section 8f #1: the tuner's native unsigned 8-bit samples, converted on the device (4x less H2D)"""
        torch = self.torch
        if on and not hasattr(self, "x_host_u8"):
            q = torch.clamp(torch.round(self.x_dev * 128.0 + 127.0), 0, 255).to(torch.uint8)
            self.x_host_u8 = torch.empty(self.n_floats, dtype=torch.uint8, pin_memory=True)
            self.x_host_u8.copy_(q)
            del q
        self.chan.setSampleFormat("u8" if on else "f32")
        self._host_in = self.x_host_u8 if on else self.x_host

    def enable_airspy_input(self, on):
        """section 8f #1: Airspy native buffers (packed 12-bit real samples at twice the complex rate, 3 bytes per
        complex sample), unpacked + DC-removed + Hilbert-transformed on the device (AirspySampleConverter).  Random ADC
        codes: this leg measures throughput, parity is tests/."""
        torch = self.torch
        if on and not hasattr(self, "x_host_airspy"):
            g = torch.Generator(device=self.dev)
            g.manual_seed(77)
            nbytes = self.n_floats // 2 * 3
            self.x_dev_airspy = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=self.dev, generator=g)
            self.x_host_airspy = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self.x_host_airspy.copy_(self.x_dev_airspy)
        self.chan.setSampleFormat("airspy_packed" if on else "f32")
        self._host_in = self.x_host_airspy if on else self.x_host
        self._dev_in = self.x_dev_airspy if on else self.x_dev

    def sanity(self):
        """decoded-vs-transmitted dibits of a few channels (first pass from reset state would be needed for an exact
        check; this is a plausibility figure for the timed workload, parity itself is tests/)"""
        if self.truth is None:
            return None
        self.step_host()
        cnt = self.cnt_host.numpy()
        out = {}
        for c in (0, 57, self.m // 2, self.m - 1):
            dec = self.sym_host.numpy()[c, :cnt[c]]
            out[str(c)] = {"symbols": int(cnt[c]), "match_after_acquisition": round(dibit_match(dec, self.truth[c]), 4)}
        return out

    def kernel_times(self, steps):
        """per-kernel device time, live, CUDA events on the launching stream (own loop so that the per-step event
        synchronisation does not perturb the throughput numbers)"""
        self.chan.enableTiming(True)
        if self.pipeline is not None:
            self.bank.enableTiming(True)
            self.pipeline.setChunks(1)        # one full-size launch per kernel, so that each time is a whole step's
        rows = []
        for _ in range(steps):
            self.step_device()
            row = {"pfb_ifft": self.chan.lastKernelMs()}
            if self.pipeline is not None:
                f, d = self.bank.lastKernelMs()
                row.update({"fir_agc": f, "psk": d})
            rows.append(row)
        self.chan.enableTiming(False)
        if self.pipeline is not None:
            self.bank.enableTiming(False)
            self.pipeline.setChunks(self.chunks)
        return {k: statistics.mean(r[k] for r in rows) for k in rows[0]}


def measure(w, args, world, dist, barrier):
    torch, L = w.torch, w.L

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(w.stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        stop.record(w.stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = start.elapsed_time(stop)
        if world > 1:
            tt = torch.tensor([ms, wall], device=w.dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, wall = tt.tolist()
        return ms, wall

    sanity = None if args.device_only else w.sanity()   # first pass over the stream, from the reset state
    launches0 = L.sdrgpu_launch_count()
    ms_dev, _ = timed(w.step_device, args.steps, args.warmup)
    launches = (L.sdrgpu_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    # host-buffer path: device time of the stream also covers the copies; wall clock is what a caller sees
    if args.device_only:   # profiling runs (ncu): only the device-resident steps
        ms_e2e_dev, wall_e2e = ms_dev, ms_dev
    else:
        ms_e2e_dev, wall_e2e = timed(w.step_host, args.steps, max(3, args.warmup))
    e2e_u8 = None
    if not args.device_only:
        w.enable_u8_input(True)
        ms_u8_dev, wall_u8 = timed(w.step_host, args.steps, 3)
        w.enable_u8_input(False)
        u8_ms = max(ms_u8_dev, wall_u8) / args.steps
        e2e_u8 = {"value": w.n_complex * world / (u8_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": w.n_floats,
                  "d2h_bytes_per_step": w.d2h, "ms_per_step": u8_ms,
                  "input": "unsigned 8-bit tuner samples (ByteSampleConverter format), converted on the device"}
    airspy = None
    if w.pipeline is None:   # channelizer only: random ADC codes would drive the demodulators of the chain off their locks
        w.enable_airspy_input(True)
        ms_a_dev, _ = timed(w.step_device, args.steps, 3)
        airspy = {"input": "Airspy packed 12-bit real samples (3 bytes per complex sample): unpack + DC removal + Hilbert "
                           "transform on the device in front of the channelizer",
                  "device_resident": {"ms_per_step": ms_a_dev / args.steps,
                                      "value": w.n_complex * world / (ms_a_dev / args.steps * 1e-3) / 1e6, "unit": UNIT}}
        if not args.device_only:
            ms_a_host, wall_a = timed(w.step_host, args.steps, 3)
            a_ms = max(ms_a_host, wall_a) / args.steps
            airspy["e2e"] = {"value": w.n_complex * world / (a_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": a_ms,
                             "h2d_bytes_per_step": w.n_floats // 2 * 3, "d2h_bytes_per_step": w.d2h}
        w.enable_airspy_input(False)
    kernels = w.kernel_times(min(args.steps, 10))
    with_sync = None
    if w.pipeline is not None:
        # section 8f #3: the same step with the P25 Phase 1 sync detector + PLL inversion feedback running in the
        # demodulator kernel
        w.bank.setSyncDetector(w.native.SYNC_P25_PHASE1)
        ms_sync, _ = timed(w.step_device, args.steps, 3)
        w.bank.setSyncDetector(w.native.SYNC_NONE)
        with_sync = {"ms_per_step": ms_sync / args.steps,
                     "value": w.n_complex * world / (ms_sync / args.steps * 1e-3) / 1e6, "unit": UNIT}
    corrected = None
    if w.pipeline is not None:
        # the usual sdrtrunk situation: every channel frequency-corrected (OneChannelOutputProcessor + Oscillator);
        # small offsets so that the demodulators keep their locks on the synthetic channels.  Measured last: the
        # selection change restarts the channel oscillators.
        w.chan.setOutputChannels([([k], 37 if k % 2 else -53) for k in range(w.m)])
        ms_corr, _ = timed(w.step_device, args.steps, 3)
        corrected = {"ms_per_step": ms_corr / args.steps,
                     "value": w.n_complex * world / (ms_corr / args.steps * 1e-3) / 1e6, "unit": UNIT}
    ms_per_step = ms_dev / args.steps
    total = w.n_complex * world
    e2e_ms = max(ms_e2e_dev, wall_e2e) / args.steps
    return {"ms_per_step": ms_per_step, "value": total / (ms_per_step * 1e-3) / 1e6,
            "e2e_ms": e2e_ms, "e2e_value": total / (e2e_ms * 1e-3) / 1e6, "launches": launches, "kernels": kernels,
            "sanity": sanity, "e2e_u8": e2e_u8, "with_sync": with_sync, "airspy": airspy, "corrected": corrected}


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sdrtrunk_b200 import native

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    native.init(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peak, peak_src = load_peaks()
    sampler = ClockSampler(local_rank)
    w = GpuWorkload(args.workload, rank, local_rank)
    if rank == 0:
        sampler.start()
    r = measure(w, args, world, dist, barrier)
    clocks = sampler.stop() if rank == 0 else None
    cfg, m, fs = w.cfg, w.m, w.cfg["fs"]

    def roofline_of(w, r):
        # dominant kernel = the one with the largest share of the step
        if w.pipeline is None:
            alg = 24.0 * w.n_complex      # 8 B read + 16 B written per input complex sample, all M bins kept
            # a step of this workload is exactly one launch of the kernel (r["launches"] / steps == 1), so its average
            # launch duration over the timed region is the step time; the per-launch time with a synchronisation after
            # every step (no overlap of one launch's tail with the next one's head) is reported beside it
            single = r["launches"] == args.steps
            k_ms = r["ms_per_step"] if single else r["kernels"]["pfb_ifft"]
            ach = alg / (k_ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": load_traffic("pfb2_kernel") if w.m == 400 else None, "kernel": "pfb2_kernel", "kernel_ms": k_ms,
                    "kernel_ms_isolated": r["kernels"]["pfb_ifft"],
                    "algorithmic_bytes_per_launch": alg, "peak_source": peak_src}
        # chain: report the serial timing-recovery kernel (latency bound) with the HBM bytes it moves, and the
        # FP32 view of the FIR that feeds it
        n_ch = w.n_blocks * m
        k_ms = r["kernels"]["psk"]
        alg = 8.0 * n_ch + n_ch * 4800.0 / 50000.0
        ach = alg / (k_ms * 1e-3) / 1e9
        fir_flop = n_ch * 72 * 4.0
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": "psk_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                "peak_source": peak_src, "note": "one warp per channel, latency bound by the per-symbol loop",
                "kernels_ms": r["kernels"],
                "fir_fp32": {"achieved_tflops": fir_flop / (r["kernels"]["fir_agc"] * 1e-3) / 1e12,
                             "peak_tflops": FP32_PEAK_TFLOPS}}

    main_roofline = roofline_of(w, r)
    h2d, d2h = w.h2d, w.d2h
    n_complex, n_blocks = w.n_complex, w.n_blocks
    extra = None
    if not args.no_extra and args.workload == "channelizer":
        del w
        torch.cuda.empty_cache()
        w2 = GpuWorkload("c4fm", rank, local_rank)
        r2 = measure(w2, args, world, dist, barrier)
        extra = {"workload": w2.cfg["desc"], "value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms_per_step"],
                 "realtime_channels": m * world * (r2["value"] / world) / (fs / 1e6),
                 "e2e": {"value": r2["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": w2.h2d,
                         "d2h_bytes_per_step": w2.d2h, "ms_per_step": r2["e2e_ms"]},
                 "e2e_u8_input": r2["e2e_u8"],
                 "with_sync_detector": r2["with_sync"],
                 "with_frequency_corrected_channels": r2["corrected"],
                 "airspy_input": r2["airspy"],
                 "gpu_launches": r2["launches"], "kernels_ms": r2["kernels"], "roofline": roofline_of(w2, r2),
                 "decode_sanity": r2["sanity"]}
        del w2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cores = os.cpu_count() or 1
    base = cpu_baseline(args.workload, cores) if not args.no_cpu_baseline else None
    if extra is not None and not args.no_cpu_baseline:
        extra["cpu_baseline"] = cpu_baseline("c4fm", cores, target_seconds=8.0)

    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "input_complex_samples_per_step_per_gpu": n_complex,
                   "channels": m, "channel_rate_hz": 2 * fs / m,
                   "l2": "per-step working set (%.0f MB in + %.0f MB out) exceeds the 126 MB L2" %
                         (8 * n_complex / 1e6, 8.0 * m * n_blocks / 1e6),
                   "sharding": "one independent tuner stream per GPU, no collective"},
        "realtime_channels": m * world * (r["value"] / world) / (fs / 1e6),
        "e2e": {"value": r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": r["e2e_ms"]},
        "e2e_u8_input": r["e2e_u8"],
        "gpu_launches": r["launches"],
        "roofline": main_roofline,
        "cpu_baseline": base,
        "clocks": clocks,
    }
    if r.get("sanity") is not None:
        line["decode_sanity"] = r["sanity"]
    if r.get("with_sync") is not None:
        line["with_sync_detector"] = r["with_sync"]
    if r.get("airspy") is not None:
        line["airspy_input"] = r["airspy"]
    if r.get("corrected") is not None:
        line["with_frequency_corrected_channels"] = r["corrected"]
    if extra is not None:
        line["chain_c4fm"] = extra
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="channelizer", choices=sorted(WORKLOADS) + ["hdqpsk_4096"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs[2] chain measurement")
    ap.add_argument("--device-only", action="store_true",
                    help="profiling aid: skip the host-buffer (e2e) pass so that every launch is a full-size one")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "hdqpsk_4096":
        run_config4(args, rank, world, local_rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
