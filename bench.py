#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 DSP hot path (contract: see the task's "Measurement" section).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic tuner I/Q (1 s of signal by default).
N > 1 is launched by torchrun, one rank per GPU; tuner streams are independent so every rank processes its own
stream (weak scaling, no data-path collective); torch.distributed is used for the barrier and the
max-over-ranks time only.

Keys of the JSON line (rank 0): value = whole-job complex MS/s with inputs resident in HBM; e2e = the same
through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region); roofline = dominant kernel
against the measured HBM peak; cpu_baseline = the oracle (CPU restatement of the reference's Java path) timed
on this box's host cores on a bounded sample.
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

WORKLOADS = {
    # BASELINE.json configs[1]: polyphase channelizer, 10 MS/s -> 400 x 25 kHz channels
    "channelizer": dict(fs=1e7, taps_per_channel=9, seconds=1.0, demod=None,
                        desc="configs[1]: polyphase channelizer 10 MS/s -> 400 x 25 kHz channels (M=400, T=9), "
                             "all bins kept, gain M, [channel][time] output"),
    # BASELINE.json configs[2]: channelizer + per-channel FIR/AGC + C4FM DQPSK timing recovery on 400 channels
    "c4fm": dict(fs=1e7, taps_per_channel=9, seconds=1.0, demod="c4fm",
                 desc="configs[2]: channelizer 10 MS/s -> 400 channels + 72-tap FIR + block AGC + DQPSK "
                      "decision-directed timing recovery (P25 Phase 1 C4FM) on all 400 channels, dibits out"),
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
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def cpu_baseline(workload, threads, target_seconds=6.0):
    """Times the oracle (kind "port": CPU restatement of the reference Java path; no JVM here) on `threads` host
    threads, one independent tuner stream per thread (the oracle releases the GIL inside ctypes calls)."""
    import oracle
    cfg = WORKLOADS[workload]
    fs = cfg["fs"]
    m = int(fs / 25000) // 2 * 2
    taps = oracle.sinc_m2_channelizer(fs / m, m, cfg["taps_per_channel"])
    # bounded sample: 2 M input samples, or 4 whole 1024-sample assembler buffers per channel for the demod chain
    n_complex = 4 * 1024 * (m // 2) if cfg["demod"] else 2000000
    x = synth_input_numpy(n_complex, m, 1)
    fir = c4fm_fir_taps() if cfg["demod"] == "c4fm" else None

    def make_state():
        st = {"chan": oracle.Channelizer(taps, m)}
        if cfg["demod"] == "c4fm":
            st["procs"] = [oracle.OneChannelOutputProcessor(50000.0, k, float(m)) for k in range(m)]
            st["chains"] = [oracle.P25Chain(oracle.C4FM, 50000.0, fir) for _ in range(m)]
        return st

    def one_pass(st):
        res = st["chan"].receive(x)
        if cfg["demod"] == "c4fm":
            for k in range(m):
                y = st["procs"][k].process(res)
                st["chains"][k].receive(y[: y.size // 2048 * 2048])

    states = [make_state() for _ in range(threads)]
    one_pass(states[0])                      # warm-up
    t0 = time.perf_counter()
    one_pass(states[0])
    single = time.perf_counter() - t0
    reps = max(1, int(target_seconds / max(single, 1e-3)))
    done = [0] * threads

    def worker(i):
        for _ in range(reps):
            one_pass(states[i])
            done[i] += 1

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = sum(done) * n_complex
    return {"value": total / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "single_thread_value": n_complex / single / 1e6,
            "sample": "%d threads x %d passes over %d complex input samples each (%s); oracle = C restatement of "
                      "the reference Java path, JVM unavailable" % (threads, reps, n_complex, workload)}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg = WORKLOADS[args.workload]
    results = []
    for _ in range(args.warmup + args.steps):
        results.append(cpu_baseline(args.workload, cores, target_seconds=3.0))
    timed = results[args.warmup:]
    value = statistics.mean(r["value"] for r in timed)
    base = timed[-1]
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"]}, "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sdrtrunk_b200 import native
    from sdrtrunk_b200.dsp import ComplexPolyphaseChannelizerM2

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    native.init(local_rank)
    L = native.lib()
    cfg = WORKLOADS[args.workload]
    fs = cfg["fs"]
    m = ComplexPolyphaseChannelizerM2.getChannelCount(fs)
    n_complex = int(fs * cfg["seconds"]) // (1024 * (m // 2)) * (1024 * (m // 2)) if cfg["demod"] else int(fs * cfg["seconds"])
    n_floats = 2 * n_complex
    n_blocks = n_complex // (m // 2)
    dev = torch.device("cuda", local_rank)

    # synthetic tuner I/Q, generated on the device (tones on bin centres + AWGN), and its pinned host copy
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    t = torch.arange(n_complex, device=dev, dtype=torch.float64)
    z = 1e-3 * torch.randn(n_complex, 2, device=dev, generator=g, dtype=torch.float32)
    for k in (3, 57, 123, 200, 277, 391):
        f = (k if k < m // 2 else k - m) / m
        ph = 2 * np.pi * ((f * t) % 1.0)
        z[:, 0] += (0.05 * torch.cos(ph)).float()
        z[:, 1] += (0.05 * torch.sin(ph)).float()
    x_dev = z.reshape(-1).contiguous()
    del t, z
    x_host = torch.empty(n_floats, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x_dev)

    stream = torch.cuda.Stream(device=dev)
    chan = ComplexPolyphaseChannelizerM2(fs, cfg["taps_per_channel"], device=local_rank, maxInputFloats=n_floats)
    chan.setStream(stream.cuda_stream)
    out_dev = torch.empty((m, 2 * n_blocks), dtype=torch.float32, device=dev)
    out_host = torch.empty((m, 2 * n_blocks), dtype=torch.float32, pin_memory=True)

    pipeline = None
    if cfg["demod"]:
        from sdrtrunk_b200.dsp import P25Bank  # noqa: F401  (lands with the bank milestone)
        pipeline = P25Bank.pipeline(chan, "c4fm", m, 50000.0, c4fm_fir_taps(), n_blocks, stream.cuda_stream)

    def step_device():
        if pipeline is None:
            chan.receiveChannels((x_dev.data_ptr(), n_floats), native.DEVICE, out_dev.data_ptr(), native.DEVICE,
                                 2 * n_blocks)
        else:
            pipeline.process_device(x_dev.data_ptr(), n_floats)

    def step_host():
        if pipeline is None:
            native.check(L.sdrgpu_chan_process(chan._h, C.c_void_p(x_host.data_ptr()), n_floats, native.HOST,
                                               C.c_void_p(out_host.data_ptr()), 2 * n_blocks, native.HOST,
                                               native.LAYOUT_CHANNELS, None))
        else:
            pipeline.process_host(x_host.data_ptr(), n_floats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        stop.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = start.elapsed_time(stop)
        if world > 1:
            tt = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, wall = tt.tolist()
        return ms, wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.sdrgpu_launch_count()
    ms_dev, _ = timed(step_device, args.steps, args.warmup)
    launches = L.sdrgpu_launch_count() - launches0
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    # host-buffer path: device time of the stream also covers the copies; wall clock is what a caller sees
    ms_e2e_dev, wall_e2e = timed(step_host, args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None

    # dominant-kernel time, live, CUDA events on the launching stream (separate loop so the per-step event
    # synchronisation does not perturb the numbers above)
    kernel_ms = []
    if pipeline is None:
        chan.enableTiming(True)
        for _ in range(args.steps):
            step_device()
            kernel_ms.append(chan.lastKernelMs())
        chan.enableTiming(False)
    else:
        kernel_ms = pipeline.kernel_ms(step_device, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_dev / args.steps
    total_samples = n_complex * world
    value = total_samples / (ms_per_step * 1e-3) / 1e6
    e2e_ms = max(ms_e2e_dev, wall_e2e) / args.steps
    e2e_value = total_samples / (e2e_ms * 1e-3) / 1e6
    peak, peak_src = load_peaks()
    if pipeline is None:
        alg_bytes = 24.0 * n_complex            # 8 B read + 16 B written per input complex sample, all M bins kept
        k_ms = statistics.mean(kernel_ms)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "kernel": "pfb_ifft_kernel<16,9>", "kernel_ms": k_ms,
                    "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
        h2d, d2h = 4 * n_floats, out_host.numel() * 4
    else:
        roofline, h2d, d2h = pipeline.roofline(kernel_ms, n_complex, peak, peak_src)

    cores = os.cpu_count() or 1
    base = cpu_baseline(args.workload, cores) if not args.no_cpu_baseline else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "input_complex_samples_per_step_per_gpu": n_complex,
                   "channels": m, "channel_rate_hz": 2 * fs / m,
                   "l2": "per-step working set (%.0f MB in + %.0f MB out) exceeds the 126 MB L2" %
                         (4 * n_floats / 1e6, out_dev.numel() * 4 / 1e6),
                   "sharding": "one independent tuner stream per GPU, no collective"},
        "realtime_channels": m * world * (value / world) / (fs / 1e6),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms},
        "gpu_launches": launches_timed,
        "roofline": roofline,
        "cpu_baseline": base,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="channelizer", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
