"""Seeded synthetic I/Q generators for the parity tests and the bench (numpy only).

Signals follow SURVEY.md section 8(d): complex tones, NBFM carriers, C4FM (4FSK, deviations
+/-600/+/-1800 Hz at 4800 sym/s), pi/4-DQPSK bursts (LSM 4800 / HDQPSK 6000 sym/s), AWGN, and a
frequency-domain multiplexer that places per-channel 50 kHz basebands on the channelizer's bin centres.
"""
import numpy as np

# Dibit.getValue(): 0 -> +1 (+45 deg), 1 -> +3 (+135), 2 -> -1 (-45), 3 -> -3 (-135)
DIBIT_TO_LEVEL = np.array([1.0, 3.0, -1.0, -3.0])


def interleave(z):
    z = np.asarray(z)
    out = np.empty(2 * z.size, np.float32)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


def deinterleave(x):
    x = np.asarray(x)
    return x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64)


def awgn(rng, n, sigma):
    return sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def tone(fs, freq, n, amplitude=1.0, phase=0.0):
    t = np.arange(n)
    return amplitude * np.exp(1j * (2 * np.pi * freq * t / fs + phase))


def nbfm(fs, n, audio_hz=1000.0, deviation=2500.0, amplitude=0.5, carrier_offset=0.0):
    t = np.arange(n) / fs
    phase = (deviation / audio_hz) * np.sin(2 * np.pi * audio_hz * t) + 2 * np.pi * carrier_offset * t
    return amplitude * np.exp(1j * phase)


def c4fm(dibits, fs=50000.0, symbol_rate=4800.0, carrier_offset=0.0, timing_phase=0.0, amplitude=1.0,
         n_samples=None, phase0=0.0):
    """Continuous-phase 4-level FM: each dibit contributes a raised-cosine (Hann) frequency pulse one
    symbol period wide whose area gives a phase step of level * pi/4 across that symbol."""
    dibits = np.asarray(dibits)
    sps = fs / symbol_rate
    if n_samples is None:
        n_samples = int(np.floor(dibits.size * sps))
    t = np.arange(n_samples) / sps - timing_phase  # in symbol units
    freq = np.zeros(n_samples)
    levels = DIBIT_TO_LEVEL[dibits]
    # pulse g(u) = 1 + cos(2*pi*u) for |u| < 1/2 (area 1 in symbol units), centred on symbol k + 1/2
    k = np.floor(t).astype(int)
    valid = (k >= 0) & (k < dibits.size)
    u = t - k - 0.5
    g = 1.0 + np.cos(2 * np.pi * u)
    freq[valid] = levels[k[valid]] * g[valid]
    phase = np.cumsum(freq) * (np.pi / 4.0) / sps
    phase += 2 * np.pi * carrier_offset * np.arange(n_samples) / fs + phase0
    return amplitude * np.exp(1j * phase)


def rrc_taps(sps, span, beta):
    n = np.arange(-span * sps, span * sps + 1) / sps
    h = np.zeros_like(n)
    for i, t in enumerate(n):
        if abs(t) < 1e-9:
            h[i] = 1.0 + beta * (4 / np.pi - 1)
        elif abs(abs(t) - 1 / (4 * beta)) < 1e-9:
            h[i] = (beta / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * beta)) +
                                          (1 - 2 / np.pi) * np.cos(np.pi / (4 * beta)))
        else:
            h[i] = (np.sin(np.pi * t * (1 - beta)) + 4 * beta * t * np.cos(np.pi * t * (1 + beta))) / \
                   (np.pi * t * (1 - (4 * beta * t) ** 2))
    return h / np.sum(h)


def dqpsk(dibits, fs=50000.0, symbol_rate=4800.0, carrier_offset=0.0, timing_phase=0.0, amplitude=1.0,
          beta=0.35, n_samples=None, phase0=0.0):
    """pi/4-DQPSK with raised-cosine-ish shaping: phase steps level*pi/4 on an impulse train that is
    low-pass interpolated in the frequency domain (arbitrary, non-integer samples per symbol)."""
    dibits = np.asarray(dibits)
    sps = fs / symbol_rate
    if n_samples is None:
        n_samples = int(np.floor(dibits.size * sps))
    sym_phase = np.cumsum(DIBIT_TO_LEVEL[dibits] * np.pi / 4.0)
    symbols = np.exp(1j * sym_phase)
    # evaluate sum_k s_k * p(t - k) with p = raised cosine, by direct windowed summation (span +-6)
    t = np.arange(n_samples) / sps - timing_phase
    out = np.zeros(n_samples, complex)
    k0 = np.floor(t).astype(int)
    for dk in range(-6, 8):
        k = k0 + dk
        valid = (k >= 0) & (k < dibits.size)
        u = t - k
        with np.errstate(divide="ignore", invalid="ignore"):
            den = 1.0 - (2 * beta * u) ** 2
            p = np.sinc(u) * np.cos(np.pi * beta * u) / den
        sing = np.abs(den) < 1e-9
        p[sing] = (np.pi / 4) * np.sinc(1 / (2 * beta))
        out[valid] += symbols[k[valid]] * p[valid]
    out *= np.exp(1j * (2 * np.pi * carrier_offset * np.arange(n_samples) / fs + phase0))
    return amplitude * out


def multiplex(channel_basebands, bins, m, n_channel_samples):
    """Places each 2x-oversampled (fs_ch = 2*fs/M) channel baseband on channelizer bin `bins[i]` of an
    M-channel wideband stream by frequency-domain zero stuffing.  Only the central half of each channel
    spectrum (+/- one channel bandwidth / 2) is kept.  Result is periodic with n_channel_samples*M/2."""
    n_ch = n_channel_samples
    assert n_ch % 4 == 0
    n_w = n_ch * m // 2
    w = np.zeros(n_w, complex)
    quarter = n_ch // 4
    j = np.concatenate([np.arange(0, quarter), np.arange(-quarter, 0)])
    for x, k in zip(channel_basebands, bins):
        xf = np.fft.fft(np.asarray(x)[:n_ch])
        centre = (k if k < m // 2 else k - m) * (n_ch // 2)
        w[(centre + j) % n_w] += xf[j % n_ch]
    return np.fft.ifft(w) * (n_w / n_ch)


def rel_rms(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.sum((a - b) ** 2) / max(np.sum(b ** 2), 1e-300)))


# FrameSync.java:27-35 (dibits MSB first)
P25_PHASE1_SYNC = 0x5575F5FF77FF
P25_PHASE2_SYNC = 0x575D57F7FF


def sync_dibits(pattern, n_bits):
    return np.array([(pattern >> (n_bits - 2 - 2 * k)) & 3 for k in range(n_bits // 2)], np.uint8)


def dibits_with_sync(rng, n_symbols, pattern, n_bits, first=40, period=180):
    """random dibits with the sync pattern written every `period` symbols"""
    d = rng.integers(0, 4, n_symbols).astype(np.uint8)
    s = sync_dibits(pattern, n_bits)
    for k in range(first, n_symbols - s.size, period):
        d[k:k + s.size] = s
    return d


def airspy_raw(x, packed=False):
    """real samples in [-1, 1) -> the Airspy's native buffer bytes: unsigned 12-bit, two bytes per sample little
    endian, or "sample packing" (two samples in three bytes, AirspySampleConverter.convertPacked)"""
    v = np.clip(np.round(np.asarray(x) * 2048.0) + 2048, 0, 4095).astype(np.uint32)
    if not packed:
        return v.astype("<u2").view(np.uint8)
    a, b = v[0::2], v[1::2]
    out = np.zeros(a.size * 3, np.uint8)
    out[0::3] = a >> 4
    out[1::3] = ((a & 0xF) << 4) | (b >> 8)
    out[2::3] = b & 0xFF
    return out


def airspy_real_signal(rng, n, tones, fs=20e6, dc=0.01, noise=1e-3):
    """real ADC stream: the complex band (fs/2 wide) sits around fs/4; tones = [(offset_hz, amplitude)]"""
    t = np.arange(n)
    x = dc + noise * rng.standard_normal(n)
    for f, a in tones:
        x = x + a * np.cos(2 * np.pi * (fs / 4 + f) * t / fs + rng.uniform(0, 2 * np.pi))
    return x
