"""ctypes binding of the CPU oracle (liboracle.so).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker / timed CPU baseline.
The product (sdrtrunk_b200) never imports it.  PARITY UNPINNED: see oracle/sdr_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "sdr_mmse_taps.h"))
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


class _Calc(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("channel_count", C.c_int),
                ("center_frequency", C.c_double), ("oversampling", C.c_double)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    sig = {
        "orc_window": (C.c_int, [C.c_int, C.c_int, _f64p]),
        "orc_kaiser": (C.c_int, [C.c_int, C.c_double, _f64p]),
        "orc_kaiser_sinc": (C.c_int, [C.c_int, C.c_double, C.c_double, _f32p]),
        "orc_evaluate": (C.c_double, [_f32p, C.c_int, C.c_double]),
        "orc_sinc_m2_channelizer": (C.c_int, [C.c_double, C.c_int, C.c_int, _f32p, C.c_int]),
        "orc_sinc_m2_synthesizer": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_int, _f32p]),
        "orc_half_band": (C.c_int, [C.c_int, C.c_int, _f32p]),
        "orc_remez_estimate_order": (C.c_int, [C.c_double] * 5),
        "orc_remez_low_pass": (C.c_int, [C.c_double] * 5 + [C.c_int, C.c_int, C.c_int, _f32p, C.c_int]),
        "orc_fft_create": (vp, [C.c_int]),
        "orc_fft_destroy": (None, [vp]),
        "orc_ifft_f32": (None, [vp, _f32p]),
        "orc_idft_f64": (None, [C.c_int, _f32p, _f32p]),
        "orc_chan_create": (vp, [_f32p, C.c_int, C.c_int]),
        "orc_chan_destroy": (None, [vp]),
        "orc_chan_receive": (C.c_int, [vp, _f32p, C.c_int, _f32p, C.c_int]),
        "orc_chan_receive_raw": (C.c_int, [vp, _f32p, C.c_int, _f32p]),
        "orc_calc_channel_indexes": (C.c_int, [C.POINTER(_Calc), C.c_longlong, C.c_int, _i32p, C.c_int]),
        "orc_calc_center_frequency_for_indexes": (C.c_longlong, [C.POINTER(_Calc), _i32p, C.c_int]),
        "orc_convert_u8": (None, [C.c_void_p, C.c_int, _f32p]),
        "orc_convert_s8": (None, [C.c_void_p, C.c_int, _f32p]),
        "orc_convert_s16le": (None, [C.c_void_p, C.c_int, _f32p]),
        "orc_get_channel": (None, [_f32p, C.c_int, C.c_int, C.c_int, _f32p]),
        "orc_apply_gain": (None, [_f32p, C.c_int, C.c_double]),
        "orc_osc_init": (None, [vp, C.c_double, C.c_double]),
        "orc_osc_mix": (None, [vp, _f32p, C.c_int]),
        "orc_one_channel_create": (vp, [C.c_double, C.c_int, C.c_double]),
        "orc_one_channel_destroy": (None, [vp]),
        "orc_one_channel_set_frequency_offset": (None, [vp, C.c_longlong]),
        "orc_one_channel_process": (None, [vp, _f32p, C.c_int, C.c_int, _f32p]),
        "orc_two_channel_create": (vp, [C.c_double, C.c_int, C.c_int, _f32p, C.c_int, C.c_double]),
        "orc_two_channel_destroy": (None, [vp]),
        "orc_two_channel_set_frequency_offset": (None, [vp, C.c_longlong]),
        "orc_two_channel_process": (None, [vp, _f32p, C.c_int, C.c_int, _f32p]),
        "orc_halfband_create": (vp, [_f32p, C.c_int]),
        "orc_halfband_destroy": (None, [vp]),
        "orc_halfband_decimate_complex": (C.c_int, [vp, _f32p, C.c_int, _f32p]),
        "orc_halfband_decimate_real": (C.c_int, [vp, _f32p, C.c_int, _f32p]),
        "orc_decimator_create": (vp, [C.c_int]),
        "orc_decimator_destroy": (None, [vp]),
        "orc_decimator_complex": (C.c_int, [vp, _f32p, C.c_int, _f32p]),
        "orc_decimator_real": (C.c_int, [vp, _f32p, C.c_int, _f32p]),
        "orc_fir_create": (vp, [_f32p, C.c_int, C.c_float]),
        "orc_fir_destroy": (None, [vp]),
        "orc_fir_filter_real": (None, [vp, _f32p, C.c_int, _f32p]),
        "orc_cfir_create": (vp, [_f32p, C.c_int, C.c_float]),
        "orc_cfir_destroy": (None, [vp]),
        "orc_cfir_filter": (None, [vp, _f32p, C.c_int, _f32p]),
        "orc_agc_block": (None, [_f32p, C.c_int, _f32p]),
        "orc_psk_create": (vp, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_float]),
        "orc_psk_destroy": (None, [vp]),
        "orc_psk_receive": (C.c_int, [vp, _f32p, C.c_int, _u8p, C.POINTER(C.c_double)]),
        "orc_psk_correct_inversion": (None, [vp, C.c_double]),
        "orc_airspy_create": (vp, []),
        "orc_airspy_destroy": (None, [vp]),
        "orc_airspy_convert": (C.c_int, [vp, _u8p, C.c_int, C.c_int, _f32p]),
        "orc_sync_create": (vp, [C.c_int, C.c_double]),
        "orc_sync_destroy": (None, [vp]),
        "orc_sync_delay": (C.c_int, [vp]),
        "orc_sync_receive": (C.c_int, [vp, C.c_int, _f64p]),
        "orc_psk_attach_sync": (None, [vp, vp]),
        "orc_p2_framer_create": (vp, [C.c_double]),
        "orc_p2_framer_destroy": (None, [vp]),
        "orc_p2_framer_receive": (C.c_int, [vp, C.c_int, _f64p]),
        "orc_p25_chain_attach_sync": (C.c_int, [vp, C.c_int, C.c_double]),
        "orc_psk_reset_pll": (None, [vp]),
        "orc_psk_get_state": (None, [vp, _f64p, _f64p, _f32p, _f32p]),
        "orc_pack_dibits": (C.c_int, [_u8p, C.c_int, _u8p]),
        "orc_p25_chain_create": (vp, [C.c_int, C.c_double, _f32p, C.c_int]),
        "orc_p25_chain_destroy": (None, [vp]),
        "orc_p25_chain_receive": (C.c_int, [vp, _f32p, C.c_int, _u8p, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


# ---------------------------------------------------------------- designers
def sinc_m2_channelizer(channel_bandwidth, channels, taps_per_channel):
    cap = channels * (taps_per_channel + 11)
    out = np.zeros(cap, np.float32)
    n = lib().orc_sinc_m2_channelizer(channel_bandwidth, channels, taps_per_channel, out.ctypes.data_as(_f32p), cap)
    if n < 0:
        raise ValueError("FilterDesignException (%d)" % n)
    return out[:n].copy()


def sinc_m2_synthesizer(channel_sample_rate, channel_bandwidth, channels, taps_per_channel):
    out = np.zeros(channels * taps_per_channel, np.float32)
    n = lib().orc_sinc_m2_synthesizer(channel_sample_rate, channel_bandwidth, channels, taps_per_channel,
                                      out.ctypes.data_as(_f32p))
    if n < 0:
        raise ValueError("FilterDesignException")
    return out[:n].copy()


def half_band(length, window):
    out = np.zeros(length, np.float32)
    n = lib().orc_half_band(length, {"hamming": 0, "blackman": 1}[window], out.ctypes.data_as(_f32p))
    if n < 0:
        raise ValueError("bad half-band length")
    return out


def remez_estimate_order(sample_rate, f1, f2, pass_ripple_db, stop_ripple_db):
    """FIRFilterSpecification.estimateFilterOrder"""
    return lib().orc_remez_estimate_order(sample_rate, f1, f2, pass_ripple_db, stop_ripple_db)


def remez_low_pass(sample_rate, pass_end, stop_start, pass_ripple_db, stop_ripple_db, order=0, odd_length=None,
                   grid_density=16):
    """FIRFilterSpecification.lowPassBuilder()...build() + FilterFactory.getTaps; None where getTaps returns null"""
    out = np.zeros(4096, np.float32)
    n = lib().orc_remez_low_pass(sample_rate, pass_end, stop_start, pass_ripple_db, stop_ripple_db, order,
                                 -1 if odd_length is None else int(bool(odd_length)), grid_density,
                                 out.ctypes.data_as(_f32p), out.size)
    if n == -2:
        raise ValueError("filter longer than 4096 taps")
    return out[:n].copy() if n > 0 else None


# the decoders' baseband filters, as the reference designs them
def c4fm_baseband_taps():
    """P25P1DecoderC4FM.getBasebandFilter (P25P1DecoderC4FM.java:136-148) at the 50 kHz channel rate"""
    return remez_low_pass(50000.0, 5100, 6500, 0.01, 0.01)


def hdqpsk_baseband_taps():
    """P25P2DecoderHDQPSK.getBasebandFilter (P25P2DecoderHDQPSK.java:155-166)"""
    return remez_low_pass(50000.0, 6500, 7200, 0.005, 0.01)


def nbfm_iq_taps(decimated_sample_rate=25000.0, channel_bandwidth=12500.0):
    """NBFMDecoder I/Q filter (NBFMDecoder.java:306-325): note sampleRate(decimatedSampleRate * 2)"""
    return remez_low_pass(decimated_sample_rate * 2, int(channel_bandwidth * .8), int(channel_bandwidth), 0.01, 0.005,
                          odd_length=True, grid_density=16)


def window(kind, length):
    out = np.zeros(length, np.float64)
    lib().orc_window({"hamming": 0, "blackman": 1}[kind], length, out.ctypes.data_as(_f64p))
    return out


def kaiser(length, attenuation):
    out = np.zeros(length, np.float64)
    lib().orc_kaiser(length, attenuation, out.ctypes.data_as(_f64p))
    return out


def evaluate(taps, frequency):
    a, p = _f32(taps)
    return lib().orc_evaluate(p, a.size, frequency)


# ---------------------------------------------------------------- tuner sample converters
def convert_samples(raw, fmt):
    """fmt: 'u8' (ByteSampleConverter), 's8' (SignedByteSampleConverter), 's16le' (ConversionUtils); raw: bytes-like"""
    a = np.ascontiguousarray(np.frombuffer(bytes(raw), dtype=np.uint8))
    n = a.size // (2 if fmt == "s16le" else 1)
    out = np.zeros(n, np.float32)
    fn = {"u8": lib().orc_convert_u8, "s8": lib().orc_convert_s8, "s16le": lib().orc_convert_s16le}[fmt]
    fn(a.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(_f32p))
    return out


# ---------------------------------------------------------------- FFT
def ifft_f32(x):
    """x: interleaved float32 [2n] -> scaled inverse FFT (float32 mixed radix)."""
    a = np.array(x, dtype=np.float32, copy=True)
    n = a.size // 2
    f = lib().orc_fft_create(n)
    lib().orc_ifft_f32(f, a.ctypes.data_as(_f32p))
    lib().orc_fft_destroy(f)
    return a


def idft_f64(x):
    a, p = _f32(x)
    out = np.zeros_like(a)
    lib().orc_idft_f64(a.size // 2, p, out.ctypes.data_as(_f32p))
    return out


# ---------------------------------------------------------------- channelizer
class Channelizer:
    """ComplexPolyphaseChannelizerM2 restated (filter bank + IFFT); results as [n_blocks, 2M] float32."""

    def __init__(self, taps, channel_count):
        a, p = _f32(taps)
        self.m = channel_count
        self._h = lib().orc_chan_create(p, a.size, channel_count)
        if not self._h:
            raise ValueError("Channel count must be an even multiple of the over-sample rate (2x)")
        self._pending = 0

    def receive(self, samples, mode="f32"):
        a, p = _f32(samples)
        max_blocks = (self._pending + a.size) // self.m + 1
        out = np.zeros((max_blocks, 2 * self.m), np.float32)
        if mode == "raw":
            n = lib().orc_chan_receive_raw(self._h, p, a.size, out.ctypes.data_as(_f32p))
        else:
            n = lib().orc_chan_receive(self._h, p, a.size, out.ctypes.data_as(_f32p), 1 if mode == "f64" else 0)
        self._pending = (self._pending + a.size) % self.m
        return out[:n]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_chan_destroy(self._h)
            self._h = None


class ChannelCalculator:
    def __init__(self, sample_rate, channel_count, center_frequency, oversampling=2.0):
        self.c = _Calc(sample_rate, channel_count, center_frequency, oversampling)

    def channel_indexes(self, frequency, bandwidth):
        idx = (C.c_int * 64)()
        n = lib().orc_calc_channel_indexes(C.byref(self.c), int(frequency), int(bandwidth), idx, 64)
        if n < 0:
            raise ValueError("IllegalArgumentException (%d)" % n)
        return list(idx[:n])

    def center_frequency_for_indexes(self, indexes):
        arr = (C.c_int * len(indexes))(*indexes)
        return lib().orc_calc_center_frequency_for_indexes(C.byref(self.c), arr, len(indexes))


class OneChannelOutputProcessor:
    def __init__(self, sample_rate, bin_index, gain):
        self._h = lib().orc_one_channel_create(sample_rate, bin_index, gain)

    def set_frequency_offset(self, offset):
        lib().orc_one_channel_set_frequency_offset(self._h, int(offset))

    def process(self, results):
        r, p = _f32(results)
        n_blocks, two_m = r.shape
        out = np.zeros(2 * n_blocks, np.float32)
        lib().orc_one_channel_process(self._h, p, n_blocks, two_m // 2, out.ctypes.data_as(_f32p))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_one_channel_destroy(self._h)
            self._h = None


class TwoChannelOutputProcessor:
    def __init__(self, sample_rate, bin1, bin2, synthesis_filter, gain):
        f, p = _f32(synthesis_filter)
        self._h = lib().orc_two_channel_create(sample_rate, bin1, bin2, p, f.size, gain)

    def set_frequency_offset(self, offset):
        lib().orc_two_channel_set_frequency_offset(self._h, int(offset))

    def process(self, results):
        r, p = _f32(results)
        n_blocks, two_m = r.shape
        out = np.zeros(2 * n_blocks, np.float32)
        lib().orc_two_channel_process(self._h, p, n_blocks, two_m // 2, out.ctypes.data_as(_f32p))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_two_channel_destroy(self._h)
            self._h = None


# ---------------------------------------------------------------- filters
class HalfBand:
    def __init__(self, coefficients):
        a, p = _f32(coefficients)
        self._h = lib().orc_halfband_create(p, a.size)
        if not self._h:
            raise ValueError("Half-band filter coefficients must be odd-length L = 4x - 1")

    def decimate_complex(self, samples):
        a, p = _f32(samples)
        out = np.zeros(a.size // 2, np.float32)
        if lib().orc_halfband_decimate_complex(self._h, p, a.size, out.ctypes.data_as(_f32p)) < 0:
            raise ValueError("Samples array length must be an integer multiple of 4")
        return out

    def decimate_real(self, samples):
        a, p = _f32(samples)
        out = np.zeros(a.size // 2, np.float32)
        if lib().orc_halfband_decimate_real(self._h, p, a.size, out.ctypes.data_as(_f32p)) < 0:
            raise ValueError("Samples array length must be an integer multiple of 2")
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_halfband_destroy(self._h)
            self._h = None


class Decimator:
    def __init__(self, rate):
        self.rate = rate
        self._h = lib().orc_decimator_create(rate)
        if not self._h:
            raise ValueError("Unsupported decimation rate: %d" % rate)

    def _run(self, fn, samples):
        a, p = _f32(samples)
        out = np.zeros(a.size if self.rate == 0 else a.size // self.rate, np.float32)
        n = fn(self._h, p, a.size, out.ctypes.data_as(_f32p))
        if n < 0:
            raise ValueError("Sample buffer length must be an integer multiple of the decimation")
        return out[:n]

    def decimate_complex(self, samples):
        return self._run(lib().orc_decimator_complex, samples)

    def decimate_real(self, samples):
        return self._run(lib().orc_decimator_real, samples)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_decimator_destroy(self._h)
            self._h = None


class RealFIR:
    def __init__(self, taps, gain=1.0):
        a, p = _f32(taps)
        self._h = lib().orc_fir_create(p, a.size, gain)

    def filter(self, samples):
        a, p = _f32(samples)
        out = np.zeros_like(a)
        lib().orc_fir_filter_real(self._h, p, a.size, out.ctypes.data_as(_f32p))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fir_destroy(self._h)
            self._h = None


class ComplexFIR:
    def __init__(self, taps, gain=1.0):
        a, p = _f32(taps)
        self._h = lib().orc_cfir_create(p, a.size, gain)

    def filter(self, samples):
        a, p = _f32(samples)
        out = np.zeros_like(a)
        lib().orc_cfir_filter(self._h, p, a.size, out.ctypes.data_as(_f32p))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_cfir_destroy(self._h)
            self._h = None


class _FM(C.Structure):
    _fields_ = [("prev_i", C.c_float), ("prev_q", C.c_float), ("gain", C.c_float)]


class _Squelch(C.Structure):
    _fields_ = [("alpha", C.c_double), ("one_minus_alpha", C.c_double), ("output", C.c_double),
                ("power", C.c_double), ("threshold", C.c_double), ("state", C.c_int),
                ("ramp_threshold", C.c_int), ("ramp_count", C.c_int), ("squelch_changed", C.c_int)]


class _SqFM(C.Structure):
    _fields_ = [("fm", _FM), ("sq", _Squelch), ("squelch_changed", C.c_int)]


class FMDemodulator:
    def __init__(self, gain=1.0):
        self.s = _FM()
        L = lib()
        L.orc_fm_init.argtypes = [C.POINTER(_FM), C.c_float]
        L.orc_fm_init.restype = None
        L.orc_fm_demodulate_buffer.argtypes = [C.POINTER(_FM), _f32p, C.c_int, _f32p]
        L.orc_fm_demodulate_buffer.restype = None
        L.orc_fm_init(C.byref(self.s), gain)

    def demodulate(self, iq):
        a, p = _f32(iq)
        out = np.zeros(a.size // 2, np.float32)
        lib().orc_fm_demodulate_buffer(C.byref(self.s), p, a.size, out.ctypes.data_as(_f32p))
        return out


class SquelchingFMDemodulator:
    def __init__(self, alpha=0.0004, threshold_db=-78.0, ramp=4):
        self.s = _SqFM()
        L = lib()
        L.orc_sqfm_init.argtypes = [C.POINTER(_SqFM), C.c_double, C.c_double, C.c_int]
        L.orc_sqfm_init.restype = None
        L.orc_sqfm_demodulate_buffer.argtypes = [C.POINTER(_SqFM), _f32p, C.c_int, _f32p]
        L.orc_sqfm_demodulate_buffer.restype = None
        L.orc_sqfm_init(C.byref(self.s), alpha, threshold_db, ramp)

    def demodulate(self, iq):
        a, p = _f32(iq)
        out = np.zeros(a.size // 2, np.float32)
        lib().orc_sqfm_demodulate_buffer(C.byref(self.s), p, a.size, out.ctypes.data_as(_f32p))
        return out


class _Osc(C.Structure):
    _fields_ = [("angle_i", C.c_float), ("angle_q", C.c_float), ("cur_i", C.c_float), ("cur_q", C.c_float)]


class Oscillator:
    """J/dsp/mixer/Oscillator.java: mixComplex(samples) = sample * current, then rotate() + fastNormalize."""

    def __init__(self, frequency, sample_rate):
        self.o = _Osc()
        lib().orc_osc_init(C.byref(self.o), float(frequency), float(sample_rate))

    def mix(self, iq):
        a = np.array(iq, np.float32, copy=True)
        lib().orc_osc_mix(C.byref(self.o), a.ctypes.data_as(_f32p), a.size)
        return a


def apply_gain(iq, gain):
    a = np.array(iq, np.float32, copy=True)
    lib().orc_apply_gain(a.ctypes.data_as(_f32p), a.size, float(gain))
    return a


def agc_block(iq):
    a, p = _f32(iq)
    out = np.zeros_like(a)
    lib().orc_agc_block(p, a.size, out.ctypes.data_as(_f32p))
    return out


# ---------------------------------------------------------------- PSK
DECISION_DIRECTED, GARDNER = 0, 1


class PSKDemodulator:
    def __init__(self, kind, sample_rate, symbol_rate, pll_bandwidth, sample_counter_gain):
        self._h = lib().orc_psk_create(kind, sample_rate, symbol_rate, pll_bandwidth, sample_counter_gain)

    def receive(self, iq, want_taps=False):
        a, p = _f32(iq)
        cap = a.size // 2 // 4 + 8
        dibits = np.zeros(cap, np.uint8)
        taps = np.zeros((cap, 6), np.float64) if want_taps else None
        n = lib().orc_psk_receive(self._h, p, a.size, dibits.ctypes.data_as(_u8p),
                                  taps.ctypes.data_as(C.POINTER(C.c_double)) if want_taps else None)
        return (dibits[:n], taps[:n]) if want_taps else dibits[:n]

    def correct_inversion(self, correction):
        lib().orc_psk_correct_inversion(self._h, correction)

    def attach_sync(self, sync):
        """dibits become dibit | event << 2 and inversion corrections are applied (orc_psk_attach_sync)"""
        self._sync = sync
        lib().orc_psk_attach_sync(self._h, sync._h if sync is not None else None)

    def state(self):
        ph, fr, sp, ds = C.c_double(), C.c_double(), C.c_float(), C.c_float()
        lib().orc_psk_get_state(self._h, C.byref(ph), C.byref(fr), C.byref(sp), C.byref(ds))
        return ph.value, fr.value, sp.value, ds.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_psk_destroy(self._h)
            self._h = None


class AirspySampleConverter:
    """AirspySampleConverter: raw 12-bit real samples -> DC removal -> Hilbert transform -> interleaved I/Q."""

    def __init__(self):
        self._h = lib().orc_airspy_create()
        self.packed = False

    def setSamplePacking(self, enabled):
        self.packed = bool(enabled)

    def convert(self, raw):
        b = np.ascontiguousarray(raw, np.uint8)
        n = b.size // 3 * 2 if self.packed else b.size // 2
        out = np.zeros(max(n, 1), np.float32)
        got = lib().orc_airspy_convert(self._h, b.ctypes.data_as(_u8p), b.size, 1 if self.packed else 0, out.ctypes.data_as(_f32p))
        return out[:got]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_airspy_destroy(self._h)
            self._h = None


SYNC_P25_PHASE1, SYNC_P25_PHASE2, SYNC_P25_PHASE2_FRAMED = 1, 2, 3
P2_EVENT_FRAGMENT, P2_EVENT_SYNC_LOSS, P2_EVENT_INVERSION, P2_EVENT_SYNCHRONIZED = 1, 2, 4, 32
SYNC_EVENT_NONE, SYNC_EVENT_SYNC, SYNC_EVENT_90_CW, SYNC_EVENT_90_CCW, SYNC_EVENT_180, SYNC_EVENT_LOST = range(6)


class SyncDetector:
    """P25P1SyncDetector / P25P2SyncDetector behind the framer's dibit delay buffer, fed every dibit."""

    def __init__(self, kind, sample_rate):
        self._h = lib().orc_sync_create(kind, sample_rate)
        if not self._h:
            raise ValueError("unknown sync detector kind")

    @property
    def delay(self):
        return lib().orc_sync_delay(self._h)

    def receive(self, dibit):
        corr = C.c_double()
        ev = lib().orc_sync_receive(self._h, int(dibit), C.byref(corr))
        return ev, corr.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sync_destroy(self._h)
            self._h = None


class P2SuperFrameDetector:
    """P25P2SuperFrameDetector with its P25P2SyncDetector: receive(dibit) -> (event bits, PLL correction)."""

    def __init__(self, sample_rate):
        self._h = lib().orc_p2_framer_create(sample_rate)

    def receive(self, dibit):
        corr = C.c_double()
        ev = lib().orc_p2_framer_receive(self._h, int(dibit), C.byref(corr))
        return ev, corr.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_p2_framer_destroy(self._h)
            self._h = None


def pack_dibits(dibits):
    d = np.ascontiguousarray(dibits, np.uint8)
    out = np.zeros(d.size // 4 + 1, np.uint8)
    n = lib().orc_pack_dibits(d.ctypes.data_as(_u8p), d.size, out.ctypes.data_as(_u8p))
    return out[:n]


C4FM, LSM, HDQPSK, DMR = 0, 1, 2, 3


class P25Chain:
    """filter -> block AGC -> DQPSK demodulator of the P25 decoder front-ends, per 1024-sample buffer."""

    def __init__(self, kind, sample_rate, fir_taps=None):
        if fir_taps is not None:
            a, p = _f32(fir_taps)
            n = a.size
        else:
            p, n = None, 0
        self._h = lib().orc_p25_chain_create(kind, sample_rate, p, n)

    def attach_sync(self, sync_kind, sample_rate):
        if lib().orc_p25_chain_attach_sync(self._h, sync_kind, sample_rate) != 0:
            raise ValueError("unknown sync detector kind")

    def receive(self, iq, want_agc=False):
        a, p = _f32(iq)
        dibits = np.zeros(a.size // 2 // 4 + 8, np.uint8)
        agc = np.zeros_like(a) if want_agc else None
        n = lib().orc_p25_chain_receive(self._h, p, a.size, dibits.ctypes.data_as(_u8p),
                                        agc.ctypes.data_as(_f32p) if want_agc else None)
        if n < 0:
            raise ValueError("length must be a multiple of 2048 floats")
        return (dibits[:n], agc) if want_agc else dibits[:n]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_p25_chain_destroy(self._h)
            self._h = None
