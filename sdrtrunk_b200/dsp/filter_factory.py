"""FilterFactory / Window: the closed-form designers the hot path needs (host side, run once).

Mirrors J/dsp/filter/FilterFactory.java (getSincM2Channelizer :808-920, getSincM2Synthesizer :755-770,
getHalfBand :1007-1036, getTaps :671-681 with FIRFilterSpecification.lowPassBuilder) and Window.WindowType; computed by
libsdrgpu's host-side design code.
"""
import ctypes as C
import enum

import numpy as np

from .. import native


class WindowType(enum.Enum):
    HAMMING = native.WINDOW_HAMMING
    BLACKMAN = native.WINDOW_BLACKMAN


class FilterFactory:
    @staticmethod
    def getSincM2Channelizer(channelBandwidth, channels, tapsPerChannel, logResults=False):
        cap = channels * (tapsPerChannel + 11)
        out = np.zeros(cap, np.float32)
        n = C.c_int(0)
        native.check(native.lib().sdrgpu_design_sinc_m2_channelizer(
            float(channelBandwidth), int(channels), int(tapsPerChannel),
            out.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return out[:n.value].copy()

    @staticmethod
    def getSincM2Synthesizer(channelSampleRate, channelBandwidth, channels, tapsPerChannel):
        cap = channels * tapsPerChannel
        out = np.zeros(cap, np.float32)
        n = C.c_int(0)
        native.check(native.lib().sdrgpu_design_sinc_m2_synthesizer(
            float(channelSampleRate), float(channelBandwidth), int(channels), int(tapsPerChannel),
            out.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return out[:n.value].copy()

    @staticmethod
    def getHalfBand(length, windowType):
        out = np.zeros(max(int(length), 1), np.float32)
        native.check(native.lib().sdrgpu_design_half_band(int(length), windowType.value,
                                                          out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    @staticmethod
    def getTaps(specification):
        """FilterFactory.getTaps(FIRFilterSpecification): Remez exchange; None where the Java returns null"""
        sp = specification
        out = np.zeros(4096, np.float32)
        n = C.c_int(0)
        status = native.lib().sdrgpu_design_remez_low_pass(
            float(sp.sampleRate), float(sp.passBandCutoff), float(sp.stopBandStart), float(sp.passBandRipple),
            float(sp.stopBandRipple), int(sp.order), -1 if sp.oddLength is None else int(bool(sp.oddLength)),
            int(sp.gridDensity), out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(n))
        if status == native.ERR_DESIGN:
            return None
        native.check(status)
        return out[:n.value].copy()


class FIRFilterSpecification:
    """FIRFilterSpecification.lowPassBuilder() ... .build() (J/dsp/filter/fir/FIRFilterSpecification.java:298-428): the
    builder's setters, same names; build() returns the specification FilterFactory.getTaps takes."""

    def __init__(self):
        self.sampleRate = 0.0
        self.order = 0
        self.oddLength = None
        self.gridDensity = 16
        self.passBandCutoff = 0.0
        self.stopBandStart = 0.0
        self.passBandRipple = 0.0
        self.stopBandRipple = 0.0

    @staticmethod
    def lowPassBuilder():
        return _LowPassBuilder()

    @staticmethod
    def estimateFilterOrder(sampleRate, frequency1, frequency2, passBandRipple, stopBandRipple):
        return native.lib().sdrgpu_design_remez_estimate_order(float(sampleRate), float(frequency1), float(frequency2),
                                                               float(passBandRipple), float(stopBandRipple))


class _LowPassBuilder:
    def __init__(self):
        self._spec = FIRFilterSpecification()
        self._amplitudes = (1.0, 0.0)

    def _set(self, name, value):
        setattr(self._spec, name, value)
        return self

    def sampleRate(self, hz):
        return self._set("sampleRate", float(hz))

    def order(self, order):
        return self._set("order", int(order))

    def oddLength(self, odd):
        return self._set("oddLength", bool(odd))

    def gridDensity(self, density):
        return self._set("gridDensity", int(density))

    def passBandCutoff(self, hz):
        return self._set("passBandCutoff", float(hz))

    def stopBandStart(self, hz):
        return self._set("stopBandStart", float(hz))

    def passBandRipple(self, db):
        return self._set("passBandRipple", float(db))

    def stopBandRipple(self, db):
        return self._set("stopBandRipple", float(db))

    def passBandAmplitude(self, amplitude):
        if float(amplitude) != 1.0:
            raise native.IllegalArgumentException("only unity pass bands are supported (the decoders' low-pass filters)")
        return self

    def stopBandAmplitude(self, amplitude):
        if float(amplitude) != 0.0:
            raise native.IllegalArgumentException("only zero stop bands are supported (the decoders' low-pass filters)")
        return self

    def build(self):
        return self._spec
