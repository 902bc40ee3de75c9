"""FilterFactory / Window: the closed-form designers the hot path needs (host side, run once).

Mirrors J/dsp/filter/FilterFactory.java (getSincM2Channelizer :808-920, getSincM2Synthesizer :755-770,
getHalfBand :1007-1036) and Window.WindowType; computed by libsdrgpu's host-side design code.
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
