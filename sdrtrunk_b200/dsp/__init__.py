"""Host-side mirror of the reference's io.github.dsheirer.dsp hot-path classes."""
from .filter_factory import FilterFactory, FIRFilterSpecification, WindowType  # noqa: F401
from .channelizer import (AirspySampleConverter, ByteSampleConverter, ChannelCalculator, ComplexPolyphaseChannelizerM2,  # noqa: F401
                          Signed16BitSampleConverter, SignedByteSampleConverter, TunerChannel)
from .bank import (Bank, ComplexFeedForwardGainControl, ComplexFIRFilter2, CostasLoop,  # noqa: F401
                   DecimationFilterFactory, Dibit, DQPSKDecisionDirectedDemodulator, DQPSKGardnerDemodulator,
                   FMDemodulator, InterpolatingSampleBuffer, Pipeline, PLLBandwidth, RealFIRFilter2,
                   SquelchingFMDemodulator)
