"""Host-side mirror of the reference's io.github.dsheirer.dsp hot-path classes."""
from .filter_factory import FilterFactory, WindowType  # noqa: F401
from .channelizer import (ChannelCalculator, ComplexPolyphaseChannelizerM2, TunerChannel)  # noqa: F401
