"""sdrtrunk_b200 -- B200-native (sm_100a) implementation of sdrtrunk's data-parallel DSP hot path.

The compute lives in libsdrgpu.so (sdrtrunk_b200/csrc, C ABI in include/sdrgpu.h).  This package is the
host-side mirror of the reference's Java interfaces for that path (same class and method names, argument
meaning and error behaviour), written in Python because no JDK exists in the build image; the Java FFM
binding a maintainer would add is in INTEGRATION.md.
"""
from . import native  # noqa: F401
from .native import (CudaError, FilterDesignException, IllegalArgumentException,  # noqa: F401
                     IllegalStateException)

__all__ = ["native", "FilterDesignException", "IllegalArgumentException", "IllegalStateException", "CudaError"]
