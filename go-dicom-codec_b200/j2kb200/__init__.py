"""j2kb200 — host-side mirror of go-dicom-codec's JPEG 2000 sample-domain path over libj2kb200.so.

The shared library (CUDA, sm_100a) is the product; this package is the thin Python face used by
the tests, bench.py and non-Go hosts.  It never computes on the CPU: without the built library
`abi.load()` raises, without a CUDA device `Context()` raises.
"""
from . import abi, shard  # noqa: F401
from .codec import (Context, J2KError, decode_quant_steps, openjpeg_quant_params,  # noqa: F401
                    quality_quant_params, runtime_quant_steps)
