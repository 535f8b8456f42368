"""j2kb200 — host-side mirror of go-dicom-codec's JPEG 2000 sample-domain path over libj2kb200.so."""
from . import abi  # noqa: F401
