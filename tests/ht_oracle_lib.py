"""ctypes wrapper around oracle/build/libht_oracle.so (the CPU checker of the HTJ2K cleanup-pass block decoder).

Test infrastructure: imported only from tests/, __graft_entry__.smoke() and bench.py's cpu legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "build", "libht_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("ht_oracle.c", "ht_vlc_src.inc")]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in srcs)
    if force or stale:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra", "-shared", "-o", LIB, srcs[0]])
    return LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HtOracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        self.lib.orc_ht_decode_block.restype = C.c_int

    def table(self, which):
        out = np.zeros(1024, np.uint16)
        n = self.lib.orc_ht_table(C.c_int(which), _p(out))
        return out[:n].copy()

    def decode_block(self, data: bytes, width, height, kmax, missing_msbs):
        buf = np.frombuffer(bytes(data) + b"\0", np.uint8)
        out = np.empty((height, width), np.int32)
        rc = self.lib.orc_ht_decode_block(_p(buf), C.c_int(len(data)), C.c_int(width), C.c_int(height), C.c_int(kmax),
                                          C.c_int(missing_msbs), _p(out))
        return rc, out

    def encode_block(self, x, missing_msbs):
        """the test-stream GENERATOR (see ht_oracle.c): bytes of a cleanup segment that decodes to x with kmax = missing_msbs + 1;
        b"" for an all-zero block"""
        a = np.ascontiguousarray(x, np.int32)
        h, w = a.shape
        out = np.zeros(w * h * 5 + 64, np.uint8)
        n = self.lib.orc_ht_encode_block(_p(a), C.c_int(w), C.c_int(h), C.c_int(missing_msbs), _p(out), C.c_int(out.size))
        assert n >= 0, n
        return out[:n].tobytes()

    def encode_ref(self, x, kmax):
        """oracle restatement of HTEncoder.Encode (htj2k/encoder.go:54-68): bytes (b"" for an empty block) or a negative code"""
        a = np.ascontiguousarray(x, np.int32)
        h, w = a.shape
        out = np.zeros(w * h * 6 + 512, np.uint8)
        n = self.lib.orc_ht_encode_ref(_p(a), C.c_int(w), C.c_int(h), C.c_int(kmax), _p(out), C.c_int(out.size))
        return out[:n].tobytes() if n >= 0 else n

    def enc_table(self, which):
        out = np.zeros(2048, np.uint16)
        self.lib.orc_ht_enc_table(C.c_int(which), _p(out))
        return out

    def decode_blocks(self, stream, offsets, lengths, kmax, mmsb, widths, heights, out_offsets, total_samples):
        """every array one entry per block; returns (block-major int32 buffer, status per block)"""
        n = len(offsets)
        stream = np.ascontiguousarray(stream, np.uint8)
        a = [np.ascontiguousarray(offsets, np.uint64), np.ascontiguousarray(lengths, np.uint32), np.ascontiguousarray(kmax, np.uint8),
             np.ascontiguousarray(mmsb, np.uint8), np.ascontiguousarray(widths, np.int32), np.ascontiguousarray(heights, np.int32),
             np.ascontiguousarray(out_offsets, np.int64)]
        out = np.zeros(total_samples, np.int32)
        status = np.zeros(max(n, 1), np.int32)
        self.lib.orc_ht_decode_blocks(_p(stream), *[_p(x) for x in a], C.c_long(n), _p(out), _p(status))
        return out, status[:n]
