"""Builds tests/emu/build/libj2kb200_emu.so: the REAL product sources (go-dicom-codec_b200/csrc)
compiled with g++ against tests/emu/cuda_runtime.h, a single-threaded stand-in for the CUDA runtime
(warps run as fibers).  Test infrastructure for the `-m "not gpu"` suite only; never a product path."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "build", "libj2kb200_emu.so")
CSRC = os.path.join(ROOT, "go-dicom-codec_b200", "csrc")


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in ("j2k_b200.cu", "j2k_kernels.cuh", "j2k_pointwise.cuh", "j2k_ring.cuh", "j2k_ht.cuh", "j2k_ht_tables.inc", "j2k_ht_enc.cuh", "j2k_ht_enc_tables.inc")]
    srcs += [os.path.join(EMU, f) for f in ("cuda_runtime.h", "emu_runtime.cpp")]
    srcs.append(os.path.join(ROOT, "include", "j2k_b200.h"))
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in srcs)
    if force or stale:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call([
            "g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-x", "c++",
            "-I", EMU, "-o", LIB, os.path.join(CSRC, "j2k_b200.cu"), os.path.join(EMU, "emu_runtime.cpp")])
    return LIB
