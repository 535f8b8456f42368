"""The C-ABI library loads without a GPU and exports every symbol include/j2k_b200.h declares
(no compute call is made here)."""
import ctypes
import os
import re

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "j2k_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(j2k_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    G.build()
    lib = ctypes.CDLL(G.LIB)
    names = declared_symbols()
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/j2k_b200.h but not exported: {missing}"
    assert lib.j2k_abi_version() == 1


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device every compute entry point fails with J2K_ERR_CUDA (no CPU path exists)."""
    import j2kb200
    from j2kb200 import abi
    try:
        import torch
        if torch.cuda.is_available():
            return  # a GPU box: covered by the -m gpu suite
    except Exception:
        pass
    try:
        j2kb200.Context()
    except j2kb200.J2KError as e:
        assert e.code == abi.J2K_ERR_CUDA
    else:
        raise AssertionError("Context() succeeded without a CUDA device")


def test_product_sources_never_load_the_oracle():
    """The product path must not import, link or execute anything under oracle/ (it is the checker)."""
    pk = os.path.join(ROOT, "go-dicom-codec_b200")
    banned = ("oracle_lib", "j2k_oracle", "ht_oracle", "libj2k_oracle", "libht_oracle", "np_mirror", "import oracle", "oracle/")
    for d, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                hits = [b for b in banned if b in text]
                assert not hits, f"{f} references the oracle: {hits}"


def test_header_compiles_as_c_and_a_c_program_round_trips(tmp_path):
    """tests/c/abi_smoke.c: a plain C caller (gcc -std=c99) linked against the CPU-emulator flavour of the library."""
    import subprocess
    import emu_lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = emu_lib.build()
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "c", "abi_smoke.c"), lib, "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "c abi ok" in out.stdout


def test_error_text_is_kept_per_context_across_threads(tmp_path):
    """tests/c/err_threads.c: thread A fails, thread B reads the message through the context (cgo goroutine migration)."""
    import subprocess
    import emu_lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = emu_lib.build()
    exe = str(tmp_path / "err_threads")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pthread", "-I", os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "c", "err_threads.c"), lib, "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "crosses threads" in out.stdout
