# Builds the CPU-emulator flavour of the product sources with AddressSanitizer and runs a parity subset under it:
# every global-memory access of the kernels is bounds-checked against the (host) allocations.  Test infrastructure only.
set -e
out=${TMPDIR:-/tmp}/j2k_asan; mkdir -p $out
root=$(cd "$(dirname "$0")/.." && pwd)
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fno-fast-math -fPIC -shared -w -x c++ \
    -I $root/tests/emu -o $out/libj2kb200_emu_asan.so $root/go-dicom-codec_b200/csrc/j2k_b200.cu $root/tests/emu/emu_runtime.cpp
cat > $out/run.py <<PY
import sys
sys.path[:0] = ["$root/tests", "$root/go-dicom-codec_b200", "$root"]
import oracle_lib, j2kb200, parity_cases as PC
orc = oracle_lib.Oracle()
ctx = j2kb200.Context(lib_path="$out/libj2kb200_emu_asan.so")
for c in [(64, 64, 1, 8, False, 3, True), (61, 47, 1, 16, True, 3, True), (64, 64, 1, 12, False, 3, False), (67, 53, 1, 8, False, 2, False),
          (48, 40, 3, 8, False, 3, True), (48, 40, 3, 8, False, 3, False), (256, 8, 1, 16, False, 2, True), (256, 8, 1, 12, True, 2, False),
          (512, 12, 1, 16, False, 2, False), (384, 10, 1, 8, False, 2, False), (136, 12, 3, 8, False, 2, False), (128, 16, 3, 8, False, 2, False),
          (128, 16, 3, 8, False, 2, True), (128, 16, 3, 16, False, 2, False), (520, 9, 1, 16, False, 3, False), (1, 1, 1, 8, False, 2, False)]:
    PC.check_pipeline(ctx, orc, *c)
PC.check_pipeline(ctx, orc, 100, 70, 1, 8, False, 3, False, tile=(48, 32))
PC.check_pipeline(ctx, orc, 70, 50, 3, 8, False, 2, False, tile=(32, 32))
PC.check_blocks(ctx, orc, 70, 50, 1, 12, 3, False, cb=(16, 8))
PC.check_blocks(ctx, orc, 48, 40, 3, 8, 2, True, tile=(32, 32), cb=(8, 8))
PC.check_blocks_roi(ctx, orc, 48, 40, 3, 8, 2, True, [4, 0, 7], tile=(32, 32), cb=(8, 8))
PC.check_blocks_roi(ctx, orc, 70, 50, 1, 12, 3, False, [9], cb=(16, 8))
PC.check_pipelined_order(ctx, orc, 64, 32, 1, 12, 3, False, 9, 4, 3)
PC.check_pipelined_order(ctx, orc, 48, 32, 3, 8, 2, True, 6, 4, 2)
PC.check_custom_mct(ctx, orc, 40, 24, 8, 2, True, "bindings")
PC.check_wavelet_api(ctx, orc, 130, 70, 5, 0, 0)
PC.check_wavelet_api(ctx, orc, 33, 17, 2, 1, 0)
print("asan run clean")
PY
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python $out/run.py
