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
# general-alignment ring variants (masked and unmasked store paths, every row phase)
for c in [(140, 20, 1, 16, False, 2, False), (150, 23, 1, 12, False, 3, True), (134, 18, 1, 8, True, 2, True), (271, 13, 1, 16, False, 4, True),
          (535, 10, 1, 16, False, 3, False), (131, 17, 1, 8, False, 3, False)]:
    PC.check_pipeline(ctx, orc, *c)
PC.check_pipeline(ctx, orc, 100, 70, 1, 8, False, 3, False, tile=(48, 32))
PC.check_pipeline(ctx, orc, 70, 50, 3, 8, False, 2, False, tile=(32, 32))
PC.check_blocks(ctx, orc, 70, 50, 1, 12, 3, False, cb=(16, 8))
PC.check_blocks(ctx, orc, 48, 40, 3, 8, 2, True, tile=(32, 32), cb=(8, 8))
PC.check_blocks_roi(ctx, orc, 48, 40, 3, 8, 2, True, [4, 0, 7], tile=(32, 32), cb=(8, 8))
PC.check_blocks_roi(ctx, orc, 70, 50, 1, 12, 3, False, [9], cb=(16, 8))
PC.check_pipelined_order(ctx, orc, 64, 32, 1, 12, 3, False, 9, 4, 3)
PC.check_pipelined_order(ctx, orc, 48, 32, 3, 8, 2, True, 6, 4, 2)
# fwd3w_kernel (one converting producer + three consumers per quad; the emulator's warp plays the four roles in turn)
PC.check_one_producer_forward(ctx, orc, 512, 44, 8, 3, 2)
PC.check_one_producer_forward(ctx, orc, 256, 64, 8, 2, 2, tile=(128, 32))
PC.check_one_producer_forward(ctx, orc, 512, 24, 16, 3, 1)
PC.check_one_producer_forward(ctx, orc, 320, 37, 8, 2, 1, chunk=8)
PC.check_custom_mct(ctx, orc, 40, 24, 8, 2, True, "bindings")
PC.check_wavelet_api(ctx, orc, 130, 70, 5, 0, 0)
PC.check_wavelet_api(ctx, orc, 33, 17, 2, 1, 0)
# HTJ2K block coder (round 2): every block shape, corrupted segments (the decoder must stay inside its buffers whatever the
# bytes say), the encoder's vector reads past the end of its streams
import ht_oracle_lib, ht_parity as HP
ht = ht_oracle_lib.HtOracle()
HP.check_fixture(ctx, ht, orc, "mono_u8_127x129", "fo_htj2k_lossless")
HP.check_fixture(ctx, ht, orc, "rgb_u8_127x129", "fo_htj2k_lossless_rpcl")
HP.check_mutations(ctx, ht, orc, "mono_u16_128x128", "fo_htj2k_lossless", rounds=4, seed=1)
HP.check_mutations(ctx, ht, orc, "rgb_u8_128x128", "fo_htj2k_lossless", rounds=3, seed=2)
for g in [(64, 64, 0, 64, 64), (70, 37, 1, 32, 32), (40, 24, 0, 4, 4), (130, 9, 0, 128, 32), (9, 130, 0, 16, 256), (260, 4, 0, 1024, 4),
          (5, 300, 0, 4, 1024), (1, 7, 0, 8, 8), (2, 2, 0, 4, 4)]:
    HP.check_random_streams(ctx, ht, orc, *g, seed=g[0] * 131 + g[1])
for g in [(64, 64, 0, 64, 64, 12, 0.7, 1, True), (75, 61, 2, 64, 64, 16, 0.9, 1, True), (33, 130, 1, 8, 512, 12, 1.0, 1, True),
          (150, 10, 0, 1024, 4, 9, 0.8, 1, True), (48, 48, 2, 64, 64, 8, 0.05, 3, False)]:
    HP.check_generated(ctx, ht, orc, *g[:7], seed=3, components=g[7], reversible=g[8])
for g in [(64, 64, 1, 8, 0, 64, 64, True), (75, 61, 1, 16, 2, 64, 64, True), (40, 40, 3, 8, 1, 16, 16, True), (33, 130, 1, 12, 1, 8, 512, True),
          (150, 10, 1, 9, 0, 1024, 4, True), (48, 48, 3, 8, 2, 64, 64, False), (1, 1, 1, 8, 0, 4, 4, True), (2, 7, 1, 8, 0, 4, 4, True)]:
    HP.check_encode(ctx, ht, orc, *g[:7], seed=5, reversible=g[7])
HP.check_encode(ctx, ht, orc, 72, 56, 1, 12, 2, 32, 32, seed=4, nframes=3)
HP.check_encode(ctx, ht, orc, 64, 48, 1, 12, 1, 64, 64, seed=6, base=6)
print("asan run clean")
PY
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python $out/run.py
