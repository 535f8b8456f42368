python tools/config_bench.py --steps 10 --only C3i
python tools/config_bench.py --steps 10 --only C5
J2K_B200_LIB=go-dicom-codec_b200/csrc/build/libj2kb200_rgb2.so python tools/config_bench.py --steps 10 --only C3i
J2K_B200_LIB=go-dicom-codec_b200/csrc/build/libj2kb200_rgb2.so python tools/config_bench.py --steps 10 --only C5
