# usage: tools/build_variant.sh NAME [extra nvcc flags]   -> csrc/build/libj2kb200_NAME.so from the working tree
name=$1; shift
cd "$(dirname "$0")/../go-dicom-codec_b200/csrc" && nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -shared -o build/libj2kb200_$name.so j2k_b200.cu "$@"
