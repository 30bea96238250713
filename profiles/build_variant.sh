# build a variant of the library with extra -D flags into profiles/_tl/lib<name>.so:  bash profiles/build_variant.sh name -DFOO -DBAR
set -e
name=$1; shift
cd "$(dirname "$0")/../cpp.audio_b200/csrc"
mkdir -p ../../profiles/_tl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c stft.cu -o ../../profiles/_tl/stft_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/_tl/lib$name.so capi.o ../../profiles/_tl/stft_$name.o rfft.o rfftw.o oscbank.o roundtrip.o pitch.o ramp.o vocoder.o -lcudart
