# build a variant of the library with extra -D flags into profiles/_tl/lib<name>.so:  bash profiles/build_variant.sh name -DFOO -DBAR
# (only stft.cu is recompiled; the other objects are the ones `make -C cpp.audio_b200/csrc` left behind)
set -e
name=$1; shift
cd "$(dirname "$0")/../cpp.audio_b200/csrc"
mkdir -p ../../profiles/_tl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c stft.cu -o ../../profiles/_tl/stft_$name.o 2> ../../profiles/_tl/stft_$name.ptxas.log
objs=$(ls *.o | grep -v '^stft\.o$')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/_tl/lib$name.so $objs ../../profiles/_tl/stft_$name.o -lcudart
