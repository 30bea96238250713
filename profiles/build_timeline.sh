# instrumented build of the analysis kernel (per-phase clock64 stamps, printed to stderr by ca_stft_analyze):
#   bash profiles/build_timeline.sh && gpurun -- 'CA_B200_LIB=/root/repo/profiles/_tl/libtl.so python profiles/prof_driver.py stft'
set -e
cd "$(dirname "$0")/../cpp.audio_b200/csrc"
mkdir -p ../../profiles/_tl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DCA_STFT_TIMELINE -c stft.cu -o ../../profiles/_tl/stft_tl.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/_tl/libtl.so capi.o ../../profiles/_tl/stft_tl.o rfft.o rfftw.o oscbank.o roundtrip.o pitch.o ramp.o vocoder.o -lcudart
