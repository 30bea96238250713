# instrumented build of the pitch tracking kernel (per-phase clock64 sums of stream 0, printed by the kernel):
#   bash profiles/build_pitch_timeline.sh && gpurun -- 'CA_B200_LIB=/root/repo/profiles/_tl/libpitchtl.so python profiles/time_roundtrip.py 32 60'
set -e
cd "$(dirname "$0")/../cpp.audio_b200/csrc"
mkdir -p ../../profiles/_tl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DCA_PITCH_TIMELINE "$@" -c pitch.cu -o ../../profiles/_tl/pitch_tl.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/_tl/libpitchtl.so capi.o stft.o rfft.o rfftw.o oscbank.o roundtrip.o ../../profiles/_tl/pitch_tl.o ramp.o vocoder.o -lcudart
