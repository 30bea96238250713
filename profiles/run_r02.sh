# Round-2 evidence run (one B200): tests, bench lines, sweeps, launch list, ncu captures. Every ncu pass follows a plain run of the
# same command that exited 0. The captures are summarised on the box (profiles/summarize.py -> gpurun_out/ncu_*_r02.json) and the
# reports deleted: gpurun brings back at most 64 MiB.
set -x
cap() {  # cap <name> <ncu kernel filter> <skip> <count> <env or ""> <driver args...>
  name=$1; filt=$2; skip=$3; cnt=$4; envs=$5; shift 5
  env $envs python profiles/prof_driver.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
    env $envs ncu --set full --clock-control none -k "$filt" -s $skip -c $cnt -f -o gpurun_out/prof_$name python profiles/prof_driver.py "$@" > gpurun_out/ncu_$name.log 2>&1
  python profiles/summarize.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_${name}_r02.json 2>> gpurun_out/ncu_$name.log
  rm -f gpurun_out/prof_$name.ncu-rep
  tail -n 1 gpurun_out/ncu_$name.log
}
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02.log 2>&1; tail -2 gpurun_out/pytest_gpu_r02.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
python profiles/sweep_rfft.py > gpurun_out/sweep_rfft_r02.jsonl 2> gpurun_out/sweep_rfft_r02.err
python profiles/time_stft_sizes.py > gpurun_out/stft_sizes_r02.txt 2>&1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/ncu_launch.log 2>&1
cap stft4096 regex:stft4096 2 1 "" stft
# oscillator bank, config 3: the default path (prep, tables, tcgen05 kernel, mma.sync fill: the 4 launches of the third render), then
# the mma.sync kernel alone and the FFMA2 form of the product
cap oscbank_umma regex:osc 8 4 "" osc
cap oscbank_mma regex:oscbank_kernel 2 1 "CA_OSC_PATH=mma" osc
cap oscbank_fp32 regex:oscbank_kernel 2 1 "CA_OSC_PATH=fp32" osc
bash profiles/time_osc_paths.sh > gpurun_out/oscbank_paths_r02.txt 2>&1
cap rfft1024 regex:rfft_warp 2 1 "" rfft 10
cap big2 regex:big2_ 4 2 "" rfft 19
# the single-pass cluster kernels (2^15: 4 CTAs per frame, 2^16: 8, 2^17: 16 with st.async delivery)
cap rfft32768_cluster regex:rfft_cluster2 2 1 "" rfft 15
cap rfft65536_cluster regex:rfft_cluster2 2 1 "" rfft 16
cap rfft131072_cluster regex:rfft_cluster2 2 1 "" rfft 17
cap roundtrip "regex:pitch_|stft64|tracks_kernel" 12 6 "" roundtrip_small
CA_RT_TIMING=1 python profiles/time_roundtrip.py 32 60 > gpurun_out/roundtrip_time_r02.txt 2>&1
python profiles/time_roundtrip.py 32 60 >> gpurun_out/roundtrip_time_r02.txt 2>&1
python profiles/time_vocoder.py > gpurun_out/vocoder_time_r02.txt 2>&1
tail -n 2 gpurun_out/bench_r02.err gpurun_out/ncu_launch.log
du -sh gpurun_out
