# Round-2 evidence run (one B200): tests, bench lines, sweeps, launch list, ncu captures. Every ncu pass follows a plain run of the
# same command that exited 0. Summaries are extracted afterwards with profiles/summarize.py into profiles/*_r02.json.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02.log 2>&1; tail -2 gpurun_out/pytest_gpu_r02.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
python profiles/sweep_rfft.py > gpurun_out/sweep_rfft_r02.jsonl 2> gpurun_out/sweep_rfft_r02.err
python profiles/time_stft_sizes.py > gpurun_out/stft_sizes_r02.txt 2>&1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/ncu_launch.log 2>&1
python profiles/prof_driver.py stft > gpurun_out/plain_stft.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stft4096 -s 2 -c 1 -f -o gpurun_out/prof_stft_r02 python profiles/prof_driver.py stft > gpurun_out/ncu_stft.log 2>&1
python profiles/prof_driver.py osc > gpurun_out/plain_osc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:oscbank_kernel -s 2 -c 1 -f -o gpurun_out/prof_osc_r02 python profiles/prof_driver.py osc > gpurun_out/ncu_osc.log 2>&1
python profiles/prof_driver.py rfft 10 > gpurun_out/plain_rfft.log 2>&1 && ncu --set full --clock-control none -k regex:rfft_warp -s 2 -c 1 -f -o gpurun_out/prof_rfft1024_r02 python profiles/prof_driver.py rfft 10 > gpurun_out/ncu_rfft.log 2>&1
python profiles/prof_driver.py rfft 19 > gpurun_out/plain_rfft19.log 2>&1 && ncu --set full --clock-control none -k regex:big2_ -s 4 -c 2 -f -o gpurun_out/prof_big2_r02 python profiles/prof_driver.py rfft 19 > gpurun_out/ncu_big2.log 2>&1
python profiles/prof_driver.py roundtrip_small > gpurun_out/plain_rts.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"pitch_|stft64|tracks_kernel" -s 12 -c 6 -f -o gpurun_out/prof_roundtrip_r02 python profiles/prof_driver.py roundtrip_small > gpurun_out/ncu_rt.log 2>&1
CA_RT_TIMING=1 python profiles/time_roundtrip.py 32 60 > gpurun_out/roundtrip_time_r02.txt 2>&1
python profiles/time_roundtrip.py 32 60 >> gpurun_out/roundtrip_time_r02.txt 2>&1
python profiles/time_vocoder.py > gpurun_out/vocoder_time_r02.txt 2>&1
tail -n 2 gpurun_out/bench_r02.err gpurun_out/ncu_launch.log gpurun_out/ncu_stft.log gpurun_out/ncu_osc.log gpurun_out/ncu_rfft.log gpurun_out/ncu_big2.log gpurun_out/ncu_rt.log
