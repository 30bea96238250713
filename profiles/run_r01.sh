# Round-1 evidence run (one B200): tests, bench lines, launch list, ncu captures. Every ncu pass follows a plain run of the
# same command that exited 0. Summaries are extracted afterwards with profiles/summarize.py into profiles/*.json.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err
python profiles/sweep_rfft.py > gpurun_out/sweep_rfft_r01.jsonl 2> gpurun_out/sweep_rfft_r01.err
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip > gpurun_out/ncu_launch.log 2>&1
python profiles/prof_driver.py stft > gpurun_out/plain_stft.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stft4096 -s 2 -c 1 -f -o gpurun_out/prof_stft_r01 python profiles/prof_driver.py stft > gpurun_out/ncu_stft.log 2>&1
python profiles/prof_driver.py osc > gpurun_out/plain_osc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:oscbank_kernel -s 2 -c 1 -f -o gpurun_out/prof_osc_r01 python profiles/prof_driver.py osc > gpurun_out/ncu_osc.log 2>&1
python profiles/prof_driver.py rfft 10 > gpurun_out/plain_rfft.log 2>&1 && ncu --set full --clock-control none -k regex:rfft_warp -s 2 -c 1 -f -o gpurun_out/prof_rfft1024_r01 python profiles/prof_driver.py rfft 10 > gpurun_out/ncu_rfft.log 2>&1
python profiles/prof_driver.py rfft 19 > gpurun_out/plain_rfft19.log 2>&1 && ncu --set full --clock-control none -k regex:big2_ -s 4 -c 2 -f -o gpurun_out/prof_big2_r01 python profiles/prof_driver.py rfft 19 > gpurun_out/ncu_big2.log 2>&1
python profiles/prof_driver.py roundtrip > gpurun_out/plain_roundtrip.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pitch_ -s 2 -c 2 -f -o gpurun_out/prof_pitch_r01 python profiles/prof_driver.py roundtrip > gpurun_out/ncu_pitch.log 2>&1
CA_RT_TIMING=1 python profiles/time_roundtrip.py 32 60 > gpurun_out/roundtrip_time.txt 2>&1
python profiles/time_vocoder.py > gpurun_out/vocoder_time.txt 2>&1
tail -n 2 gpurun_out/bench_r01.err gpurun_out/ncu_launch.log gpurun_out/ncu_stft.log gpurun_out/ncu_osc.log gpurun_out/ncu_rfft.log
