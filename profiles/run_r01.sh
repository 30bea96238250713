set -x
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err
python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_launch.log 2>&1
python profiles/prof_driver.py stft > gpurun_out/plain_stft.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stft4096 -s 2 -c 1 -o gpurun_out/prof_stft_r01c python profiles/prof_driver.py stft > gpurun_out/ncu_stft.log 2>&1
python profiles/prof_driver.py osc > gpurun_out/plain_osc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:oscbank_kernel -s 2 -c 1 -o gpurun_out/prof_osc_r01c python profiles/prof_driver.py osc > gpurun_out/ncu_osc.log 2>&1
tail -n 2 gpurun_out/bench_r01.err gpurun_out/ncu_launch.log
