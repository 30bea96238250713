# Round-2 evidence refresh after the stft4096p_kernel changes (frame position without divisions, sign-bit peak mask, packed |X|^2):
# ncu capture of the kernel first (traffic.json, which bench.py quotes, is rebuilt from it on the box), then the bench lines, the
# fused-analysis sweep and the launch list. Every ncu pass follows a plain run of the same command that exited 0.
set -x
python profiles/prof_driver.py stft > gpurun_out/plain_stft4096.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:stft4096 -s 2 -c 1 -f -o gpurun_out/prof_stft4096 python profiles/prof_driver.py stft > gpurun_out/ncu_stft4096.log 2>&1
python profiles/summarize.py gpurun_out/prof_stft4096.ncu-rep > gpurun_out/ncu_stft4096_r02.json
python profiles/top_sass.py gpurun_out/prof_stft4096.ncu-rep 40 > gpurun_out/top_sass_stft4096_r02.txt 2>&1
rm -f gpurun_out/prof_stft4096.ncu-rep
cp gpurun_out/ncu_stft4096_r02.json profiles/ncu_stft4096_r02.json && python profiles/make_traffic.py && cp profiles/traffic.json gpurun_out/traffic.json
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
python profiles/time_stft_sizes.py > gpurun_out/stft_sizes_r02.txt 2>&1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/ncu_launch.log 2>&1
python __graft_entry__.py > /dev/null 2>&1; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
tail -n 2 gpurun_out/bench_r02.err gpurun_out/ncu_launch.log
head -c 600 gpurun_out/bench_r02.json
