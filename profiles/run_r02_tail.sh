set -x
cap() {
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
cap rfft32768_cluster regex:rfft_cluster2 2 1 "" rfft 15
cap rfft65536_cluster regex:rfft_cluster2 2 1 "" rfft 16
cap rfft131072_cluster regex:rfft_cluster2 2 1 "" rfft 17
python profiles/time_vocoder.py 32 60 > gpurun_out/vocoder_time_r02.txt 2>&1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip --skip-e2e-spectrum > gpurun_out/ncu_launch.log 2>&1
tail -n 2 gpurun_out/bench_r02.err
