# full GPU check: every gpu-marked test, then the headline bench (with the CPU legs) and the reference arm
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python -c "
import json;d=json.load(open('gpurun_out/bench_full.json'));print(round(d['value']/1e6,1),'Mfr/s', round(d['ms_per_step'],3),'ms', 'e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3)); print(d.get('cpu_baseline')); print(d['extra'])"
