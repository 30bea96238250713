# quick GPU check used while tuning the analysis kernel
python -m pytest tests/test_gpu_stft.py -x -q -m gpu > gpurun_out/pytest_stft.log 2>&1; tail -3 gpurun_out/pytest_stft.log
python bench.py --steps 20 --warmup 3 --skip-synth --skip-cpu > gpurun_out/b.json 2> gpurun_out/b.err; tail -2 gpurun_out/b.err
python -c "
import json;d=json.load(open('gpurun_out/b.json'));print(round(d['value']/1e6,1),'Mfr/s', round(d['ms_per_step'],3),'ms', 'e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3))"
CA_STFT_WARPS=4 CA_B200_LIB=/root/repo/profiles/_tl/libtl.so python profiles/prof_driver.py stft 2>&1 | tail -3
CA_B200_LIB=/root/repo/profiles/_tl/libtl.so python profiles/prof_driver.py stft 2>&1 | tail -3
