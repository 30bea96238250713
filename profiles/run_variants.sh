set -x
python -m pytest tests/test_gpu_stft.py -x -q -m gpu 2>&1 | tail -5
for v in 4 10 11 12 13; do
  CA_STFT_VARIANT=$v python bench.py --steps 10 --warmup 3 --skip-synth --skip-cpu > gpurun_out/sw_$v.json 2> gpurun_out/sw_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_$v.json'));print('variant $v', round(d['value']/1e6,1),'Mfr/s', round(d['ms_per_step'],3),'ms', 'e2e', round(d['e2e']['value']/1e6,1))"
done
