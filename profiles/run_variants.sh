# A/B runs of library variants built by profiles/build_variant.sh
for v in "$@"; do
CA_B200_LIB=/root/repo/profiles/_tl/lib$v.so python bench.py --steps 20 --warmup 3 --skip-synth --skip-cpu > gpurun_out/b_$v.json 2> gpurun_out/b_$v.err; tail -2 gpurun_out/b_$v.err
python -c "
import json;d=json.load(open('gpurun_out/b_$v.json'));print('$v', round(d['value']/1e6,1),'Mfr/s', round(d['ms_per_step'],3),'ms', 'frac', round(d['roofline']['frac'],3))"
done
