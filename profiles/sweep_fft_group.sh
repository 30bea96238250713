# work-buffer size of the three-pass FFT path (N >= 65536): does an L2-resident work buffer keep the passes off HBM?
for mb in 512 128 96 64 48 32 16; do
  echo "group ${mb} MB"
  CA_FFT_GROUP_MB=$mb python profiles/sweep_rfft.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d.get('n_fft',0) >= 65536: print('  ', d['n_fft'], round(d['ms'],3), 'ms', round(d['gbs']), 'GB/s', round(d['frac_hbm'],3))"
done
