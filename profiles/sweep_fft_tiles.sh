# tile shapes of the two-pass FFT path (N >= 65536): columns per CTA in pass 1 x rows per CTA in pass 2
for cw in 16 32; do for g in 4 8; do
  echo "CW $cw G $g"
  CA_FFT_CW=$cw CA_FFT_G=$g python profiles/sweep_rfft.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d.get('n_fft',0) >= 65536: print('  ', d['n_fft'], round(d['ms'],3), 'ms', round(d['gbs']), 'GB/s', round(d['frac_hbm'],3))"
done; done
