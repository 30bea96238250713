# persistent two-pass FFT (N >= 65536): tile shapes and ring size against the two-launch form
run() { python profiles/sweep_rfft.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d.get('n_fft',0) >= 65536: print('  ', d['n_fft'], round(d['ms'],3), 'ms', round(d['gbs']), 'GB/s', round(d['frac_hbm'],3), 'err %.2e' % d['err'])"; }
echo "two launches"; CA_FFT_TWO_LAUNCHES=1 timeout 120 bash -c "$(declare -f run); run"
for g in 4 8; do for ring in 16 32 64; do
  echo "persistent G $g ring ${ring} MB"; CA_FFT_G=$g CA_FFT_RING_MB=$ring timeout 120 bash -c "$(declare -f run); run"
done; done
