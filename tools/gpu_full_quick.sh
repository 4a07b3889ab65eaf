timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in c4_qpsk_alamouti2x2 c2_qam64_flat_rayleigh ofdm1024_qam64_mimo2x2_tdl; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 10 > gpurun_out/q/$w.json 2>gpurun_out/q/$w.err
  python - "$w" <<'PY'
import json,sys
d=json.load(open('gpurun_out/q/%s.json'%sys.argv[1]))
print('%-34s value %.4g  kernel_ms %.3f frac %.3f fused %.4g  clocks %s'%(sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['fused_rng']['value'], d['clocks']))
PY
done
