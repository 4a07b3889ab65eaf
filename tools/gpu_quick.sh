# quick GPU iteration: the OFDM/TDL parity tests, then short device-timed bench lines (no CPU leg)
mkdir -p gpurun_out/q
timeout 600 python -m pytest tests/test_gpu_ofdm_tdl.py -m gpu -x -q 2>&1 | tail -${TAILN:-5}
for w in ofdm1024_qam64_mimo2x2_tdl c3_ofdm1024_qam64_siso_tdl c5_ofdm2048_qam256_mimo4x4_tdl; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 5 > gpurun_out/q/$w.json 2>gpurun_out/q/$w.err
  python - "$w" <<'PY'
import json,sys
d=json.load(open('gpurun_out/q/%s.json'%sys.argv[1]))
print('%-34s value %.4g  kernel_ms %.3f  fused %.4g  e2e %.4g clocks %s'%(sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['fused_rng']['value'], (d.get('e2e') or {}).get('value',0), d['clocks']))
PY
done
