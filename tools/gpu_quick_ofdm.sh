# OFDM/TDL parity tests + device-timed bench of the three OFDM workloads (--quick)
mkdir -p gpurun_out/q
timeout 900 python -m pytest tests/test_gpu_ofdm_tdl.py tests/test_gpu_precision.py tests/test_torch_ops.py -m gpu -x -q 2>&1 | tail -${TAILN:-8}
for w in ${WL:-ofdm1024_qam64_mimo2x2_tdl c3_ofdm1024_qam64_siso_tdl c5_ofdm2048_qam256_mimo4x4_tdl}; do
  timeout 300 python bench.py --workload $w --quick --steps 10 > gpurun_out/q/$w.json 2>gpurun_out/q/$w.err
  python - "$w" <<'PY'
import json,sys
d=json.load(open('gpurun_out/q/%s.json'%sys.argv[1]))
print('%-34s value %.4g  kernel_ms %.3f  frac %.3f fused %.4g clocks %s'%(sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['fused_rng']['value'], d['clocks']))
PY
done
