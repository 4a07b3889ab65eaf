# tcgen05 H_k path: parity tests that reach it, then headline bench (timeouts guard against a hung barrier)
mkdir -p gpurun_out/q
timeout 120 python - <<'PY'
import numpy as np, bench
w = bench.WORKLOADS['ofdm1024_qam64_mimo2x2_tdl']
link = bench.make_link(w)
c1 = link.run(300, first_unit=7)
link.params.reserved = 2            # CUDA-core H_k
c0 = link.run(300, first_unit=7)
print('tc', c1, 'cuda-core', c0, 'symbol drift', int(c1[0]) - int(c0[0]))
d = link.draw(7, 300)
link.params.reserved = 0
cs = link.run(300, first_unit=7, draws=d)
print('stream tc', cs, 'equal to fused tc:', np.array_equal(cs, c1))
PY
echo "probe rc=$?"
timeout 600 python -m pytest tests/test_gpu_ofdm_tdl.py tests/test_gpu_precision.py -m gpu -x -q 2>&1 | tail -6
timeout 200 python bench.py --workload ofdm1024_qam64_mimo2x2_tdl --quick --steps 10 > gpurun_out/q/h.json 2>gpurun_out/q/h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q/h.json'))
print('headline value %.4g kernel_ms %.3f fused %.4g %s'%(d['value'], d['roofline']['kernel_ms'], d['fused_rng']['value'], d['clocks']))
PY
