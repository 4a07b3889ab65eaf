# A/B two (or more) builds of the library on the SAME box: usage  bash tools/gpu_ab.sh libA.so libB.so ...
# (build variants with pyphysim_b200._build.build(lib_path=..., obj_dir=..., extra_flags=...); B200PHY_LIB selects one)
for i in 1 2 3; do
for lib in "$@"; do
  B200PHY_LIB=$PWD/$lib python bench.py --quick --steps 10 ${ABW:+--workload $ABW} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'stream %.4g' % d['value'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'fused %.4g' % d['fused_rng']['value'])"
done; done
