set -x
O=gpurun_out/${TAG:-r2c}; mkdir -p $O
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"; tail -5 $O/bench_full.err
python - <<PY
import json
d=json.loads(open('$O/bench_full.json').read())
def show(k,v):
    print(k, 'value %.4g'%v['value'], 'frac %.3f'%v['roofline']['frac'], 'fused %.4g'%v['fused_rng']['value'], 'e2e %.4g (%.2f of fused, overhead %.3f ms)'%(v['e2e']['value'], v['e2e']['frac_of_device_fused_rate'], v['e2e']['runner_overhead_ms_per_snr_point']), 'e2e_stream %.4g'%v.get('e2e_stream',{}).get('value',0), v['roofline']['kernel'])
show('default', d)
for k,v in d.get('configs',{}).items(): show(k,v)
print('parity', d.get('parity')); print('f64', d.get('f64')); print('cpu', d.get('cpu_baseline')); print('issue', d.get('issue')); print(d['roofline'].get('traffic_note'))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
