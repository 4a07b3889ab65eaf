set -x
O=gpurun_out/${TAG:-r2e}; mkdir -p $O
N=${NG:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/multigpu_check_${N}gpu.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W" $O/multigpu_check_${N}gpu.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('$O/bench_${N}gpu.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.4g'%d['value'], 'fused %.4g'%d['fused_rng']['value'], 'e2e %.4g'%d['e2e']['value'], 'overhead ms %.3f'%d['e2e']['runner_overhead_ms_per_snr_point'], 'e2e_stream %.4g'%d['e2e_stream']['value'])
for k,v in d.get('configs',{}).items(): print(k, 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $O/bench_ref_${N}gpu.json 2> $O/bench_ref_${N}gpu.err; wc -l $O/bench_ref_${N}gpu.json
