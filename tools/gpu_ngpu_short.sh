# N-GPU check + bench without the reference arm (budget-friendly): NG=8 TAG=r2q bash tools/gpu_ngpu_short.sh
O=gpurun_out/${TAG:-r2q}; mkdir -p $O
N=${NG:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/multigpu_check_${N}gpu.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W" $O/multigpu_check_${N}gpu.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('$O/bench_${N}gpu.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.4g'%d['value'], 'fused %.4g'%d['fused_rng']['value'], 'e2e %.4g'%d['e2e']['value'], 'overhead ms %.3f'%d['e2e']['runner_overhead_ms_per_snr_point'])
for k,v in d.get('configs',{}).items(): print(k, 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'])
PY
