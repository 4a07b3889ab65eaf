mkdir -p gpurun_out/q
for i in 1 2; do timeout 200 python bench.py --workload ofdm1024_qam64_mimo2x2_tdl --quick --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('TC stream %.4g' % d['value'], 'fused %.4g' % d['fused_rng']['value'], d['roofline']['kernel'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ofdm_tdl_pair_kernel -s 4 -c 1 -o gpurun_out/q/tc22 -f python bench.py --workload ofdm1024_qam64_mimo2x2_tdl --steps 1 --warmup 3 --units 5920 --quick > gpurun_out/q/ncu_tc22.log 2>&1
python tools/ncu_phases.py gpurun_out/q/tc22.ncu-rep 0 > gpurun_out/q/tc22_phases.txt; python tools/ncu_phase_time.py gpurun_out/q/tc22.ncu-rep 0 > gpurun_out/q/tc22_time.txt
python tools/ncu_summary.py gpurun_out/q/tc22.ncu-rep > gpurun_out/q/tc22_metrics.csv
cat gpurun_out/q/tc22_phases.txt; cat gpurun_out/q/tc22_time.txt; grep -E "time_duration|inst_executed|issue_active|warps_active|tensor|wavefronts_mem_shared.sum,|bank_conflicts|fma_cycles" gpurun_out/q/tc22_metrics.csv
