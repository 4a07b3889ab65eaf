# round-2 evidence run on one B200: tests, smoke, full bench (+ reference arm), launch list, --set full captures of every
# BASELINE kernel (stream + fused), tensor-core microbenchmarks.  Everything lands in gpurun_out/$TAG.
set -x
O=gpurun_out/${TAG:-r2h}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --units 20000 --no-cpu > $O/ncu_launch.log 2>&1
cap() {  # workload kernel-regex units tag keep-report
  timeout 800 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 5 -o $O/$4 -f python bench.py --workload $1 --steps 1 --warmup 3 --units $3 --quick > $O/ncu_$4.log 2>&1
  # launch 0 of the report = stream mode (kernel-only timed step), launch 4 = fused RNG (timed step)
  python tools/ncu_to_json.py $O/$4.ncu-rep $3 '#0' > $O/ncu_$1.json
  python tools/ncu_to_json.py $O/$4.ncu-rep $3 '#4' > $O/ncu_$1_fused.json
  python tools/ncu_summary.py $O/$4.ncu-rep > $O/$4_ncu_metrics.csv
  python tools/ncu_phases.py $O/$4.ncu-rep 0 > $O/$4_stream_phases.txt; python tools/ncu_phase_time.py $O/$4.ncu-rep 0 > $O/$4_stream_time.txt
  python tools/ncu_phases.py $O/$4.ncu-rep 4 > $O/$4_fused_phases.txt; python tools/ncu_phase_time.py $O/$4.ncu-rep 4 > $O/$4_fused_time.txt
  ls -la $O/$4.ncu-rep
  if [ "$5" != keep ]; then rm -f $O/$4.ncu-rep; fi
}
cap ofdm1024_qam64_mimo2x2_tdl ofdm_tdl_pair_kernel 5920 h22 keep
cap c3_ofdm1024_qam64_siso_tdl ofdm_tdl_fpair_kernel 11840 c3
cap c5_ofdm2048_qam256_mimo4x4_tdl ofdm_tdl_pair_kernel 1776 c5
cap c2_qam64_flat_rayleigh siso_flat_kernel 100000000 c2
cap c4_qpsk_alamouti2x2 alamouti22_kernel 20000000 c4
mkdir -p $O/tc
./build_mb/mb_tc all > $O/tc/microbench_tc.jsonl; cat $O/tc/microbench_tc.jsonl
for g in hk fir gram; do
ncu --clock-control none --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active \
  --log-file $O/tc/ncu_$g.csv ./build_mb/mb_tc $g > /dev/null 2>&1
done
ls -la $O
