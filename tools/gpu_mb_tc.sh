# tensor-core microbenchmarks (tools/microbench_tc.cu, prebuilt into build_mb/mb_tc) + ncu pipe utilisation
mkdir -p gpurun_out/tc
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/tc/clocks.csv &
SMI=$!
./build_mb/mb_tc all | tee gpurun_out/tc/microbench_tc.jsonl
kill $SMI
for g in hk fir gram; do
ncu --clock-control none --csv --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active \
  --log-file gpurun_out/tc/ncu_$g.csv ./build_mb/mb_tc $g > /dev/null 2>&1
done
tail -3 gpurun_out/tc/clocks.csv
