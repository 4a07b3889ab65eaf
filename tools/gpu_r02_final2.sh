# after the Box-Muller change: full tests, full bench (+ reference arm), launch list, headline / C2 / C4 / C3 / C5 captures (JSON + CSV only)
set -x
O=gpurun_out/${TAG:-r2i}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
cap() {  # workload kernel-regex units tag
  timeout 800 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 5 -o $O/$4 -f python bench.py --workload $1 --steps 1 --warmup 3 --units $3 --quick > $O/ncu_$4.log 2>&1
  python tools/ncu_to_json.py $O/$4.ncu-rep $3 '#0' > $O/ncu_$1.json
  python tools/ncu_to_json.py $O/$4.ncu-rep $3 '#4' > $O/ncu_$1_fused.json
  python tools/ncu_summary.py $O/$4.ncu-rep > $O/$4_ncu_metrics.csv
  python tools/ncu_phases.py $O/$4.ncu-rep 0 > $O/$4_stream_phases.txt; python tools/ncu_phase_time.py $O/$4.ncu-rep 0 > $O/$4_stream_time.txt
  python tools/ncu_phases.py $O/$4.ncu-rep 4 > $O/$4_fused_phases.txt; python tools/ncu_phase_time.py $O/$4.ncu-rep 4 > $O/$4_fused_time.txt
  rm -f $O/$4.ncu-rep
}
cap ofdm1024_qam64_mimo2x2_tdl ofdm_tdl_pair_kernel 5920 h22
cap c3_ofdm1024_qam64_siso_tdl ofdm_tdl_fpair_kernel 11840 c3
cap c5_ofdm2048_qam256_mimo4x4_tdl ofdm_tdl_pair_kernel 1776 c5
cap c2_qam64_flat_rayleigh siso_flat_kernel 100000000 c2
cap c4_qpsk_alamouti2x2 alamouti22_kernel 20000000 c4
# the bench reads profiles/ncu_*.json: use the fresh captures for this run's traffic / issue objects
cp $O/ncu_*.json profiles/
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --units 20000 --no-cpu > $O/ncu_launch.log 2>&1
ls -la $O
