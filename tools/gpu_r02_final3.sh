# last pass of round 2: full GPU suite, smoke, C3 re-capture (frame-pair kernel changed), full bench + reference arm
set -x
O=gpurun_out/${TAG:-r2n}; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 800 ncu --set full --clock-control none --import-source on -k regex:ofdm_tdl_fpair_kernel -s 4 -c 5 -o $O/c3 -f python bench.py --workload c3_ofdm1024_qam64_siso_tdl --steps 1 --warmup 3 --units 11840 --quick > $O/ncu_c3.log 2>&1
python tools/ncu_to_json.py $O/c3.ncu-rep 11840 '#0' > $O/ncu_c3_ofdm1024_qam64_siso_tdl.json
python tools/ncu_to_json.py $O/c3.ncu-rep 11840 '#4' > $O/ncu_c3_ofdm1024_qam64_siso_tdl_fused.json
python tools/ncu_summary.py $O/c3.ncu-rep > $O/c3_ncu_metrics.csv
for i in 0 4; do m=stream; [ $i = 4 ] && m=fused; python tools/ncu_phases.py $O/c3.ncu-rep $i > $O/c3_${m}_phases.txt; python tools/ncu_phase_time.py $O/c3.ncu-rep $i > $O/c3_${m}_time.txt; done
rm -f $O/c3.ncu-rep
cp $O/ncu_c3_*.json profiles/
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
ls -la $O
