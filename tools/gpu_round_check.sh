set -x
O=gpurun_out/${TAG:-r1c}; mkdir -p $O

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
timeout 400 python bench.py > $O/bench_headline.json 2> $O/bench_headline.err; tail -c 600 $O/bench_headline.json
for w in c3_ofdm1024_qam64_siso_tdl c5_ofdm2048_qam256_mimo4x4_tdl c2_qam64_flat_rayleigh c4_qpsk_alamouti2x2; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --units 20000 --no-cpu > $O/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:ofdm_tdl_pair_kernel -s 4 -c 1 -o $O/pair22_full -f python bench.py --steps 1 --warmup 3 --units 5920 --no-cpu > $O/ncu_full.log 2>&1
ls -la $O
