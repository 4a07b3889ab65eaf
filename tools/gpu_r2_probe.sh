set -x
O=gpurun_out/${TAG:-r2a}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
grep -h "^parity\|^c3_\|^ofdm1024\|^c5_" $O/pytest_gpu.log | head -60
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
timeout 400 python bench.py --quick > $O/bench_headline.json 2> $O/bench_headline.err; tail -c 1500 $O/bench_headline.json
for w in c3_ofdm1024_qam64_siso_tdl c5_ofdm2048_qam256_mimo4x4_tdl c2_qam64_flat_rayleigh c4_qpsk_alamouti2x2; do
  timeout 300 python bench.py --workload $w --quick --steps 5 > $O/bench_$w.json 2> $O/bench_$w.err; tail -c 700 $O/bench_$w.json
done
