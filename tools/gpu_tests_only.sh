set -x
O=gpurun_out/${TAG:-r2b}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -30 $O/pytest_gpu.log
grep -h "^parity\|^c3_\|^ofdm1024\|^c5_\|^\.*parity" $O/pytest_gpu.log | head -80
