set -x
O=gpurun_out/${TAG:-r2d}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_flat_links.py tests/test_gpu_facade.py -m gpu -q -x > $O/pytest_flat.log 2>&1; tail -5 $O/pytest_flat.log
bash tools/gpu_bench_full.sh
