# compute-sanitizer memcheck + racecheck over tools/sanitize_smoke.py, then the new TMA alignment test
mkdir -p gpurun_out/san
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/san/memcheck.log 2>&1; tail -4 gpurun_out/san/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/san/racecheck.log 2>&1; tail -4 gpurun_out/san/racecheck.log
timeout 600 python -m pytest tests/test_gpu_ofdm_tdl.py -m gpu -x -q -k "tma or fused_equals" 2>&1 | tail -4
