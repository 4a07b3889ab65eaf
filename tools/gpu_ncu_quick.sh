# one --set full capture of the headline kernel (5920 frames = 10 waves of 148 x 4... CTAs), report left in gpurun_out/q/
mkdir -p gpurun_out/q
W=${1:-ofdm1024_qam64_mimo2x2_tdl}; K=${2:-ofdm_tdl_pair_kernel}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o gpurun_out/q/cap -f python bench.py --workload $W --steps 1 --warmup 3 --units ${3:-5920} --no-cpu > gpurun_out/q/ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/q/cap.ncu-rep | head -12
