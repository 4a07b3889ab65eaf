# --set full captures of launches [S, S+C) matching kernel regex K of one bench workload (stream + fused legs);
# usage: bash tools/gpu_ncu_multi.sh WORKLOAD KERNEL_REGEX UNITS SKIP COUNT TAG
mkdir -p gpurun_out/q
W=${1:-ofdm1024_qam64_mimo2x2_tdl}; K=${2:-ofdm_tdl_pair_kernel}; U=${3:-5920}; S=${4:-4}; C=${5:-5}; TAG=${6:-cap}
timeout 800 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/q/$TAG -f python bench.py --workload $W --steps 1 --warmup 3 --units $U --quick > gpurun_out/q/ncu_$TAG.log 2>&1
tail -2 gpurun_out/q/ncu_$TAG.log
ls -la gpurun_out/q/$TAG.ncu-rep
