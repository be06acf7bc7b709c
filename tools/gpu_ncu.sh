TAG=${1:-rX}
W=${2:-qft_n15}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_$W python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_$W.log 2>&1
