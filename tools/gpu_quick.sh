# short 1-GPU call: GPU parity tests, bench lines for the n=15 workloads, one ncu --set full capture of the sweeps of
# one workload (default qft_n15).  Usage: tools/gpu_quick.sh TAG [ncu-workload] [more bench workloads...]
TAG=${1:-rX}; W=${2:-qft_n15}; shift; shift
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for w in qft_n15 bv_n15 hlayer_n15 random_c1c2_n15 single_n15 "$@"; do python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
if [ "$W" != "none" ]; then
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_$W python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_$W.log 2>&1
fi
python tools/show_bench.py ${TAG}
