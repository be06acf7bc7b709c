# one 1-GPU box call: GPU parity tests, smoke, bench lines (default + other workloads + reference arm),
# ncu launch list and ncu --set full captures of the sweep kernel.  Usage: tools/gpu_round.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
(time python bench.py) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
for w in bv_n15 random_c1c2_n15 single_n15 hlayer_n15 vqe_uccsd_n8 adder_n10; do python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_qft_n15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
for W in qft_n15 random_c1c2_n15; do
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_$W python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_$W.log 2>&1
done
python tools/show_bench.py ${TAG}
