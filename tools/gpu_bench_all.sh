TAG=${1:-rX}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
(time python bench.py) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
for w in bv_n15 vqe_uccsd_n8 adder_n10 random_c1c2_n15; do python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
