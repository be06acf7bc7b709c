TAG=${1:-r2ref}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/${TAG}_mem.txt; nproc >> gpurun_out/${TAG}_mem.txt
(time timeout 900 python bench.py --impl reference --cpu-full-size --steps 1 --warmup 0) > gpurun_out/${TAG}_bench_reference_full.json 2> gpurun_out/${TAG}_bench_reference_full.err
cat gpurun_out/${TAG}_mem.txt; tail -3 gpurun_out/${TAG}_bench_reference_full.err; cut -c1-600 gpurun_out/${TAG}_bench_reference_full.json
