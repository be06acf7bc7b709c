TAG=${1:-r2z}
mkdir -p gpurun_out
export DMB_JIT_CACHE=off DMB_SEGV_TRACE=1 DMB_JIT_VERBOSE=1
for i in 1 2; do
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --workload random_c1c2_n15 > gpurun_out/${TAG}_bench_${i}.json 2> gpurun_out/${TAG}_bench_${i}.err; echo "run $i rc=$?"
  tail -60 gpurun_out/${TAG}_bench_${i}.err
done
