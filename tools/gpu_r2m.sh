TAG=${1:-r2m}
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for w in hstride_n15 single_n15; do
  for d in 1 0; do
  DMB_DIRECT_STORE=$d $B --workload $w > gpurun_out/${TAG}_bench_ds${d}_$w.json 2> gpurun_out/${TAG}_bench_ds${d}_$w.err
  done
  DMB_TMA=0 $B --workload $w > gpurun_out/${TAG}_bench_notma_$w.json 2> gpurun_out/${TAG}_bench_notma_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.2f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "sweeps", d["config"]["sweeps_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
