# round 2, call D (1 GPU): resident-CTA experiment on the FP64-bound workload, new default bench line.  Usage: tools/gpu_r2d.sh TAG
TAG=${1:-r2d}
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
for g in 2 3; do
  DMB_GRID_PER_SM=$g $B --workload random_c1c2_n15 > gpurun_out/${TAG}_bench_grid${g}_random_c1c2_n15.json 2> gpurun_out/${TAG}_bench_grid${g}_random_c1c2_n15.err
  DMB_GRID_PER_SM=$g $B --workload qft_n15 > gpurun_out/${TAG}_bench_grid${g}_qft_n15.json 2> gpurun_out/${TAG}_bench_grid${g}_qft_n15.err
done
(time timeout 900 python bench.py) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -c 600 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.2f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "hbm %.3f fp64 %.3f"%(d["roofline"]["frac_hbm"], d["roofline"]["frac_fp64"]), "e2e %.2f"%d["e2e"]["ms_per_step"])
        for x in d.get("extra_workloads", []): print("   ", json.dumps(x)[:400])
        if "parity" in d: print("   parity", json.dumps(d["parity"])[:600])
    except Exception as e:
        print(f, "FAILED", e)
PY
