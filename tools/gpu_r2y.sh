TAG=${1:-r2y}
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
for i in 1 2 3; do
  for w in bv_n15 random_c1c2_n15; do
    $B --workload $w > gpurun_out/${TAG}_bench_${i}_$w.json 2> gpurun_out/${TAG}_bench_${i}_$w.err; echo "run $i $w rc=$?"
  done
done
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], "first %.1f"%d["e2e"]["first_call_ms"], "res %.1f"%d["e2e"]["resident_state"]["ms_per_step"], "cont %.1f"%d["continued_state"]["ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
