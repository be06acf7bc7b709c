TAG=${1:-r2x}
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-first-call"
for w in qft_n15 bv_n15 hlayer_n15 random_c1c2_n15; do
  for lf in 1 0; do
  DMB_LIGHT_FIRST=$lf $B --workload $w > gpurun_out/${TAG}_bench_lf${lf}_$w.json 2> gpurun_out/${TAG}_bench_lf${lf}_$w.err
  done
done
(time timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "sweeps", d["config"]["sweeps_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], d["jit"]["sweeps_specialised"], "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
