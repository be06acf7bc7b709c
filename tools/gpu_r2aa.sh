TAG=${1:-r2aa}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --jit 1 --workload qft_n15 > gpurun_out/${TAG}_bench_tiered.json 2> gpurun_out/${TAG}_bench_tiered.err; echo "rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --workload qft_n15 > gpurun_out/${TAG}_bench_wait.json 2> gpurun_out/${TAG}_bench_wait.err; echo "rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], "first %.1f"%d["e2e"]["first_call_ms"], "res %.1f"%d["e2e"]["resident_state"]["ms_per_step"], "cont %.1f"%d["continued_state"]["ms_per_step"], {k: d["jit"][k] for k in ("mode","sweeps_specialised","sweeps_pending","compiled_kernels","compile_ms_total")})
    except Exception as e:
        print(f, "FAILED", e)
PY
