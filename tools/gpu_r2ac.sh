TAG=${1:-r2ac}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for w in vqe_uccsd_n8 adder_n10; do
$B --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], d["e2e"]["host_call_ms"], "first %.1f"%d["e2e"]["first_call_ms"], "warm %.2f"%d["e2e"]["repeated_circuit"]["ms_per_step"], {k: d["jit"][k] for k in ("mode","sweeps_specialised","sweeps_pending","compiled_kernels","compile_ms_total")}, "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
