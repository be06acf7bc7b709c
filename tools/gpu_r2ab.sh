TAG=${1:-r2ab}
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-first-call"
for w in vqe_uccsd_n8 adder_n10; do
DMB_JIT_MIN_BITS=0 $B --workload $w > gpurun_out/${TAG}_bench_jit_$w.json 2> gpurun_out/${TAG}_bench_jit_$w.err; echo "rc=$?"
$B --jit 0 --workload $w > gpurun_out/${TAG}_bench_nojit_$w.json 2> gpurun_out/${TAG}_bench_nojit_$w.err; echo "rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], "warm %.2f"%d["e2e"]["repeated_circuit"]["ms_per_step"], {k: d["jit"][k] for k in ("mode","sweeps_specialised","sweeps_pending","compiled_kernels","compile_ms_total")}, "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
