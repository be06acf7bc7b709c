# one 1-GPU box call (round 2): GPU tests, smoke, the default bench line (with extras) and the reference arm as the driver
# runs them, a few workload lines, the ncu launch list and ncu --set full captures of the (specialised) sweep kernel.
# Usage: tools/gpu_round2.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
(time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
(time timeout 900 python bench.py) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
for w in hlayer_n15 single_n15; do timeout 300 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-extra > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_qft_n15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_launches.log 2>&1
for W in qft_n15 random_c1c2_n15; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmb_jit_sweep -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_$W python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-first-call > gpurun_out/${TAG}_ncu_full_$W.log 2>&1
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        if d.get("impl") == "reference": print("reference", d["value"], d["cpu_baseline"]); continue
        print(d["config"]["workload"], "ms/step %.3f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], "first", d["e2e"]["first_call_ms"], "cont", d["continued_state"]["ms_per_step"], {k: d["jit"][k] for k in d["jit"] if k != "what"})
        for x in d.get("extra_workloads", []): print("   extra", x.get("workload"), x.get("ms_per_step"), x.get("roofline", {}).get("frac"), x.get("failed"))
        if "parity" in d: print("   parity", d["parity"]["max_abs_err"], d["parity"]["ok"])
    except Exception as e:
        print(f, "FAILED", e)
PY
