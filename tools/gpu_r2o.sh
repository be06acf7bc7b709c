TAG=${1:-r2o}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -x -q) > gpurun_out/${TAG}_pytest_jit.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_jit.log
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for w in qft_n15 bv_n15 hlayer_n15 random_c1c2_n15; do
  for j in 2 0; do
  DMB_JIT=$j $B --workload $w > gpurun_out/${TAG}_bench_jit${j}_$w.json 2> gpurun_out/${TAG}_bench_jit${j}_$w.err
  done
done
$B --workload vqe_uccsd_n8 > gpurun_out/${TAG}_bench_vqe_uccsd_n8.json 2> gpurun_out/${TAG}_bench_vqe_uccsd_n8.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "launches", d["gpu_launches"], "e2e %.2f"%d["e2e"]["ms_per_step"], "warm %.2f"%d["e2e"]["repeated_circuit"]["ms_per_step"], "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
