TAG=${1:-r2k}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for w in qft_n15 hlayer_n15 random_c1c2_n15 bv_n15; do
  for d in 1 0; do
  DMB_DIRECT_STORE=$d $B --workload $w > gpurun_out/${TAG}_bench_ds${d}_$w.json 2> gpurun_out/${TAG}_bench_ds${d}_$w.err
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_qft_n15 python bench.py --workload qft_n15 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_full_qft_n15.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.2f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
